"""Host-side logic of the fused-head paths (no GPU): graph pattern matching in GraphAgent, CPU fall-through of the
fused flags, the conv split-backward function without a sink, hostmem's topology parsing."""
import os

import pytest

torch = pytest.importorskip("torch")


def _agents():
    from distributed_rl_b200.agent import GraphAgent
    from distributed_rl_b200.apex import default_apex_model
    from distributed_rl_b200.r2d2 import default_r2d2_model
    from distributed_rl_b200.impala import default_impala_model
    return GraphAgent, default_apex_model, default_r2d2_model, default_impala_model


def test_dueling_pattern_is_found_in_apex_and_r2d2_not_in_impala():
    GraphAgent, apex_m, r2d2_m, impala_m = _agents()
    a = GraphAgent(apex_m())
    (g, d), = a._dueling.items()
    assert (d["adv"], d["val"], d["out"]) == ("module02", "module02_1", "module04")
    assert set(d["inner"]) == {"module02", "module02_1", "module03", "module03_1"}
    r = GraphAgent(r2d2_m())
    (g, d), = r._dueling.items()
    assert (d["adv"], d["val"], d["out"]) == ("module03", "module03_1", "module05")
    assert GraphAgent(impala_m())._dueling == {}


def test_pattern_rejects_lookalikes():
    GraphAgent, apex_m, _, _ = _agents()
    m = apex_m()
    m["module03_1"]["prevNodeNames"] = ["module02_1"]          # mean over the value head: not the dueling combine
    assert GraphAgent(m)._dueling == {}
    m = apex_m()
    m["module02"]["output"] = True                             # an inner node is also a network output
    assert GraphAgent(m)._dueling == {}
    m = apex_m()
    m["module02_1"]["fSize"] = [512, 2]                        # "value" head with two outputs
    assert GraphAgent(m)._dueling == {}


def test_fused_flags_fall_through_on_cpu_and_split_backward_matches_autograd():
    GraphAgent, apex_m, _, _ = _agents()
    torch.manual_seed(0)
    m = GraphAgent(apex_m())
    x = torch.rand(3, 4, 84, 84)
    ref = m([x])[0]
    m.fused_dueling_tail = m.dense_3xtf32 = True               # CUDA-only paths: must not be taken for CPU tensors
    with m.packed_heads_cache():
        m.prepack_heads()
        out = m([x])[0]
    assert torch.equal(out, ref)
    # _ConvSplitBackward without an active sink == plain conv backward
    from distributed_rl_b200.agent import _ConvSplitBackward
    w = torch.randn(8, 4, 3, 3, requires_grad=True)
    xi = torch.randn(2, 4, 9, 9, requires_grad=True)
    gy = torch.randn(2, 8, 4, 4)
    _ConvSplitBackward.apply(xi, w, (2, 2), (0, 0)).backward(gy)
    gx, gw = xi.grad.clone(), w.grad.clone()
    xi.grad = w.grad = None
    torch.nn.functional.conv2d(xi, w, None, (2, 2), (0, 0)).backward(gy)
    torch.testing.assert_close(gx, xi.grad)
    torch.testing.assert_close(gw, w.grad)


def test_hostmem_cpulist_parsing_and_unbound_fallback():
    from distributed_rl_b200 import hostmem
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()
    assert hostmem.gpu_node_cpus("cuda:0") is None or isinstance(hostmem.gpu_node_cpus("cuda:0"), set)
    before = os.sched_getaffinity(0)
    with hostmem.on_gpu_node("cuda:0") as bound:               # no GPU here: must be a no-op
        assert bound in (False, True)
    assert os.sched_getaffinity(0) == before


def test_flat_grads_lay_sibling_head_gradients_back_to_back():
    """optim.flat_grads pre-allocates every .grad in one buffer, largest first, so that the two dense heads' first-layer
    gradients (cfg/ape_x.json:52-71) are consecutive row blocks: linear._stacked_rows then hands the weight-gradient
    GEMM ONE (1024, 3136) output view.  Reversed or non-adjacent tensors must be refused (the caller falls back to
    temporaries + add)."""
    import torch
    from distributed_rl_b200.agent import GraphAgent
    from distributed_rl_b200.apex import default_apex_model
    from distributed_rl_b200.linear import _stacked_rows
    from distributed_rl_b200.optim import flat_grads
    m = GraphAgent(default_apex_model())
    flat = flat_grads(m.getParameters())
    assert all(p.grad is not None and p.grad.shape == p.shape and p.grad.stride() == p.stride() for p in m.parameters())
    assert sum(p.numel() for p in m.parameters()) <= flat.numel() < sum(p.numel() for p in m.parameters()) + 4
    wa, wv = m.module02.MLP_1.weight, m.module02_1.MLP_1.weight
    joint = _stacked_rows([wa.grad, wv.grad])
    assert joint is not None and joint.shape == (1024, 3136)
    joint[:512].fill_(1.0)
    joint[512:].fill_(2.0)
    assert float(wa.grad.min()) == 1.0 == float(wa.grad.max()) and float(wv.grad.min()) == 2.0 == float(wv.grad.max())
    others = [p for p in m.parameters() if p is not wa and p is not wv]
    assert all(float(p.grad.abs().max()) == 0.0 for p in others)            # nothing else was touched
    assert _stacked_rows([wv.grad, wa.grad]) is None                         # wrong order
    assert _stacked_rows([wa.grad, m.module02.MLP_2.weight.grad]) is None    # different widths
    assert _stacked_rows([torch.zeros(4, 8), torch.zeros(4, 8)]) is None     # separate allocations
