"""GPU tests of the actor-facing edge: the three Replay ingest threads against a Redis stand-in with REAL
list semantics (tests/fake_redis.py) and the three Learner.run loops' publication / reward drain / checkpoint
cadence (APE_X/Learner.py:140-262, R2D2/Learner.py:217-339, IMPALA/Learner.py:274-297)."""
import os
import pickle
import time

import numpy as np
import pytest

from fake_redis import FakeRedis
from test_wire_cpu import _r2d2_record

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _wait(cond, timeout=20.0):
    t0 = time.time()
    while not cond():
        if time.time() - t0 > timeout:
            return False
        time.sleep(0.005)
    return True


def _apex_rec(rng, prio):
    return [rng.integers(0, 256, (4, 84, 84), dtype=np.uint8), int(rng.integers(6)), float(rng.standard_normal()),
            rng.integers(0, 256, (4, 84, 84), dtype=np.uint8), bool(rng.random() < 0.3), float(prio)]


def test_apex_ingest_thread_reads_each_record_once_and_stops():
    """A lone pending record (actors RPUSH one at a time, APE_X/Player.py:258) must land exactly once: the
    reference's LTRIM -1 0 leaves a 1-element list in place.  Also: the thread can be stopped and joined
    (`_stop` would shadow threading.Thread's own method)."""
    from distributed_rl_b200 import apex
    conn = FakeRedis()
    cfg = apex.ApexConfig(BATCHSIZE=4, REPLAY_MEMORY_LEN=64, BUFFER_SIZE=2, LEARNER_DEVICE="cuda:0")
    mem = apex.Replay(cfg, conn)
    mem.start()
    rng = np.random.default_rng(0)
    recs = [_apex_rec(rng, 0.5 + i) for i in range(6)]
    conn.rpush("experience", pickle.dumps(recs[0]))
    assert _wait(lambda: len(mem.store) == 1)
    time.sleep(0.1)                                   # ~50 more polls: the record must not be re-read
    assert len(mem.store) == 1 and mem.total_frame == 1
    for r in recs[1:]:
        conn.rpush("experience", pickle.dumps(r))
    assert _wait(lambda: mem.total_frame == 6)
    time.sleep(0.05)
    assert len(mem.store) == 6 and conn.llen("experience") == 0
    torch.cuda.synchronize()
    np.testing.assert_array_equal(mem.store.field_view("state")[:6].cpu().numpy(), np.stack([r[0] for r in recs]))
    np.testing.assert_allclose(mem.store.priorities(0, 6).cpu().numpy(), [r[5] for r in recs], rtol=0)
    assert mem.cond is True                           # > BUFFER_SIZE
    batch = mem.sample()
    assert batch is not False and batch[0].shape == (4, 4, 84, 84)
    mem.lock = True                                   # eviction handshake: served by the ingest thread
    assert _wait(lambda: mem.lock is False)
    mem.stop()
    mem.join(timeout=5)
    assert not mem.is_alive()


def test_r2d2_ingest_thread_and_run_loop(tmp_path):
    from distributed_rl_b200 import r2d2
    conn = FakeRedis()
    conn.set("Start", b"stale"); conn.rpush("reward", pickle.dumps(123.0))      # leftovers of a previous run
    T, MEM, B = 8, 4, 2
    cfg = r2d2.R2D2Config(BATCHSIZE=B, FIXED_TRAJECTORY=T, MEM=MEM, UNROLL_STEP=2, REPLAY_MEMORY_LEN=16, BUFFER_SIZE=3,
                          TARGET_FREQUENCY=20, LOG_W=str(tmp_path / "weight"))
    torch.manual_seed(0)
    L = r2d2.Learner(cfg, connect=conn)
    assert conn.get("Start") is None and conn.llen("reward") == 0                # R2D2/Learner.py:54,63-64
    assert L.memory.is_alive()
    rng = np.random.default_rng(3)
    recs = [_r2d2_record(rng, T, bool(i % 3 == 0)) for i in range(6)]
    conn.rpush("experience", pickle.dumps(recs[0]))
    assert _wait(lambda: len(L.memory.store) == 1)
    time.sleep(0.05)
    assert len(L.memory.store) == 1
    for r in recs[1:]:
        conn.rpush("experience", pickle.dumps(r))
    assert _wait(lambda: L.memory.total_frame == 6)
    torch.cuda.synchronize()
    st = L.memory.store
    for i, r in enumerate(recs):
        np.testing.assert_array_equal(st.field_view("state")[i, 3].cpu().numpy(), r[1 + 9])
        assert float(st.field_view("notdone")[i]) == float(not r[-2])
    np.testing.assert_allclose(st.priorities(0, 6).cpu().numpy(), np.asarray([r[-1] for r in recs], np.float32))
    for v in (2.0, 4.0):
        conn.rpush("reward", pickle.dumps(v))
    before = [p.detach().clone() for p in L.model.parameters()]
    assert L.run(max_steps=50, log_every=25) == 50
    for p in L._publishers:
        p.poll(block=True)
    assert pickle.loads(conn.get("Start")) is True
    sd = pickle.loads(conn.get("state_dict"))
    assert list(sd) == list(L.model.state_dict()) and all(not v.is_cuda for v in sd.values())
    assert pickle.loads(conn.get("count")) in (1, -25, 0)                        # step - 50 (sic, :293)
    assert set(pickle.loads(conn.get("target_state_dict"))) == set(sd)
    assert conn.llen("reward") == 0 and L.last_log["step"] == 50                # drained at step 25 (mean 3.0) and 50
    assert any((a != b).any().item() for a, b in zip(before, L.model.parameters()))
    ck = torch.load(os.path.join(cfg.LOG_W, "weight.pth"))
    assert set(ck) == set(sd)
    L.memory.stop(); L.memory.join(timeout=5)
    assert not L.memory.is_alive()


def test_impala_ingest_thread_and_run_loop(tmp_path):
    from distributed_rl_b200 import impala
    conn = FakeRedis()
    T, B = 4, 2
    cfg = impala.ImpalaConfig(BATCHSIZE=B, UNROLL_STEP=T, REPLAY_MEMORY_LEN=8, BUFFER_SIZE=3,
                              LOG_W=str(tmp_path / "weight"))
    torch.manual_seed(0)
    L = impala.Learner(cfg, connect=conn)
    assert L._memory.is_alive()
    rng = np.random.default_rng(4)
    recs = [[rng.integers(0, 256, (T + 1, 28224), dtype=np.uint8), rng.integers(0, 6, (T, 1)),
             rng.uniform(0.05, 0.9, (T, 1)).astype(np.float32), rng.standard_normal(T), int(i % 2)] for i in range(5)]
    conn.rpush("trajectory", pickle.dumps(recs[0]))
    assert _wait(lambda: len(L._memory) == 1)
    time.sleep(0.05)
    assert len(L._memory) == 1                                                   # lone rollout read once
    for r in recs[1:]:
        conn.rpush("trajectory", pickle.dumps(r))
    assert _wait(lambda: len(L._memory) == 5)
    torch.cuda.synchronize()
    st = L._memory.store
    np.testing.assert_array_equal(st.field_view("state")[:5].cpu().numpy(), np.stack([r[0] for r in recs]))
    np.testing.assert_array_equal(st.field_view("action")[:5].cpu().numpy(), np.stack([r[1][:, 0] for r in recs]))
    np.testing.assert_array_equal(st.field_view("done")[:5].cpu().numpy().ravel(), [float(r[4]) for r in recs])
    # uniform sampling WITHOUT replacement (random.sample, baseline/utils.py:310-315)
    L._memory.bufferSave(2)
    seen = torch.cat([b[4] for b in L._memory.deque])
    assert len(L._memory.deque) == 2 and seen.numel() == 4
    L._memory.deque.clear()
    with pytest.raises(ValueError):
        L._memory.bufferSave(3)                                                  # 6 > 5 rollouts
    assert L.run(max_steps=101) == 101
    for p in L._publishers:
        p.poll(block=True)
    params = pickle.loads(conn.get("params"))
    assert isinstance(params, tuple) and len(params) == 1                        # IMPALA/Learner.py:268-272
    assert list(params[0]) == list(L.model.state_dict())
    assert 0 <= pickle.loads(conn.get("Count")) <= 100
    assert os.path.isfile(os.path.join(cfg.LOG_W, "weight.pth"))                 # :290-297, every 100 steps
    L._memory.stop(); L._memory.join(timeout=5)


def test_apex_run_loop_control_plane(tmp_path):
    """Stale keys wiped at start (:41-43), `reward` drained and logged every log_every steps (:219-253), a
    checkpoint written (:256-262), eviction request served when no ingest thread runs."""
    from distributed_rl_b200 import apex
    conn = FakeRedis()
    conn.set("Start", b"stale"); conn.rpush("experience", b"junk-from-a-previous-run")
    cfg = apex.ApexConfig(BATCHSIZE=16, REPLAY_MEMORY_LEN=1024, BUFFER_SIZE=0, TARGET_FREQUENCY=30,
                          LEARNER_DEVICE="cuda:0", CUDNN_BENCHMARK=False, LOG_W=str(tmp_path / "w"))
    L = apex.Learner(cfg, connect=conn, start_replay=False)
    assert conn.get("Start") is None and conn.llen("experience") == 0
    st = L.memory.store
    st.fill_hash(1024, seed=1)
    st.build(torch.rand(1024, device="cuda") + 0.1)
    for v in (-3.0, 5.0):
        conn.rpush("reward", pickle.dumps(v))
    assert L.run(max_steps=60, log_every=20) == 60
    for p in L._publishers:
        p.poll(block=True)
    assert conn.llen("reward") == 0
    assert L.last_log["step"] == 60 and L.last_log["reward"] == -21.0            # drained at step 20 (mean 1.0)
    assert np.isfinite(L.last_log["norm"]) and L.last_log["norm"] > 0
    assert L.memory.lock is False
    assert set(torch.load(os.path.join(cfg.LOG_W, "weight.pth"))) == set(L.model.state_dict())
    assert pickle.loads(conn.get("count")) in (1, 0)


def test_standalone_replay_server_round_trip():
    """ReplayServer (APE_X/ReplayServer.py) serving pickled minibatches over the reference's Redis keys,
    Replay_Server (APE_X/ReplayMemory.py:170-257) consuming them, priorities flowing back through `update`,
    FLAG_BATCH / FLAG_ENOUGH / FLAG_REMOVE honoured — one process, two connections' worth of keys in one stand-in."""
    from distributed_rl_b200 import apex
    from distributed_rl_b200.replay_server import ReplayServer, Replay_Server
    conn = FakeRedis()
    cfg = apex.ApexConfig(BATCHSIZE=4, REPLAY_MEMORY_LEN=64, BUFFER_SIZE=8, LEARNER_DEVICE="cuda:0")
    srv = ReplayServer(cfg, conn, conn, m=3)
    cli = Replay_Server(cfg, conn, conn)
    assert pickle.loads(conn.get("FLAG_BATCH")) is False
    rng = np.random.default_rng(0)
    recs = [_apex_rec(rng, 0.5 + 0.01 * i) for i in range(20)]
    for r in recs[:6]:
        conn.rpush("experience", pickle.dumps(r))
    st = srv.serve_once()
    assert st["ingested"] == 6 and st["batches_queued"] == 0 and conn.llen("BATCH") == 0    # below BUFFER_SIZE
    for r in recs[6:]:
        conn.rpush("experience", pickle.dumps(r))
    st = srv.serve_once()
    assert st["ingested"] == 14 and conn.llen("BATCH") == 3
    st = srv.serve_once()
    assert pickle.loads(conn.get("FLAG_BATCH")) is True
    cli.poll_once()
    assert conn.llen("BATCH") == 0 and len(cli.deque) == 3 and pickle.loads(conn.get("FLAG_ENOUGH")) is False
    batch = cli.sample()
    s, a, r, ns, d, w, idx = batch
    assert s.shape == (4, 4, 84, 84) and ns.shape == (4, 4, 84, 84) and len(a) == 4 and w.shape == (4,)
    ii = idx.numpy()
    np.testing.assert_array_equal(s, np.stack([recs[i][0] for i in ii]))           # the served rows are the records
    np.testing.assert_array_equal(a, [recs[i][1] for i in ii])
    # a learner can train on it as is (host arrays, reference signature) and send priorities back
    L = apex.Learner(cfg, connect=None, start_replay=False)
    info, prio, idx2, _ = L.train(batch)
    cli.update(list(idx2), prio.cpu().numpy())
    cli.idx += [0] * 1000; cli.vals.append(np.full(1000, 0.25, np.float32))         # force the > 1000 flush
    cli.poll_once()
    assert conn.llen("update") == 1
    assert srv.serve_once()["updates_applied"] == 1004
    torch.cuda.synchronize()
    pr = srv.store.priorities(0, 20).cpu().numpy()
    assert pr[0] == np.float32(0.25)
    for i, p in zip(idx2.numpy()[::-1], prio.cpu().numpy()[::-1]):
        if i != 0:
            assert pr[i] == p
            break
    assert cli.sample() is not False and cli.sample() is not False and cli.sample() is False
    cli.lock = True; cli.poll_once()
    assert pickle.loads(conn.get("FLAG_REMOVE")) is True
