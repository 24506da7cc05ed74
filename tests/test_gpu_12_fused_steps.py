"""GPU tests of the resident learner steps of R2D2 and IMPALA (`fused_step`: sample -> conv_1 over the sampled
sequences' frames read IN PLACE in the replay payload -> ... -> write-back) against the reference-signature
`train` on the same minibatch staged the way R2D2/ReplayMemory.py:53-122 / IMPALA/ReplayMemory.py:30-54 build it."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.mark.parametrize("pool", [0, 16])
def test_r2d2_fused_step_equals_train_on_the_staged_batch(pool):
    """Same weights, same sum-tree, same device RNG stream: fused_step() (frames read in place, row = slot_row*T + t;
    with PAYLOAD_POOL the slot's row is slot % pool) and sample() -> train() -> update() give the same priorities,
    the same tree and the same parameter update."""
    from distributed_rl_b200 import r2d2
    T, MEM, B, N = 12, 4, 4, 64
    res = []
    for fused in (True, False):
        cfg = r2d2.R2D2Config(BATCHSIZE=B, FIXED_TRAJECTORY=T, MEM=MEM, UNROLL_STEP=3, REPLAY_MEMORY_LEN=N,
                              BUFFER_SIZE=0, PAYLOAD_POOL=pool)
        torch.manual_seed(0)
        L = r2d2.Learner(cfg)
        with torch.no_grad():
            for p in L.target_model.parameters():
                p.add_(0.02 * torch.randn(p.shape, device=p.device))
        init = [p.detach().clone() for p in L.model.parameters()]
        rng = np.random.default_rng(5)
        n = pool or N
        cols = [rng.integers(0, 256, size=(n, T, 4, 84, 84), dtype=np.uint8),
                rng.integers(0, 6, size=(n, T)).astype(np.int32), rng.standard_normal((n, T)).astype(np.float32),
                (rng.standard_normal((n, 512)) * 0.1).astype(np.float32),
                (rng.standard_normal((n, 512)) * 0.1).astype(np.float32), (rng.random(n) > 0.3).astype(np.float32)]
        L.memory.pool.push(cols, np.ones(n, np.float32))
        L.memory.store.build(torch.from_numpy(rng.uniform(0.1, 1, N).astype(np.float32)).cuda())
        L.memory.store.seed(77, 0)
        if fused:
            out = L.fused_step()
            prio, idx = out["prio"], out["idx"]
        else:
            batch = L.memory.sample()
            info, prio, idx = L.train(batch)
            L.memory.update(idx, prio)
        torch.cuda.synchronize()
        res.append((idx.clone(), prio.clone(), L.memory.store.priorities().clone(),
                    [p.detach().clone() for p in L.model.parameters()], init))
    (i0, p0, t0, w0, init), (i1, p1, t1, w1, _) = res
    assert torch.equal(i0, i1)
    np.testing.assert_allclose(p0.cpu().numpy(), p1.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(t0.cpu().numpy(), t1.cpu().numpy(), rtol=1e-5, atol=1e-7)
    for a, b, w in zip(w0, w1, init):
        assert _rel(a - w, b - w) <= 1e-3


def test_impala_fused_step_equals_train_on_the_staged_batch():
    from distributed_rl_b200 import impala
    T, B, n = 6, 4, 16
    res = []
    for fused in (True, False):
        cfg = impala.ImpalaConfig(BATCHSIZE=B, UNROLL_STEP=T, REPLAY_MEMORY_LEN=n, BUFFER_SIZE=0)
        torch.manual_seed(1)
        L = impala.Learner(cfg)
        init = [p.detach().clone() for p in L.model.parameters()]
        rng = np.random.default_rng(2)
        L._memory.push_arrays(rng.integers(0, 256, size=(n, T + 1, 28224), dtype=np.uint8),
                              rng.integers(0, 6, size=(n, T)).astype(np.int32),
                              rng.uniform(0.05, 0.9, size=(n, T)).astype(np.float32),
                              rng.standard_normal((n, T)).astype(np.float32), (rng.random(n) > 0.3).astype(np.float32))
        L._memory._rng.manual_seed(9)
        if fused:
            L.fused_step()
        else:
            L.train(L._memory.sample(), 0)
        torch.cuda.synchronize()
        res.append((L.last["vtarget"].clone(), L.last["advantage"].clone(), L.last["criticLoss"].clone(),
                    [p.detach().clone() for p in L.model.parameters()], init))
    (v0, a0, c0, w0, init), (v1, a1, c1, w1, _) = res
    np.testing.assert_allclose(v0.cpu().numpy(), v1.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a0.cpu().numpy(), a1.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(c0), float(c1), rtol=1e-5)
    for a, b, w in zip(w0, w1, init):
        assert _rel(a - w, b - w) <= 1e-3


def test_impala_draw_skips_slots_reserved_by_an_ingest_in_flight():
    """Uniform sampling WITHOUT replacement over the ring's valid region only (random.sample, baseline/utils.py:310-315)."""
    from distributed_rl_b200 import impala
    T, n = 2, 32
    cfg = impala.ImpalaConfig(BATCHSIZE=4, UNROLL_STEP=T, REPLAY_MEMORY_LEN=n, BUFFER_SIZE=0)
    mem = impala.Replay(cfg)
    rng = np.random.default_rng(0)
    def cols(k):
        return [torch.from_numpy(rng.integers(0, 256, size=(k, T + 1, 28224), dtype=np.uint8)).pin_memory(),
                torch.zeros(k, T, dtype=torch.int32).pin_memory(), torch.full((k, T), 0.5).pin_memory(),
                torch.zeros(k, T).pin_memory(), torch.ones(k).pin_memory()]
    mem.push_arrays(*cols(n))                       # ring full, head back at 0
    mem.push_arrays(*cols(5))                       # head = 5
    mem.store.push_begin(cols(8), 8)                # slots 5..12 reserved, copy in flight
    seen = set()
    for _ in range(20):
        idx = mem.draw(24).cpu().numpy()
        assert len(set(idx)) == 24                  # no replacement
        seen |= set(idx)
    assert seen == set(range(32)) - set(range(5, 13))
    with pytest.raises(ValueError):
        mem.draw(25)                                # only 24 kept rollouts are sampleable right now
    mem.store.push_commit(torch.ones(8))
    assert len(set(mem.draw(32).cpu().numpy())) == 32
