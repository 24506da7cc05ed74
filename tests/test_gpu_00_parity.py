"""GPU parity tests: CUDA path (through the C ABI) vs the numpy oracle and the
golden vectors recorded from the reference.  Integer / index work is bit-exact;
fp32 outputs are compared to the tolerance written beside each assert (contract:
1e-5, BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def R():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from distributed_rl_b200 import replay
    return replay


def _dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _pow2ceil(n):
    return 1 << max(0, (n - 1).bit_length())


def _rand_prios(rng, n, alpha=0.6):
    return ((np.abs(rng.standard_normal(n)).clip(max=1) + 1e-7) ** alpha).astype(np.float32)


# --------------------------------------------------------------------------- #
# sum-tree                                                                      #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("n", [1, 2, 5, 777, 2048, 3000, 4096, 65536, 100000])
def test_tree_build_sample_matches_oracle_bit_exact(R, n):
    rng = np.random.default_rng(n)
    p = _rand_prios(rng, n)
    u = rng.random(512)
    rep = R.DeviceReplay(n, fields=())
    rep.build(_dev(p))
    idx, prob, w = rep.sample(512, beta=0.4, u01=_dev(u))
    t = O.SumTreeOracle(_pow2ceil(n)); t.build(p)
    oidx, _ = t.sample(u)
    ow, oprob, omaxw = O.is_weights(p[oidx], t.total, t.min_priority, n, 0.4)
    np.testing.assert_array_equal(idx.cpu().numpy(), oidx)           # indices: bit-exact
    np.testing.assert_array_equal(prob.cpu().numpy(), oprob)         # fp32 division: bit-exact
    np.testing.assert_allclose(w.cpu().numpy(), ow, rtol=2.4e-7)     # <= 2 ulp (fp64 pow, rounded once)
    st = rep.stats(0.4).cpu().numpy()
    assert st[0] == t.total                                           # fp64 root: bit-exact
    assert np.float32(st[1]) == t.min_priority
    np.testing.assert_allclose(st[2], omaxw, rtol=2.4e-7)
    np.testing.assert_array_equal(rep.priorities(0, n).cpu().numpy(), p)
    rep.close()


@pytest.mark.parametrize("tag", ["pow2", "ragged", "tiny", "one"])
def test_tree_sample_matches_reference_sumtree_golden(R, golden, tag):
    """vs baseline/sumtree.py SumTree.prioritized_sample outputs (arbitrary priorities)."""
    g = golden("tree")
    p = g[f"st_{tag}_prios"]
    rep = R.DeviceReplay(len(p), fields=())
    rep.build(_dev(p))
    idx, _, _ = rep.sample(len(g[f"st_{tag}_u01"]), u01=_dev(g[f"st_{tag}_u01"]))
    np.testing.assert_array_equal(idx.cpu().numpy(), g[f"st_{tag}_idx"])
    assert rep.stats().cpu().numpy()[0] == float(g[f"st_{tag}_total"])
    if f"st_{tag}_upd_idx" in g:
        rep.update(_dev(g[f"st_{tag}_upd_idx"]), _dev(g[f"st_{tag}_upd_val"]))
        np.testing.assert_array_equal(rep.priorities(0, len(p)).cpu().numpy().astype(np.float64),
                                      g[f"st_{tag}_leaves_after"])
        assert rep.stats().cpu().numpy()[0] == float(g[f"st_{tag}_total_after"])
        idx2, _, _ = rep.sample(len(g[f"st_{tag}_u01_after"]), u01=_dev(g[f"st_{tag}_u01_after"]))
        np.testing.assert_array_equal(idx2.cpu().numpy(), g[f"st_{tag}_idx_after"])
    rep.close()


@pytest.mark.parametrize("tag", ["4k", "64k"])
def test_per_sample_matches_reference_per_golden(R, golden, tag):
    """vs baseline/PER.py PER.sample + APE_X/ReplayMemory.py:65-67 (dyadic priorities)."""
    g = golden("tree")
    p = g[f"per_{tag}_prios"]
    rep = R.DeviceReplay(len(p), fields=())
    rep.build(_dev(p))
    idx, prob, w = rep.sample(512, beta=0.4, u01=_dev(g[f"per_{tag}_u01"]))
    np.testing.assert_array_equal(idx.cpu().numpy(), g[f"per_{tag}_idx"])       # bit-exact
    np.testing.assert_array_equal(prob.cpu().numpy(), g[f"per_{tag}_prob"])     # bit-exact
    np.testing.assert_allclose(w.cpu().numpy(), g[f"per_{tag}_weight"], rtol=5e-7)  # torch powf is 1-ulp
    np.testing.assert_allclose(rep.stats(0.4).cpu().numpy()[2], g[f"per_{tag}_max_weight"], rtol=3e-7)
    rep.update(_dev(g[f"per_{tag}_upd_idx"]), _dev(g[f"per_{tag}_upd_val"]))
    np.testing.assert_array_equal(rep.priorities(0, len(p)).cpu().numpy(), g[f"per_{tag}_prios_after"])
    rep.close()


@pytest.mark.parametrize("n,nupd", [(5, 64), (3000, 1000), (65536, 8192), (65536, 1), (1 << 18, 100000)])
def test_tree_update_duplicates_last_writer_wins(R, n, nupd):
    rng = np.random.default_rng(n + nupd)
    p = _rand_prios(rng, n)
    rep = R.DeviceReplay(n, fields=())
    rep.build(_dev(p))
    t = O.SumTreeOracle(_pow2ceil(n)); t.build(p)
    for rnd in range(3):   # repeated: tag/mark scratch must self-clean
        ui = rng.integers(0, n, size=nupd)
        if nupd > 8:
            ui[-4:] = ui[0]
        uv = _rand_prios(rng, nupd)
        rep.update(_dev(ui), _dev(uv))
        t.update(ui, uv)
        u = rng.random(256)
        idx, _, _ = rep.sample(256, u01=_dev(u))
        np.testing.assert_array_equal(idx.cpu().numpy(), t.sample(u)[0])
        st = rep.stats().cpu().numpy()
        assert st[0] == t.total and np.float32(st[1]) == t.min_priority
    np.testing.assert_array_equal(rep.priorities(0, n).cpu().numpy().astype(np.float64), t.leaves())
    rep.close()


def test_ring_push_evict_matches_model(R):
    """PER.push / remove_to_fit semantics on a ring with stable slot ids."""
    cap = 1000
    rng = np.random.default_rng(5)
    fields = (R.Field("x", torch.uint8, (48,)), R.Field("a", torch.int32, ()))
    rep = R.DeviceReplay(cap, fields=fields)
    model = O.RingModel(cap)
    store = np.zeros((cap, 48), np.uint8); act = np.zeros(cap, np.int32)
    for step, n in enumerate([300, 500, 400, 1000, 7]):
        x = rng.integers(0, 256, size=(n, 48), dtype=np.uint8)
        a = rng.integers(0, 6, size=n).astype(np.int32)
        p = _rand_prios(rng, n)
        host = step % 2 == 0
        rep.push([torch.from_numpy(x) if host else _dev(x), torch.from_numpy(a) if host else _dev(a)],
                 torch.from_numpy(p) if host else _dev(p))
        slots = model.push(p)
        store[slots] = x; act[slots] = a
        assert len(rep) == model.size and rep.head == model.head
        np.testing.assert_array_equal(rep.priorities().cpu().numpy(), model.prios)
        if step == 2:
            rep.evict(250); model.evict(250)
            assert len(rep) == model.size
            np.testing.assert_array_equal(rep.priorities().cpu().numpy(), model.prios)
        t = O.SumTreeOracle(1024); t.build(model.prios); t.size = model.size
        u = rng.random(128)
        idx, prob, w = rep.sample(128, beta=0.4, u01=_dev(u))
        oidx, _ = t.sample(u)
        np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
        ow, _, _ = O.is_weights(model.prios[oidx], t.total, t.min_priority, model.size, 0.4)
        np.testing.assert_allclose(w.cpu().numpy(), ow, rtol=2.4e-7)
        out = rep.gather(idx)
        np.testing.assert_array_equal(out["x"].cpu().numpy(), store[oidx])
        np.testing.assert_array_equal(out["a"].cpu().numpy(), act[oidx])
    rep.close()


def test_pipelined_ingest_matches_ring_model(R):
    """reserve / copy (ingest stream) / commit == PER.push on the ring; reserved slots are unsampleable."""
    cap, n = 512, 96
    rng = np.random.default_rng(8)
    fields = (R.Field("x", torch.uint8, (64,)), R.Field("a", torch.int32, ()))
    rep = R.DeviceReplay(cap, fields=fields)
    model = O.RingModel(cap)
    store = np.zeros((cap, 64), np.uint8)
    for step in range(9):      # wraps the ring once
        x = torch.from_numpy(rng.integers(0, 256, size=(n, 64), dtype=np.uint8)).pin_memory()
        a = torch.from_numpy(rng.integers(0, 6, size=n).astype(np.int32)).pin_memory()
        p = _rand_prios(rng, n)
        rep.push_begin([x, a], n)
        slots = (model.head + np.arange(n)) % cap
        reserved = rep.priorities().cpu().numpy()
        assert (reserved[slots] == 0).all()                      # retired before the copy may land
        if model.size:
            idx, _, _ = rep.sample(256)
            assert not np.isin(idx.cpu().numpy(), slots).any()   # never sampled while being overwritten
        rep.push_commit(torch.from_numpy(p))
        model.push(p); store[slots] = x.numpy()
        assert len(rep) == model.size and rep.head == model.head
        np.testing.assert_array_equal(rep.priorities().cpu().numpy(), model.prios)
        idx, _, _ = rep.sample(128)
        np.testing.assert_array_equal(rep.gather(idx)["x"].cpu().numpy(), store[idx.cpu().numpy()])
    rep.close()


def test_single_call_pipelined_ingest_matches_ring_model(R):
    """b2rl_replay_ingest_pipelined: call k publishes batch k-1 and starts the copy of batch k; the ring, the
    priorities and the payload equal PER.push applied one call later; slots in flight are never sampled."""
    cap, n = 512, 96
    rng = np.random.default_rng(18)
    fields = (R.Field("x", torch.uint8, (64,)), R.Field("a", torch.int32, ()))
    rep = R.DeviceReplay(cap, fields=fields)
    model = O.RingModel(cap)
    store = np.zeros((cap, 64), np.uint8)
    bufs = [(torch.empty(n, 64, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory(),
             torch.empty(n, dtype=torch.float32).pin_memory()) for _ in range(2)]
    pending = None
    for step in range(11):     # wraps the ring twice
        x, a, p = bufs[step & 1]
        x.copy_(torch.from_numpy(rng.integers(0, 256, size=(n, 64), dtype=np.uint8)))
        a.copy_(torch.from_numpy(rng.integers(0, 6, size=n).astype(np.int32)))
        p.copy_(torch.from_numpy(_rand_prios(rng, n)))
        rep.ingest_pipelined([x, a], p)
        if pending is not None:                 # the previous batch is published by this call
            slots_prev, x_prev, p_prev = pending
            model.push(p_prev); store[slots_prev] = x_prev
        slots = (model.head + np.arange(n)) % cap
        pr = rep.priorities().cpu().numpy()
        assert (pr[slots] == 0).all()                                  # retired before the copy may land
        keep = np.setdiff1d(np.arange(cap), slots)
        np.testing.assert_array_equal(pr[keep], model.prios[keep])
        assert len(rep) == min(model.size, cap - n) and rep.head == model.head
        if model.size:
            idx, _, _ = rep.sample(256)
            assert not np.isin(idx.cpu().numpy(), slots).any()         # never sampled while being overwritten
            np.testing.assert_array_equal(rep.gather(idx)["x"].cpu().numpy(), store[idx.cpu().numpy()])
        pending = (slots, x.numpy().copy(), p.numpy().copy())
    rep.ingest_pipelined(None)                  # flush: publish the last batch
    slots_prev, x_prev, p_prev = pending
    model.push(p_prev); store[slots_prev] = x_prev
    assert len(rep) == model.size and rep.head == model.head
    np.testing.assert_array_equal(rep.priorities().cpu().numpy(), model.prios)
    idx, _, _ = rep.sample(512)
    np.testing.assert_array_equal(rep.gather(idx)["x"].cpu().numpy(), store[idx.cpu().numpy()])
    rep.close()


def test_empty_replay_sampling_is_an_error(R):
    from distributed_rl_b200._lib import B2RLError
    rep = R.DeviceReplay(16, fields=())
    with pytest.raises(B2RLError):
        rep.sample(4)
    rep.close()


def test_philox_device_rng_matches_restatement(R):
    rep = R.DeviceReplay(4096, fields=())
    p = _rand_prios(np.random.default_rng(0), 4096)
    rep.build(_dev(p))
    seed = 0x1234ABCD5678
    u = rep.philox_uniforms(seed, 7, 1000).cpu().numpy()
    np.testing.assert_array_equal(u, O.philox_u01(seed, 7, 1000))
    t = O.SumTreeOracle(4096); t.build(p)
    idx, _, _ = rep.sample_counter(seed, 7, 1000)          # stateless device-drawn uniforms
    np.testing.assert_array_equal(idx.cpu().numpy(), t.sample(u)[0])
    rep.seed(seed, 7)                                       # device-resident stream: 600 + 400 draws
    i1, _, _ = rep.sample(600); i2, _, _ = rep.sample(400)
    np.testing.assert_array_equal(torch.cat([i1, i2]).cpu().numpy(), t.sample(u)[0])
    assert 0.45 < u.mean() < 0.55
    rep.close()


def test_sampling_distribution_chi2(R):
    """Proportional sampling: chi-square of 2^20 device-RNG draws over 64 bins vs p / sum(p)."""
    n = 1 << 14
    p = _rand_prios(np.random.default_rng(9), n)
    rep = R.DeviceReplay(n, fields=())
    rep.build(_dev(p))
    draws = 1 << 20
    idx, _, _ = rep.sample(draws)
    counts = np.bincount(idx.cpu().numpy() // (n // 64), minlength=64).astype(np.float64)
    expect = p.astype(np.float64).reshape(64, -1).sum(1) / p.astype(np.float64).sum() * draws
    chi2 = ((counts - expect) ** 2 / expect).sum()
    assert chi2 < 120.0, chi2      # 63 dof: P(chi2 > 120) ~ 2e-5
    rep.close()


# --------------------------------------------------------------------------- #
# gather                                                                        #
# --------------------------------------------------------------------------- #
def test_gather_apex_layout_hash_roundtrip(R):
    cap = 4096
    rep = R.DeviceReplay(cap, fields=R.APEX_FIELDS)
    rep.fill_hash(cap, seed=0xB200)
    rng = np.random.default_rng(3)
    for n in (1, 7, 32, 512, 1500):
        idx = rng.integers(0, cap, size=n)
        idx[-1] = cap - 1; idx[0] = 0
        out = rep.gather(_dev(idx))
        for fi, f in enumerate(R.APEX_FIELDS):
            want = O.hash_rows(fi, idx, f.nbytes, 0xB200)
            got = out[f.name].contiguous().view(torch.uint8).reshape(n, -1).cpu().numpy()
            np.testing.assert_array_equal(got, want, err_msg=f"{f.name} n={n}")
    rep.close()


def test_gather_long_rows_are_chunked(R):
    """R2D2-like rows (multi-chunk: 5 x 28 224 B = 141 120 B) and an odd-sized field."""
    cap = 64
    fields = (R.Field("seq", torch.uint8, (5, 4, 84, 84)), R.Field("odd", torch.uint8, (13,)),
              R.Field("h", torch.float32, (512,)))
    rep = R.DeviceReplay(cap, fields=fields)
    rep.fill_hash(cap, seed=7)
    idx = np.random.default_rng(1).integers(0, cap, size=37)
    out = rep.gather(_dev(idx))
    for fi, f in enumerate(fields):
        want = O.hash_rows(fi, idx, f.nbytes, 7)
        got = out[f.name].contiguous().view(torch.uint8).reshape(len(idx), -1).cpu().numpy()
        np.testing.assert_array_equal(got, want, err_msg=f.name)
    rep.close()


# --------------------------------------------------------------------------- #
# targets                                                                       #
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("case", ["b32", "b8"])
def test_apex_target_vs_reference_golden(R, golden, case):
    g = golden("apex")
    out = R.apex_target(_dev(g[f"{case}_q_s"]), _dev(g[f"{case}_qn_online"]), _dev(g[f"{case}_qn_target"]),
                        _dev(g[f"{case}_action"]), _dev(g[f"{case}_reward"]),
                        _dev(1.0 - g[f"{case}_done"].astype(np.float32)), _dev(g[f"{case}_weight"]),
                        float(g[f"{case}_gamma_n"]), float(g[f"{case}_alpha"]))
    np.testing.assert_allclose(out["prio"].cpu().numpy(), g[f"{case}_new_priority"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out["grad_q"].cpu().numpy(), g[f"{case}_grad_q"], rtol=1e-5, atol=1e-8)
    sc = out["scalars"].cpu().numpy()
    np.testing.assert_allclose(sc[1], g[f"{case}_mean_value"], atol=1e-5)
    np.testing.assert_allclose(sc[2], g[f"{case}_mean_weight"], atol=1e-5)


@pytest.mark.parametrize("B,A", [(1, 6), (32, 6), (512, 6), (4096, 18), (777, 3)])
def test_apex_target_vs_oracle(R, B, A):
    rng = np.random.default_rng(B * 31 + A)
    q = rng.standard_normal((B, A)).astype(np.float32)
    qo = rng.standard_normal((B, A)).astype(np.float32)
    qt = rng.standard_normal((B, A)).astype(np.float32)
    if B > 4:
        qo[3, :] = 0.25                      # ties: first max wins
    a = rng.integers(0, A, size=B)
    r = np.clip(rng.standard_normal(B), -1, 1).astype(np.float32)
    nd = (rng.random(B) > 0.1).astype(np.float32)
    w = rng.uniform(0.1, 1, size=B).astype(np.float32)
    out = R.apex_target(_dev(q), _dev(qo), _dev(qt), _dev(a), _dev(r), _dev(nd), _dev(w), 0.99 ** 3, 0.6)
    tgt, td, prio, gq, info = O.apex_target(q, qo, qt, a, r, nd, w, 0.99 ** 3, 0.6)
    np.testing.assert_array_equal(out["target"].cpu().numpy(), tgt)     # fp32 op-by-op: bit-exact
    np.testing.assert_array_equal(out["td"].cpu().numpy(), td)
    np.testing.assert_allclose(out["prio"].cpu().numpy(), prio, rtol=2.4e-7)
    np.testing.assert_array_equal(out["grad_q"].cpu().numpy(), gq)
    sc = out["scalars"].cpu().numpy()
    np.testing.assert_allclose(sc, [info["loss"], info["mean_value"], info["mean_weight"]], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("case", ["s0", "s1"])
def test_r2d2_target_vs_reference_golden(R, golden, case):
    g = golden("r2d2")
    out = R.r2d2_target(_dev(g[f"{case}_q"]), _dev(g[f"{case}_q_target"]), _dev(g[f"{case}_action"]),
                        _dev(g[f"{case}_reward"]), _dev(g[f"{case}_notdone"], torch.float32),
                        _dev(g[f"{case}_weight"]), int(g["n_step"]), float(g["gamma"]), float(g["alpha"]))
    np.testing.assert_allclose(out["prio"].cpu().numpy(), g[f"{case}_new_priority"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out["grad_q"].cpu().numpy(), g[f"{case}_grad_q"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(out["scalars"].cpu().numpy()[1], g[f"{case}_mean_value"], atol=1e-5)


@pytest.mark.parametrize("L,B,A,n,rescale", [(8, 4, 6, 5, True), (60, 64, 6, 5, True), (40, 32, 6, 5, False),
                                             (200, 3, 4, 3, True), (3, 2, 2, 1, True)])
def test_r2d2_target_vs_oracle(R, L, B, A, n, rescale):
    rng = np.random.default_rng(L * 1000 + B)
    q = (rng.standard_normal((L, B, A)) * 3).astype(np.float32)
    qt = (rng.standard_normal((L, B, A)) * 3).astype(np.float32)
    a = rng.integers(0, A, size=(L - 1, B))
    r = rng.standard_normal((L - 1, B)).astype(np.float32)
    nd = (rng.random(B) > 0.3).astype(np.float32)
    w = rng.uniform(0.1, 1, size=B).astype(np.float32)
    out = R.r2d2_target(_dev(q), _dev(qt), _dev(a), _dev(r), _dev(nd), _dev(w), n, 0.997, 0.9, rescale)
    tgt, td, prio, gq, info = O.r2d2_target(q, qt, a, r, nd, w, n, 0.997, 0.9, rescale)
    np.testing.assert_array_equal(out["target"].cpu().numpy(), tgt)     # bit-exact (IEEE sqrt/div)
    np.testing.assert_array_equal(out["td"].cpu().numpy(), td)
    np.testing.assert_allclose(out["prio"].cpu().numpy(), prio, rtol=5e-7)
    np.testing.assert_array_equal(out["grad_q"].cpu().numpy(), gq)
    np.testing.assert_allclose(out["scalars"].cpu().numpy(), [info["loss"], info["mean_value"]],
                               rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("case", ["c1", "c2"])
def test_vtrace_vs_reference_golden(R, golden, case):
    g = golden("impala")
    gamma, lam, cbar, pbar = [float(x) for x in g[f"{case}_params"]]
    vt, adv = R.vtrace(_dev(g[f"{case}_pi_a"]), _dev(g[f"{case}_mu_a"]), _dev(g[f"{case}_value"]),
                       _dev(g[f"{case}_bootstrap"]), _dev(g[f"{case}_reward"]), gamma, lam, cbar, pbar)
    np.testing.assert_allclose(vt.cpu().numpy(), g[f"{case}_vtarget"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(adv.cpu().numpy(), g[f"{case}_advantage"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("T,B", [(20, 1024), (20, 33), (1, 5), (100, 7)])
def test_vtrace_vs_oracle(R, T, B):
    rng = np.random.default_rng(T * 7 + B)
    pi = rng.uniform(0.02, 0.95, size=(T, B)).astype(np.float32)
    mu = rng.uniform(0.05, 0.9, size=(T, B)).astype(np.float32)
    v = rng.standard_normal((T, B)).astype(np.float32)
    boot = (rng.standard_normal(B) * (rng.random(B) > 0.3)).astype(np.float32)
    r = rng.standard_normal((T, B)).astype(np.float32)
    vt, adv = R.vtrace(_dev(pi), _dev(mu), _dev(v), _dev(boot), _dev(r), 0.99, 1.0, 1.0, 1.0)
    ovt, oadv, _ = O.vtrace(pi, mu, v, boot, r, 0.99, 1.0, 1.0, 1.0)
    # expf/logf differ from numpy's by a few ulp; the recursion keeps that below the contract
    np.testing.assert_allclose(vt.cpu().numpy(), ovt, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(adv.cpu().numpy(), oadv, rtol=1e-5, atol=1e-5)
