"""Host-side wire logic (distributed_rl_b200/wire.py): atomic list drain (the reference's LTRIM -1 0 quirk),
stale-key wipe, reward drain, and the record decoders against records built exactly like the actors build them
(APE_X/Player.py:252-261, R2D2/Player.py:38-63,312-319, IMPALA/Player.py:97-114)."""
import pickle

import numpy as np
import pytest
import torch

from distributed_rl_b200 import wire
from fake_redis import FakeRedis


def test_fake_redis_has_real_ltrim_semantics():
    r = FakeRedis()
    r.rpush("k", b"a")
    p = r.pipeline(); p.lrange("k", 0, -1); p.ltrim("k", -1, 0)
    assert p.execute()[0] == [b"a"]
    assert r.llen("k") == 1            # LTRIM -1 0 on a 1-element list keeps the element (start = stop = 0)
    r.rpush("k", b"b", b"c")
    p = r.pipeline(); p.lrange("k", 0, -1); p.ltrim("k", -1, 0); p.execute()
    assert r.llen("k") == 0            # start (2) > stop (0): emptied


@pytest.mark.parametrize("n", [1, 2, 7])
def test_drain_takes_each_record_exactly_once(n):
    r = FakeRedis()
    for i in range(n):
        r.rpush("experience", pickle.dumps(i))
    got = wire.drain(r, "experience")
    assert [pickle.loads(b) for b in got] == list(range(n))
    assert wire.drain(r, "experience") == []        # nothing is read twice — also for a lone record
    r.rpush("experience", pickle.dumps(99))
    assert [pickle.loads(b) for b in wire.drain(r, "experience")] == [99]


def test_wipe_and_reward_drain():
    r = FakeRedis()
    r.set("Start", b"1"); r.rpush("experience", b"x"); r.set("state_dict", b"y")
    assert wire.wipe_stale_keys(r) == 3 and r.scan()[1] == []
    assert wire.drain_rewards(r) == (-21.0, 0)      # APE_X/Learner.py:229-230
    for v in (1.0, 3.0, 8.0):
        r.rpush("reward", pickle.dumps(v))
    assert wire.drain_rewards(r) == (4.0, 3)
    assert r.llen("reward") == 0


def test_decode_apex_record():
    rng = np.random.default_rng(0)
    recs = [[rng.integers(0, 256, (4, 84, 84), dtype=np.uint8), int(rng.integers(6)), float(rng.standard_normal()),
             rng.integers(0, 256, (4, 84, 84), dtype=np.uint8), bool(i % 2), 0.1 + i] for i in range(3)]
    blobs = [pickle.dumps(r) for r in recs]
    out = {"s": np.zeros((4, 4, 84, 84), np.uint8), "ns": np.zeros((4, 4, 84, 84), np.uint8),
           "a": np.zeros(4, np.int32), "r": np.zeros(4, np.float32), "d": np.zeros(4, np.uint8),
           "p": np.zeros(4, np.float32)}
    wire.decode_apex([pickle.loads(b) for b in blobs], out)
    for i, r in enumerate(recs):
        assert np.array_equal(out["s"][i], r[0]) and np.array_equal(out["ns"][i], r[3])
        assert out["a"][i] == r[1] and out["r"][i] == np.float32(r[2]) and out["d"][i] == r[4]
        assert out["p"][i] == np.float32(r[5])


def _r2d2_record(rng, T, done):
    """np.array(traj_) of R2D2/Player.py LocalBuffer.get_traj (:38-63) + np.append(priority) (:314)."""
    traj = [(torch.from_numpy(rng.standard_normal((1, 1, 512)).astype(np.float32)),
             torch.from_numpy(rng.standard_normal((1, 1, 512)).astype(np.float32)))]
    for _ in range(T):
        traj += [rng.integers(0, 256, (4, 84, 84), dtype=np.uint8), int(rng.integers(6)), float(rng.standard_normal())]
    traj.append(done)
    arr = np.empty(len(traj), dtype=object)
    for i, x in enumerate(traj):
        arr[i] = x
    return np.append(arr, float(rng.random()) + 0.1)


def test_decode_r2d2_record_follows_reference_indexing():
    rng = np.random.default_rng(1)
    T = 6
    recs = [_r2d2_record(rng, T, d) for d in (False, True)]
    recs = [pickle.loads(pickle.dumps(r)) for r in recs]
    (s, a, rw, h0, h1, nd), p = wire.decode_r2d2(recs, T)
    assert s.shape == (2, T, 4, 84, 84) and a.shape == (2, T) and h0.shape == (2, 512)
    for i, r in enumerate(recs):
        # R2D2/ReplayMemory.py:70-88: state_idx = 1+3t, action_idx = 2+3t, reward_idx = 3+3t, done = exp[-2]
        for t in range(T):
            assert np.array_equal(s[i, t], r[1 + 3 * t]) and a[i, t] == r[2 + 3 * t]
            assert rw[i, t] == np.float32(r[3 + 3 * t])
        assert nd[i] == float(not r[-2]) and p[i] == np.float32(r[-1])
        assert np.array_equal(h0[i], r[0][0].numpy().ravel()) and np.array_equal(h1[i], r[0][1].numpy().ravel())


def test_decode_impala_record():
    rng = np.random.default_rng(2)
    T = 5
    recs = []
    for flag in (0, 1):
        recs.append([rng.integers(0, 256, (T + 1, 28224), dtype=np.uint8),
                     rng.integers(0, 6, (T, 1)), rng.uniform(0.05, 0.9, (T, 1)).astype(np.float32),
                     rng.standard_normal(T), flag])
    s, a, mu, rw, d = wire.decode_impala([pickle.loads(pickle.dumps(r)) for r in recs], T)
    for i, r in enumerate(recs):
        assert np.array_equal(s[i], r[0]) and np.array_equal(a[i], r[1][:, 0])
        assert np.array_equal(mu[i], r[2][:, 0]) and np.array_equal(rw[i], r[3].astype(np.float32))
        assert d[i] == float(r[4])
