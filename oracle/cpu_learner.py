"""TEST / BASELINE INFRASTRUCTURE — CPU port of the reference's Ape-X learner loop.

Used only by bench.py (`cpu_baseline` and `--impl reference`) and by tests: it
is the thing the GPU path is timed AGAINST, never part of the product path.
The reference is pure Python (PyTorch CPU + NumPy + pickle), so the port is
Python too and keeps the reference's data structures and costs:

  * flat fp32 priority vector sampled with torch.distributions.Categorical
    (baseline/PER.py:92-116), two extra O(N) passes for max_weight (:129-133)
  * records are pickled `[s, a, R_n, s', done, prio]` blobs in a Python list;
    a minibatch is built by deepcopy + pickle.loads + np.stack for 16 batches
    at a time (APE_X/ReplayMemory.py:61-116; the ragged np.array needs
    dtype=object on numpy >= 1.24)
  * Learner.train on the CPU: fp32 /255 conversion, three forwards of the
    dueling DQN, double-DQN n-step target, clipped TD, priority in NumPy,
    IS-weighted loss, per-tensor grad-norm loop, centered RMSprop
    (APE_X/Learner.py:55-138)
  * priority write-back `prior_torch[np.array(idx)] = vals` (baseline/PER.py:36-42)

To fit host RAM at N = 2^20 the pickled payload is a pool of `pool` distinct
records addressed `slot % pool`; the priority vector has the
full N entries, so sampling / max_weight / update cost what they cost in the
reference.
"""
from __future__ import annotations

import pickle
import time
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn


class DuelingDQN(nn.Module):
    """cfg/ape_x.json:37-88: conv 8x8s4-32, 4x4s2-64, 3x3s1-64 (no bias), two
    bias-free heads 3136-512-{A,1}, Q = (A + V) - mean(A)."""

    def __init__(self, actions=6):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(4, 32, 8, 4, bias=False), nn.ReLU(),
                                  nn.Conv2d(32, 64, 4, 2, bias=False), nn.ReLU(),
                                  nn.Conv2d(64, 64, 3, 1, bias=False), nn.ReLU(), nn.Flatten())
        self.adv = nn.Sequential(nn.Linear(3136, 512, bias=False), nn.ReLU(), nn.Linear(512, actions, bias=False))
        self.val = nn.Sequential(nn.Linear(3136, 512, bias=False), nn.ReLU(), nn.Linear(512, 1, bias=False))

    def forward(self, x):
        f = self.conv(x)
        a, v = self.adv(f), self.val(f)
        return (a + v) - a.mean(dim=-1, keepdim=True)


class CpuApexLearner:
    def __init__(self, n_slots: int, batch: int, m: int = 16, pool: int = 2048, actions: int = 6,
                 alpha: float = 0.6, beta: float = 0.4, unroll: int = 3, seed: int = 0, threads: int | None = None):
        if threads:
            torch.set_num_threads(threads)
        self.N, self.B, self.m, self.A = n_slots, batch, m, actions
        self.alpha, self.beta, self.unroll = alpha, beta, unroll
        rng = np.random.default_rng(0xB200 + seed)
        self.pool = []
        for i in range(pool):
            rec = [rng.integers(0, 256, size=(4, 84, 84), dtype=np.uint8), int(rng.integers(0, actions)),
                   float(np.clip(rng.standard_normal(), -1, 1)),
                   rng.integers(0, 256, size=(4, 84, 84), dtype=np.uint8), bool(rng.random() < 0.02), 1.0]
            self.pool.append(pickle.dumps(rec))
        self.npool = pool
        self.prior = torch.from_numpy(
            ((np.abs(rng.standard_normal(n_slots)).clip(max=1) + 1e-7) ** alpha).astype(np.float32))
        torch.manual_seed(seed)
        self.model, self.target = DuelingDQN(actions), DuelingDQN(actions)
        self.optim = torch.optim.RMSprop(self.model.parameters(), lr=6.25e-5, eps=1.5e-7, alpha=0.95,
                                         momentum=0, centered=True)
        self.deque = []
        self.pend_idx, self.pend_val = [], []

    # APE_X/ReplayMemory.py:61-116 (+ baseline/PER.py:92-116, 129-133)
    def buffer(self):
        n = self.B * self.m
        prob = self.prior / torch.sum(self.prior)
        idx = torch.distributions.categorical.Categorical(prob).sample([n])
        blobs = deepcopy([self.pool[int(i) % self.npool] for i in idx])
        s_prob = prob[idx]
        weight = (1 / (self.N * s_prob)) ** self.beta
        prob2 = self.prior / torch.sum(self.prior)
        max_w = float(((self.N * prob2) ** -self.beta).max().numpy())
        weight /= max_w
        exp = np.array([pickle.loads(b) for b in blobs], dtype=object)
        state = np.stack(exp[:, 0], 0)
        next_state = np.stack(exp[:, 3], 0)
        action, reward, done = exp[:, 1], exp[:, 2], exp[:, 4]
        for k in range(self.m):
            sl = slice(k * self.B, (k + 1) * self.B)
            self.deque.append([state[sl], action[sl], reward[sl], next_state[sl], done[sl], weight[sl], idx[sl]])

    # APE_X/Learner.py:55-138
    def train(self, tr):
        state, action, reward, next_state, done, weight, idx = tr
        weight = weight.clone().float()
        s = torch.tensor(state).float() / 255.
        ns = torch.tensor(next_state).float() / 255.
        act = [self.A * i + int(a) for i, a in enumerate(action)]
        r = torch.tensor(reward.astype(np.float32)).float()
        nd = torch.tensor([float(not d) for d in done.astype(bool)]).float()
        q = self.model(s)
        with torch.no_grad():
            qt = self.target(ns)
            qn = self.model(ns)
            a_star = qn.argmax(dim=-1).cpu().numpy()
            sel = [self.A * i + a for i, a in enumerate(a_star)]
            nxt = qt.view(-1)[sel] * nd
        q_sa = q.view(-1)[act]
        target = r + 0.99 ** self.unroll * nxt
        td = torch.clamp(target - q_sa, -1, 1)
        prio = (np.abs(td.detach().cpu().numpy()) + 1e-7) ** self.alpha
        loss = torch.mean(weight * td ** 2) * 0.5
        loss.backward()
        norm = 0
        for p in self.model.parameters():
            norm += p.grad.data.norm(2)
        norm = norm ** .5
        self.optim.step()
        self.optim.zero_grad()
        return prio, idx, float(target.mean()), float(norm)

    # APE_X/ReplayMemory.py:43-59 + baseline/PER.py:36-42
    def update(self, idx, prio):
        self.pend_idx += list(idx)
        self.pend_val.append(prio)

    def flush_updates(self):
        if not self.pend_idx:
            return
        vals = np.concatenate(self.pend_val, 0)
        self.prior[np.array(self.pend_idx)] = torch.tensor(vals).float()
        self.pend_idx.clear(); self.pend_val.clear()

    def cycle(self):
        """One reference cycle: buffer() for m minibatches, m train steps, write-back."""
        t0 = time.perf_counter()
        self.buffer()
        t1 = time.perf_counter()
        for _ in range(self.m):
            prio, idx, _, _ = self.train(self.deque.pop(0))
            self.update(idx, prio)
        t2 = time.perf_counter()
        self.flush_updates()
        t3 = time.perf_counter()
        return {"transitions": self.B * self.m, "t_buffer": t1 - t0, "t_train": t2 - t1, "t_update": t3 - t2,
                "t_total": t3 - t0}
