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


# =========================================================================== #
# The reference's OWN sum-tree as a priority store ("also time the same loop with
# SumTree / PrioritizedMemory substituted for PER") — baseline/sumtree.py:4-140, baseline/utils.py:328-360.
# Cost model kept: one heap-allocated Python node per tree node, Python floats, one recursive call
# per level for every write / find, root doubling on append.
# =========================================================================== #
class _PNode:
    __slots__ = ("value", "left", "right")

    def __init__(self, value=0.0, left=None, right=None):
        self.value, self.left, self.right = value, left, right


def _pt_write(node, lo, hi, key, value):          # Node._write :29-41 (+ _expand, _reduce)
    if hi - lo == 1:
        node.value, node.left, node.right = value, None, None
        return
    if node.left is None and node.right is None:
        node.left, node.right = _PNode(), _PNode()
    mid = (lo + hi) // 2
    if key < mid:
        _pt_write(node.left, lo, mid, key, value)
    else:
        _pt_write(node.right, mid, hi, key, value)
    node.value = sum([node.left.value, node.right.value])


def _pt_find(node, lo, hi, pos):                  # Node._find :53-62
    if hi - lo == 1:
        return lo
    mid = (lo + hi) // 2
    lv = node.left.value if node.left is not None else 0.0
    if pos < lv:
        return _pt_find(node.left, lo, mid, pos)
    return _pt_find(node.right, mid, hi, pos - lv)


def _pt_get(node, lo, hi, key):                   # Node._get :43-51
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if key < mid:
            node, hi = node.left, mid
        else:
            node, lo = node.right, mid
    return node.value


class PointerSumTree:
    def __init__(self):
        self.length, self.root, self.hi = 0, None, 0

    def append(self, v):                          # TreeQueue.append :81-95
        if self.length == 0:
            self.root, self.hi, self.length = _PNode(v), 1, 1
            return
        if self.hi == self.length:
            self.root = _PNode(self.root.value, self.root, _PNode())
            self.hi *= 2
        _pt_write(self.root, 0, self.hi, self.length, v)
        self.length += 1

    def __setitem__(self, i, v):
        _pt_write(self.root, 0, self.hi, i, v)

    def prioritized_sample(self, n, rng):         # SumTree.prioritized_sample :128-140
        ixs, vals = [], []
        for _ in range(n):
            ix = _pt_find(self.root, 0, self.hi, rng.uniform(0.0, self.root.value))
            ixs.append(ix); vals.append(_pt_get(self.root, 0, self.hi, ix))
        return ixs, vals


class CpuApexSumTreeLearner(CpuApexLearner):
    """CpuApexLearner with the flat `PER` store swapped for `PrioritizedMemory` (SumTree) — same payload
    handling and `train`; sampling = n recursive descents, IS weights from the returned priorities, write-back =
    one recursive `_write` per index (baseline/utils.py:341-350)."""

    def __init__(self, n_slots, batch, **kw):
        super().__init__(n_slots, batch, **kw)
        self.tree = PointerSumTree()
        for v in self.prior.numpy().tolist():
            self.tree.append(v)
        self.nprng = np.random.RandomState(0)
        self.min_prior = float(self.prior.min())

    def buffer(self):
        n = self.B * self.m
        ixs, vals = self.tree.prioritized_sample(n, self.nprng)
        blobs = deepcopy([self.pool[i % self.npool] for i in ixs])
        total = self.tree.root.value
        prob = torch.tensor(vals, dtype=torch.float32) / total
        weight = (1 / (self.N * prob)) ** self.beta
        weight /= float((self.N * self.min_prior / total) ** -self.beta)
        idx = torch.tensor(ixs)
        exp = np.array([pickle.loads(b) for b in blobs], dtype=object)
        state = np.stack(exp[:, 0], 0)
        next_state = np.stack(exp[:, 3], 0)
        action, reward, done = exp[:, 1], exp[:, 2], exp[:, 4]
        for k in range(self.m):
            sl = slice(k * self.B, (k + 1) * self.B)
            self.deque.append([state[sl], action[sl], reward[sl], next_state[sl], done[sl], weight[sl], idx[sl]])

    def flush_updates(self):
        if not self.pend_idx:
            return
        vals = np.concatenate(self.pend_val, 0)
        for i, v in zip(self.pend_idx, vals):     # PrioritizedMemory.update_priorities :347-350
            self.tree[int(i)] = float(v)
        self.pend_idx.clear(); self.pend_val.clear()


# =========================================================================== #
# R2D2 (BASELINE.json configs[2]) — R2D2/ReplayMemory.py:53-122, R2D2/Learner.py:76-215        #
# =========================================================================== #
def _h(x, eps=1e-3):      # value_transform, R2D2/Learner.py:22-27
    return torch.sign(x) * (torch.sqrt(torch.abs(x) + 1) - 1) + eps * x


def _h_inv(x, eps=1e-3):  # value_inv_transform, :30-35
    return torch.sign(x) * (((torch.sqrt(1 + 4 * eps * (torch.abs(x) + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1)


class R2D2Net(nn.Module):
    """cfg/r2d2.json:33-103: the Ape-X conv stack (no bias) -> LSTM(3136, 512) -> bias-free dueling heads
    512-512-{A,1}; Q = (A + V) - mean(A).  Time-major (T, B) sequences, state carried between calls."""

    def __init__(self, actions=6):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(4, 32, 8, 4, bias=False), nn.ReLU(),
                                  nn.Conv2d(32, 64, 4, 2, bias=False), nn.ReLU(),
                                  nn.Conv2d(64, 64, 3, 1, bias=False), nn.ReLU(), nn.Flatten())
        self.lstm = nn.LSTM(3136, 512, 1)
        self.adv = nn.Sequential(nn.Linear(512, 512, bias=False), nn.ReLU(), nn.Linear(512, actions, bias=False))
        self.val = nn.Sequential(nn.Linear(512, 512, bias=False), nn.ReLU(), nn.Linear(512, 1, bias=False))
        self.state = None

    def forward(self, x, T, B):
        f = self.conv(x).view(T, B, -1)
        y, self.state = self.lstm(f, self.state)
        y = y.reshape(T * B, -1)
        a, v = self.adv(y), self.val(y)
        return (a + v) - a.mean(dim=-1, keepdim=True)

    def detach_state(self):
        self.state = tuple(s.detach() for s in self.state)


class CpuR2D2Learner:
    def __init__(self, n_slots: int, batch: int, m: int = 16, pool: int = 32, T: int = 80, mem: int = 20,
                 unroll: int = 5, gamma: float = 0.997, alpha: float = 0.9, beta: float = 0.4, actions: int = 6,
                 seed: int = 0, threads: int | None = None):
        if threads:
            torch.set_num_threads(threads)
        self.N, self.B, self.m, self.T, self.MEM, self.n, self.A = n_slots, batch, m, T, mem, unroll, actions
        self.gamma, self.alpha, self.beta = gamma, alpha, beta
        rng = np.random.default_rng(0xB200 + 2 + seed)
        self.pool = []
        for _ in range(pool):        # records as R2D2/Player.py LocalBuffer.get_traj builds them (:38-63, :312-319)
            traj = [(torch.from_numpy((rng.standard_normal((1, 1, 512)) * 0.1).astype(np.float32)),
                     torch.from_numpy((rng.standard_normal((1, 1, 512)) * 0.1).astype(np.float32)))]
            for _t in range(T):
                traj += [rng.integers(0, 256, size=(4, 84, 84), dtype=np.uint8), int(rng.integers(0, actions)),
                         float(rng.standard_normal())]
            traj.append(bool(rng.random() < 0.3))
            arr = np.empty(len(traj) + 1, dtype=object)
            for i, x in enumerate(traj):
                arr[i] = x
            arr[-1] = 1.0
            self.pool.append(pickle.dumps(arr))
        self.npool = pool
        self.prior = torch.from_numpy(
            ((np.abs(rng.standard_normal(n_slots)).clip(max=1) + 1e-7) ** alpha).astype(np.float32))
        torch.manual_seed(seed)
        self.model, self.target = R2D2Net(actions), R2D2Net(actions)
        self.optim = torch.optim.Adam(self.model.parameters(), lr=1e-4, eps=1e-3)
        self.deque, self.pend_idx, self.pend_val = [], [], []
        L = T - mem
        self.action_idx = torch.tensor([actions * i for i in range(batch * L)])                # :59
        self.action_idx_np = np.array([actions * i for i in range(batch * (L - 1))])           # :60

    def buffer(self):                  # R2D2/ReplayMemory.py:53-122
        n, T = self.B * self.m, self.T
        prob = self.prior / torch.sum(self.prior)
        idx = torch.distributions.categorical.Categorical(prob).sample([n])
        blobs = deepcopy([self.pool[int(i) % self.npool] for i in idx])
        s_prob = prob[idx]
        weight = (1 / (self.N * s_prob)) ** self.beta
        prob2 = self.prior / torch.sum(self.prior)
        weight /= float(((self.N * prob2) ** -self.beta).max().numpy())
        exps = [pickle.loads(b) for b in blobs]
        state_idx = [1 + i * 3 for i in range(T)]
        action_idx = [2 + i * 3 for i in range(T)]
        reward_idx = [3 + i * 3 for i in range(T)]
        state = np.stack([np.stack(e[state_idx], 0) for e in exps], 0)
        action = np.array([e[action_idx].astype(np.int32) for e in exps])
        reward = np.array([e[reward_idx].astype(np.float32) for e in exps])
        done = np.array([float(not e[-2]) for e in exps])
        h0 = torch.cat([e[0][0] for e in exps], 1)
        h1 = torch.cat([e[0][1] for e in exps], 1)
        for k in range(self.m):
            sl = slice(k * self.B, (k + 1) * self.B)
            self.deque.append([(h0[:, sl], h1[:, sl]), state[sl], action[sl], reward[sl], done[sl], weight[sl], idx[sl]])

    def train(self, tr):               # R2D2/Learner.py:76-215 (action slice [MEM:-1])
        (h0, h1), state, action, reward, done, weight, idx = tr
        T, MEM, B, A, n = self.T, self.MEM, self.B, self.A, self.n
        L = T - MEM
        weight = weight.clone().float()
        self.model.state = (h0.contiguous(), h1.contiguous())
        self.target.state = (h0.contiguous(), h1.contiguous())
        st = torch.tensor(state).float() / 255.
        sv = st.permute(1, 0, 2, 3, 4).contiguous()
        burn = sv[:MEM].contiguous().view(-1, 4, 84, 84)
        trunc = sv[MEM:].contiguous().view(-1, 4, 84, 84)
        with torch.no_grad():
            self.model(burn, MEM, B); self.target(burn, MEM, B)
            self.model.detach_state(); self.target.detach_state()
        act = np.transpose(action, (1, 0))[MEM:-1].reshape(-1)
        act = self.action_idx_np + act
        rew = np.transpose(reward.astype(np.float32), (1, 0))[MEM:-1]
        q = self.model(trunc, L, B).view(-1)
        sel = q[act]
        qd = q.detach().view(-1, A)
        with torch.no_grad():
            qt = self.target(trunc, L, B).view(-1)
            nxt = qt[self.action_idx + qd.argmax(-1)].view(L, B)
            tv = _h_inv(nxt[n:-1].contiguous())
            rewards = np.zeros((L - n - 1, B))
            remainder = [nxt[-1].numpy() * done]
            for i in range(n):
                rewards += self.gamma ** i * rew[i:L - n - 1 + i]
                remainder.append(rew[-(i + 2)] + self.gamma * remainder[i])
            rewards = torch.tensor(rewards).float()
            remainder = remainder[::-1]; remainder.pop()
            remainder = torch.tensor(np.array(remainder)).float()
            target = _h(torch.cat((rewards + self.gamma ** n * tv, remainder), 0).view(-1)).detach()
        td = target - sel
        tdp = abs(np.reshape(td.detach().numpy(), (L - 1, -1)))
        prio = (tdp.max(0) * 0.9 + 0.1 * tdp.mean(0)) ** self.alpha
        loss = torch.mean(weight.view(-1, 1) * (td.view(L - 1, -1).permute(1, 0).contiguous() ** 2)) * 0.5
        loss.backward()
        norm = 0
        for p in self.model.parameters():
            norm += p.grad.data.norm(2)
        torch.nn.utils.clip_grad_norm_(list(self.model.parameters()), 40)
        self.optim.step(); self.optim.zero_grad()
        return prio, idx, float(sel.mean().detach()), float(norm ** .5)

    def cycle(self):
        t0 = time.perf_counter()
        self.buffer()
        t1 = time.perf_counter()
        for _ in range(self.m):
            prio, idx, _, _ = self.train(self.deque.pop(0))
            self.pend_idx += list(idx); self.pend_val.append(prio)
        t2 = time.perf_counter()
        self.prior[np.array(self.pend_idx)] = torch.tensor(np.concatenate(self.pend_val, 0)).float()
        self.pend_idx.clear(); self.pend_val.clear()
        t3 = time.perf_counter()
        return {"transitions": self.B * self.m * self.T, "sequences": self.B * self.m, "t_buffer": t1 - t0,
                "t_train": t2 - t1, "t_update": t3 - t2, "t_total": t3 - t0}


# =========================================================================== #
# IMPALA (BASELINE.json configs[3]) — IMPALA/ReplayMemory.py:30-54, IMPALA/Learner.py:70-266    #
# =========================================================================== #
class ImpalaNet(nn.Module):
    """cfg/impala.json:24-52: conv 8x8s4-16, 4x4s2-32 (no bias) -> MLP 2592-256-(A+1); last column = V."""

    def __init__(self, actions=6):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(4, 16, 8, 4, bias=False), nn.ReLU(),
                                 nn.Conv2d(16, 32, 4, 2, bias=False), nn.ReLU(), nn.Flatten(),
                                 nn.Linear(2592, 256, bias=False), nn.ReLU(), nn.Linear(256, actions + 1, bias=False))

    def forward(self, x):
        return self.net(x)


class CpuImpalaLearner:
    def __init__(self, n_slots: int, batch: int, m: int = 8, pool: int = 256, T: int = 20, actions: int = 6,
                 gamma: float = 0.99, seed: int = 0, threads: int | None = None):
        if threads:
            torch.set_num_threads(threads)
        import random
        self.random = random.Random(seed)
        self.N, self.B, self.m, self.T, self.A, self.gamma = n_slots, batch, m, T, actions, gamma
        rng = np.random.default_rng(0xB200 + 3 + seed)
        self.pool = [pickle.dumps([rng.integers(0, 256, size=(T + 1, 28224), dtype=np.uint8),
                                   rng.integers(0, actions, size=(T, 1)),
                                   rng.uniform(0.05, 0.9, size=(T, 1)).astype(np.float32),
                                   rng.standard_normal(T).astype(np.float32), int(rng.random() > 0.3)])
                     for _ in range(pool)]
        self.memory = [self.pool[i % pool] for i in range(n_slots)]     # references to pooled blobs (no copies)
        torch.manual_seed(seed)
        self.model = ImpalaNet(actions)
        self.optim = torch.optim.RMSprop(self.model.parameters(), lr=6e-4, weight_decay=0, eps=1e-5, alpha=0.99)
        self.deque = []
        self.c_value, self.p_value = torch.tensor(1.0), torch.tensor(1.0)

    def bufferSave(self):              # IMPALA/ReplayMemory.py:30-54
        m = self.m
        tr = self.random.sample(self.memory, self.B * m)
        tr = np.array([pickle.loads(b) for b in tr], dtype=object)
        state = np.uint8(np.stack(tr[:, 0], axis=1))
        action = np.concatenate(list(tr[:, 1]), axis=1)
        policy = np.concatenate(list(tr[:, 2]), axis=1)
        reward = np.stack(tr[:, 3], axis=1)
        done = np.float32(np.array(tr[:, 4], dtype=np.float32))
        for s, a, p, r, d in zip(np.split(state, m, 1), np.split(action, m, 1), np.split(policy, m, 1),
                                 np.split(reward, m, 1), np.split(done, m)):
            self.deque.append((s, a, p, r, d))

    def _fwd(self, state, action):     # Learner.forward, :70-83
        out = self.model(state)
        pol = torch.softmax(out[:, :self.A], dim=-1)
        ind = (torch.arange(0, len(pol)) * self.A + action[:, 0]).long()
        return pol.view(-1)[ind], out[:, -1:]

    def train(self, transition):       # IMPALA/Learner.py:121-266 without the TensorBoard / Redis tail
        T, B, A, g = self.T, self.B, self.A, self.gamma
        with torch.no_grad():
            sb, action, policy, reward, done = [torch.tensor(x) for x in transition]
            done = done.view(-1, 1)
            sb = sb.float(); sb /= torch.tensor(255).float()
            reward = reward.view(T, B, 1)
            sb = sb.view(T + 1, B, 4, 84, 84)
            last = sb[-1]
            sb = sb[:-1].view(-1, 4, 84, 84)
            est = self.model(last)[:, -1:] * done
            ab = action.view(-1, 1)
            lp, lv = self._fwd(sb, ab)
            log_ratio = torch.log(lp.view(-1, 1)) - torch.log(policy.view(-1, 1))
            lv = lv.view(T, B, 1)
            vmt = torch.zeros((T, B, 1)).float()
            ratio = torch.exp(log_ratio).view(T, B, 1)
            for i in reversed(range(T)):                          # the Python V-trace loop, :176-200
                if i == T - 1:
                    vmt[i] += reward[i] + g * est - lv[i]
                else:
                    td = reward[i] + g * lv[i + 1] - lv[i]
                    cr = torch.min(self.c_value, ratio[i])
                    vmt[i] += td * cr + g * (1.0 * cr) * vmt[i + 1]
            vt = lv + vmt
            nvt = torch.cat((vt, est.unsqueeze(0)), 0)[1:]
            adv = ((reward + g * nvt).view(-1, 1) - lv.view(-1, 1)) * torch.min(self.p_value, ratio).view(-1, 1)
            vt = vt.view(-1, 1)
        out = self.model(sb.detach())                              # calLoss, :95-119
        pol = torch.softmax(out[:, :A], dim=-1)
        logp = torch.log(pol)
        ent = -torch.sum(pol * logp, -1, keepdim=True)
        ind = (torch.arange(0, B * T) * A + ab[:, 0]).long()
        sel = logp.view(-1)[ind].view(-1, 1)
        obj = torch.mean(sel * adv + 0.01 * ent)
        critic = torch.mean((out[:, -1] - vt[:, 0]).pow(2)) / 2
        self.optim.zero_grad()
        (-obj + critic).backward()
        torch.nn.utils.clip_grad_norm_(list(self.model.parameters()), 40)
        self.optim.step()
        return float(critic.detach())

    def cycle(self):
        t0 = time.perf_counter()
        self.bufferSave()
        t1 = time.perf_counter()
        for _ in range(self.m):
            self.train(self.deque.pop(0))
        t2 = time.perf_counter()
        return {"transitions": self.B * self.m * self.T, "rollouts": self.B * self.m, "t_buffer": t1 - t0,
                "t_train": t2 - t1, "t_update": 0.0, "t_total": t2 - t0}
