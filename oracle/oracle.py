"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's learner hot path.

This file is the *checker*, never the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product
path (distributed_rl_b200/) must never route through it.

Parity status: PINNED.  Every function here is checked against golden vectors
recorded by executing the unmodified reference in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py).
The reference itself ships no tests or golden vectors.

Citations are relative to the reference root (seungju-k1m/Distributed_RL @ f248548).

Numerical conventions shared with the CUDA kernels:
  * sum-tree nodes are fp64, node = fl64(left + right): identical to
    baseline/sumtree.py Node._reduce (:26-31, builtin sum over two floats).
  * every fp32 op is individually rounded (no FMA contraction) — numpy does
    that by construction, the kernels use __fmul_rn/__fadd_rn.
  * pow() is evaluated in fp64 from the fp32 operands and rounded once to
    fp32 ("powcr").  torch-CPU (Sleef, 1 ulp) and numpy powf differ from it by
    <= 1 fp32 ulp, far inside the 1e-5 contract; GPU and oracle agree exactly.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64


def powcr(x: np.ndarray, e: float) -> np.ndarray:
    """fp32 x ** fp32(e), evaluated in fp64 and rounded once to fp32."""
    return np.power(np.asarray(x, F32).astype(F64), F64(F32(e))).astype(F32)


# --------------------------------------------------------------------------- #
# Sum-tree (baseline/sumtree.py)                                               #
# --------------------------------------------------------------------------- #
class SumTreeOracle:
    """Implicit-heap restatement of baseline/sumtree.py SumTree.

    The reference grows a pointer tree by doubling its root (TreeQueue.append
    :81-95) so that after L appends the bounds are (0, 2^ceil(log2 L)) and
    never-written children read as 0.0 (Node._find :58).  That is exactly a
    zero-padded complete binary tree over `capacity = 2^k` leaves:
    node[1] is the root, node[i] = node[2i] + node[2i+1], leaf j is node[cap+j].
    A min-tree over the same shape carries min(priority) for the IS-weight
    normaliser (baseline/PER.py:129-133 recomputes it with an O(N) pass).
    """

    def __init__(self, capacity: int):
        assert capacity >= 1 and (capacity & (capacity - 1)) == 0, "capacity must be 2^k"
        self.cap = capacity
        self.sum = np.zeros(2 * capacity, F64)
        self.min = np.full(2 * capacity, np.inf, F32)
        self.size = 0

    # bulk build == the state after `extend(prios)` (TreeQueue.extend :97-99);
    # every internal node is recomputed from its children (Node._reduce :21-27),
    # so the final state does not depend on the write order.
    def build(self, prios) -> None:
        p = np.asarray(prios)
        n = p.shape[0]
        assert n <= self.cap
        self.size = n
        self.sum[:] = 0.0
        self.min[:] = np.inf
        self.sum[self.cap:self.cap + n] = p.astype(F64)
        p32 = p.astype(F32)
        self.min[self.cap:self.cap + n] = np.where(p32 > 0, p32, F32(np.inf))  # p == 0 <=> empty slot
        lvl = self.cap // 2
        while lvl >= 1:
            i = np.arange(lvl, 2 * lvl)
            self.sum[i] = self.sum[2 * i] + self.sum[2 * i + 1]
            self.min[i] = np.minimum(self.min[2 * i], self.min[2 * i + 1])
            lvl //= 2

    @property
    def total(self) -> float:
        return float(self.sum[1])

    @property
    def min_priority(self) -> np.float32:
        return self.min[1]

    def find(self, pos: float) -> int:
        """Node._find (:53-62): `pos < left ? left : (pos -= left, right)`.
        The `right == 0` guard only acts when pos has rounded up to the subtree
        total, where the reference dereferences a None child and raises."""
        i = 1
        while i < self.cap:
            left = self.sum[2 * i]
            if pos < left or self.sum[2 * i + 1] == 0.0:
                i = 2 * i
            else:
                pos = pos - left
                i = 2 * i + 1
        return i - self.cap

    def sample(self, u01: np.ndarray):
        """SumTree.prioritized_sample (:128-140) driven by explicit uniforms:
        np.random.uniform(0.0, root) == 0.0 + (root - 0.0) * random_sample()."""
        u01 = np.asarray(u01, F64)
        root = self.sum[1]
        idx = np.empty(u01.shape[0], np.int64)
        for k in range(u01.shape[0]):
            idx[k] = self.find(root * u01[k])
        return idx, self.sum[self.cap + idx].copy()

    def update(self, idx, vals) -> None:
        """PrioritizedMemory.update_priorities (baseline/utils.py:347-350) /
        Tree.update (baseline/PER.py:36-42): sequential writes, so for a
        duplicated index the LAST value wins; parents re-reduced on the path."""
        idx = np.asarray(idx, np.int64)
        vals = np.asarray(vals)
        for j, v in zip(idx, vals):
            self.sum[self.cap + j] = F64(v)
            self.min[self.cap + j] = F32(v) if F32(v) > 0 else F32(np.inf)
        touched = np.unique(idx + self.cap)
        while touched.size and touched[0] > 1:
            touched = np.unique(touched // 2)
            self.sum[touched] = self.sum[2 * touched] + self.sum[2 * touched + 1]
            self.min[touched] = np.minimum(self.min[2 * touched], self.min[2 * touched + 1])

    def leaves(self) -> np.ndarray:
        return self.sum[self.cap:self.cap + self.size].copy()


# --------------------------------------------------------------------------- #
# Flat prioritized store (baseline/PER.py) — explicit-uniform replay rule      #
# --------------------------------------------------------------------------- #
def per_sample_flat(prior32: np.ndarray, u01: np.ndarray):
    """Restatement of PER.sample (baseline/PER.py:92-116) for explicit uniforms.

    PER.sample -> Categorical(prob).sample -> torch.multinomial(prob, n, True)
    (third-party: PyTorch CPU kernel, version unpinned by the reference;
    torch 2.11.0 here).  Verified rule: with q = p/sum(p),
    q2 = q/sum(q), c = sequential fp32 cumsum(q2), c /= c[-1], c[-1] = 1, the
    draw for uniform u (fp64) is the first j with c[j] >= u.
    The two fp32 sums use torch's own CPU reduction order, so they are taken
    from torch here; everything else is numpy.
    """
    import torch

    p = torch.from_numpy(np.ascontiguousarray(prior32, F32))
    q = p / torch.sum(p)
    q2 = q / q.sum(-1, keepdim=True)
    c = np.cumsum(q2.numpy(), dtype=F32)  # np.add.accumulate: strictly sequential fp32
    c = (c / c[-1]).astype(F32)
    c[-1] = F32(1.0)
    idx = np.searchsorted(c.astype(F64), np.asarray(u01, F64), side="left").astype(np.int64)
    return idx, q.numpy()[idx]


def is_weights(prior32_sampled, total64, min_prior32, n: int, beta: float):
    """IS weights as the GPU computes them; follows APE_X/ReplayMemory.py:65-67
    and PER.max_weight (baseline/PER.py:129-133):
        prob = p / sum(p);  w = (1 / (n * prob)) ** BETA / max_j (n * prob_j) ** -beta
    x -> x**-beta is decreasing, so the max over j is attained at min_j p_j.
    sum(p) is the fp64 tree root rounded once to fp32."""
    s32 = F32(total64)
    n32 = F32(n)
    prob = (np.asarray(prior32_sampled, F32) / s32).astype(F32)
    w = powcr((F32(1.0) / (n32 * prob).astype(F32)).astype(F32), beta)
    min_prob = F32(F32(min_prior32) / s32)
    max_w = powcr(np.array([n32 * min_prob], F32), -beta)[0]
    return (w / max_w).astype(F32), prob, max_w


# --------------------------------------------------------------------------- #
# Ape-X target / TD / priority (APE_X/Learner.py:85-121)                       #
# --------------------------------------------------------------------------- #
def apex_target(q_s, qn_online, qn_target, action, reward, notdone, weight,
                gamma_n: float, alpha: float):
    """Double-DQN n-step target, clipped TD error, new priority, loss and
    dLoss/dQ(s,.) for the reference Ape-X learner.

      a*   = argmax_a Q(s',a)                               (:90)
      y    = r + 0.99**UNROLL_STEP * Qbar(s',a*) * (1-done) (:93-103)
      d    = clamp(y - Q(s,a), -1, 1)                        (:105-106)
      p    = (|d| + 1e-7) ** ALPHA                           (:108-110)
      loss = 0.5 * mean(w * d^2)                             (:112-114)
    grad_q is what autograd would deliver to `action_value` (:78): zero except
    at (b, a_b), where it is -w_b * d_b / B inside the clamp range
    (torch.clamp passes the gradient on [-1, 1] inclusive).
    """
    q_s = np.asarray(q_s, F32); qn_online = np.asarray(qn_online, F32)
    qn_target = np.asarray(qn_target, F32)
    B, A = q_s.shape
    action = np.asarray(action, np.int64)
    reward = np.asarray(reward, F32); notdone = np.asarray(notdone, F32)
    weight = np.asarray(weight, F32)
    rows = np.arange(B)
    a_star = np.argmax(qn_online, axis=1)
    nxt = (qn_target[rows, a_star] * notdone).astype(F32)
    target = (reward + (F32(gamma_n) * nxt).astype(F32)).astype(F32)
    td_raw = (target - q_s[rows, action]).astype(F32)
    td = np.clip(td_raw, F32(-1), F32(1)).astype(F32)
    prio = powcr((np.abs(td) + F32(1e-7)).astype(F32), alpha)
    wtd2 = (weight * (td * td).astype(F32)).astype(F32)
    loss = F32(F32(wtd2.astype(F64).sum() / B) * F32(0.5))
    inside = (td_raw >= -1) & (td_raw <= 1)
    g = np.where(inside, -(weight * td).astype(F32) / F32(B), F32(0)).astype(F32)
    grad_q = np.zeros((B, A), F32)
    grad_q[rows, action] = g
    info = {"mean_value": F32(target.astype(F64).mean()),
            "mean_weight": F32(weight.astype(F64).mean()), "loss": loss}
    return target, td, prio, grad_q, info


# --------------------------------------------------------------------------- #
# R2D2 sequence target / priority (R2D2/Learner.py:110-198)                    #
# --------------------------------------------------------------------------- #
def value_transform(x, eps=1e-3):
    """h(x), R2D2/Learner.py:22-27, fp32 op by op."""
    x = np.asarray(x, F32)
    return (np.sign(x) * ((np.sqrt((np.abs(x) + F32(1)).astype(F32)) - F32(1)).astype(F32))
            + (F32(eps) * x).astype(F32)).astype(F32)


def value_inv_transform(x, eps=1e-3):
    """h^-1(x), R2D2/Learner.py:30-35, fp32 op by op (python scalars 4*eps and
    2*eps are formed in fp64 first, then cast, as torch does)."""
    x = np.asarray(x, F32)
    inner = ((np.abs(x) + F32(1)).astype(F32) + F32(eps)).astype(F32)
    s = np.sqrt((F32(1) + (F32(4 * eps) * inner).astype(F32)).astype(F32))
    t = ((s - F32(1)).astype(F32) / F32(2 * eps)).astype(F32)
    return (np.sign(x) * ((t * t).astype(F32) - F32(1)).astype(F32)).astype(F32)


def r2d2_target(q, q_target, action, reward, notdone, weight,
                n_step: int, gamma: float, alpha: float, rescale: bool = True):
    """R2D2 n-step double-Q sequence targets over the training window.

    Shapes (time-major, window length L = FIXED_TRAJECTORY - MEM):
      q, q_target : (L, B, A) fp32   online / target net outputs (:121, :132)
      action      : (L-1, B) int     actions   [MEM:-1]  (a-note 1: the shipped
                                     slice at :111 only runs when MEM == T/2,
                                     where it equals [MEM:-1])
      reward      : (L-1, B) fp32    rewards   [MEM:-1]  (:120)
      notdone     : (B,)             1 - done  (R2D2/ReplayMemory.py:86, fp64)
      weight      : (B,) fp32
    Follows :134-192 literally, including the tail `remainder` recursion that
    is seeded with the un-inverted bootstrap and indexes reward[-(i+2)] (:146-157).
    """
    q = np.asarray(q, F32); q_target = np.asarray(q_target, F32)
    L, B, A = q.shape
    action = np.asarray(action, np.int64); reward = np.asarray(reward, F32)
    notdone64 = np.asarray(notdone, F64); weight = np.asarray(weight, F32)
    n = n_step
    tt, bb = np.meshgrid(np.arange(L - 1), np.arange(B), indexing="ij")
    sel = q[tt, bb, action]                                   # (L-1, B)  :123
    amax = np.argmax(q, axis=2)                               # (L, B)    :134
    t2, b2 = np.meshgrid(np.arange(L), np.arange(B), indexing="ij")
    nmv = q_target[t2, b2, amax]                              # (L, B)    :137-140
    target_value = nmv[n:L - 1]                               # (L-1-n, B) :142
    if rescale:
        target_value = value_inv_transform(target_value)      # :143-144
    rewards = np.zeros((L - n - 1, B), F64)                   # :145
    rem = [nmv[L - 1].astype(F64) * notdone64]                # :146-148
    for i in range(n):                                        # :149-153
        rewards += (F32(gamma ** i) * reward[i:L - n - 1 + i]).astype(F32)
        rem.append(reward[-(i + 2)].astype(F64) + gamma * rem[i])
    rewards32 = rewards.astype(F32)                           # :154
    rem = rem[::-1]; rem.pop()                                # :155-156
    remainder = np.stack(rem, 0).astype(F32)                  # :157
    target = (rewards32 + (F32(gamma ** n) * target_value).astype(F32)).astype(F32)  # :161
    target = np.concatenate([target, remainder], 0)           # :162
    if rescale:
        target = value_transform(target)                      # :165-166
    td = (target - sel).astype(F32)                           # :172
    atd = np.abs(td)
    mx = atd.max(0)
    mean = (atd.astype(F64).sum(0) / (L - 1)).astype(F32)
    mixed = ((mx * F32(0.9)).astype(F32) + (F32(0.1) * mean).astype(F32)).astype(F32)  # :180
    prio = powcr(mixed, alpha)                                # :181
    denom = F32(B * (L - 1))
    loss = F32(F32(((weight[None, :] * (td * td).astype(F32)).astype(F32)).astype(F64).sum() / denom)
               * F32(0.5))                                    # :189-191
    g = (-(weight[None, :] * td).astype(F32) / denom).astype(F32)
    grad_q = np.zeros((L, B, A), F32)
    grad_q[tt, bb, action] = g
    info = {"mean_value": F32(sel.astype(F64).mean()), "loss": loss}
    return target, td, prio, grad_q, info


# --------------------------------------------------------------------------- #
# IMPALA V-trace (IMPALA/Learner.py:141-215)                                   #
# --------------------------------------------------------------------------- #
def vtrace(pi_a, mu_a, value, bootstrap, reward,
           gamma: float, c_lambda: float, c_bar: float, p_bar: float):
    """V-trace targets and policy-gradient advantages, as the reference
    computes them (quirks kept: the last step is NOT rho-clipped (:177-185),
    c_bar is used for both the delta weight and the trace (:192-197), no
    intra-rollout done masking).

      pi_a, mu_a, value, reward : (T, B) fp32;  bootstrap : (B,) = V(s_T)*done (:143)
    returns Vtarget (T, B) (:202) and advantage (T, B) (:207-212).
    """
    pi_a = np.asarray(pi_a, F32); mu_a = np.asarray(mu_a, F32)
    value = np.asarray(value, F32); reward = np.asarray(reward, F32)
    boot = np.asarray(bootstrap, F32)
    T, B = value.shape
    g = F32(gamma)
    ratio = np.exp((np.log(pi_a) - np.log(mu_a)).astype(F32)).astype(F32)   # :151-174
    vmt = np.zeros((T, B), F32)
    for i in reversed(range(T)):
        if i == T - 1:
            vmt[i] = ((reward[i] + (g * boot).astype(F32)).astype(F32) - value[i]).astype(F32)
        else:
            td = ((reward[i] + (g * value[i + 1]).astype(F32)).astype(F32) - value[i]).astype(F32)
            cr = np.minimum(F32(c_bar), ratio[i])
            cs = (F32(c_lambda) * cr).astype(F32)
            vmt[i] = ((td * cr).astype(F32)
                      + ((g * cs).astype(F32) * vmt[i + 1]).astype(F32)).astype(F32)
    vtarget = (value + vmt).astype(F32)
    nxt = np.concatenate([vtarget[1:], boot[None, :]], 0)
    atarget = (reward + (g * nxt).astype(F32)).astype(F32)
    pt = np.minimum(F32(p_bar), ratio)
    adv = ((atarget - value).astype(F32) * pt).astype(F32)
    return vtarget, adv, ratio


# --------------------------------------------------------------------------- #
# Synthetic payload hash + device RNG (no reference counterpart; restated so    #
# that full-size GPU runs can be verified without materialising them on host)   #
# --------------------------------------------------------------------------- #
def lowbias32(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.uint32).copy()
    x ^= x >> np.uint32(16); x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15); x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def hash_rows(field_index: int, slots: np.ndarray, row_bytes: int, seed: int) -> np.ndarray:
    """Bytes that b2rl_replay_fill_hash writes into rows `slots` of field `field_index`."""
    slots = np.asarray(slots, np.int64)
    words = (row_bytes + 3) // 4
    with np.errstate(over="ignore"):
        s = (slots.astype(np.uint32) * np.uint32(2654435761))[:, None]
        w = (np.arange(words, dtype=np.uint32) * np.uint32(2246822519))[None, :]
        fs = np.uint32((field_index * 0x9E3779B9) & 0xFFFFFFFF)
        v = lowbias32(np.uint32(seed & 0xFFFFFFFF) ^ fs ^ s ^ w)
    b = v.astype("<u4").view(np.uint8).reshape(len(slots), words * 4)
    return b[:, :row_bytes]


def philox_u01(seed: int, offset: int, n: int) -> np.ndarray:
    """Philox4x32-10 uniforms as k_tree_sample draws them: counter (ctr,0), key = seed."""
    ctr = np.arange(offset, offset + n, dtype=np.uint64)
    c0 = (ctr & np.uint64(0xFFFFFFFF)); c1 = (ctr >> np.uint64(32))
    c2 = np.zeros(n, np.uint64); c3 = np.zeros(n, np.uint64)
    k0 = np.uint64(seed & 0xFFFFFFFF); k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & mask
        c0, c1, c2, c3 = n0 & mask, n1, n2 & mask, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    x = (c1 << np.uint64(32)) | c0
    return (x & np.uint64((1 << 53) - 1)).astype(np.float64) * (1.0 / 9007199254740992.0)


class RingModel:
    """Host model of the device ring (b2rl_replay_push / _evict): stable slot ids,
    FIFO overwrite — the idiomatic replacement of PER.remove_to_fit's renumbering
    (baseline/PER.py:118-127)."""

    def __init__(self, capacity: int):
        self.capacity = capacity
        self.prios = np.zeros(capacity, F32)
        self.size = 0
        self.head = 0

    def push(self, prios):
        prios = np.asarray(prios, F32)
        slots = (self.head + np.arange(len(prios))) % self.capacity
        self.prios[slots] = prios
        self.head = int((self.head + len(prios)) % self.capacity)
        self.size = min(self.capacity, self.size + len(prios))
        return slots

    def evict(self, delta):
        tail = (self.head - self.size) % self.capacity
        slots = (tail + np.arange(delta)) % self.capacity
        self.prios[slots] = 0
        self.size -= delta
        return slots
