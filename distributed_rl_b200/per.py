"""`PER` and `PrioritizedMemory` with the reference's method surface
(baseline/PER.py:48-133, baseline/utils.py:328-360) over a DeviceReplay.

Differences forced by device residency: records have a fixed
field layout instead of being arbitrary pickles, and `sample` hands back the
gathered fields as CUDA tensors where the reference returns pickled blobs.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch

from . import replay as R


class _PriorityView:
    """`PER.priority` (baseline/PER.py:12-45): exposes prior_torch / update / __len__."""

    def __init__(self, store: R.DeviceReplay):
        self._s = store

    @property
    def prior_torch(self) -> torch.Tensor:
        return self._s.priorities(0, len(self._s))

    def update(self, idx, vals):
        self._s.update(_as_index(idx, self._s.device), torch.as_tensor(np.asarray(vals, np.float32)))

    def __len__(self):
        return len(self._s)      # the reference returns len(self.prior) == 0 forever (:16,:44-45)


def _as_index(idx, device):
    if torch.is_tensor(idx):
        return idx.to(device=device, dtype=torch.int64)
    if len(idx) and torch.is_tensor(idx[0]):
        return torch.stack([i.reshape(()) for i in idx]).to(device=device, dtype=torch.int64)
    return torch.as_tensor(np.asarray(idx, np.int64)).to(device)


class PER:
    def __init__(self, maxlen=1000, max_value=1.0, beta=0.4, fields=R.APEX_FIELDS, device="cuda:0",
                 decode=None):
        self.beta, self.maxlen, self.max_value = beta, maxlen, max_value
        self.store = R.DeviceReplay(maxlen, fields, device)
        self.priority = _PriorityView(self.store)
        self.memory = self.store               # len(per.memory) works like the reference's list
        self._decode = decode or _decode_apex_record

    def push(self, d):
        """d: list of pickled records whose last element is the priority (:69-75)."""
        if not d:
            return
        cols, prios = self._decode([pickle.loads(b) for b in d])
        self.store.push(cols, prios)

    def __getitem__(self, idx):
        i = torch.as_tensor([int(idx)], device=self.store.device)
        return {k: v[0] for k, v in self.store.gather(i).items()}

    def __len__(self):
        return len(self.store)

    def update(self, idx: list, vals: np.ndarray):
        assert isinstance(vals, np.ndarray)     # same contract as :87-88
        assert isinstance(idx, list)
        self.priority.update(idx, vals)

    def sample(self, batch_size):
        """-> (fields dict of CUDA tensors, prob fp32[n], idx int64[n])  (:92-116)"""
        idx, prob, _ = self.store.sample(batch_size, beta=self.beta)
        return self.store.gather(idx), prob, idx

    def sample_with_weights(self, batch_size):
        """sample + the IS weights of APE_X/ReplayMemory.py:65-67 in the same launch."""
        idx, prob, w = self.store.sample(batch_size, beta=self.beta)
        return self.store.gather(idx), prob, idx, w

    def remove_to_fit(self):
        """FIFO drop of everything beyond `maxlen` (:118-127) = b2rl_replay_evict of the oldest records.  The
        ring overwrites its oldest slot on push, so with capacity == maxlen there is normally nothing to drop."""
        over = len(self.store) - self.maxlen
        if over > 0:
            self.store.evict(over)

    @property
    def max_weight(self) -> float:
        return float(self.store.stats(self.beta)[2].item())


def _decode_apex_record(recs):
    s = np.stack([np.asarray(r[0], np.uint8) for r in recs])
    ns = np.stack([np.asarray(r[3], np.uint8) for r in recs])
    a = np.asarray([int(r[1]) for r in recs], np.int32)
    rw = np.asarray([float(r[2]) for r in recs], np.float32)
    d = np.asarray([bool(r[4]) for r in recs], np.uint8)
    p = np.asarray([float(r[-1]) for r in recs], np.float32)
    return [s, ns, a, rw, d], p


class PrioritizedMemory:
    """baseline/utils.py:328-360 on the device sum-tree (the reference's SumTree semantics)."""

    def __init__(self, capacity, fields=R.APEX_FIELDS, device="cuda:0"):
        self.capacity = capacity
        self.store = R.DeviceReplay(capacity, fields, device)

    def push(self, transitions, priorities):
        """transitions: sequence of per-field arrays (n, *shape), one per field."""
        self.store.push(list(transitions), np.asarray(priorities, np.float32))

    def sample(self, batch_size, u01=None):
        idx, _, _ = self.store.sample(batch_size, u01=u01)
        prios = self.store.priorities()[idx] if batch_size else torch.empty(0, device=self.store.device)
        return self.store.gather(idx), prios, idx

    def update_priorities(self, indices, priorities):
        self.store.update(_as_index(indices, self.store.device),
                          torch.as_tensor(np.asarray(priorities, np.float32)))

    def remove_to_fit(self):
        """popleft() until len <= capacity (baseline/utils.py:352-357) = evict the oldest records."""
        over = len(self.store) - self.capacity
        if over > 0:
            self.store.evict(over)

    def __len__(self):
        return len(self.store)

    def total_prios(self) -> float:
        return float(self.store.stats()[0].item())
