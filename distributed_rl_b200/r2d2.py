"""R2D2 learner side with the reference's surface (R2D2/ReplayMemory.py:23-185,
R2D2/Learner.py:39-339): sequences of FIXED_TRAJECTORY steps with a stored LSTM
state live in HBM; burn-in, double-Q n-step targets with value rescaling, the
mixed max/mean sequence priority and dLoss/dQ come from one kernel
(b2rl_r2d2_target) instead of a D2H hop into NumPy fp64 (R2D2/Learner.py:145-181).

SURVEY §8a-note 1: the shipped action slice `action[FIXED_TRAJECTORY-MEM:-1]`
(:111) only runs when MEM == T/2; `[MEM:-1]` (what :120 uses for the rewards)
is used here, identical whenever the reference runs at all.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field

import torch

from . import replay as R
from .agent import GraphAgent
from .apex import make_optimizer, _MemoryView, _Conv1Gathered


def default_r2d2_model() -> dict:
    """cfg/r2d2.json:33-103 (configuration values)."""
    return {
        "module00": {"netCat": "CNN2D", "iSize": 4, "nLayer": 4, "fSize": [8, 4, 3, -1], "nUnit": [32, 64, 64],
                     "padding": [0, 0, 0], "stride": [4, 2, 1], "act": ["relu", "relu", "relu"],
                     "BN": [False] * 4, "linear": True, "input": [0], "prior": 0},
        "module01": {"netCat": "ViewV2", "prevNodeNames": ["module00"], "input": [1], "prior": 1},
        "module02": {"netCat": "LSTMNET", "hiddenSize": 512, "nLayer": 1, "iSize": 3136, "device": "cpu",
                     "FlattenMode": True, "return_hidden": False, "prior": 2, "prevNodeNames": ["module01"]},
        "module03": {"netCat": "MLP", "iSize": 512, "nLayer": 2, "fSize": [512, 6], "act": ["relu", "linear"],
                     "BN": [False] * 3, "prior": 3, "prevNodeNames": ["module02"]},
        "module03_1": {"netCat": "MLP", "iSize": 512, "nLayer": 2, "fSize": [512, 1], "act": ["relu", "linear"],
                       "BN": [False] * 3, "prior": 3, "prevNodeNames": ["module02"]},
        "module04": {"netCat": "Add", "prior": 4, "prevNodeNames": ["module03", "module03_1"]},
        "module04_1": {"netCat": "Mean", "prior": 4, "prevNodeNames": ["module03"]},
        "module05": {"netCat": "Substract", "prior": 5, "prevNodeNames": ["module04", "module04_1"],
                     "output": True},
    }


@dataclass
class R2D2Config:
    BATCHSIZE: int = 32
    ACTION_SIZE: int = 6
    ALPHA: float = 0.9
    BETA: float = 0.4
    GAMMA: float = 0.997
    UNROLL_STEP: int = 5
    FIXED_TRAJECTORY: int = 80
    MEM: int = 20
    USE_RESCALING: bool = True
    REPLAY_MEMORY_LEN: int = 10000
    BUFFER_SIZE: int = 1000
    TARGET_FREQUENCY: int = 2500
    LEARNER_DEVICE: str = "cuda:0"
    REDIS_SERVER: str = "localhost"
    LOG_W: str | None = None
    OPTIM_INFO: dict = field(default_factory=lambda: {"name": "adam", "lr": 1e-4, "eps": 0.001})
    MODEL: dict = field(default_factory=default_r2d2_model)
    PAYLOAD_POOL: int = 0        # > 0: the sum-tree has REPLAY_MEMORY_LEN slots but only this many distinct sequences
                                 #      are stored, slot s reading row s % PAYLOAD_POOL (2^20 x 2.26 MB = 2.4 TB
                                 #      does not fit HBM).  Benchmark-only; ingest needs 0.
    FUSED_CONV1: bool = True     # conv_1 of every frame through libb2rl's tcgen05 kernel (gather fused)
    FUSED_HEADS: bool = True     # dueling heads: 3xTF32 tcgen05 GEMM + fused tail kernels (csrc/gemm.cu, csrc/dueling.cu)

    @staticmethod
    def from_configuration():
        import configuration as C
        names = ("BATCHSIZE", "ACTION_SIZE", "ALPHA", "BETA", "GAMMA", "UNROLL_STEP", "FIXED_TRAJECTORY", "MEM",
                 "USE_RESCALING", "REPLAY_MEMORY_LEN", "BUFFER_SIZE", "TARGET_FREQUENCY", "LEARNER_DEVICE",
                 "REDIS_SERVER", "OPTIM_INFO", "MODEL")
        return R2D2Config(LOG_W=getattr(C, "LOG_W", None), **{k: getattr(C, k) for k in names})


class Replay(threading.Thread):
    """R2D2/ReplayMemory.py Replay: batch = [(h0, h1), s, a, r, notdone, w, idx] (:118-120)."""

    def __init__(self, cfg: R2D2Config | None = None, connect=None):
        super().__init__(daemon=True)
        self.cfg = cfg or R2D2Config.from_configuration()
        self.device = torch.device(self.cfg.LEARNER_DEVICE)
        fields = R.r2d2_fields(self.cfg.FIXED_TRAJECTORY)
        if self.cfg.PAYLOAD_POOL:
            self.store = R.DeviceReplay(self.cfg.REPLAY_MEMORY_LEN, (), self.device)          # priorities only
            self.pool = R.DeviceReplay(self.cfg.PAYLOAD_POOL, fields, self.device)           # the stored sequences
        else:
            self.store = R.DeviceReplay(self.cfg.REPLAY_MEMORY_LEN, fields, self.device)
            self.pool = self.store
        self.memory = _MemoryView(self.store, self.cfg.BETA)
        self.connect, self.cond, self.lock = connect, False, False
        self.deque, self.total_frame = [], 0
        self._lock = threading.Lock()       # the reference's update() is unlocked (:48-51); the handle needs it
        self._stop_evt = threading.Event()

    def push_arrays(self, s, a, r, h0, h1, notdone, p):
        with self._lock:
            self.store.push([s, a, r, h0, h1, notdone], p)
        self.total_frame += int(torch.as_tensor(p).numel())

    def push_records(self, blobs) -> None:
        """PER.push for the actors' pickled sequences (R2D2/Player.py:312-319): decoded once on the host with
        the indexing of R2D2/ReplayMemory.py:70-88, one batched H2D copy + leaf writes."""
        if not blobs:
            return
        import pickle
        from .wire import decode_r2d2
        cols, p = decode_r2d2([pickle.loads(b) for b in blobs], self.cfg.FIXED_TRAJECTORY)
        self.push_arrays(*cols, p)

    def stop(self) -> None:
        self._stop_evt.set()

    def run(self):
        """R2D2/ReplayMemory.py:141-178: drain `experience`, push, serve the eviction request."""
        if self.connect is None:
            return
        import time
        from .wire import drain
        while not self._stop_evt.is_set():
            data = drain(self.connect, "experience")
            if data:
                self.push_records(data)
                self.cond = len(self.store) > self.cfg.BUFFER_SIZE
            if self.lock:
                self._evict_on_request()
            if not data:
                time.sleep(0.002)

    def _evict_on_request(self) -> None:
        if len(self.store) >= self.cfg.REPLAY_MEMORY_LEN:       # :165-173
            with self._lock:
                self.deque.clear()
                over = len(self.store) - self.cfg.REPLAY_MEMORY_LEN
                if over > 0:
                    self.store.evict(over)
        self.lock = False

    def buffer(self, m: int = 1):
        B = self.cfg.BATCHSIZE
        with self._lock:
            idx, _, w = self.store.sample(B * m, beta=self.cfg.BETA)
            b = self.pool.gather(self.rows_of(idx))
        for k in range(m):
            sl = slice(k * B, (k + 1) * B)
            h0 = b["h0"][sl].unsqueeze(0).contiguous()     # (1, B, 512) like torch.cat(..., 1) at :87-88
            h1 = b["h1"][sl].unsqueeze(0).contiguous()
            self.deque.append([(h0, h1), b["state"][sl], b["action"][sl], b["reward"][sl], b["notdone"][sl],
                               w[sl], idx[sl]])

    def rows_of(self, idx: torch.Tensor) -> torch.Tensor:
        """payload row of each sampled slot (identity unless PAYLOAD_POOL)."""
        return idx % self.cfg.PAYLOAD_POOL if self.cfg.PAYLOAD_POOL else idx

    def sample(self):
        if not self.deque:
            if len(self.store) <= self.cfg.BUFFER_SIZE:
                return False
            self.buffer(1)
        return self.deque.pop(0)

    def update(self, idx, vals):
        if isinstance(idx, (list, tuple)):
            idx = torch.stack([torch.as_tensor(i) for i in idx])
        with self._lock:
            self.store.update(torch.as_tensor(idx).to(self.device), torch.as_tensor(vals).to(self.device))


class Learner:
    def __init__(self, cfg: R2D2Config | None = None, connect=None, start_replay: bool = True, writer=None):
        self.cfg = cfg or R2D2Config.from_configuration()
        self.device = torch.device(self.cfg.LEARNER_DEVICE)
        self.model = GraphAgent(self.cfg.MODEL).to(self.device)
        self.target_model = GraphAgent(self.cfg.MODEL).to(self.device)
        for m in (self.model, self.target_model):
            m.dense_3xtf32 = m.fused_dueling_tail = bool(self.cfg.FUSED_HEADS) and self.device.type == "cuda"
        self.optim = make_optimizer(self.cfg.OPTIM_INFO, self.model.getParameters())
        self.connect = connect
        self.memory = Replay(self.cfg, connect)
        self.writer = writer
        if connect is not None:
            if start_replay:
                self.memory.start()                              # R2D2/Learner.py:46-48
            from .wire import wipe_stale_keys
            wipe_stale_keys(connect)                             # :54,63-64

    def train(self, transition, t=0):
        c = self.cfg
        T, MEM, B, A = c.FIXED_TRAJECTORY, c.MEM, c.BATCHSIZE, c.ACTION_SIZE
        L = T - MEM
        (h0, h1), state, action, reward, notdone, weight, idx = transition
        dev = self.device
        weight = torch.as_tensor(weight).to(dev, torch.float32)
        h0, h1 = h0.to(dev), h1.to(dev)
        self.model.setCellState((h0, h1))                       # :86-87
        self.target_model.setCellState((h0, h1))
        state = torch.as_tensor(state).to(dev)
        fused = c.FUSED_CONV1 and state.dtype == torch.uint8 and self.model.first_conv_node() is not None
        if fused:
            frames = state.contiguous().view(B * T, 4, 84, 84)
            q, q_target = self._forward_fused(frames, self._time_major_rows(None, T, B), T, MEM, B, A)
        else:
            state = state.float() / 255.0                       # :89-90
            sv = state.permute(1, 0, 2, 3, 4).contiguous()      # time-major, :93
            burn = sv[:MEM].reshape(-1, 4, 84, 84)
            window = sv[MEM:].reshape(-1, 4, 84, 84)
            with torch.no_grad():                               # burn-in, :99-104
                shape = torch.tensor([MEM, B, -1])
                self.model.forward([burn, shape])
                self.target_model.forward([burn, shape])
                self.model.detachCellState()
                self.target_model.detachCellState()
            shape = torch.tensor([L, B, -1])
            q = self.model.forward([window, shape])[0].view(L, B, A)              # :121
            with torch.no_grad():
                q_target = self.target_model.forward([window, shape])[0].view(L, B, A)   # :132
        act = torch.as_tensor(action).to(dev, torch.int64).t()[MEM:-1].contiguous()      # (L-1, B)
        rew = torch.as_tensor(reward).to(dev, torch.float32).t()[MEM:-1].contiguous()
        nd = torch.as_tensor(notdone).to(dev, torch.float32).contiguous()
        out = R.r2d2_target(q.detach().contiguous(), q_target.contiguous(), act, rew, nd, weight,
                            c.UNROLL_STEP, c.GAMMA, c.ALPHA, c.USE_RESCALING)
        q.backward(out["grad_q"])                               # == loss.backward(), :189-192
        info = self.step()
        info["mean_value"] = out["scalars"][1]
        info["loss"] = out["scalars"][0]
        return info, out["prio"], idx

    def _time_major_rows(self, seq_rows, T, B):
        """Frame-table rows of the (t, b) frames in time-major order: row = seq_row[b] * T + t
        (seq_rows None: the batch itself is the table, sequence b = rows b*T ... b*T+T-1)."""
        dev = self.device
        if not hasattr(self, "_t_idx"):
            self._t_idx = torch.arange(T, device=dev).view(T, 1)
            self._b_idx = torch.arange(B, device=dev).view(1, B)
        base = self._b_idx if seq_rows is None else seq_rows.view(1, B)
        return (base * T + self._t_idx).reshape(-1).contiguous()

    def _forward_fused(self, frames, tm_rows, T, MEM, B, A):
        """The forward passes with conv_1 on the tensor cores, reading `frames` (a uint8 (rows, 4, 84, 84) table:
        the staged batch or the replay payload itself) IN PLACE: the (b, t) -> time-major reordering and the
        uint8 -> /255 conversion are folded into the kernel's gather (`tm_rows`)."""
        dev = self.device
        if not hasattr(self, "_pack2"):
            self._conv_name = self.model.first_conv_node()
            c_out = getattr(self.model, self._conv_name).conv_1.out_channels
            self._pack2 = R.Conv1Pack(2, dev, c_out)
            self._pack1 = R.Conv1Pack(1, dev, c_out)
        w_on = getattr(self.model, self._conv_name).conv_1.weight
        w_tg = getattr(self.target_model, self._conv_name).conv_1.weight
        self._pack2.pack(0, w_on); self._pack2.pack(1, w_tg); self._pack1.pack(0, w_on)
        rows_burn, rows_win = tm_rows[:MEM * B], tm_rows[MEM * B:]
        L = T - MEM
        with torch.no_grad():                                   # burn-in, :99-104
            y_on, y_tg = R.conv1_fused(frames, rows_burn, self._pack2, relu=True)
            shape = torch.tensor([MEM, B, -1])
            self.model.forward_from_conv1(y_on, True, [shape])
            self.target_model.forward_from_conv1(y_tg, True, [shape])
            self.model.detachCellState()
            self.target_model.detachCellState()
            y_on_w, y_tg_w = R.conv1_fused(frames, rows_win, self._pack2, relu=False)
        shape = torch.tensor([L, B, -1])
        mf = torch.contiguous_format
        y = _Conv1Gathered.apply(w_on, frames, rows_win, self._pack1, mf, None, y_on_w)
        q = self.model.forward_from_conv1(y, False, [shape])[0].view(L, B, A)             # :121
        with torch.no_grad():
            q_target = self.target_model.forward_from_conv1(torch.relu(y_tg_w), True, [shape])[0].view(L, B, A)
        return q, q_target

    def fused_step(self):
        """One learner step with everything resident: sample B sequence slots + IS weights from the sum-tree,
        gather only the small per-sequence fields (a, r, h0, h1, notdone: 5 KB of the 2.26 MB), run conv_1 over the
        sequences' frames IN PLACE in the replay payload (row = slot_row * T + t), target / priority kernel,
        backward, clip + Adam, priority write-back.  No host round trip (R2D2/Learner.py:235-274 in one call)."""
        c = self.cfg
        T, MEM, B, A = c.FIXED_TRAJECTORY, c.MEM, c.BATCHSIZE, c.ACTION_SIZE
        mem = self.memory
        st, pool = mem.store, mem.pool
        if not hasattr(self, "_small"):
            self._small = pool.alloc_batch(B, ("action", "reward", "h0", "h1", "notdone"))
            self._frames = pool.field_view("state").view(-1, 4, 84, 84)
        idx, _, w = st.sample(B, beta=c.BETA, want_prob=False)
        rows = mem.rows_of(idx)
        b = pool.gather(rows, self._small)
        h0, h1 = b["h0"].unsqueeze(0), b["h1"].unsqueeze(0)
        self.model.setCellState((h0, h1))
        self.target_model.setCellState((h0, h1))
        q, q_target = self._forward_fused(self._frames, self._time_major_rows(rows, T, B), T, MEM, B, A)
        act = b["action"].to(torch.int64).t()[MEM:-1].contiguous()
        rew = b["reward"].t()[MEM:-1].contiguous()
        out = R.r2d2_target(q.detach().contiguous(), q_target.contiguous(), act, rew, b["notdone"], w,
                            c.UNROLL_STEP, c.GAMMA, c.ALPHA, c.USE_RESCALING)
        q.backward(out["grad_q"])
        info = self.step()
        st.update(idx, out["prio"])
        return {"scalars": out["scalars"], "p_norm": info["p_norm"], "prio": out["prio"], "idx": idx}

    def step(self):
        """R2D2/Learner.py:200-215: norm, clip at 40, Adam."""
        params = self.model.getParameters()
        grads = [p.grad for p in params if p.grad is not None]
        p_norm = torch.stack(torch._foreach_norm(grads, 2)).sum().sqrt()
        torch.nn.utils.clip_grad_norm_(params, 40, foreach=True)
        self.optim.step()
        self.optim.zero_grad(set_to_none=False)
        return {"p_norm": p_norm}

    @property
    def state_dict(self):
        return {k: v.cpu() for k, v in self.model.state_dict().items()}

    @property
    def target_state_dict(self):
        return {k: v.cpu() for k, v in self.target_model.state_dict().items()}

    def run(self, max_steps=None, log_every: int = 500):
        """R2D2/Learner.py:217-339: wait for BUFFER_SIZE sequences, announce `Start`, then per step sample ->
        train -> priority write-back; hard target sync every TARGET_FREQUENCY steps (+ `target_state_dict`),
        `state_dict` / `count` (= step - 50, sic :293) every 25 steps, and every 500 steps the eviction request,
        the `reward` drain + log line and a checkpoint.  Publication is asynchronous (ParamPublisher)."""
        import pickle
        import time
        from . import wire
        from .publish import ParamPublisher
        while len(self.memory.memory) <= self.cfg.BUFFER_SIZE:
            time.sleep(0.05)
        if self.connect is not None:                                     # :227-234
            self.connect.set("state_dict", pickle.dumps(self.state_dict))
            self.connect.set("count", pickle.dumps(1))
            self.connect.set("target_state_dict", pickle.dumps(self.target_state_dict))
            self.connect.set("Start", pickle.dumps(True))
        pub = ParamPublisher(self.model, self.connect, "state_dict", "count")
        pub_t = ParamPublisher(self.target_model, self.connect, "target_state_dict", None)
        ckpt_path = wire.checkpoint_path(self.cfg.LOG_W)
        ckpt = ParamPublisher(self.model, None, None, None,
                              on_ready=lambda sd, step: torch.save(sd, ckpt_path)) if ckpt_path else None
        self._publishers = (pub, pub_t) + ((ckpt,) if ckpt else ())
        step, acc, t0 = 0, None, time.time()
        self.last_log = None
        while max_steps is None or step < max_steps:
            batch = self.memory.sample()
            if batch is False:
                time.sleep(0.002)
                continue
            info, prio, idx = self.train(batch)
            step += 1
            if step % log_every == 0:
                self.memory.lock = True                                  # :266-268
                if self.connect is None or not self.memory.is_alive():
                    self.memory._evict_on_request()
            if not self.memory.lock:
                self.memory.update(idx, prio)                            # :271-274
            tot = torch.stack([info["mean_value"].reshape(()), info["p_norm"].reshape(())])
            acc = tot if acc is None else acc + tot
            if step % self.cfg.TARGET_FREQUENCY == 0:                    # :283-286
                self.target_model.updateParameter(self.model, 1)
                pub_t.snapshot(step)
            if step % 25 == 0:                                           # :288-293
                pub.snapshot(step - 50)
            for p in self._publishers:
                p.poll()
            if step % log_every == 0:                                    # :296-339
                reward, n_rew = wire.drain_rewards(self.connect) if self.connect is not None else (-21.0, 0)
                mean_value, norm = (acc / log_every).tolist()
                dt = (time.time() - t0) / log_every
                self.last_log = {"step": step, "mean_value": mean_value, "norm": norm, "reward": reward,
                                 "time_per_step": dt}
                print(f"step:{step} // mean_value:{mean_value:.3f} // norm: {norm:.3f} // REWARD:{reward:.3f} // "
                      f"NUM_MEMORY:{len(self.memory.memory)} // MAX_WEIGHT:{self.memory.memory.max_weight:.3f} // "
                      f"TIME:{dt:.5f}")
                if self.writer is not None:
                    if n_rew:
                        self.writer.add_scalar("Reward", reward, step)
                    self.writer.add_scalar("value", mean_value, step)
                    self.writer.add_scalar("norm", norm, step)
                if ckpt is not None:
                    ckpt.snapshot(step)
                acc, t0 = None, time.time()
        return step
