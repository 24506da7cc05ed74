"""IMPALA learner side with the reference's surface (IMPALA/ReplayMemory.py:14-85,
IMPALA/Learner.py:17-297).  Rollouts (T+1 frames, T actions / behaviour probs /
rewards, done) live in HBM; sampling is uniform WITHOUT replacement per call like
baseline/utils.py ReplayMemory.sample (:310-315); the V-trace backward scan
(:176-200, a Python loop of ~10 tiny kernels per step in the reference) is one
launch of b2rl_vtrace."""
from __future__ import annotations

import threading
from dataclasses import dataclass, field

import torch

from . import replay as R
from .agent import GraphAgent
from .apex import make_optimizer, _Conv1Gathered


def default_impala_model() -> dict:
    """cfg/impala.json:24-52 (configuration values)."""
    return {
        "module00": {"netCat": "CNN2D", "iSize": 4, "nLayer": 3, "fSize": [8, 4, -1], "nUnit": [16, 32],
                     "padding": [0, 0], "stride": [4, 2], "act": ["relu", "relu"], "BN": [False] * 3,
                     "linear": True, "input": [0], "prior": 0},
        "module01": {"netCat": "MLP", "iSize": 2592, "nLayer": 2, "fSize": [256, 7], "act": ["relu", "linear"],
                     "BN": [False, False], "prior": 1, "prevNodeNames": ["module00"], "output": True},
    }


@dataclass
class ImpalaConfig:
    BATCHSIZE: int = 32
    ACTION_SIZE: int = 6
    GAMMA: float = 0.99
    C_LAMBDA: float = 1
    C_VALUE: float = 1.0
    P_VALUE: float = 1.0
    ENTROPY_R: float = 0.01
    UNROLL_STEP: int = 20
    REPLAY_MEMORY_LEN: int = 10000
    BUFFER_SIZE: int = 9999
    LEARNER_DEVICE: str = "cuda:0"
    REDIS_SERVER: str = "localhost"
    LOG_W: str | None = None
    OPTIM_INFO: dict = field(default_factory=lambda: {"name": "rmsprop", "lr": 6e-4, "decay": 0})
    MODEL: dict = field(default_factory=default_impala_model)
    FUSED_CONV1: bool = True     # conv_1 (4 -> 16 channels) of every frame through libb2rl's tcgen05 kernel
    DENSE_3XTF32: bool = True    # the 2592 -> 256 layer as a 3xTF32 tcgen05 GEMM (csrc/gemm.cu) instead of an fp32 SIMT sgemm

    @staticmethod
    def from_configuration():
        import configuration as C
        names = ("BATCHSIZE", "ACTION_SIZE", "GAMMA", "C_LAMBDA", "C_VALUE", "P_VALUE", "ENTROPY_R", "UNROLL_STEP",
                 "REPLAY_MEMORY_LEN", "BUFFER_SIZE", "LEARNER_DEVICE", "REDIS_SERVER", "OPTIM_INFO", "MODEL")
        return ImpalaConfig(LOG_W=getattr(C, "LOG_W", None), **{k: getattr(C, k) for k in names})


class Replay(threading.Thread):
    """IMPALA/ReplayMemory.py Replay: batch = (s[T+1,B,28224], a[T,B], mu[T,B], r[T,B], done[B])."""

    def __init__(self, cfg: ImpalaConfig | None = None, connect=None):
        super().__init__(daemon=True)
        self.cfg = cfg or ImpalaConfig.from_configuration()
        self.device = torch.device(self.cfg.LEARNER_DEVICE)
        self.store = R.DeviceReplay(self.cfg.REPLAY_MEMORY_LEN, R.impala_fields(self.cfg.UNROLL_STEP), self.device)
        self.deque = []
        self._connect = connect
        self._lock = threading.Lock()
        self._stop_evt = threading.Event()
        self._rng = torch.Generator(device=self.device)        # uniform sampling stream (random.sample in the reference)
        self._rng.manual_seed(0x1A9A1A)

    def push_arrays(self, s, a, mu, r, done):
        n = torch.as_tensor(done).numel()
        with self._lock:
            self.store.push([s, a, mu, r, done], torch.ones(n))     # uniform replay: unit priorities

    def push_records(self, blobs) -> None:
        """ReplayMemory.push (baseline/utils.py:305-309) for the actors' pickled rollouts
        (IMPALA/Player.py:183-190): FIFO ring, oldest rollouts overwritten beyond REPLAY_MEMORY_LEN."""
        if not blobs:
            return
        import pickle
        from .wire import decode_impala
        self.push_arrays(*decode_impala([pickle.loads(b) for b in blobs], self.cfg.UNROLL_STEP))

    def stop(self) -> None:
        self._stop_evt.set()

    def run(self):
        """IMPALA/ReplayMemory.py:56-76: drain `trajectory`, push.  Batches are assembled on demand."""
        if self._connect is None:
            return
        import time
        from .wire import drain
        while not self._stop_evt.is_set():
            data = drain(self._connect, "trajectory")
            if data:
                self.push_records(data)
            else:
                time.sleep(0.002)

    def bufferSave(self, m: int = 1):
        """IMPALA/ReplayMemory.py:30-54 with random.sample's no-replacement semantics."""
        B = self.cfg.BATCHSIZE
        with self._lock:
            idx = self.draw(B * m)
            b = self.store.gather(idx)
        for k in range(m):
            sl = slice(k * B, (k + 1) * B)
            self.deque.append((b["state"][sl].transpose(0, 1).contiguous(), b["action"][sl].t().contiguous(),
                               b["mu"][sl].t().contiguous(), b["reward"][sl].t().contiguous(), b["done"][sl]))

    def draw(self, n: int) -> torch.Tensor:
        """n distinct slots, uniformly (random.sample over the list of kept rollouts, baseline/utils.py:310-315).
        The kept rollouts are the ring's valid region [head - size, head): slots reserved for an ingest in flight
        are outside it and are never read while they are being overwritten."""
        size, cap, head = self.store._sizes()
        if n > size:
            raise ValueError("Sample larger than population")      # what random.sample raises
        k = torch.randperm(size, device=self.device, generator=self._rng)[:n]
        tail = (head - size) % cap
        return (k + tail) % cap if tail else k

    def sample(self):
        if not self.deque:
            if len(self.store) <= self.cfg.BUFFER_SIZE:
                return False
            self.bufferSave(1)
        return self.deque.pop(0)

    def __len__(self):
        return len(self.store)


class Learner:
    def __init__(self, cfg: ImpalaConfig | None = None, connect=None, start_replay: bool = True):
        self.cfg = cfg or ImpalaConfig.from_configuration()
        self.device = torch.device(self.cfg.LEARNER_DEVICE)
        self.model = GraphAgent(self.cfg.MODEL).to(self.device)
        self.model.dense_3xtf32 = bool(self.cfg.DENSE_3XTF32) and self.device.type == "cuda"
        self.mOptim = make_optimizer(self.cfg.OPTIM_INFO, self.model.getParameters())
        self._connect = connect
        self._memory = Replay(self.cfg, connect)
        if connect is not None and start_replay:
            self._memory.start()                                # IMPALA/Learner.py:27-28
        self.last = {}

    def forward(self, state, action):
        """IMPALA/Learner.py:70-83: pi(a|s) of the taken action and V(s)."""
        out = self.model.forward([state])[0]
        A = self.cfg.ACTION_SIZE
        policy = torch.softmax(out[:, :A], dim=-1)
        pi_a = policy.gather(1, action.view(-1, 1).long())[:, 0]
        return pi_a, out[:, -1]

    def train(self, transition, step=0):
        c = self.cfg
        T, B = c.UNROLL_STEP, c.BATCHSIZE
        dev = self.device
        state, action, mu, reward, done = [torch.as_tensor(x).to(dev) for x in transition]
        fused = c.FUSED_CONV1 and state.dtype == torch.uint8 and self.model.first_conv_node() is not None
        if fused:   # the staged batch is the frame table: rows already are time-major
            frames = state.contiguous().view((T + 1) * B, 4, 84, 84)
            self._train_core(frames, None, action, mu, reward, done, step)
        else:
            self._train_core(state, "staged", action, mu, reward, done, step)

    def fused_step(self, step=0):
        """One learner step with everything resident: draw B rollouts uniformly without replacement
        (random.sample, baseline/utils.py:310-315), gather only a / mu / r / done (244 B of the 593 KB rollout),
        run conv_1 over the rollouts' (T+1) frames IN PLACE in the replay payload (row = slot * (T+1) + t,
        time-major), V-trace kernel, loss, backward, clip + RMSprop."""
        c = self.cfg
        T, B = c.UNROLL_STEP, c.BATCHSIZE
        mem = self._memory
        st = mem.store
        if not hasattr(self, "_small"):
            self._small = st.alloc_batch(B, ("action", "mu", "reward", "done"))
            self._frames = st.field_view("state").view(-1, 4, 84, 84)
            self._t_idx = torch.arange(T + 1, device=self.device).view(T + 1, 1)
        idx = mem.draw(B)
        b = st.gather(idx, self._small)
        rows = (idx.view(1, B) * (T + 1) + self._t_idx).reshape(-1).contiguous()
        self._train_core(self._frames, rows, b["action"].t().contiguous(), b["mu"].t().contiguous(),
                         b["reward"].t().contiguous(), b["done"], step)
        return self.last

    def _train_core(self, frames, rows, action, mu, reward, done, step):
        """IMPALA/Learner.py:121-235 on a uint8 frame table read in place (`rows`: time-major frame rows, None =
        all rows in order) or, with rows == "staged", on a staged (T+1, B, 28224) batch through PyTorch's conv_1."""
        c = self.cfg
        T, B, A = c.UNROLL_STEP, c.BATCHSIZE, c.ACTION_SIZE
        dev = self.device
        fused = not isinstance(rows, str)
        with torch.no_grad():
            if fused:
                # one launch: conv_1 of all (T+1)*B frame stacks, uint8 -> /255 folded in, no fp32 staging
                if not hasattr(self, "_pack1"):
                    self._conv_name = self.model.first_conv_node()
                    self._pack1 = R.Conv1Pack(1, dev, getattr(self.model, self._conv_name).conv_1.out_channels)
                w1 = getattr(self.model, self._conv_name).conv_1.weight
                self._pack1.pack(0, w1)
                y_all = R.conv1_fused(frames, rows, self._pack1, relu=False)[0]
                y_seq, y_last = y_all[:T * B], y_all[T * B:]
                out_last = self.model.forward_from_conv1(y_last, False)[0]
                out_seq = self.model.forward_from_conv1(y_seq, False)[0]
            else:
                s = frames.float().div_(255.0).view(T + 1, B, 4, 84, 84)        # :131-140
                last, seq = s[-1], s[:-1].reshape(-1, 4, 84, 84)
                out_last = self.model.forward([last])[0]
                out_seq = self.model.forward([seq])[0]
            boot = (out_last[:, -1] * done.float().view(-1)).contiguous()       # :143
            policy = torch.softmax(out_seq[:, :A], dim=-1)                      # forward(), :70-83
            pi_a = policy.gather(1, action.reshape(-1, 1).long())[:, 0]
            value = out_seq[:, -1]
            vt, adv = R.vtrace(pi_a.view(T, B).contiguous(), mu.float().view(T, B).contiguous(),
                               value.view(T, B).contiguous(), boot, reward.float().view(T, B).contiguous(),
                               c.GAMMA, c.C_LAMBDA, c.C_VALUE, c.P_VALUE)       # :151-215 in one launch
        # calLoss (:95-119): second forward with grad (conv_1's output is reused: same weights, same frames)
        if fused:
            seq_frames = frames[:T * B] if rows is None else frames
            seq_rows = None if rows is None else rows[:T * B]
            y = _Conv1Gathered.apply(w1, seq_frames, seq_rows, self._pack1, torch.contiguous_format, None, y_seq)
            out = self.model.forward_from_conv1(y, False)[0]
        else:
            out = self.model.forward([seq])[0]
        logp = torch.log_softmax(out[:, :A], dim=-1)
        p = logp.exp()
        entropy = -(p * logp).sum(-1, keepdim=True)
        sel = logp.gather(1, action.reshape(-1, 1).long())
        obj_actor = torch.mean(sel * adv.view(-1, 1) + c.ENTROPY_R * entropy)
        critic = torch.mean((out[:, -1] - vt.view(-1)).pow(2)) / 2
        self.mOptim.zero_grad(set_to_none=False)
        (-obj_actor + critic).backward()                                         # :223-225
        self.step(step)
        self.last = {"objActor": obj_actor.detach(), "criticLoss": critic.detach(), "vtarget": vt, "advantage": adv}

    def step(self, step=0):
        """IMPALA/Learner.py:258-266: clip at 40, RMSprop."""
        self.model.clippingNorm(40)
        self.mOptim.step()

    def state_dict(self):
        return ({k: v.cpu() for k, v in self.model.state_dict().items()},)

    def run(self, max_steps=None):
        """IMPALA/Learner.py:274-297: per step sample -> train, publish `params` (the 1-tuple of the state dict)
        and `Count` EVERY step (:286-287), checkpoint every 100 steps (:290-297).  Publication is asynchronous:
        the weights are snapshot on the learner stream and SET once their D2H copy has landed; if the previous
        snapshot is still in flight this step's is skipped (the actors poll every 400 env steps anyway)."""
        import time
        from . import wire
        from .publish import ParamPublisher
        while len(self._memory) <= self.cfg.BUFFER_SIZE:
            time.sleep(0.05)
        pub = ParamPublisher(self.model, self._connect, "params", "Count", wrap=lambda sd: (sd,))
        ckpt_path = wire.checkpoint_path(self.cfg.LOG_W)
        ckpt = ParamPublisher(self.model, None, None, None,
                              on_ready=lambda sd, step: torch.save(sd, ckpt_path)) if ckpt_path else None
        self._publishers = (pub,) + ((ckpt,) if ckpt else ())
        t = 0
        while max_steps is None or t < max_steps:
            tr = self._memory.sample()
            if tr is False:
                time.sleep(0.2)
                continue
            self.train(tr, t)
            pub.snapshot(t)
            if ckpt is not None and (t + 1) % 100 == 0:
                ckpt.snapshot(t)
            for p in self._publishers:
                p.poll()
            t += 1
        return t
