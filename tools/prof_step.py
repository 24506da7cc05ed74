"""Warm (back-to-back, non-ncu) per-kernel breakdown of the eager learner step via torch.profiler (CUPTI).
Usage: python tools/prof_step.py [--cublas-dense]   -> prints kernels sorted by total device time per step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_rl_b200.apex import ApexConfig, Learner  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
N, B = 1 << 20, 512
cfg = ApexConfig(BATCHSIZE=B, REPLAY_MEMORY_LEN=N, BUFFER_SIZE=0, LEARNER_DEVICE=str(dev),
                 DENSE_3XTF32="--cublas-dense" not in sys.argv)
torch.manual_seed(0)
learner = Learner(cfg, connect=None, start_replay=False)
st = learner.memory.store
st.fill_hash(N)
st.build((torch.rand(N, device=dev) + 1e-3) ** 0.6)
for _ in range(10):
    learner.fused_step(use_graph=False)
torch.cuda.synchronize()
STEPS = 20
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        learner.fused_step(use_graph=False)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / STEPS, e.count / STEPS) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"# sum of kernel time per step: {tot:.1f} us over {sum(r[2] for r in rows):.0f} launches")
for k, t, n in rows[:45]:
    print(f"{t:8.1f} {100 * t / tot:5.1f}% {n:5.1f}  {k[:110]}")
