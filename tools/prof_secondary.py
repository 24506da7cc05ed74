"""Warm per-kernel breakdown (torch.profiler / CUPTI) of the resident R2D2 or IMPALA learner step.
Usage: python tools/prof_secondary.py r2d2|impala   -> kernels sorted by device time per step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1] if len(sys.argv) > 1 else "r2d2"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
torch.backends.cudnn.benchmark = True
g = torch.Generator(device=dev); g.manual_seed(1)
if which == "r2d2":
    from distributed_rl_b200 import r2d2
    N, P, T, B = 1 << 20, 1 << 11, 80, 64
    cfg = r2d2.R2D2Config(BATCHSIZE=B, REPLAY_MEMORY_LEN=N, BUFFER_SIZE=0, PAYLOAD_POOL=P, FIXED_TRAJECTORY=T, MEM=20)
    torch.manual_seed(0)
    L = r2d2.Learner(cfg)
    pool, tree = L.memory.pool, L.memory.store
    pool.fill_hash(P, seed=3)
    pool.field_view("action").copy_(torch.randint(0, 6, (P, T), device=dev, generator=g, dtype=torch.int32))
    pool.field_view("reward").copy_(torch.randn(P, T, device=dev, generator=g))
    pool.field_view("h0").copy_(torch.randn(P, 512, device=dev, generator=g) * 0.1)
    pool.field_view("h1").copy_(torch.randn(P, 512, device=dev, generator=g) * 0.1)
    pool.field_view("notdone").fill_(1.0)
    tree.build(torch.rand(N, device=dev) + 0.01)
else:
    from distributed_rl_b200 import impala
    cap, T, B = 1 << 13, 20, 1024
    cfg = impala.ImpalaConfig(BATCHSIZE=B, REPLAY_MEMORY_LEN=cap, BUFFER_SIZE=0, UNROLL_STEP=T)
    torch.manual_seed(0)
    L = impala.Learner(cfg)
    st = L._memory.store
    st.fill_hash(cap, seed=4)
    st.field_view("action").copy_(torch.randint(0, 6, (cap, T), device=dev, generator=g, dtype=torch.int32))
    st.field_view("mu").copy_(torch.rand(cap, T, device=dev, generator=g) * 0.85 + 0.05)
    st.field_view("reward").copy_(torch.randn(cap, T, device=dev, generator=g))
    st.field_view("done").fill_(1.0)
    st.build(torch.ones(cap, device=dev))
for _ in range(5):
    L.fused_step()
torch.cuda.synchronize()
STEPS = 5
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        L.fused_step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / STEPS, e.count / STEPS) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"# {which}: sum of kernel time per step: {tot:.1f} us over {sum(r[2] for r in rows):.0f} launches")
for k, t, n in rows[:30]:
    print(f"{t:9.1f} {100 * t / tot:5.1f}% {n:6.1f}  {k[:120]}")
