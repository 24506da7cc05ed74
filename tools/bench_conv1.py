"""CUDA-event times of the fused gather + conv_1 kernels (forward, 1 and 2 networks; weight gradient)
at the bench shape: 512 sampled frame stacks out of a 2^16-slot Ape-X payload."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_rl_b200 import replay as R

dev = torch.device("cuda:0")
N = 1 << 16
rep = R.DeviceReplay(N, fields=R.APEX_FIELDS, device=dev)
rep.fill_hash(N)
rep.build(torch.rand(N, device=dev) + 0.1)
w = torch.empty(32, 4, 8, 8, device=dev).uniform_(-0.06, 0.06)
idxs = [rep.sample(512)[0].clone() for _ in range(8)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=40):
    for i in range(5):
        fn(i)
    ts = []
    for i in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for nn_ in (2, 1):
    pack = R.Conv1Pack(nn_, dev)
    for i in range(nn_):
        pack.pack(i, w)
    f = rep.field_view("next_state")
    med, mn = timed(lambda i: R.conv1_fused(f, idxs[i % 8], pack, relu=True))
    print(f"conv1_fused n_nets={nn_} n=512: median {med:.1f} us  min {mn:.1f} us (L2 flushed between launches)")
if hasattr(R, "conv1_wgrad"):
    gy = torch.randn(512, 32, 20, 20, device=dev).contiguous(memory_format=torch.channels_last)
    f = rep.field_view("state")
    try:
        med, mn = timed(lambda i: R.conv1_wgrad(f, idxs[i % 8], gy))
        print(f"conv1_wgrad n=512: median {med:.1f} us  min {mn:.1f} us")
    except Exception as e:  # signature differs between builds: report, do not hide
        print("conv1_wgrad not timed:", repr(e))
