"""Eager launches of the sum-tree kernels for `ncu --set full` (no CUDA graph): build at 2^23, then at 2^20
sample 512 / update 512 scattered / update 65536 scattered / sample 2^20."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_rl_b200 import replay as R

big = R.DeviceReplay(1 << 23, fields=(), device="cuda:0")
pb = torch.rand(1 << 23, device="cuda") + 0.01
for _ in range(3):
    big.build(pb)
torch.cuda.synchronize()
big.close(); del pb
N = 1 << 20
rep = R.DeviceReplay(N, fields=(R.Field("a", torch.int32, ()),), device="cuda:0")
rep.build(torch.rand(N, device="cuda") + 0.01)
ui = torch.randint(0, N, (512,), device="cuda"); uv = torch.rand(512, device="cuda") + 0.01
ul = torch.randint(0, N, (65536,), device="cuda"); vl = torch.rand(65536, device="cuda") + 0.01
for _ in range(3):
    rep.sample(512)
    rep.update(ui, uv)
    rep.update(ul, vl)
    rep.sample(1 << 20)
torch.cuda.synchronize()
