"""Per-role cycle counters of the fused conv_1 wgrad kernel (B2RL_CONV1_DBG=1) at the bench shape."""
import os, sys
os.environ["B2RL_CONV1_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_rl_b200 import replay as R
n, rows = 512, 4096
fr = torch.randint(0, 256, (rows, 4, 84, 84), dtype=torch.uint8, device="cuda")
idx = torch.randint(0, rows, (n,), device="cuda")
gy = (torch.randn(n, 32, 20, 20, device="cuda") * (torch.rand(n, 32, 20, 20, device="cuda") > 0.5)).contiguous(memory_format=torch.channels_last)
for _ in range(3):
    R.conv1_wgrad(fr, idx, gy)
torch.cuda.synchronize()
