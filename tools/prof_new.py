"""Short driver for ncu captures of the round-1 late kernels at the bench shapes:
k_gemm_tf32x3 (fwd / dgrad / wgrad shapes of the heads), k_conv1_wgrad<32>, k_dueling_*."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_rl_b200 import linear as L, replay as R
dev = "cuda:0"
x = torch.randn(512, 3136, device=dev, requires_grad=True)
w1 = (torch.randn(512, 3136, device=dev) * 0.02).requires_grad_()
w2 = (torch.randn(512, 3136, device=dev) * 0.02).requires_grad_()
wa = (torch.randn(6, 512, device=dev) * 0.05).requires_grad_()
wv = (torch.randn(1, 512, device=dev) * 0.05).requires_grad_()
fr = torch.randint(0, 256, (4096, 4, 84, 84), dtype=torch.uint8, device=dev)
idx = torch.randint(0, 4096, (512,), device=dev)
gy = (torch.randn(512, 32, 20, 20, device=dev) * (torch.rand(512, 32, 20, 20, device=dev) > 0.5)).contiguous(memory_format=torch.channels_last)
for _ in range(3):
    h = L.linear3x(x, [w1, w2])
    q = L.dueling_tail(h, wa, wv)
    q.square().sum().backward()
    R.conv1_wgrad(fr, idx, gy)
torch.cuda.synchronize()
print("done")
