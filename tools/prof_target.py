"""Short driver for ncu captures: a few launches of each hand-written kernel at the
bench sizes (N = 2^20 slots, Ape-X payload, batch 512 and 8192)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_rl_b200 import replay as R  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
N = 1 << log2n
rep = R.DeviceReplay(N, fields=R.APEX_FIELDS, device=dev)
rep.fill_hash(N)
p = (torch.randn(N, device=dev).abs().clamp(max=1) + 1e-7) ** 0.6
rep.build(p)
for n in (512, 8192):
    out = rep.alloc_batch(n)
    for it in range(4):
        idx, prob, w = rep.sample(n)
        rep.gather(idx, out)
        rep.update(idx, torch.rand(n, device=dev) + 0.01)
torch.cuda.synchronize()
print("done")
