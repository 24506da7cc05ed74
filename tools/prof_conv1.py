"""ncu target: a few launches of the fused gather+conv1 kernel at bench size."""
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
for nn_ in (2, 1):
    pack = R.Conv1Pack(nn_, dev)
    for i in range(nn_):
        pack.pack(i, w)
    for it in range(3):
        idx = rep.sample(512)[0]
        R.conv1_fused(rep.field_view("next_state"), idx, pack, relu=True)
torch.cuda.synchronize()
print("done")
