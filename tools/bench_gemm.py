"""Time the 3xTF32 GEMM pieces against cuBLAS fp32 on the learner's head shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_rl_b200 import linear as L

def t(fn, it=50):
    for _ in range(5): fn()
    flush = torch.empty(64 << 20, device="cuda", dtype=torch.float32)
    e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in e:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in e)
    return ts[len(ts) // 2]

for (M, N, K) in [(512, 1024, 3136), (512, 3136, 1024), (1024, 3136, 512)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda")
    a = L.split_pack(x, False, False); b = L.split_pack(w, False, True)
    print(M, N, K, "cublas fp32 %.1f us" % t(lambda: x @ w.T),
          "pack A %.1f" % t(lambda: L.split_pack(x, False, False)),
          "pack B %.1f" % t(lambda: L.split_pack(w, False, True)),
          "packT B %.1f" % t(lambda: L.split_pack(w.T.contiguous(), True, True)),
          "gemm %.1f" % t(lambda: L.gemm_packed(a, b, M, N, K)),
          "linear3x %.1f" % t(lambda: L.linear3x(x, w)))


def warm(fn, reps=20):
    """graph of `reps` back-to-back launches, operands L2-warm (the learner step's situation)"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / reps)
    return best


print("warm (graph x20), multicast env =", os.environ.get("B2RL_GEMM_MULTICAST", "default(on)"))
for (M, N, K) in [(512, 1024, 3136), (512, 3136, 1024), (1024, 3136, 512)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda")
    a = L.split_pack(x, False, False); b = L.split_pack(w, False, True)
    out = torch.empty(M, N, device="cuda")
    print(M, N, K, "gemm+reduce warm %.2f us" % warm(lambda: L.gemm_packed(a, b, M, N, K, out=out)))
