"""Experiment: fp32 SIMT vs TF32 vs 3xTF32-split for the 3136->512 linear layers (B=512)."""
import torch, time
dev = "cuda"
torch.manual_seed(0)
B, K, N = 512, 3136, 512
x = torch.randn(B, K, device=dev); w = torch.randn(N, K, device=dev) * 0.02; gy = torch.randn(B, N, device=dev)

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / reps

def split(t):
    hi = (t.view(torch.int32) & -8192).view(torch.float32)   # keep 10 mantissa bits
    return hi, t - hi

def fwd_bwd_plain():
    y = x @ w.t(); dx = gy @ w; dw = gy.t() @ x
    return y, dx, dw

def mm3(a, b):   # a @ b with 3xTF32
    ah, al = split(a); bh, bl = split(b)
    out = ah @ bh
    out.addmm_(ah, bl); out.addmm_(al, bh)
    return out

def fwd_bwd_3x():
    return mm3(x, w.t().contiguous() if False else w.t()), mm3(gy, w), mm3(gy.t(), x)

ref = [t.double() for t in (x, w, gy)]
yr = ref[0] @ ref[1].t(); dxr = ref[2] @ ref[1]; dwr = ref[2].t() @ ref[0]
def err(a, b): return ((a.double() - b).abs().max() / b.abs().max()).item()

torch.backends.cuda.matmul.allow_tf32 = False
t_fp32 = timeit(fwd_bwd_plain); e_fp32 = [err(a, b) for a, b in zip(fwd_bwd_plain(), (yr, dxr, dwr))]
torch.backends.cuda.matmul.allow_tf32 = True
t_tf32 = timeit(fwd_bwd_plain); e_tf32 = [err(a, b) for a, b in zip(fwd_bwd_plain(), (yr, dxr, dwr))]
t_3x = timeit(fwd_bwd_3x); e_3x = [err(a, b) for a, b in zip(fwd_bwd_3x(), (yr, dxr, dwr))]
print(f"fp32 SIMT  {t_fp32:7.1f} us  err {e_fp32}")
print(f"TF32       {t_tf32:7.1f} us  err {e_tf32}")
print(f"3xTF32     {t_3x:7.1f} us  err {e_3x}")
