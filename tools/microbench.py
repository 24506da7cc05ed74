"""Per-kernel timings (CUDA events, L2 flushed between iterations) at the
BASELINE.json sizes.  Diagnostic tool; bench.py is the judged harness."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_rl_b200 import replay as R  # noqa: E402


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--payload", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    N = 1 << args.log2n
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fields = R.APEX_FIELDS if args.payload else ()
    rep = R.DeviceReplay(N, fields=fields, device=dev)
    if args.payload:
        rep.fill_hash(N)
    p = (torch.randn(N, device=dev).abs().clamp(max=1) + 1e-7) ** 0.6
    res = {"N": N}
    med, best = timeit(lambda: rep.build(p), flush=flush)
    res["build_us"] = med; res["build_GBs_8N"] = 8 * N / med / 1e3
    for n in (512, 8192, 65536, 1 << 20):
        idx = torch.empty(n, dtype=torch.int64, device=dev); pr = torch.empty(n, device=dev); w = torch.empty(n, device=dev)
        med, best = timeit(lambda: rep.sample(n, out=(idx, pr, w)), flush=flush)
        res[f"sample_{n}_us"] = med
        res[f"sample_{n}_GBs_88B"] = 88 * n / med / 1e3
        vals = torch.rand(n, device=dev) + 0.01
        med, best = timeit(lambda: rep.update(idx, vals), flush=flush)
        res[f"update_{n}_us"] = med
        res[f"update_{n}_GBs_164B"] = 164 * n / med / 1e3
    if args.payload:
        for n in (512, 8192):
            idx, _, _ = rep.sample(n)
            out = rep.alloc_batch(n)
            for mode in ("bulk", "ldg"):
                os.environ["B2RL_GATHER"] = mode  # read once per process; second value needs a fresh run
                med, best = timeit(lambda: rep.gather(idx, out), flush=flush)
                res[f"gather_{n}_us"] = med
                res[f"gather_{n}_GBs_alg"] = 56457 * n / med / 1e3
                break
    B, A = 512, 6
    q = [torch.randn(B, A, device=dev) for _ in range(3)]
    a = torch.randint(0, A, (B,), device=dev); r = torch.randn(B, device=dev); nd = torch.ones(B, device=dev); w = torch.rand(B, device=dev)
    out = R.apex_target(*q, a, r, nd, w, 0.97, 0.6)
    med, _ = timeit(lambda: R.apex_target(*q, a, r, nd, w, 0.97, 0.6, out=out))
    res["apex_target_512_us"] = med
    L, Bq = 60, 64
    q2 = [torch.randn(L, Bq, A, device=dev) for _ in range(2)]
    a2 = torch.randint(0, A, (L - 1, Bq), device=dev); r2 = torch.randn(L - 1, Bq, device=dev)
    med, _ = timeit(lambda: R.r2d2_target(q2[0], q2[1], a2, r2, torch.ones(Bq, device=dev), torch.rand(Bq, device=dev), 5, 0.997, 0.9))
    res["r2d2_target_64x60_us"] = med
    T, Bv = 20, 1024
    x = [torch.rand(T, Bv, device=dev) * 0.8 + 0.1 for _ in range(4)]
    med, _ = timeit(lambda: R.vtrace(x[0], x[1], x[2], torch.rand(Bv, device=dev), x[3], 0.99, 1.0, 1.0, 1.0))
    res["vtrace_20x1024_us"] = med
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/microbench_{os.environ.get('B2RL_GATHER', 'bulk')}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
