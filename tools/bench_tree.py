"""Microbenchmark of the sum-tree kernels (csrc/tree.cu): CUDA-graph of `reps` launches, CUDA events.
    python tools/bench_tree.py [log2n ...]   -> JSON on stdout"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_rl_b200 import replay as R


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / reps)
    return round(best, 2)


def main():
    out = {}
    for lg in [int(x) for x in sys.argv[1:]] or [20, 23]:
        N = 1 << lg
        rep = R.DeviceReplay(N, fields=(R.Field("a", torch.int32, ()),), device="cuda:0")
        p = torch.rand(N, device="cuda") + 0.01
        rep.build(p)
        r = {}
        for n in (512, 8192, 65536, 1 << 20):
            oi = torch.empty(n, dtype=torch.int64, device="cuda"); ow = torch.empty(n, device="cuda")
            r[f"sample_{n}"] = timed(lambda: rep.sample(n, want_prob=False, out=(oi, None, ow)), 10)
        oi = torch.empty(512, dtype=torch.int64, device="cuda"); ow = torch.empty(512, device="cuda")
        oa = torch.empty(512, dtype=torch.int32, device="cuda")
        r["sample_fetch_512"] = timed(lambda: rep.sample_fetch(512, 0.4, oi, ow, {"a": oa}), 10)
        for n in (512, 1024, 8192, 65536):
            ui = torch.randint(0, N, (n,), device="cuda"); uv = torch.rand(n, device="cuda") + 0.01
            r[f"update_scattered_{n}"] = timed(lambda: rep.update(ui, uv), 10)
        x = torch.zeros(512, dtype=torch.int32, device="cuda"); pr = torch.rand(512, device="cuda") + 0.01
        r["push_512_dev(ring update + D2D)"] = timed(lambda: rep.push([x], pr), 10)
        x2 = torch.zeros(65536, dtype=torch.int32, device="cuda"); pr2 = torch.rand(65536, device="cuda") + 0.01
        r["push_65536_dev(large ring update + D2D)"] = timed(lambda: rep.push([x2], pr2), 10)
        r["build"] = timed(lambda: rep.build(p), 10)
        r["build_GBs_on_8N"] = round(8 * N / (r["build"] * 1e-6) / 1e9, 1)
        out[f"2^{lg}"] = r
        rep.close(); del p
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
