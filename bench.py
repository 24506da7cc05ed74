#!/usr/bin/env python
"""bench.py — learner transitions/s for the Ape-X hot path (sample + gather +
target + priority update, inside a full learner step) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Ape-X DQN, 2^20-slot
device-resident sum-tree per GPU, synthetic (4,84,84) uint8 frame stacks (59.2 GB
payload per GPU), batch 512 per GPU.  One "step" = one learner step:
  tree sample (512) -> IS weights -> TMA gather of a, r, done (s and s' are read by the fused
  gather + conv_1 kernels straight from the replay payload) -> Q(s), Q(s'), Qbar(s') ->
  fused double-DQN n-step target / clipped TD / priority / dLoss/dQ -> backward ->
  centered RMSprop -> tree priority write-back
N > 1: one process per GPU, replay sharded (2^20 slots each, weak scaling),
NCCL all-reduce of the gradients (AVG) and of the max IS weight (MAX) — the
only inter-GPU traffic.

`value` = transitions/s with everything resident in HBM, the whole step replayed
as one CUDA graph (with parallel branches: the three forward passes, the weight
gradients, the tree update).  `e2e` = the same loop through the public Python API
with HOST buffers: every step ingests 512 new transitions from pinned host memory
(Replay.begin_ingest / commit_ingest -> b2rl_replay_reserve / copy_payload / commit,
the copy overlapping the step) and reads the step's scalars back.

--impl reference times the CPU port of the reference learner loop
(oracle/cpu_learner.py; the reference is pure Python and /root/reference does not
exist on the GPU box) on the host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "learner transitions/sec (sample+target+prio-update)"
UNIT = "transitions/s"
ALG_BYTES_PER_TRANSITION_GATHER = 2 * 28224 + 4 + 4 + 1   # 56 457 B read per transition


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=20, help="log2 of replay slots per GPU")
    ap.add_argument("--batch", type=int, default=512, help="batch per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nchw", action="store_true", help="keep the network in NCHW (default: channels_last)")
    ap.add_argument("--unfused-conv1", action="store_true", help="stage the batch and let cuDNN run conv_1")
    ap.add_argument("--cudnn-conv1-wgrad", action="store_true", help="conv_1 weight gradient through a staged fp32 copy + cuDNN instead of csrc/conv1_wgrad.cu")
    ap.add_argument("--inline-wgrad", action="store_true", help="weight gradients inline in backward instead of on a side stream")
    ap.add_argument("--serial-forwards", action="store_true", help="the three forward passes of a step on one stream")
    ap.add_argument("--unfused-tail", action="store_true", help="dueling tail as separate PyTorch ops instead of csrc/dueling.cu")
    ap.add_argument("--cublas-dense", action="store_true", help="dense heads as cuBLAS fp32 GEMMs instead of the 3xTF32 tcgen05 kernel (csrc/gemm.cu)")
    ap.add_argument("--torch-optim", action="store_true", help="torch.optim.RMSprop instead of the fused kernel")
    ap.add_argument("--no-cudnn-benchmark", action="store_true", help="leave cuDNN's algorithm choice to its heuristics")
    ap.add_argument("--blaslt", action="store_true", help="route fp32 GEMMs through cuBLASLt")
    ap.add_argument("--tf32-matmul", action="store_true",
                    help="INFORMATIONAL ONLY: let the dense heads use TF32 like cuDNN's convolutions already do "
                         "(PyTorch's default, which the reference runs, is fp32 matmul; the headline keeps fp32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=48)
    return ap.parse_args()


# --------------------------------------------------------------------------- #
# clocks                                                                        #
# --------------------------------------------------------------------------- #
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if "Active" in v and "Not" not in v:
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------- #
# reference arm / cpu baseline                                                  #
# --------------------------------------------------------------------------- #
def run_cpu_port(n_slots, batch, steps, warmup, threads=None):
    """`steps` train steps of the CPU port, sampled/assembled `m` at a time like the reference."""
    import torch
    from oracle.cpu_learner import CpuApexLearner

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = max(1, min(16, steps))
    L = CpuApexLearner(n_slots, batch, m=m, pool=2048, threads=cores)
    cycles = max(1, (steps + m - 1) // m)
    for _ in range(max(0, (warmup + m - 1) // m)):
        L.cycle()
    tot_t, tot_n, parts = 0.0, 0, {"t_buffer": 0.0, "t_train": 0.0, "t_update": 0.0}
    for _ in range(cycles):
        r = L.cycle()
        tot_t += r["t_total"]; tot_n += r["transitions"]
        for k in parts:
            parts[k] += r[k]
    return {"value": tot_n / tot_t, "seconds": tot_t, "transitions": tot_n, "cores": cores, "m": m,
            "cycles": cycles, "parts": {k: v / cycles for k, v in parts.items()}}


def best_cpu_port(n_slots, batch, steps, warmup):
    """The reference leaves torch's intra-op thread count at its default (= all cores).  On a
    many-core host that oversubscribes the small convolutions badly (measured: 64 tr/s at 128
    threads vs 3200 tr/s at 32), so a 2-step probe picks the fastest of a few thread counts and the
    bounded sample is then timed at that count (reported as `cores`)."""
    ncpu = os.cpu_count() or 1
    tried = {}
    for th in sorted({min(8, ncpu), min(32, ncpu), ncpu}):
        tried[th] = round(run_cpu_port(n_slots, batch, 2, 1, threads=th)["value"], 1)
    best_th = max(tried, key=tried.get)
    best = run_cpu_port(n_slots, batch, steps, warmup, threads=best_th)
    best["tried_threads"] = tried
    return best


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 1 << args.log2n
    steps = max(1, min(args.steps, 48))  # bounded sample: <= 48 train steps of 512 (about 10 s of CPU work)
    warm = min(args.warmup, 1)
    r = best_cpu_port(n, args.batch, steps, warm)
    sample = (f"{r['cycles']} cycle(s) of {r['m']} train steps x batch {args.batch} at N=2^{args.log2n} "
              f"priorities (payload pool of 2048 pickled records), after {warm} warm-up step(s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["cycles"] * r["m"], "warmup": warm, "ms_per_step": 1e3 * r["seconds"] / (r["cycles"] * r["m"]),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample,
                         "parts_s_per_cycle": r["parts"], "tried_threads_tr_per_s": r["tried_threads"],
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"Ape-X DQN learner step, 2^{args.log2n}-slot device sum-tree per GPU, synthetic (4,84,84) "
                        f"uint8 frame stacks, batch={args.batch} per GPU (BASELINE.json configs[1])",
            "slots_per_gpu": 1 << args.log2n, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "record_bytes": ALG_BYTES_PER_TRANSITION_GATHER,
            "parallelism": f"replay-sharded dp{world}" if world > 1 else "single GPU",
            "l2": "inputs >> L2: every step gathers random rows of a 59 GB payload (no L2 flush needed)",
            "network": "dueling DQN of cfg/ape_x.json; conv_1 forward and weight gradient fused with the gather on tcgen05 "
                       "(int8 digits, fp32-exact), dense heads as 3xTF32 tcgen05 GEMMs at fp32 accuracy, fused dueling tail; "
                       "conv_2/conv_3 in cuDNN at PyTorch's default precision (TF32 convs) = what the reference runs"}


# --------------------------------------------------------------------------- #
# our arm                                                                       #
# --------------------------------------------------------------------------- #
def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from distributed_rl_b200 import _lib, replay as R
    from distributed_rl_b200.apex import ApexConfig, Learner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # library knobs for the PyTorch remainder of the network (no precision change: fp32 matmul, TF32 conv)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    if args.cudnn_conv1_wgrad:
        from distributed_rl_b200.apex import _Conv1Gathered
        _Conv1Gathered.fused_wgrad = False
    if args.blaslt:
        torch.backends.cuda.preferred_blas_library("cublaslt")
    if args.tf32_matmul:
        torch.backends.cuda.matmul.allow_tf32 = True
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    N, B = 1 << args.log2n, args.batch
    cfg = ApexConfig(BATCHSIZE=B, REPLAY_MEMORY_LEN=N, BUFFER_SIZE=0, LEARNER_DEVICE=str(dev),
                     CHANNELS_LAST=not args.nchw, FUSED_CONV1=not args.unfused_conv1,
                     FUSED_OPTIM=not args.torch_optim, DENSE_3XTF32=not args.cublas_dense, FUSED_DUELING_TAIL=not args.unfused_tail, PARALLEL_FORWARDS=not args.serial_forwards, DEFERRED_WGRAD=not args.inline_wgrad)
    torch.manual_seed(0)
    learner = Learner(cfg, connect=None, start_replay=False)
    if world > 1:   # identical initial weights on every rank
        for p in list(learner.model.parameters()) + list(learner.target_model.parameters()):
            dist.broadcast(p.data, 0)
        learner.enable_data_parallel()
    store = learner.memory.store
    # ---- pre-fill: synthetic frames by counter hash, typed scalars, priorities ----
    store.fill_hash(N, seed=0xB200 + rank)
    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 1 + rank)
    store.field_view("action").copy_(torch.randint(0, cfg.ACTION_SIZE, (N,), device=dev, generator=g, dtype=torch.int32))
    store.field_view("reward").copy_(torch.randn(N, device=dev, generator=g).clamp_(-1, 1))
    store.field_view("done").copy_((torch.rand(N, device=dev, generator=g) < 0.02).to(torch.uint8))
    prios = (torch.randn(N, device=dev, generator=g).abs().clamp(max=1) + 1e-7) ** cfg.ALPHA
    store.build(prios)
    store.seed(1234 + rank, 0)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: the whole step as one CUDA graph --------------------------
    use_graph = not args.no_graph
    for _ in range(max(3, args.warmup)):     # first call builds (3 eager warm-ups + capture)
        learner.fused_step(use_graph=use_graph)
    per_step_launches = learner.launches_per_step
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = learner.fused_step(use_graph=use_graph)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clock_info = clocks.stop() if clocks else None
    value = B * world * args.steps / (ms / 1e3)
    scal = out["scalars"].tolist()

    # ---- dominant hand-written kernels, each timed alone with CUDA events (graph of `reps`
    #      launches on distinct index sets -> no Python launch overhead, no L2 reuse of the rows) ----
    reps = 20
    idxs = [store.sample(B, beta=cfg.BETA, want_prob=False)[0] for _ in range(reps)]

    def time_graph(fn):
        for i in range(3):
            fn(idxs[i])
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(reps):
                fn(idxs[i])
        gr.replay(); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); gr.replay(); t1.record(); torch.cuda.synchronize()
        del gr
        return t0.elapsed_time(t1) * 1e3 / reps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)" if peaks else "fallback 6650"
    outb = store.alloc_batch(B)
    gather_us = time_graph(lambda ix: store.gather(ix, outb))
    g_ach = ALG_BYTES_PER_TRANSITION_GATHER * B / (gather_us * 1e-6) / 1e9
    kernels = {"k_gather_bulk": {"launch_us": gather_us, "algorithmic_bytes_per_launch": ALG_BYTES_PER_TRANSITION_GATHER * B,
                                 "achieved_GBs": g_ach, "frac": g_ach / peak,
                                 "note": "whole minibatch (s, s', a, r, done) staged in one launch; a copy moves 2x its "
                                         "algorithmic read bytes, so frac <= 0.5 for an unfused gather"}}
    roofline = {"kernel": "k_gather_bulk — TMA bulk gather of one minibatch", "bound": "hbm", "achieved": g_ach,
                "peak": peak, "unit": "GB/s", "frac": g_ach / peak, "peak_source": peak_src, "traffic": None,
                "launch_us": gather_us, "algorithmic_bytes_per_launch": ALG_BYTES_PER_TRANSITION_GATHER * B}
    if cfg.FUSED_CONV1 and learner._conv1_ready():
        # fused gather + conv_1 for online+target nets over s': reads B frame stacks, writes 2 x B x (20,20,32) fp32
        alg = B * (28224 + 2 * 400 * 32 * 4)
        out2 = torch.empty((2, B, 20, 20, 32), device=dev)
        nsf = store.field_view("next_state")
        c_us = time_graph(lambda ix: R.conv1_fused(nsf, ix, learner._pack2, relu=True, out=out2))
        c_ach = alg / (c_us * 1e-6) / 1e9
        ops = 2.0 * (B * 400) * 256 * 256        # useful MACs x2 incl. the 4 weight digits (N=256 columns)
        kernels["k_conv1_fused<2>"] = {"launch_us": c_us, "algorithmic_bytes_per_launch": alg, "achieved_GBs": c_ach,
                                       "frac": c_ach / peak, "int8_TOPS": ops / (c_us * 1e-6) / 1e12,
                                       "note": "reads 28 224 B per sampled s' directly from the replay payload "
                                               "(no staging copy) and writes conv_1 activations of both networks"}
        roofline = {"kernel": "k_conv1_fused<2> — fused TMA gather + im2col + tcgen05 conv_1 (online+target) of s'",
                    "bound": "hbm", "achieved": c_ach, "peak": peak, "unit": "GB/s", "frac": c_ach / peak,
                    "peak_source": peak_src, "traffic": None, "launch_us": c_us, "algorithmic_bytes_per_launch": alg,
                    "note": "algorithmic bytes = sampled frames read (SURVEY §8d: 28 224 B per frame stack) + the two "
                            "fp32 NHWC activation maps written; currently epilogue/issue-bound, not HBM-bound "
                            "(tensor pipe 21.6 % active, profiles/r01_conv1.md)"}
    if cfg.FUSED_CONV1 and learner._conv1_ready():
        # fused gather + conv_1 weight gradient: reads B frame stacks + dL/dy (B x 400 x 32 fp32), writes 8192 floats
        alg_w = B * (28224 + 400 * 32 * 4)
        gyw = torch.randn(B, 32, 20, 20, device=dev).contiguous(memory_format=torch.channels_last)
        sf = store.field_view("state")
        w_us = time_graph(lambda ix: R.conv1_wgrad(sf, ix, gyw))
        w_ach = alg_w / (w_us * 1e-6) / 1e9
        kernels["k_conv1_wgrad"] = {"launch_us": w_us, "algorithmic_bytes_per_launch": alg_w, "achieved_GBs": w_ach,
                                    "frac": w_ach / peak, "int8_TOPS": 2.0 * (B * 400) * 256 * 128 / (w_us * 1e-6) / 1e12,
                                    "note": "launch_us covers k_conv1_wgrad + k_conv1_wgrad_reduce; reads each sampled s "
                                            "and its dL/dy once (dL/dy twice: scale pre-scan), no fp32 staging of the frames"}
    if cfg.DENSE_3XTF32:
        # the dominant kernel by device time: the 3xTF32 GEMM of the fused 3136 -> 2x512 heads (forward shape)
        from distributed_rl_b200 import linear as LIN
        Mg, Ng, Kg = B, 1024, 3136
        xa = LIN.split_pack(torch.randn(Mg, Kg, device=dev), False, False)
        wb = LIN.split_pack(torch.randn(Ng, Kg, device=dev) * 0.02, False, True)
        og = torch.empty(Mg, Ng, device=dev)
        g_us = time_graph(lambda ix: LIN.gemm_packed(xa, wb, Mg, Ng, Kg, out=og))
        tpeak = float(peaks.get("bf16_tflops", 1719.3))
        alg_fl = 2.0 * Mg * Ng * Kg
        t_ach = alg_fl / (g_us * 1e-6) / 1e12
        kernels["k_gemm_tf32x3"] = {"launch_us": g_us, "algorithmic_flops_per_launch": alg_fl, "achieved_TFLOPs": t_ach,
                                    "frac": t_ach / tpeak, "tf32_TFLOPs_executed": 3 * t_ach,
                                    "frac_of_tf32_peak_est": 3 * t_ach / (tpeak / 2),
                                    "note": "launch_us covers k_gemm_tf32x3 + k_splitk_reduce for x[512x3136] @ W[1024x3136]^T; "
                                            "algorithmic flops = the fp32 GEMM (2MNK); the kernel executes 3 TF32 products per "
                                            "term pair, and TF32 dense peak is half the measured bf16 peak"}
        roofline = {"kernel": "k_gemm_tf32x3 — fp32-accurate dense heads as 3xTF32 tcgen05 GEMM (largest share of the step)",
                    "bound": "tensor", "achieved": t_ach, "peak": tpeak, "unit": "TFLOP/s", "frac": t_ach / tpeak,
                    "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback 1719.3",
                    "traffic": None, "launch_us": g_us, "algorithmic_flops_per_launch": alg_fl,
                    "note": "achieved counts the fp32 GEMM's 2MNK flops once; the tensor pipe executes 3x that in TF32 "
                            "(tf32_TFLOPs_executed), whose dense peak is bf16/2 — see kernels[k_gemm_tf32x3]"}
    roofline["kernels"] = kernels
    prof = os.path.join(REPO, "profiles", "r01_traffic.json")
    if os.path.isfile(prof):
        try:
            tr = json.load(open(prof))
            key = next((k for k in ("k_gemm_tf32x3", "k_conv1_fused<2>", "k_gather_bulk") if k in kernels and k in tr),
                       None)
            roofline["traffic"] = tr.get(key) if key else None
        except Exception:
            pass

    # ---- e2e: public API, host buffers in, scalars out -----------------------------------
    from distributed_rl_b200.hostmem import pinned_like, pinned_empty   # pinned pages on the GPU's NUMA node
    pin = lambda t: pinned_like(t, dev)
    rng = np.random.default_rng(7 + rank)
    hs = pin(torch.from_numpy(rng.integers(0, 256, size=(B, 4, 84, 84), dtype=np.uint8)))
    hns = pin(torch.from_numpy(rng.integers(0, 256, size=(B, 4, 84, 84), dtype=np.uint8)))
    ha = pin(torch.from_numpy(rng.integers(0, 6, size=B).astype(np.int32)))
    hr = pin(torch.from_numpy(np.clip(rng.standard_normal(B), -1, 1).astype(np.float32)))
    hd = pin(torch.from_numpy((rng.random(B) < 0.02).astype(np.uint8)))
    hp = pin(torch.ones(B, dtype=torch.float32))
    h2d = sum(t.numel() * t.element_size() for t in (hs, hns, ha, hr, hd, hp))
    host_scal = [pinned_empty(3, torch.float32, dev) for _ in range(2)]
    d2h_stream = torch.cuda.Stream(dev)
    d2h_done = [torch.cuda.Event(), torch.cuda.Event()]
    step_done = torch.cuda.Event()
    seen = {"n": 0, "loss": 0.0}

    # Pipelined ingest: the copy of the NEXT 512 transitions runs on the ingest stream while the
    # current learner step computes; every step still moves its own 29 MB H2D inside the timed region.
    # Every step's scalars (loss, mean target, mean weight) are read back to pinned host memory on a
    # D2H stream and consumed by the host one step later, so the host never idles the GPU.
    learner.memory.begin_ingest(hs, hns, ha, hr, hd)

    def e2e_step(i=[0]):
        k = i[0] & 1
        if i[0] > 0:   # the previous step's 12-byte read must leave `scalars` before the graph rewrites it
            torch.cuda.current_stream(dev).wait_event(d2h_done[k ^ 1])
        learner.memory.commit_ingest(hp)                            # copy done -> priorities -> sampleable
        learner.memory.begin_ingest(hs, hns, ha, hr, hd)            # H2D of the next 512 transitions (async)
        o = learner.fused_step(use_graph=use_graph)
        step_done.record(torch.cuda.current_stream(dev))
        if i[0] > 0:                                                # consume the PREVIOUS step's result
            d2h_done[k ^ 1].synchronize()
            seen["n"] += 1; seen["loss"] = float(host_scal[k ^ 1][0])
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(step_done)
            host_scal[k].copy_(o["scalars"], non_blocking=True)     # D2H of this step's loss / mean target / mean w
            d2h_done[k].record(d2h_stream)
        i[0] += 1

    # Warm-up: the PCIe link reaches its full rate only after ~0.2 s of sustained traffic (tools/h2d_probe.py:
    # 25 -> 47 -> 55 GB/s over the first three 150-copy bursts), and it idles during the device-resident
    # region above, so the steady-state loop is entered with enough untimed steps to move ~5 GB first.
    # The host thread that drives the loop is bound to the GPU's NUMA node (numactl --cpunodebind in a
    # deployment): every step makes ~10 driver calls whose doorbell writes cross the socket interconnect otherwise.
    from distributed_rl_b200.hostmem import on_gpu_node
    with on_gpu_node(dev) as bound:
        for _ in range(max(3, args.warmup) + 170):
            e2e_step()
        barrier()
        k2 = max(10, args.steps // 2)
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        th0 = time.perf_counter()
        for _ in range(k2):
            e2e_step()
        host_ms = (time.perf_counter() - th0) * 1e3
        s1.record()
    d2h_done[0].synchronize(); d2h_done[1].synchronize()   # the last step's result has been read too
    barrier()
    ms2 = s0.elapsed_time(s1)
    if world > 1:
        t = torch.tensor([ms2], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    e2e = {"value": B * world * k2 / (ms2 / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 12, "steps": k2,
           "h2d_GBs": h2d * k2 / (ms2 / 1e3) / 1e9, "host_ms_per_step": host_ms / k2,
           "host_thread_bound_to_gpu_numa_node": bool(bound),
           "what": "Replay.commit_ingest + begin_ingest (512 new transitions from pinned host, H2D on the ingest "
                   "stream overlapping the step) + Learner.fused_step() + per-step D2H of the step's scalars to pinned "
                   "host memory (consumed by the host one step later)"}

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = best_cpu_port(N, B, args.cpu_steps, 1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "tried_threads_tr_per_s": r["tried_threads"], "host_cpus": os.cpu_count(),
               "sample": f"{r['cycles']} cycle(s) x {r['m']} train steps x batch {B} at N=2^{args.log2n} priorities "
                         f"(pool of 2048 pickled records), {r['seconds']:.1f} s of CPU work",
               "parts_s_per_cycle": r["parts"]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cpu,
                "e2e": e2e, "gpu_launches": int(per_step_launches * args.steps), "clocks": clock_info,
                "cuda_graph": use_graph, "fused_gather_conv1": bool(cfg.FUSED_CONV1), "fused_optimizer": bool(cfg.FUSED_OPTIM), "tf32_matmul": bool(args.tf32_matmul), "dense_3xtf32": bool(cfg.DENSE_3XTF32), "fused_conv1_wgrad": not args.cudnn_conv1_wgrad, "fused_dueling_tail": bool(cfg.FUSED_DUELING_TAIL), "parallel_forwards": bool(cfg.PARALLEL_FORWARDS), "deferred_wgrad": bool(cfg.DEFERRED_WGRAD), "last_step": {"loss": scal[0], "mean_target": scal[1], "mean_weight": scal[2]}}
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if world > 1:
        # CUDA graphs that hold NCCL kernels make destroy_process_group()/interpreter teardown hang:
        # drop them, drain the device, leave together, and exit without running destructors.
        learner._graph = None
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
