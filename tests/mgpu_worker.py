"""Worker of tests/test_gpu_20_multigpu.py — run under torchrun with >= 2 ranks, one GPU each.

Checks SURVEY.md §8e's parity statement on real GPUs over NCCL:
  1. every rank samples B transitions locally from its own shard with IS weights normalised by the GLOBAL max
     weight (one MAX all-reduce); per-shard indices / weights equal a single-process run on that shard;
  2. the gradient after the flat-bucket all-reduce (AVG) equals the mean of the per-shard gradients = the gradient
     of the loss over the concatenated batch, computed by ONE process on all shards (<= 1e-5 relative, per tensor);
  3. in the step loop the MAX all-reduce runs one step behind: step k's IS weights use the maximum reduced during
     step k-1 (the reference's own max_weight is up to 16 minibatches stale, APE_X/ReplayMemory.py:61-67).
Prints MGPU_OK on every rank when all checks hold."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from distributed_rl_b200 import apex, dist as D

    B, N = 64, 8192
    cfg = apex.ApexConfig(BATCHSIZE=B, REPLAY_MEMORY_LEN=N, BUFFER_SIZE=0, LEARNER_DEVICE=str(dev),
                          CUDNN_BENCHMARK=False)

    def make_learner():
        torch.manual_seed(0)                                  # identical weights everywhere
        L = apex.Learner(cfg, connect=None, start_replay=False)
        g = torch.Generator(device=dev); g.manual_seed(7)
        with torch.no_grad():
            for p in L.target_model.parameters():
                p.add_(0.01 * torch.randn(p.shape, device=dev, generator=g))
        return L

    def fill(L, shard):
        st = L.memory.store
        st.fill_hash(N, seed=100 + shard)
        g = torch.Generator(device=dev); g.manual_seed(200 + shard)
        st.field_view("action").copy_(torch.randint(0, 6, (N,), device=dev, generator=g, dtype=torch.int32))
        st.field_view("reward").copy_(torch.randn(N, device=dev, generator=g).clamp_(-1, 1))
        st.field_view("done").copy_((torch.rand(N, device=dev, generator=g) < 0.1).to(torch.uint8))
        st.build((torch.randn(N, device=dev, generator=g).abs().clamp(max=1) + 1e-7) ** 0.6 * (1.0 + 0.5 * shard))
        st.seed(500 + shard, 0)

    def one_step_grads(L, max_w):
        st = L.memory.store
        assert L._conv1_ready()
        idx, _, w = st.sample(B, beta=cfg.BETA, want_prob=False, max_w=max_w)
        b = st.gather(idx, st.alloc_batch(B, ("action", "reward", "done")))
        L._forward_backward_fused(idx, b["action"].to(torch.int64), b["reward"], b["done"], w)
        torch.cuda.synchronize()
        return idx.clone(), w.clone(), [p.grad.detach().clone() for p in L.model.getParameters()]

    # ---- data-parallel run: one shard per rank ---------------------------------------------------
    L = make_learner(); fill(L, rank)
    L.enable_data_parallel()
    local_max = L.memory.store.max_weight(cfg.BETA).clone()
    glob = local_max.clone()
    dist.all_reduce(glob, op=dist.ReduceOp.MAX)
    allmax = [torch.empty_like(local_max) for _ in range(world)]
    dist.all_gather(allmax, local_max)
    assert float(glob) == max(float(m) for m in allmax)
    if world > 1:
        assert len({float(m) for m in allmax}) > 1, "shards must differ for the test to mean anything"
    idx_dp, w_dp, g_dp = one_step_grads(L, glob)

    # ---- single-process reference: every shard on THIS GPU, no collective ---------------------------
    g_sum, g_own = None, None
    for shard in range(world):
        Lr = make_learner(); fill(Lr, shard)
        idx_r, w_r, g_r = one_step_grads(Lr, glob)
        if shard == rank:
            assert torch.equal(idx_r, idx_dp) and torch.equal(w_r, w_dp)     # local sampling is shard-local
            g_own = g_r
        g_sum = g_r if g_sum is None else [a + b for a, b in zip(g_sum, g_r)]
        del Lr
    names = [n for n, _ in L.model.named_parameters()]
    worst, bad = 0.0, []
    for n, a, b, own in zip(names, g_dp, g_sum, g_own):
        r = rel(a, b / world)
        worst = max(worst, r)
        if r > 1e-5:
            bad.append(f"{n}: vs mean {r:.3e}, vs own shard {rel(a, own):.3e}, vs sum {rel(a, b):.3e}")
    assert not bad, "rank %d: %s" % (rank, "; ".join(bad))
    # the late slice (conv stack) goes through libb2rl's peer-memory kernel when the ranks share a node: it sums in
    # rank order, so every rank must hold BIT-identical averaged gradients, and no peer may have timed out
    peer = bool(getattr(L, "peer_allreduce", False))
    if peer:
        late = L._bucket._late.clone()
        every = [torch.empty_like(late) for _ in range(world)]
        dist.all_gather(every, late)
        assert all(torch.equal(every[0], e) for e in every), "peer all-reduce: ranks disagree bitwise"
        assert int(L._bucket._peer.error.item()) == 0
        big = L._bucket._peer_big
        if big is not None:                                   # the heads: reduce-scatter + all-gather kernel
            heads = L._bucket._early.clone()
            every = [torch.empty_like(heads) for _ in range(world)]
            dist.all_gather(every, heads)
            assert all(torch.equal(every[0], e) for e in every), "peer all-reduce (heads): ranks disagree bitwise"
            assert int(big.error.item()) == 0
        elif os.environ.get("B2RL_PEER_ALLREDUCE_BIG"):
            raise AssertionError("peer-memory all-reduce of the heads was requested but is not active")
    elif not os.environ.get("B2RL_NO_PEER_ALLREDUCE"):
        raise AssertionError("peer-memory all-reduce expected on a single-node NCCL run")

    # ---- step loop: the IS-weight normaliser is the MAX reduced during the previous step ------------------
    st = L.memory.store
    prev_use = None
    for k in range(3):
        torch.cuda.synchronize()
        lm = st.max_weight(cfg.BETA).clone()
        gm = lm.clone(); dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        stats = st.stats(cfg.BETA).clone()
        prios = st.priorities().clone()
        n_valid = float(len(st))
        out = L.fused_step(use_graph=False)
        torch.cuda.synchronize()
        used = float(glob) if prev_use is None else prev_use         # first step: reduced synchronously before it
        if k == 0:
            used = float(gm)                                          # fused_step's own first-step reduce
        idx, w = L._cur["idx"], L._cur["w"]
        prob = (prios[idx] / stats[0].float())
        w_exp = ((1.0 / (n_valid * prob.double())) ** cfg.BETA).float() / used
        assert rel(w, w_exp) <= 1e-6, (k, rel(w, w_exp))
        assert float(L._max_w_use) == float(gm)                      # reduced behind this step, used by the next
        prev_use = float(gm)
    if peer:
        assert int(L._bucket._peer.error.item()) == 0
    print(f"MGPU_OK rank {rank}/{world} worst_grad_rel {worst:.2e} peer_allreduce {peer} heads {bool(peer and L._bucket._peer_big is not None)}", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
