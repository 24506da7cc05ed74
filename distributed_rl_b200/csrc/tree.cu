// Device-resident sum-tree: bulk build, proportional sampling + IS weights,
// deterministic batched priority update, stats.
//
// Semantics follow the reference's own sum-tree (baseline/sumtree.py) exactly:
// fp64 nodes, node = fl64(left + right) (Node._reduce :21-27), descent
// `pos < left ? left : (pos -= left, right)` (Node._find :53-62).  The flat
// fp32 store the live learners use (baseline/PER.py) agrees with it whenever
// its own fp32 cumulative sums are exact (tests/golden/tree.npz, dyadic cases).
//
// HBM/L2 layout: implicit heap, level k occupies [2^k, 2^(k+1))
// so both children of a node share one 16-byte aligned pair -> one LDG.128 per
// level of the descent.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

#include <cub/block/block_scan.cuh>

namespace b2rl {

// ----------------------------------------------------------------------------
// Bulk build
// ----------------------------------------------------------------------------
constexpr int BUILD_CHUNK = 2048;   // leaves per CTA (11 levels resolved in SMEM)
constexpr int BUILD_THREADS = 256;
constexpr int BUILD_PER_THREAD = BUILD_CHUNK / 2 / BUILD_THREADS;  // 4

// Each CTA owns leaves [c*2048, (c+1)*2048): converts the fp32 priorities to
// fp64 leaves, reduces the bottom levels of its subtree in shared memory and
// writes every node out; its subtree root lands at heap index cap2/2048 + c.
__global__ void __launch_bounds__(BUILD_THREADS)
k_build_bottom(const float* __restrict__ prios, int64_t n_valid, TreeNode* __restrict__ node, int64_t cap2) {
  __shared__ double s_sum[BUILD_CHUNK];
  __shared__ float s_min[BUILD_CHUNK];
  const int64_t leaf0 = (int64_t)blockIdx.x * BUILD_CHUNK;
  const int nleaf = (int)min((int64_t)BUILD_CHUNK, cap2 - leaf0);  // cap2 < 2048 -> one CTA
  for (int i = threadIdx.x; i < nleaf; i += BUILD_THREADS) {
    const int64_t j = leaf0 + i;
    const float p = (j < n_valid) ? prios[j] : 0.0f;
    const float m = (p > 0.0f) ? p : INFINITY;
    s_sum[i] = (double)p;
    s_min[i] = m;
    st_node(node + cap2 + j, (double)p, m);
  }
  __syncthreads();
  int per = 2;  // leaves per node of the level being produced
  for (int w = nleaf / 2; w >= 1; w >>= 1, per <<= 1) {
    const int64_t base = (cap2 + leaf0) / per;  // heap index of this CTA's first node on the level
    double ts[BUILD_PER_THREAD];
    float tm[BUILD_PER_THREAD];
    int c = 0;
    for (int i = threadIdx.x; i < w; i += BUILD_THREADS, ++c) {
      ts[c] = s_sum[2 * i] + s_sum[2 * i + 1];
      tm[c] = fminf(s_min[2 * i], s_min[2 * i + 1]);
    }
    __syncthreads();
    c = 0;
    for (int i = threadIdx.x; i < w; i += BUILD_THREADS, ++c) {
      s_sum[i] = ts[c];
      s_min[i] = tm[c];
      st_node(node + base + i, ts[c], tm[c]);
    }
    __syncthreads();
  }
}

// Levels above the 2048-leaf subtrees: `m` = cap2/2048 nodes on the deepest of
// them.  One CTA, level-synchronous (at most 12 levels for cap2 = 2^23).
__global__ void __launch_bounds__(1024)
k_build_top(TreeNode* __restrict__ node, int64_t m) {
  for (int64_t w = m / 2; w >= 1; w >>= 1) {
    for (int64_t i = threadIdx.x; i < w; i += blockDim.x) {
      const int64_t nd = w + i;
      const TreeNode a = ld_node(node + 2 * nd), b = ld_node(node + 2 * nd + 1);
      st_node(node + nd, a.s + b.s, fminf(a.m, b.m));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_node(node, 0.0, INFINITY);
}

// ----------------------------------------------------------------------------
// Sampling + importance weights
// ----------------------------------------------------------------------------
constexpr int SAMPLE_THREADS = 128;

__global__ void __launch_bounds__(SAMPLE_THREADS)
k_tree_sample(const TreeNode* __restrict__ node, int64_t cap2,
              int levels, const double* __restrict__ u01, uint64_t seed, uint64_t rng_offset,
              const uint64_t* __restrict__ rng_state, int64_t n, const float* __restrict__ n_valid_dev, float beta,
              const float* __restrict__ max_w_ext, int64_t* __restrict__ idx_out,
              float* __restrict__ prob_out, float* __restrict__ w_out) {
  const int64_t k = (int64_t)blockIdx.x * SAMPLE_THREADS + threadIdx.x;
  if (rng_state) {
    // device-resident Philox stream: every block reads {seed, counter}; the LAST block to have done
    // so advances the counter by n (and re-arms the ticket), so no separate launch is needed.
    seed = rng_state[0];
    rng_offset = rng_state[1];
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned int* ticket = reinterpret_cast<unsigned int*>(const_cast<uint64_t*>(rng_state) + 2);
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        const_cast<uint64_t*>(rng_state)[1] = rng_offset + (uint64_t)n;
        *ticket = 0u;
      }
    }
  }
  if (k >= n) return;
  const TreeNode rootn = ld_node(node + 1);
  const double root = rootn.s;
  const double u = u01 ? u01[k] : philox_u01(seed, rng_offset + (uint64_t)k);
  double pos = __dmul_rn(root, u);  // np.random.uniform(0, root) == root * random_sample()
  int64_t i = 1;
  for (int l = 0; l < levels; ++l) {
    const double cl = node[2 * i].s, cr = node[2 * i + 1].s;   // both children live in one 32-byte sector
    // Node._find: left iff pos < left.  The `cr == 0` guard only matters when
    // pos rounds up to the subtree total (the reference dereferences None there).
    const bool left = (pos < cl) || (cr == 0.0);
    if (!left) pos = __dsub_rn(pos, cl);
    i = 2 * i + (left ? 0 : 1);
  }
  const int64_t j = i - cap2;
  idx_out[k] = j;
  if (prob_out == nullptr && w_out == nullptr) return;
  // APE_X/ReplayMemory.py:65-67, baseline/PER.py:98,129-133 — fp32 op by op.
  const float s32 = (float)root;
  const float p = (float)node[i].s;
  const float prob = __fdiv_rn(p, s32);
  if (prob_out) prob_out[k] = prob;
  if (w_out) {
    const float n_valid = *n_valid_dev;   // current number of valid slots (stream-ordered, not a launch constant)
    const float w_un = powcr(__fdiv_rn(1.0f, __fmul_rn(n_valid, prob)), beta);
    float max_w;
    if (max_w_ext) {
      max_w = *max_w_ext;
    } else {
      const float min_prob = __fdiv_rn(rootn.m, s32);
      max_w = powcr(__fmul_rn(n_valid, min_prob), -beta);
    }
    w_out[k] = __fdiv_rn(w_un, max_w);
  }
}

__global__ void k_rng_seed(uint64_t* __restrict__ rng_state, uint64_t seed, uint64_t ctr) {
  rng_state[0] = seed; rng_state[1] = ctr; rng_state[2] = 0;
}

__global__ void k_philox_uniforms(uint64_t seed, uint64_t off, int64_t n, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = philox_u01(seed, off + (uint64_t)k);
}

__global__ void k_set_n_valid(float* __restrict__ n_valid_dev, float v) { *n_valid_dev = v; }

__global__ void k_tree_stats(const TreeNode* __restrict__ node, const float* __restrict__ n_valid_dev, float beta,
                             double* __restrict__ out, float* __restrict__ max_w_out) {
  const float n_valid = *n_valid_dev;
  const TreeNode r = ld_node(node + 1);
  const double root = r.s;
  const float s32 = (float)root;
  const float mn = r.m;
  const float mw = powcr(__fmul_rn(n_valid, __fdiv_rn(mn, s32)), -beta);
  if (out) { out[0] = root; out[1] = (double)mn; out[2] = (double)mw; }
  if (max_w_out) *max_w_out = mw;
}

__global__ void k_tree_leaves(const TreeNode* __restrict__ node, int64_t cap2, int64_t start, int64_t n,
                              float* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = (float)node[cap2 + start + k].s;
}

// ----------------------------------------------------------------------------
// Batched priority update — deterministic last-writer-wins, no fp atomics.
//   A  tag[idx[k]] = max(tag, k+1)                     (who is the last writer?)
//   B  winners write their leaf and mark every ancestor with the side they
//      come from (bit0 = left child touched, bit1 = right child touched)
//   C  winners climb; a node expects popc(mark&3) arrivals, and only the LAST
//      arriver recomputes node = left + right from the (now final) children
//      and continues upward.
// tag[] and mark[] return to zero, so the sequence is CUDA-graph replayable.
// The final tree equals the reference's sequential writes because every
// Node._reduce recomputes the node from its children (state is path-independent).
// ----------------------------------------------------------------------------
constexpr int UPD_THREADS = 256;

__device__ __forceinline__ int64_t upd_index(const int64_t* idx, int64_t ring_start, int64_t capacity,
                                             int64_t k) {
  if (idx) return idx[k];
  int64_t j = ring_start + k;
  return j >= capacity ? j - capacity : j;
}

__global__ void __launch_bounds__(UPD_THREADS)
k_update_tag(const int64_t* __restrict__ idx, int64_t ring_start, int64_t capacity, int64_t n,
             uint32_t* __restrict__ tag) {
  const int64_t k = (int64_t)blockIdx.x * UPD_THREADS + threadIdx.x;
  if (k >= n) return;
  const int64_t j = upd_index(idx, ring_start, capacity, k);
  if (j < 0 || j >= capacity) return;  // out-of-range indices are ignored
  atomicMax(tag + j, (uint32_t)(k + 1));
}

__global__ void __launch_bounds__(UPD_THREADS)
k_update_write(const int64_t* __restrict__ idx, int64_t ring_start, int64_t capacity,
               const float* __restrict__ vals, float const_val, int64_t n,
               const uint32_t* __restrict__ tag, TreeNode* __restrict__ tree,
               int32_t* __restrict__ mark, int64_t cap2, int levels) {
  const int64_t k = (int64_t)blockIdx.x * UPD_THREADS + threadIdx.x;
  if (k >= n) return;
  const int64_t j = upd_index(idx, ring_start, capacity, k);
  if (j < 0 || j >= capacity) return;
  if (tag[j] != (uint32_t)(k + 1)) return;  // a later k wrote the same slot
  const float p = vals ? vals[k] : const_val;
  int64_t node = cap2 + j;
  st_node(tree + node, (double)p, (p > 0.0f) ? p : INFINITY);
  for (int l = 0; l < levels; ++l) {
    const int bit = (node & 1) ? 2 : 1;
    node >>= 1;
    const int old = atomicOr(mark + node, bit);
    if (old & bit) break;  // another winner below the same child already marks the rest of the path
  }
}

__global__ void __launch_bounds__(UPD_THREADS)
k_update_climb(const int64_t* __restrict__ idx, int64_t ring_start, int64_t capacity, int64_t n,
               uint32_t* __restrict__ tag, TreeNode* __restrict__ tree,
               int32_t* __restrict__ mark, int64_t cap2, int levels) {
  const int64_t k = (int64_t)blockIdx.x * UPD_THREADS + threadIdx.x;
  if (k >= n) return;
  const int64_t j = upd_index(idx, ring_start, capacity, k);
  if (j < 0 || j >= capacity) return;
  if (tag[j] != (uint32_t)(k + 1)) return;
  tag[j] = 0u;  // self-clean (only the winner touches it in this kernel)
  int64_t node = cap2 + j;
  for (int l = 0; l < levels; ++l) {
    node >>= 1;
    __threadfence();                               // publish my child before announcing
    const int old = atomicAdd(mark + node, 4);     // arrivals live above the two side bits
    if ((old >> 2) + 1 < __popc(old & 3)) return;  // the other touched child arrives later
    __threadfence();                               // acquire: see the other subtree's writes
    const TreeNode a = ld_node_cg(tree + 2 * node), b = ld_node_cg(tree + 2 * node + 1);
    st_node(tree + node, a.s + b.s, fminf(a.m, b.m));
    mark[node] = 0;                                // self-clean: nobody else visits this node now
  }
}

// ----------------------------------------------------------------------------
// Small-batch update (n <= 512 per launch): ONE CTA, no atomics, no scratch.
//   1. bitonic sort of (leaf id, batch position) -> duplicates are adjacent and the
//      last occurrence (the winner) is the last of its run
//   2. every thread prefetches the `levels` sibling nodes of its path in one go
//      (independent addresses: one L2 round trip instead of one per level)
//   3. level by level in shared memory: threads sharing a node form a contiguous
//      group [lo, hi); the sibling node is either the adjacent group (touched by
//      this batch -> take its fresh value from SMEM) or untouched (-> prefetched
//      value); parent = left + right; groups merge; the group leader writes the
//      node back.  ~150 cycles per level instead of two fenced L2 atomics.
// Same final state as the sequential reference writes (path-independent reduce).
// ----------------------------------------------------------------------------
constexpr int US_THREADS = 512;
constexpr int US_MAX_LEVELS = 24;

__global__ void __launch_bounds__(US_THREADS, 1)
k_update_sorted(const int64_t* __restrict__ idx, int64_t ring_start, int64_t capacity,
                const float* __restrict__ vals, float const_val, int n, TreeNode* __restrict__ tree,
                int64_t cap2, int levels, float* __restrict__ n_valid_dev, float n_valid_new) {
  __shared__ uint64_t s_key[US_THREADS];
  if (n_valid_dev != nullptr && threadIdx.x == 0) *n_valid_dev = n_valid_new;   // ring size after this ingest step
  __shared__ uint32_t s_leaf[US_THREADS];
  __shared__ float s_valf[US_THREADS];
  __shared__ double s_sum[US_THREADS];
  __shared__ float s_min[US_THREADS];
  __shared__ uint16_t s_lo[US_THREADS], s_hi[US_THREADS], s_end[US_THREADS];

  const int t = threadIdx.x;
  // key = leaf << 9 | batch position: ascending sort puts duplicates of a leaf next to
  // each other with the LAST occurrence last; invalid entries (all ones) go to the end.
  uint64_t key = ~0ULL;
  if (t < n) {
    const int64_t j = upd_index(idx, ring_start, capacity, t);
    if (j >= 0 && j < capacity) key = ((uint64_t)j << 9) | (uint64_t)t;
  }
  // block-wide bitonic sort, one key per thread: strides < 32 by warp shuffle, the rest via SMEM
#pragma unroll
  for (int size = 2; size <= US_THREADS; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      uint64_t other;
      if (stride >= 32) {
        s_key[t] = key;
        __syncthreads();
        other = s_key[t ^ stride];
        __syncthreads();
      } else {
        other = __shfl_xor_sync(0xffffffffu, key, stride);
      }
      const bool up = (t & size) == 0;
      const bool lower = (t & stride) == 0;
      const bool take_min = (lower == up);
      const uint64_t mn = key < other ? key : other, mx = key < other ? other : key;
      key = take_min ? mn : mx;
    }
  }
  const bool valid = key != ~0ULL;
  const uint32_t leaf = valid ? (uint32_t)(key >> 9) : 0xFFFFFFFFu;
  s_leaf[t] = leaf;
  s_valf[t] = valid ? (vals ? vals[(int)(key & 511u)] : const_val) : 0.0f;
  const int nvalid = __syncthreads_count(valid);  // also publishes s_leaf / s_valf

  // prefetch the sibling of every node on my path (depth `levels` = leaves ... depth 1)
  double pre_sum[US_MAX_LEVELS];
  float pre_min[US_MAX_LEVELS];
#pragma unroll
  for (int l = 0; l < US_MAX_LEVELS; ++l) {
    if (valid && l < levels) {
      const int64_t node = ((cap2 + (int64_t)leaf) >> l) ^ 1;   // heap index of the sibling
      const TreeNode sib = ld_node_cg(tree + node);             // one 16-byte request per level
      pre_sum[l] = sib.s;
      pre_min[l] = sib.m;
    } else { pre_sum[l] = 0.0; pre_min[l] = INFINITY; }
  }

  // runs of equal leaves: [lo, hi)
  const bool head = valid && (t == 0 || s_leaf[t - 1] != leaf);
  const bool tail = valid && (t == nvalid - 1 || s_leaf[t + 1] != leaf);
  int lo;
  {
    using Scan = cub::BlockScan<int, US_THREADS>;
    __shared__ typename Scan::TempStorage scan_tmp;
    Scan(scan_tmp).InclusiveScan(head ? t : 0, lo, cub::Max());
  }
  if (tail) s_end[lo] = (uint16_t)(t + 1);
  __syncthreads();
  int hi = valid ? s_end[lo] : 0;
  const float vwin = valid ? s_valf[hi - 1] : 0.0f;          // last writer wins
  double cur_sum = (double)vwin;
  float cur_min = (vwin > 0.0f) ? vwin : INFINITY;
  if (valid && t == lo) st_node(tree + cap2 + leaf, cur_sum, cur_min);
#pragma unroll
  for (int l = 0; l < US_MAX_LEVELS; ++l) {
    if (l >= levels) break;
    s_sum[t] = cur_sum; s_min[t] = cur_min; s_lo[t] = (uint16_t)lo; s_hi[t] = (uint16_t)hi;
    __syncthreads();
    if (valid) {
      const uint32_t node = leaf >> l;          // index within its level
      double sib_sum = pre_sum[l];
      float sib_min = pre_min[l];
      if ((node & 1u) == 0u) {
        if (hi < nvalid && (s_leaf[hi] >> l) == node + 1u) {
          sib_sum = s_sum[hi]; sib_min = s_min[hi]; hi = s_hi[hi];
        }
      } else {
        if (lo > 0 && (s_leaf[lo - 1] >> l) == node - 1u) {
          sib_sum = s_sum[lo - 1]; sib_min = s_min[lo - 1]; lo = s_lo[lo - 1];
        }
      }
      cur_sum = cur_sum + sib_sum;              // fl64(left + right); + is commutative
      cur_min = fminf(cur_min, sib_min);
    }
    __syncthreads();
    if (valid && t == lo) {
      st_node(tree + ((cap2 + (int64_t)leaf) >> (l + 1)), cur_sum, cur_min);
    }
  }
}

}  // namespace b2rl

using namespace b2rl;

static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

int b2rl::publish_size(b2rl_replay* h, cudaStream_t st) {
  k_set_n_valid<<<1, 1, 0, st>>>(h->n_valid_dev, (float)h->size);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

// idx_dev == nullptr means the ring range [ring_start, ring_start+n) (mod capacity);
// vals_dev == nullptr means the constant `const_val`.
int b2rl_tree_update_impl(b2rl_replay* h, const int64_t* idx_dev, int64_t ring_start,
                          const float* vals_dev, float const_val, int64_t n, cudaStream_t st,
                          bool publish_size_too) {
  if (n == 0) return publish_size_too ? publish_size(h, st) : B2RL_OK;
  static int force_atomic = -1;
  if (force_atomic < 0) {
    const char* e = getenv("B2RL_UPDATE");
    force_atomic = (e && e[0] == 'a') ? 1 : 0;   // B2RL_UPDATE=atomic forces the scalable path
  }
  if (!force_atomic && n <= 2 * US_THREADS && h->levels <= US_MAX_LEVELS) {
    // chunks are applied in stream order, so last-writer-wins also holds across chunks
    for (int64_t off = 0; off < n; off += US_THREADS) {
      const int m = (int)((n - off < US_THREADS) ? (n - off) : US_THREADS);
      const bool last = off + US_THREADS >= n;
      k_update_sorted<<<1, US_THREADS, 0, st>>>(idx_dev ? idx_dev + off : nullptr,
                                                (ring_start + off) % h->capacity, h->capacity,
                                                vals_dev ? vals_dev + off : nullptr, const_val, m, h->node,
                                                h->cap2, h->levels,
                                                (publish_size_too && last) ? h->n_valid_dev : nullptr,
                                                (float)h->size);
      count_launch();
    }
    B2RL_CHECK_LAUNCH();
    return B2RL_OK;
  }
  const unsigned g = grid_for(n, UPD_THREADS);
  k_update_tag<<<g, UPD_THREADS, 0, st>>>(idx_dev, ring_start, h->capacity, n, h->tag);
  k_update_write<<<g, UPD_THREADS, 0, st>>>(idx_dev, ring_start, h->capacity, vals_dev, const_val, n,
                                            h->tag, h->node, h->mark, h->cap2, h->levels);
  k_update_climb<<<g, UPD_THREADS, 0, st>>>(idx_dev, ring_start, h->capacity, n, h->tag, h->node,
                                            h->mark, h->cap2, h->levels);
  count_launch(3);
  B2RL_CHECK_LAUNCH();
  return publish_size_too ? publish_size(h, st) : B2RL_OK;
}

extern "C" int b2rl_tree_build(b2rl_replay* h, const float* prios_dev, int64_t n, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(n >= 0 && n <= h->capacity, "n out of range");
  B2RL_REQUIRE(n == 0 || prios_dev != nullptr, "null priorities");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chunks = (h->cap2 + BUILD_CHUNK - 1) / BUILD_CHUNK;
  k_build_bottom<<<(unsigned)chunks, BUILD_THREADS, 0, st>>>(prios_dev, n, h->node, h->cap2);
  count_launch();
  if (h->cap2 > BUILD_CHUNK) {
    k_build_top<<<1, 1024, 0, st>>>(h->node, h->cap2 / BUILD_CHUNK);
    count_launch();
  }
  B2RL_CHECK_LAUNCH();
  h->size = n;
  h->head = (n == h->capacity) ? 0 : n;
  return publish_size(h, st);
}

extern "C" int b2rl_tree_sample(b2rl_replay* h, const double* u01_dev, uint64_t seed,
                                uint64_t rng_offset, int64_t n, float beta, const float* max_w_dev,
                                int64_t* idx_out_dev, float* prob_out_dev, float* w_out_dev,
                                void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(n >= 0, "negative n");
  B2RL_REQUIRE(h->size > 0, "sampling from an empty replay");
  B2RL_REQUIRE(n == 0 || idx_out_dev != nullptr, "null idx_out");
  if (n == 0) return B2RL_OK;
  DeviceGuard g(h->device);
  k_tree_sample<<<grid_for(n, SAMPLE_THREADS), SAMPLE_THREADS, 0, (cudaStream_t)stream>>>(
      h->node, h->cap2, h->levels, u01_dev, seed, rng_offset, nullptr, n, h->n_valid_dev, beta,
      max_w_dev, idx_out_dev, prob_out_dev, w_out_dev);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_replay_seed(b2rl_replay* h, uint64_t seed, uint64_t counter, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  DeviceGuard g(h->device);
  k_rng_seed<<<1, 1, 0, (cudaStream_t)stream>>>(h->rng_dev, seed, counter);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_tree_sample_stream(b2rl_replay* h, int64_t n, float beta, const float* max_w_dev,
                                       int64_t* idx_out_dev, float* prob_out_dev, float* w_out_dev,
                                       void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(n >= 0, "negative n");
  B2RL_REQUIRE(h->size > 0, "sampling from an empty replay");
  B2RL_REQUIRE(n == 0 || idx_out_dev != nullptr, "null idx_out");
  if (n == 0) return B2RL_OK;
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  k_tree_sample<<<grid_for(n, SAMPLE_THREADS), SAMPLE_THREADS, 0, st>>>(
      h->node, h->cap2, h->levels, nullptr, 0, 0, h->rng_dev, n, h->n_valid_dev, beta,
      max_w_dev, idx_out_dev, prob_out_dev, w_out_dev);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_philox_uniforms(uint64_t seed, uint64_t rng_offset, int64_t n, double* out_dev,
                                    void* stream) {
  B2RL_REQUIRE(n >= 0 && (n == 0 || out_dev), "bad arguments");
  if (n == 0) return B2RL_OK;
  k_philox_uniforms<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, rng_offset, n, out_dev);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_tree_update(b2rl_replay* h, const int64_t* idx_dev, const float* vals_dev,
                                int64_t n, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(n >= 0, "negative n");
  B2RL_REQUIRE(n == 0 || (idx_dev && vals_dev), "null idx/vals");
  B2RL_REQUIRE(n < (int64_t)0xFFFFFFFFLL, "batch too large");
  DeviceGuard g(h->device);
  return b2rl_tree_update_impl(h, idx_dev, 0, vals_dev, 0.0f, n, (cudaStream_t)stream, false);
}

extern "C" int b2rl_tree_stats(b2rl_replay* h, float beta, double* stats_out_dev, float* max_w_out_dev,
                               void* stream) {
  B2RL_REQUIRE(h != nullptr && (stats_out_dev != nullptr || max_w_out_dev != nullptr), "null argument");
  DeviceGuard g(h->device);
  k_tree_stats<<<1, 1, 0, (cudaStream_t)stream>>>(h->node, h->n_valid_dev, beta, stats_out_dev,
                                                  max_w_out_dev);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_tree_leaves(b2rl_replay* h, int64_t start, int64_t n, float* out_dev, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(start >= 0 && n >= 0 && start + n <= h->capacity, "range out of bounds");
  if (n == 0) return B2RL_OK;
  B2RL_REQUIRE(out_dev != nullptr, "null out");
  DeviceGuard g(h->device);
  k_tree_leaves<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(h->node, h->cap2, start, n, out_dev);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}
