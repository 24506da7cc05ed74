// Shared declarations for libb2rl (sm_100a).  Built with -fmad=false so every
// fp32/fp64 operation is individually rounded, which is what makes the kernels
// bit-comparable with the numpy oracle (oracle/oracle.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/b2rl.h"

namespace b2rl {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define B2RL_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      b2rl::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return B2RL_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define B2RL_CHECK_LAUNCH()                                                          \
  do {                                                                               \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) {                                                         \
      b2rl::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return B2RL_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define B2RL_REQUIRE(cond, msg)                                                      \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      b2rl::set_error("%s:%d invalid argument: %s", __FILE__, __LINE__, msg);        \
      return B2RL_ERR_INVALID;                                                       \
    }                                                                                \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != target) cudaSetDevice(prev);
  }
  int target = -1;
};

// ---- device helpers --------------------------------------------------------

// x^e for fp32 operands: evaluated in fp64, rounded once ("powcr").
__device__ __forceinline__ float powcr(float x, float e) {
  return (float)pow((double)x, (double)e);
}

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// Philox4x32-10 (Salmon et al. 2011), counter = (ctr_lo, ctr_hi, 0, 0), key = seed.
__host__ __device__ __forceinline__ void philox4x32_10(uint64_t ctr, uint64_t seed, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0u, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53U * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57U * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9U; k1 += 0xBB67AE85U;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from one Philox block: (x & (2^53-1)) * 2^-53 (the
// reference's CPU generator is mt19937 with the same 53-bit mask construction).
__host__ __device__ __forceinline__ double philox_u01(uint64_t seed, uint64_t ctr) {
  uint32_t r[4];
  philox4x32_10(ctr, seed, r);
  uint64_t x = ((uint64_t)r[1] << 32) | r[0];
  return (double)(x & ((1ULL << 53) - 1)) * (1.0 / 9007199254740992.0);
}

}  // namespace b2rl

// The sum-tree as the kernels see it (tree.cu): a binary tree of depth `levels` stored sparsely — every 4th
// level only.  Stored level 0 = leaves (fp32 priorities, 0 = empty slot); stored level k (1 <= k < G) = binary
// depth levels-4k as fp64 sums + fp32 mins over valid leaves; stored level G = the root.  The levels in between
// are recomputed in registers with the binary tree's own pairwise association (bit-identical values).
constexpr int B2RL_TREE_MAX_GROUPS = 8;   // levels <= 32
struct TreeView {
  float* leaf;                              // [cap2]
  double* sum;                              // stored level k at sum + off[k]
  float* minv;                              // stored level k at minv + off[k]
  int64_t off[B2RL_TREE_MAX_GROUPS + 1];    // off[0] unused
  int64_t cap2;                             // leaves = 2^levels >= max(capacity, 2)
  int levels;
  int G;                                    // ceil(levels / 4) stored internal levels
  int top_bits;                             // binary levels spanned by the top group: levels - 4(G-1), in 1..4
};

struct b2rl_replay;
namespace b2rl {
// Stream-ordered publication of the host-side `size` to n_valid_dev (tree.cu).
int publish_size(b2rl_replay* h, cudaStream_t st);
}  // namespace b2rl

// The opaque handle.
struct b2rl_replay {
  int device = 0;
  int64_t capacity = 0;   // requested slots
  int64_t cap2 = 0;       // tree leaves = 2^levels >= max(capacity, 2)
  int levels = 0;
  int n_fields = 0;
  int64_t field_bytes[B2RL_MAX_FIELDS] = {0};
  uint8_t* field[B2RL_MAX_FIELDS] = {nullptr};
  TreeView tree = {};         // leaves + sparse fp64 levels (owned: tree.leaf, tree.sum, tree.minv)
  uint32_t* tag = nullptr;    // [cap2]   last-writer tags of the large scattered update, self-cleaning
  int64_t* scratch_idx = nullptr;  // [capacity] ring indices for push/evict
  float* scratch_val = nullptr;    // [capacity]
  uint64_t* rng_dev = nullptr;     // [3] device-resident Philox stream {seed, counter, ticket}
  unsigned int* build_ticket = nullptr;   // [1] last-CTA-done counter of the bulk build, self re-arming
  float* n_valid_dev = nullptr;    // [1] (float)size, rewritten in stream order whenever size changes: the
                                   //     sampling / stats kernels read it, so a captured graph never bakes it in
  int64_t size = 0;       // valid slots
  int64_t head = 0;       // next slot to write
  int64_t reserved = 0;   // slots zeroed by b2rl_replay_reserve and not yet committed
  // b2rl_replay_ingest_pipelined: library-owned copy stream + events, device staging of the pending priorities
  cudaStream_t ingest_stream = nullptr;
  cudaEvent_t ev_reserved = nullptr, ev_copied = nullptr;
  float* pipe_prios = nullptr;
  int64_t pipe_cap = 0;   // floats allocated at pipe_prios
  int64_t pipe_n = 0;     // records of the batch whose copy is in flight (0: none)
};
