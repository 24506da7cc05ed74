// Mean all-reduce of a SMALL gradient slice across the data-parallel learner ranks of one node, over NVLink /
// NVSwitch peer memory, as one kernel per rank.
//
// The reference has a single learner process and no collective at all (replay-sharded data
// parallelism is new work).  Per step the ranks average two gradient slices: the dense heads' (12.9 MB, launched
// early from the weight-gradient lane, NCCL) and what is left when backward ends — the convolution stack's
// 0.3 MB.  That second one sits on the critical path between backward and the optimizer, and is latency-bound: an
// NCCL all-reduce of it costs 12-19 us (RING_LL at 2 ranks; two of them were 30 us of the 45 us tail in
// profiles/r02_timeline_2gpu.txt).  Here every rank
//   0. copies its slice into its own symmetric staging buffer (parity = step & 1) and publishes a per-CTA flag
//      to every peer (st.release.sys after __threadfence_system),
//   1. waits for the same CTA's flag of every peer (ld.acquire.sys, bounded spin),
//   2. reads every rank's staged slice through the peer mapping, adds them IN RANK ORDER (so every rank forms the
//      bit-identical sum), scales by 1/world and writes its own gradient memory.
// No trailing barrier: a staging buffer is rewritten two steps later, and the heads' all-reduce of the step in
// between orders that write after every peer's read.  Flags only ever grow (epoch per CTA), nothing is reset.
#include "common.cuh"

namespace b2rl {
namespace peer {

constexpr int THREADS = 256;
constexpr int MAX_CTAS = 64;       // flag slots per (rank, peer)
constexpr int MAX_WORLD = 16;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

struct Params {
  const uint64_t* stage;    // [world] device pointers: rank r's staging buffer (2 x cap floats), peer-mapped
  const uint64_t* flags;    // [world] device pointers: rank r's flag pad (world x MAX_CTAS uint32), peer-mapped
  int32_t rank, world;
  int64_t cap;              // floats per staging parity
  int64_t n;                // floats to reduce (multiple of 4)
  float* data;              // this rank's slice: read, then overwritten with the mean
  uint32_t* epoch;          // [MAX_CTAS] per-CTA launch counters (device, private to this rank)
  uint32_t* error;          // set to 1 if a peer's flag did not arrive within the spin bound
};

__global__ void __launch_bounds__(THREADS)
k_peer_allreduce_mean(const __grid_constant__ Params P) {
  __shared__ uint32_t s_epoch;
  __shared__ const float4* s_src[MAX_WORLD];
  const int c = blockIdx.x, G = gridDim.x;
  if (threadIdx.x == 0) s_epoch = P.epoch[c] + 1u;
  __syncthreads();
  const uint32_t e = s_epoch;
  const int64_t quads = P.n >> 2;
  const int64_t q0 = quads * c / G, q1 = quads * (c + 1) / G;     // this CTA's float4 range
  float4* mine = reinterpret_cast<float4*>(P.stage[P.rank]) + (int64_t)(e & 1u) * (P.cap >> 2);
  float4* data = reinterpret_cast<float4*>(P.data);
  // 0. stage my slice where the peers can read it
  for (int64_t q = q0 + threadIdx.x; q < q1; q += THREADS) mine[q] = data[q];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < P.world) {
    const int p = threadIdx.x;
    s_src[p] = reinterpret_cast<const float4*>(P.stage[p]) + (int64_t)(e & 1u) * (P.cap >> 2);
    if (p != P.rank) {
      st_release_sys(reinterpret_cast<uint32_t*>(P.flags[p]) + P.rank * MAX_CTAS + c, e);
      // 1. wait for peer p's slice c of this step (bounded: ~2 s of polling, then flag the error and go on)
      const uint32_t* f = reinterpret_cast<const uint32_t*>(P.flags[P.rank]) + p * MAX_CTAS + c;
      long long spins = 0;
      while ((int32_t)(ld_acquire_sys(f) - e) < 0) {
        if (++spins > (1LL << 26)) { *P.error = 1u; break; }
        __nanosleep(20);
      }
    }
  }
  __syncthreads();
  // 2. sum in rank order (identical on every rank), scale, write my gradients
  const float inv = 1.0f / (float)P.world;
  for (int64_t q = q0 + threadIdx.x; q < q1; q += THREADS) {
    float4 a = ld_peer(s_src[0] + q);
    for (int p = 1; p < P.world; ++p) {
      const float4 b = ld_peer(s_src[p] + q);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    data[q] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  }
  if (threadIdx.x == 0) P.epoch[c] = e;
}

}  // namespace peer
}  // namespace b2rl

using namespace b2rl;

extern "C" int32_t b2rl_peer_allreduce_max_ctas(void) { return peer::MAX_CTAS; }

extern "C" int b2rl_peer_allreduce_mean(const uint64_t* stage_ptrs_dev, const uint64_t* flag_ptrs_dev, int32_t rank,
                                        int32_t world, int64_t stage_cap_floats, float* data_dev, int64_t n,
                                        uint32_t* epoch_dev, uint32_t* error_dev, void* stream) {
  B2RL_REQUIRE(stage_ptrs_dev && flag_ptrs_dev && data_dev && epoch_dev && error_dev, "null argument");
  B2RL_REQUIRE(world >= 2 && world <= peer::MAX_WORLD && rank >= 0 && rank < world, "2..16 ranks");
  B2RL_REQUIRE(n >= 4 && (n % 4) == 0 && n <= stage_cap_floats && (stage_cap_floats % 4) == 0,
               "n must be a multiple of 4 within the staging capacity");
  B2RL_REQUIRE(((uintptr_t)data_dev % 16) == 0, "data must be 16-byte aligned");
  peer::Params P{stage_ptrs_dev, flag_ptrs_dev, rank, world, stage_cap_floats, n, data_dev, epoch_dev, error_dev};
  const int64_t quads = n >> 2;
  int64_t g = (quads + 4 * peer::THREADS - 1) / (4 * peer::THREADS);     // ~4 float4 per thread
  if (g < 1) g = 1;
  if (g > peer::MAX_CTAS) g = peer::MAX_CTAS;
  peer::k_peer_allreduce_mean<<<(unsigned)g, peer::THREADS, 0, (cudaStream_t)stream>>>(P);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}
