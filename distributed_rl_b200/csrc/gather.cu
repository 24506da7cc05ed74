// Payload path: minibatch gather (HBM -> SMEM -> HBM through the TMA engine's
// 1-D bulk copies), ingest (push), eviction and the synthetic counter-hash fill.
//
// Replaces Replay.buffer's deepcopy + pickle.loads + np.stack
// (APE_X/ReplayMemory.py:61-116, baseline/PER.py:113) — the dominant CPU cost
// of the reference path.
#include "common.cuh"

namespace b2rl {

// ----------------------------------------------------------------------------
// mbarrier / bulk-copy PTX wrappers (sm_90+; SASS: UBLKCP / SYNCS)
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------
// Bulk gather.  Work item = (field, sample k, chunk c); a chunk is at most CHUNK
// bytes of one row.  Thread 0 of each CTA drives a ring of GATHER_SMEM / CHUNK stages:
//   load(item) : cp.async.bulk global -> smem, completes on mbarrier[stage]
//   store(item): cp.async.bulk smem -> global (bulk_group)
// The SMs only issue descriptors; the payload never touches the register file.
// CHUNK = 14 KiB (half a frame stack, 16 stages) measured best: stores of the first
// chunks overlap the loads of the later ones, and the single driving thread is not
// yet issue-bound (8 KiB chunks were 10 % slower, 28 KiB chunks 2 % slower).
// Warp 1 copies the scalar fields (action, reward, done, ...) of the CTA's share of
// the samples, so one launch assembles the whole minibatch.
// ----------------------------------------------------------------------------
constexpr int GATHER_THREADS = 64;
constexpr int GATHER_SMEM = 229376;  // 224 KiB ring; stages = GATHER_SMEM / CHUNK

struct GatherField {
  const uint8_t* src;   // field base
  uint8_t* dst;         // output base
  int64_t row_bytes;    // multiple of 16 for bulk fields
  int32_t chunks;       // ceil(row_bytes / CHUNK)
  int32_t pad;
};
struct GatherParams {
  GatherField f[B2RL_MAX_FIELDS];    // bulk (TMA) fields
  GatherField s[B2RL_MAX_FIELDS];    // small fields, copied by warp 1
  int32_t n_fields;
  int32_t n_small;
  int64_t n;            // samples
  int64_t capacity;
  int64_t items_per_sample;  // sum of chunks over bulk fields
  int64_t total_items;
};

// Walks the CTA's contiguous item range (sample-major, then field, then chunk) without divisions.
template <int CHUNK>
struct ItemCursor {
  int64_t k, row;
  int32_t f, c;
  __device__ __forceinline__ void init(const GatherParams& P, const int64_t* __restrict__ idx, int64_t item) {
    k = item / P.items_per_sample;
    int32_t r = (int32_t)(item - k * P.items_per_sample);
    f = 0;
    while (r >= P.f[f].chunks) { r -= P.f[f].chunks; ++f; }
    c = r;
    row = clamp_row(P, idx[k]);
  }
  static __device__ __forceinline__ int64_t clamp_row(const GatherParams& P, int64_t r) {
    return r < 0 ? 0 : (r >= P.capacity ? P.capacity - 1 : r);
  }
  __device__ __forceinline__ void get(const GatherParams& P, const uint8_t*& src, uint8_t*& dst,
                                      uint32_t& bytes) const {
    const int64_t off = (int64_t)c * CHUNK;
    const int64_t rem = P.f[f].row_bytes - off;
    bytes = (uint32_t)(rem < CHUNK ? rem : CHUNK);
    src = P.f[f].src + row * P.f[f].row_bytes + off;
    dst = P.f[f].dst + k * P.f[f].row_bytes + off;
  }
  __device__ __forceinline__ void next(const GatherParams& P, const int64_t* __restrict__ idx, bool more) {
    if (++c == P.f[f].chunks) {
      c = 0;
      if (++f == P.n_fields) {
        f = 0;
        ++k;
        if (more) row = clamp_row(P, idx[k]);
      }
    }
  }
};

template <int CHUNK, int LAG>
__global__ void __launch_bounds__(GATHER_THREADS, 1)
k_gather_bulk(const __grid_constant__ GatherParams P, const int64_t* __restrict__ idx) {
  constexpr int STAGES = GATHER_SMEM / CHUNK;
  static_assert(STAGES <= 32 && STAGES > LAG + 1, "ring geometry");
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[STAGES];

  if (threadIdx.x >= 32) {
    // ---- warp 1: scalar fields of samples [k0, k1) ------------------------------------
    const int lane = threadIdx.x - 32;
    const int64_t per = (P.n + gridDim.x - 1) / gridDim.x;
    const int64_t k0 = (int64_t)blockIdx.x * per;
    const int64_t k1 = (k0 + per < P.n) ? k0 + per : P.n;
    for (int f = 0; f < P.n_small; ++f) {
      const int64_t rb = P.s[f].row_bytes;
      if ((rb & 3) == 0) {
        const int64_t words = rb >> 2;
        for (int64_t u = lane; u < (k1 - k0) * words; u += 32) {
          const int64_t k = k0 + u / words, w = u % words;
          int64_t row = idx[k];
          row = row < 0 ? 0 : (row >= P.capacity ? P.capacity - 1 : row);
          reinterpret_cast<uint32_t*>(P.s[f].dst)[k * words + w] =
              reinterpret_cast<const uint32_t*>(P.s[f].src)[row * words + w];
        }
      } else {
        for (int64_t u = lane; u < (k1 - k0) * rb; u += 32) {
          const int64_t k = k0 + u / rb, b = u % rb;
          int64_t row = idx[k];
          row = row < 0 ? 0 : (row >= P.capacity ? P.capacity - 1 : row);
          P.s[f].dst[k * rb + b] = P.s[f].src[row * rb + b];
        }
      }
    }
    return;
  }
  if (threadIdx.x != 0 || P.total_items == 0) return;  // a single thread drives the copy engine
  for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  // items of this CTA: a contiguous range, so consecutive items share idx[] cache lines
  const int64_t per_cta = (P.total_items + gridDim.x - 1) / gridDim.x;
  const int64_t first = (int64_t)blockIdx.x * per_cta;
  if (first >= P.total_items) return;
  const int64_t my_items = (P.total_items - first < per_cta) ? P.total_items - first : per_cta;
  uint32_t phase_bits = 0;  // bit s = parity to wait for on stage s

  ItemCursor<CHUNK> ld, stc;   // load cursor runs ahead of the store cursor
  ld.init(P, idx, first);
  stc = ld;
  int64_t loaded = 0;
  // prologue: fill the ring
  const int64_t pre = my_items < STAGES ? my_items : STAGES;
  for (; loaded < pre; ++loaded) {
    const uint8_t* src; uint8_t* dst; uint32_t bytes;
    ld.get(P, src, dst, bytes);
    mbar_expect_tx(&bar[loaded], bytes);
    bulk_g2s(smem + (size_t)loaded * CHUNK, src, bytes, &bar[loaded]);
    ld.next(P, idx, loaded + 1 < my_items);
  }
  int s = 0;            // stage of item t
  int rs = 0;           // stage to recycle next (item t - LAG)
  for (int64_t t = 0; t < my_items; ++t) {
    const uint8_t* src; uint8_t* dst; uint32_t bytes;
    stc.get(P, src, dst, bytes);
    mbar_wait(&bar[s], (phase_bits >> s) & 1u);
    phase_bits ^= (1u << s);
    bulk_s2g(dst, smem + (size_t)s * CHUNK, bytes);
    bulk_commit();
    stc.next(P, idx, t + 1 < my_items);
    if (++s == STAGES) s = 0;
    // refill the stage used LAG items ago once its store has drained SMEM
    if (t >= LAG) {
      if (loaded < my_items) {
        bulk_wait_read<LAG>();   // all but the newest LAG store groups have finished reading SMEM
        const uint8_t* nsrc; uint8_t* ndst; uint32_t nbytes;
        ld.get(P, nsrc, ndst, nbytes);
        mbar_expect_tx(&bar[rs], nbytes);
        bulk_g2s(smem + (size_t)rs * CHUNK, nsrc, nbytes, &bar[rs]);
        ++loaded;
        ld.next(P, idx, loaded < my_items);
      }
      if (++rs == STAGES) rs = 0;
    }
  }
  bulk_wait_all();
}

// Generic fallback / small fields: one thread per (sample, 4-byte word or byte).
__global__ void __launch_bounds__(256)
k_gather_small(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t row_bytes,
               const int64_t* __restrict__ idx, int64_t n, int64_t capacity) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((row_bytes & 3) == 0) {
    const int64_t words = row_bytes >> 2;
    if (t >= n * words) return;
    const int64_t k = t / words, w = t - k * words;
    int64_t row = idx[k];
    row = row < 0 ? 0 : (row >= capacity ? capacity - 1 : row);
    reinterpret_cast<uint32_t*>(dst)[k * words + w] =
        reinterpret_cast<const uint32_t*>(src)[row * words + w];
  } else {
    if (t >= n * row_bytes) return;
    const int64_t k = t / row_bytes, b = t - k * row_bytes;
    int64_t row = idx[k];
    row = row < 0 ? 0 : (row >= capacity ? capacity - 1 : row);
    dst[k * row_bytes + b] = src[row * row_bytes + b];
  }
}

// LDG.128/STG.128 reference implementation of the big-row gather (kept for the
// A/B comparison in profiles/, selectable with B2RL_GATHER=ldg).
__global__ void __launch_bounds__(256)
k_gather_ldg(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t row_bytes,
             const int64_t* __restrict__ idx, int64_t n, int64_t capacity) {
  const int64_t k = blockIdx.x;
  if (k >= n) return;
  int64_t row = idx[k];
  row = row < 0 ? 0 : (row >= capacity ? capacity - 1 : row);
  const int4* s = reinterpret_cast<const int4*>(src + row * row_bytes);
  int4* d = reinterpret_cast<int4*>(dst + k * row_bytes);
  const int64_t nv = row_bytes >> 4;
  for (int64_t i = threadIdx.x; i < nv; i += 4 * 256) {
    int4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * 256 < nv) v[u] = __ldg(s + i + u * 256);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * 256 < nv) d[i + u * 256] = v[u];
  }
}

// ----------------------------------------------------------------------------
// Synthetic fill: word w of slot s of field f =
//   lowbias32(seed ^ f*0x9E3779B9 ^ (uint32)s*2654435761 ^ (uint32)w*2246822519)
// (tail bytes of a row whose size is not a multiple of 4 take the low bytes).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fill_hash(uint8_t* __restrict__ base, int64_t row_bytes, int64_t n_rows, uint32_t seed, uint32_t fsalt) {
  const int64_t words_per_row = (row_bytes + 3) >> 2;
  const int64_t total = n_rows * words_per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = t / words_per_row, w = t - s * words_per_row;
    const uint32_t v = lowbias32(seed ^ fsalt ^ ((uint32_t)s * 2654435761U) ^ ((uint32_t)w * 2246822519U));
    uint8_t* p = base + s * row_bytes + w * 4;
    const int64_t left = row_bytes - w * 4;
    if (left >= 4 && ((row_bytes & 3) == 0)) {
      *reinterpret_cast<uint32_t*>(p) = v;
    } else {
      for (int b = 0; b < 4 && b < left; ++b) p[b] = (uint8_t)(v >> (8 * b));
    }
  }
}

}  // namespace b2rl

using namespace b2rl;

int b2rl_tree_update_impl(b2rl_replay* h, const int64_t* idx_dev, int64_t ring_start,
                          const float* vals_dev, float const_val, int64_t n, cudaStream_t st,
                          bool publish_size_too);

static int gather_chunk() {
  static int chunk = -1;
  if (chunk < 0) {
    const char* e = getenv("B2RL_GATHER_CHUNK");   // 28672 (8 stages) | 14336 (16) | 7168 (32 -> 28 used)
    chunk = e ? atoi(e) : 14336;
    if (chunk != 28672 && chunk != 14336 && chunk != 8192) chunk = 14336;
  }
  return chunk;
}

static int gather_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("B2RL_GATHER");
    mode = (e && e[0] == 'l') ? 1 : 0;  // "ldg" -> 1, default bulk/TMA -> 0
  }
  return mode;
}

extern "C" int b2rl_replay_gather(b2rl_replay* h, const int64_t* idx_dev, int64_t n,
                                  void* const* out_fields_dev, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(n >= 0, "negative n");
  if (n == 0) return B2RL_OK;
  B2RL_REQUIRE(idx_dev != nullptr && out_fields_dev != nullptr, "null argument");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;

  GatherParams P{};
  P.n = n;
  P.capacity = h->capacity;
  int nb = 0, ns = 0;
  for (int f = 0; f < h->n_fields; ++f) {
    uint8_t* out = (uint8_t*)out_fields_dev[f];
    if (out == nullptr) continue;
    const int64_t rb = h->field_bytes[f];
    const bool big = rb >= 1024;
    const bool aligned16 = (rb % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)h->field[f] % 16 == 0);
    if (big && aligned16 && gather_mode() == 0) {
      P.f[nb].src = h->field[f];
      P.f[nb].dst = out;
      P.f[nb].row_bytes = rb;
      P.f[nb].chunks = (int32_t)((rb + gather_chunk() - 1) / gather_chunk());
      P.items_per_sample += P.f[nb].chunks;
      ++nb;
    } else if (big && aligned16) {
      k_gather_ldg<<<(unsigned)n, 256, 0, st>>>(h->field[f], out, rb, idx_dev, n, h->capacity);
      count_launch();
    } else if (!big && ((rb % 4 != 0) || (((uintptr_t)out % 4 == 0) && ((uintptr_t)h->field[f] % 4 == 0)))) {
      P.s[ns].src = h->field[f];
      P.s[ns].dst = out;
      P.s[ns].row_bytes = rb;
      ++ns;
    } else {
      const int64_t units = (rb % 4 == 0 && (uintptr_t)out % 4 == 0) ? n * (rb / 4) : n * rb;
      k_gather_small<<<(unsigned)((units + 255) / 256), 256, 0, st>>>(h->field[f], out, rb, idx_dev, n,
                                                                     h->capacity);
      count_launch();
    }
  }
  if (nb > 0 || ns > 0) {
    P.n_fields = nb;
    P.n_small = ns;
    P.total_items = P.items_per_sample * n;
    static int sms[64] = {0};
    static bool attr_set[64] = {false};
    const int dev = h->device;
    const size_t smem_bytes = (size_t)GATHER_SMEM;
    if (!attr_set[dev & 63]) {
      B2RL_CUDA(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
      B2RL_CUDA(cudaFuncSetAttribute(k_gather_bulk<28672, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
      B2RL_CUDA(cudaFuncSetAttribute(k_gather_bulk<14336, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
      B2RL_CUDA(cudaFuncSetAttribute(k_gather_bulk<8192, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes));
      attr_set[dev & 63] = true;
    }
    int64_t grid = sms[dev & 63];
    const int64_t work = P.total_items > n ? P.total_items : n;
    if (grid > work) grid = work;
    switch (gather_chunk()) {
      case 28672: k_gather_bulk<28672, 1><<<(unsigned)grid, GATHER_THREADS, smem_bytes, st>>>(P, idx_dev); break;
      case 8192:  k_gather_bulk<8192, 6><<<(unsigned)grid, GATHER_THREADS, smem_bytes, st>>>(P, idx_dev); break;
      default:    k_gather_bulk<14336, 3><<<(unsigned)grid, GATHER_THREADS, smem_bytes, st>>>(P, idx_dev); break;
    }
    count_launch();
  }
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_replay_fill_hash(b2rl_replay* h, int64_t n, uint32_t seed, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(n >= 0 && n <= h->capacity, "n out of range");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  for (int f = 0; f < h->n_fields; ++f) {
    const int64_t words = n * ((h->field_bytes[f] + 3) / 4);
    if (words == 0) continue;
    int64_t blocks = (words + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_fill_hash<<<(unsigned)blocks, 256, 0, st>>>(h->field[f], h->field_bytes[f], n, seed,
                                                  (uint32_t)f * 0x9E3779B9U);
    count_launch();
  }
  B2RL_CHECK_LAUNCH();
  h->size = n;
  h->head = (n == h->capacity) ? 0 : n;
  return publish_size(h, st);
}

extern "C" int b2rl_replay_push(b2rl_replay* h, const void* const* fields_src, const float* prios,
                                int64_t n, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(n >= 0 && n <= h->capacity, "n out of range (0..capacity)");
  if (n == 0) return B2RL_OK;
  B2RL_REQUIRE(prios != nullptr, "null priorities");
  B2RL_REQUIRE(h->n_fields == 0 || fields_src != nullptr, "null fields");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t head = h->head;
  const int64_t first = (head + n <= h->capacity) ? n : (h->capacity - head);  // before the wrap
  for (int f = 0; f < h->n_fields; ++f) {
    const uint8_t* src = (const uint8_t*)fields_src[f];
    if (src == nullptr) continue;
    const int64_t rb = h->field_bytes[f];
    B2RL_CUDA(cudaMemcpyAsync(h->field[f] + head * rb, src, (size_t)(first * rb), cudaMemcpyDefault, st));
    if (first < n)
      B2RL_CUDA(cudaMemcpyAsync(h->field[f], src + first * rb, (size_t)((n - first) * rb),
                                cudaMemcpyDefault, st));
  }
  B2RL_CUDA(cudaMemcpyAsync(h->scratch_val, prios, (size_t)n * sizeof(float), cudaMemcpyDefault, st));
  h->size = (h->size + n > h->capacity) ? h->capacity : h->size + n;   // published by the update kernel itself
  int rc = b2rl_tree_update_impl(h, nullptr, head, h->scratch_val, 0.0f, n, st, true);
  if (rc != B2RL_OK) return rc;
  h->head = (head + n) % h->capacity;
  return B2RL_OK;
}

extern "C" int b2rl_replay_reserve(b2rl_replay* h, int64_t n, int64_t* start_slot, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(n >= 1 && n <= h->capacity, "n out of range (1..capacity)");
  B2RL_REQUIRE(h->reserved == 0, "a reservation is already pending (call b2rl_replay_commit first)");
  DeviceGuard g(h->device);
  const int64_t overwritten = h->size + n - h->capacity;     // records that become unsampleable now
  if (overwritten > 0) h->size -= overwritten;
  int rc = b2rl_tree_update_impl(h, nullptr, h->head, nullptr, 0.0f, n, (cudaStream_t)stream, overwritten > 0);
  if (rc != B2RL_OK) return rc;
  h->reserved = n;
  if (start_slot) *start_slot = h->head;
  return B2RL_OK;
}

extern "C" int b2rl_replay_copy_payload(b2rl_replay* h, const void* const* fields_src, int64_t start_slot,
                                        int64_t n, void* stream) {
  B2RL_REQUIRE(h != nullptr && fields_src != nullptr, "null argument");
  B2RL_REQUIRE(n >= 1 && n <= h->capacity && start_slot >= 0 && start_slot < h->capacity, "range out of bounds");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t first = (start_slot + n <= h->capacity) ? n : (h->capacity - start_slot);
  for (int f = 0; f < h->n_fields; ++f) {
    const uint8_t* src = (const uint8_t*)fields_src[f];
    if (src == nullptr) continue;
    const int64_t rb = h->field_bytes[f];
    B2RL_CUDA(cudaMemcpyAsync(h->field[f] + start_slot * rb, src, (size_t)(first * rb), cudaMemcpyDefault, st));
    if (first < n)
      B2RL_CUDA(cudaMemcpyAsync(h->field[f], src + first * rb, (size_t)((n - first) * rb), cudaMemcpyDefault, st));
  }
  return B2RL_OK;
}

extern "C" int b2rl_replay_commit(b2rl_replay* h, const float* prios, int64_t n, void* stream) {
  B2RL_REQUIRE(h != nullptr && prios != nullptr, "null argument");
  B2RL_REQUIRE(n >= 1 && n == h->reserved, "commit size must equal the pending reservation");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  B2RL_CUDA(cudaMemcpyAsync(h->scratch_val, prios, (size_t)n * sizeof(float), cudaMemcpyDefault, st));
  h->size = (h->size + n > h->capacity) ? h->capacity : h->size + n;
  int rc = b2rl_tree_update_impl(h, nullptr, h->head, h->scratch_val, 0.0f, n, st, true);
  if (rc != B2RL_OK) return rc;
  h->head = (h->head + n) % h->capacity;
  h->reserved = 0;
  return B2RL_OK;
}

// One call per learner iteration for a steady ingest: publish the batch whose host->device copy was started by
// the PREVIOUS call (the learner stream waits for that copy's event, then writes its priorities: the records
// become sampleable), then retire the slots of the NEXT batch and start its copy on the library's own copy
// stream — which therefore overlaps whatever the caller enqueues on `stream` next (the learner step).
// The priorities travel with the payload and wait in device memory, so no host buffer has to outlive its copy
// beyond the next call.  fields_src == NULL flushes: publishes the pending batch and starts nothing.
extern "C" int b2rl_replay_ingest_pipelined(b2rl_replay* h, const void* const* fields_src, const float* prios_src,
                                            int64_t n, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(fields_src == nullptr || (n >= 1 && n <= h->capacity && prios_src != nullptr), "bad batch");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->ingest_stream == nullptr) {
    B2RL_CUDA(cudaStreamCreateWithFlags(&h->ingest_stream, cudaStreamNonBlocking));
    B2RL_CUDA(cudaEventCreateWithFlags(&h->ev_reserved, cudaEventDisableTiming));
    B2RL_CUDA(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
  }
  if (h->pipe_n > 0) {                      // 1. publish the batch in flight
    B2RL_REQUIRE(h->reserved == h->pipe_n, "pipelined ingest mixed with reserve/commit");
    B2RL_CUDA(cudaStreamWaitEvent(st, h->ev_copied, 0));
    const int64_t m = h->pipe_n;
    h->size = (h->size + m > h->capacity) ? h->capacity : h->size + m;
    int rc = b2rl_tree_update_impl(h, nullptr, h->head, h->pipe_prios, 0.0f, m, st, true);
    if (rc != B2RL_OK) return rc;
    h->head = (h->head + m) % h->capacity;
    h->reserved = 0;
    h->pipe_n = 0;
    // the copy stream must not overwrite pipe_prios before this update has read it
    B2RL_CUDA(cudaEventRecord(h->ev_reserved, st));
    B2RL_CUDA(cudaStreamWaitEvent(h->ingest_stream, h->ev_reserved, 0));
  }
  if (fields_src == nullptr) return B2RL_OK;
  B2RL_REQUIRE(h->reserved == 0, "a reservation is already pending (call b2rl_replay_commit first)");
  if (h->pipe_cap < n) {                    // (re)grow the priority staging; rare, synchronous
    B2RL_CUDA(cudaStreamSynchronize(h->ingest_stream));
    B2RL_CUDA(cudaStreamSynchronize(st));
    if (h->pipe_prios) cudaFree(h->pipe_prios);
    h->pipe_prios = nullptr;
    B2RL_CUDA(cudaMalloc((void**)&h->pipe_prios, sizeof(float) * (size_t)n));
    h->pipe_cap = n;
  }
  // 2. retire the slots about to be overwritten (they can no longer be sampled)
  const int64_t overwritten = h->size + n - h->capacity;
  if (overwritten > 0) h->size -= overwritten;
  int rc = b2rl_tree_update_impl(h, nullptr, h->head, nullptr, 0.0f, n, st, overwritten > 0);
  if (rc != B2RL_OK) return rc;
  h->reserved = n;
  h->pipe_n = n;
  // 3. payload + priorities on the copy stream, behind the retirement
  B2RL_CUDA(cudaEventRecord(h->ev_reserved, st));
  B2RL_CUDA(cudaStreamWaitEvent(h->ingest_stream, h->ev_reserved, 0));
  const int64_t start = h->head;
  const int64_t first = (start + n <= h->capacity) ? n : (h->capacity - start);
  for (int f = 0; f < h->n_fields; ++f) {
    const uint8_t* src = (const uint8_t*)fields_src[f];
    if (src == nullptr) continue;
    const int64_t rb = h->field_bytes[f];
    B2RL_CUDA(cudaMemcpyAsync(h->field[f] + start * rb, src, (size_t)(first * rb), cudaMemcpyDefault, h->ingest_stream));
    if (first < n)
      B2RL_CUDA(cudaMemcpyAsync(h->field[f], src + first * rb, (size_t)((n - first) * rb), cudaMemcpyDefault,
                                h->ingest_stream));
  }
  B2RL_CUDA(cudaMemcpyAsync(h->pipe_prios, prios_src, (size_t)n * sizeof(float), cudaMemcpyDefault, h->ingest_stream));
  B2RL_CUDA(cudaEventRecord(h->ev_copied, h->ingest_stream));
  return B2RL_OK;
}

extern "C" int b2rl_replay_evict(b2rl_replay* h, int64_t delta, void* stream) {
  B2RL_REQUIRE(h != nullptr, "null handle");
  B2RL_REQUIRE(delta >= 0 && delta <= h->size, "delta out of range (0..size)");
  if (delta == 0) return B2RL_OK;
  DeviceGuard g(h->device);
  // oldest record lives at (head - size) mod capacity
  int64_t tail = h->head - h->size;
  if (tail < 0) tail += h->capacity;
  h->size -= delta;
  return b2rl_tree_update_impl(h, nullptr, tail, nullptr, 0.0f, delta, (cudaStream_t)stream, true);
}
