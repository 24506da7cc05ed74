// Fused gather + first convolution of the Q-network on the 5th-gen tensor cores.
//
//   y[k, oy, ox, co] = relu?( (1/255) * sum_{c,ky,kx} W[co, c, ky, kx] * frame[idx[k]][c, 4oy+ky, 4ox+kx] )
//
// for the 8x8 / stride-4 / 4->32 channel, bias-free conv_1 of cfg/ape_x.json and
// cfg/r2d2.json (baseline/baseNetwork.py:165-172), evaluated for up to two networks
// (online + target) in ONE pass over the sampled uint8 frame stacks.  It replaces,
// for the consumer of the gather, the staging copy + fp32 conversion + cuDNN conv
// (APE_X/Learner.py:61-67,78,85,87): the sampled rows go HBM -> SMEM (TMA bulk copy)
// -> im2col in SMEM -> tcgen05.mma -> TMEM -> registers -> NHWC fp32 activations, and
// the uint8 frames are never written back to HBM.
//
// Arithmetic: pixels 0..255 are exact in bf16.  fp32 weights are split into
// three bf16 terms, W = hi + mid + lo (24 significand bits, exact to ~2^-24 |W|); the MMA runs
// in kind::f16 (bf16 x bf16 -> fp32) and all three terms accumulate into the SAME fp32
// accumulator in TMEM (same A tile, three B descriptors), so the epilogue reads one accumulator
// per output channel.  Every product is exact in fp32, so the result matches an fp32
// convolution to a few ulp — tighter than the TF32 cuDNN path the reference would run.
// (A first version used kind::i8 with four 7-bit digits in separate accumulators: exact, but
// the epilogue then reads 4x the TMEM and became the bottleneck — profiles/r01_conv1.md.)
//
// Warp roles per CTA (persistent, one CTA per SM, 12 warps):
//   warp 0      TMA loader: weights once, then one 28 224-byte frame stack per item
//   warp 1      MMA issuer (one elected thread): 24 x tcgen05.mma (K = 16) per half tile
//   warp 2      TMEM allocator
//   warps 4-7   im2col producers: SMEM frame (u8) -> bf16, 128B-swizzled K-major A half tile
//   warps 8-11  epilogue: tcgen05.ld -> x 1/255 -> ReLU -> global (NHWC fp32)
#include "common.cuh"

namespace b2rl {
namespace conv1 {

constexpr int C_IN = 4, HW = 84, KS = 8, STRIDE = 4, OHW = 20, C_OUT = 32;
constexpr int K_TOTAL = C_IN * KS * KS;            // 256
constexpr int FRAME_BYTES = C_IN * HW * HW;        // 28 224
constexpr int POS = OHW * OHW;                     // 400 output positions per frame stack
constexpr int TILE_M = 128;
constexpr int TILES = (POS + TILE_M - 1) / TILE_M; // 4 (the last one has 16 valid rows)
constexpr int NSPLIT = 3;                          // bf16 terms per fp32 weight
constexpr int CHUNK_K = 64;                        // bf16 elements per 128-byte swizzle row = one input channel
constexpr int HALF_BYTES = TILE_M * 2 * 128;       // A stage = 2 channels x 128 rows x 128 B = 32 768
constexpr int A_CHUNK_BYTES = TILE_M * 128;        // 16 384
constexpr int RAW_STRIDE = 28288;                  // FRAME_BYTES rounded up to 128
constexpr int THREADS = 384;

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(sptr(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   sptr(dst)), "l"(src), "r"(bytes), "r"(sptr(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sptr(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc),
      "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle SMEM matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (= 1 for swizzled K-major) | [32,46) SBO >> 4
//   (= 1024 B between 8-row groups) | [46,48) version = 1 | [61,64) layout = SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a_format BF16 (1) @7,
// b_format BF16 (1) @10, K-major A and B, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// Byte offset of bf16 element (row n, k) inside a K-major SW128 operand with `rows` rows:
// [chunk = k/64][n/8][n%8][16-byte unit ^ (n%8)][byte]
__host__ __device__ __forceinline__ int sw128_offset_bf16(int rows, int n, int k) {
  const int j = k >> 6, kb = (k & 63) * 2;
  return j * rows * 128 + (n >> 3) * 1024 + (n & 7) * 128 + ((((kb >> 4) ^ (n & 7))) << 4) + (kb & 15);
}

// ---- weight packing: fp32 [32][256] -> three bf16 terms, hi + mid + lo == W to ~2^-24 |W| ----
// Layout of the packed operand: [term j][chunk c = input channel][row = net*32 + co][128 B swizzled]
__global__ void __launch_bounds__(K_TOTAL)
k_conv1_pack(const float* __restrict__ w, int net, int n_nets, uint16_t* __restrict__ bq, float* __restrict__ scale) {
  const int co = blockIdx.x, k = threadIdx.x;
  const int rows = n_nets * C_OUT;
  float r = w[co * K_TOTAL + k];
  if (k == 0) scale[net * C_OUT + co] = 1.0f / 255.0f;   // input normalisation, applied in the epilogue
#pragma unroll
  for (int j = 0; j < NSPLIT; ++j) {
    // round-to-nearest-even fp32 -> bf16 (finite inputs)
    uint32_t u = __float_as_uint(r);
    u += 0x7FFFu + ((u >> 16) & 1u);
    const uint16_t h = (uint16_t)(u >> 16);
    bq[((size_t)j * rows * K_TOTAL * 2 + sw128_offset_bf16(rows, net * C_OUT + co, k)) >> 1] = h;
    r = r - __uint_as_float((uint32_t)h << 16);           // exact: the remainder fits in fp32
  }
}

struct Params {
  const uint8_t* frames;     // field base: rows of FRAME_BYTES
  const int64_t* idx;        // sampled rows, or nullptr for rows 0..n-1
  int64_t n;                 // frame stacks to process
  int64_t capacity;          // rows in `frames` (indices are clamped)
  const uint16_t* bq;        // packed bf16 weight terms [3][4 chunks][n_nets*32 rows][128 B]
  const float* scale;        // [n_nets][32] = 1/255
  float* out;                // [n_nets][n][400][32] fp32 (NHWC)
  int relu;
};

template <int N_NETS>
__global__ void __launch_bounds__(THREADS, 1)
k_conv1_fused(const __grid_constant__ Params P) {
  constexpr int N_TOTAL = N_NETS * C_OUT;               // MMA N: 32 or 64
  constexpr int B_TERM_BYTES = N_TOTAL * K_TOTAL * 2;   // one bf16 term of all networks: 16 / 32 KiB
  constexpr int B_BYTES = NSPLIT * B_TERM_BYTES;        // 48 / 96 KiB
  constexpr int B_CHUNK_BYTES = N_TOTAL * 128;          // one input channel of one term
  constexpr int A_STAGES = (N_NETS == 1) ? 3 : 2;       // 32 KiB half tiles; SMEM budget 227 KiB
  constexpr uint32_t TMEM_COLS = (2 * N_TOTAL < 32) ? 32 : 2 * N_TOTAL;   // double-buffered accumulator
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B atoms must be 1024-byte aligned in the shared window: align by hand (1 KiB slack reserved)
  uint8_t* smem = smem_raw + ((1024u - (sptr(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;
  uint8_t* sA = smem + B_BYTES;
  uint8_t* sRaw = sA + A_STAGES * HALF_BYTES;
  __shared__ __align__(8) uint64_t b_full, raw_full[2], raw_empty[2], a_full[A_STAGES], a_empty[A_STAGES],
      t_full[2], t_empty[2];
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&b_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 128); }
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sptr(&s_tmem)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int64_t first = blockIdx.x, stride = gridDim.x;

  if (warp == 0) {
    // ------------------------------ TMA loader ------------------------------
    if (lane == 0) {
      mbar_expect_tx(&b_full, B_BYTES);
      for (int off = 0; off < B_BYTES; off += 16384) bulk_g2s(sB + off, (const uint8_t*)P.bq + off, 16384, &b_full);
      int it = 0;
      for (int64_t k = first; k < P.n; k += stride, ++it) {
        const int s = it & 1;
        mbar_wait(&raw_empty[s], ((it >> 1) & 1) ^ 1);
        int64_t row = P.idx ? P.idx[k] : k;
        row = row < 0 ? 0 : (row >= P.capacity ? P.capacity - 1 : row);
        mbar_expect_tx(&raw_full[s], FRAME_BYTES);
        bulk_g2s(sRaw + s * RAW_STRIDE, P.frames + row * FRAME_BYTES, FRAME_BYTES, &raw_full[s]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(N_TOTAL);
      mbar_wait(&b_full, 0);
      tc_fence_after();
      const uint32_t b_base = sptr(sB);
      int ah = 0;   // A half-tile counter
      int at = 0;   // tile counter (accumulator ring)
      for (int64_t k = first; k < P.n; k += stride) {
        for (int t = 0; t < TILES; ++t, ++at) {
          const int acc = at & 1;
          mbar_wait(&t_empty[acc], ((at >> 1) & 1) ^ 1);
          const uint32_t d = tmem + (uint32_t)(acc * N_TOTAL);
#pragma unroll
          for (int h = 0; h < 2; ++h, ++ah) {
            const int stage = ah % A_STAGES;
            mbar_wait(&a_full[stage], (ah / A_STAGES) & 1);
            tc_fence_after();
            const uint32_t a_base = sptr(sA + stage * HALF_BYTES);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = make_desc(a_base + cc * A_CHUNK_BYTES + ks * 32);
#pragma unroll
                for (int j = 0; j < NSPLIT; ++j) {
                  const uint64_t bd =
                      make_desc(b_base + j * B_TERM_BYTES + (2 * h + cc) * B_CHUNK_BYTES + ks * 32);
                  tc_mma_bf16(d, ad, bd, idesc, (h | cc | ks | j) ? 1u : 0u);
                }
              }
            }
            tc_commit(&a_empty[stage]);   // SMEM stage reusable once these MMAs have read it
          }
          tc_commit(&t_full[acc]);        // accumulator complete
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // --------------------------- im2col producers ---------------------------
    const int r_local = threadIdx.x - 128;       // one A-tile row per thread
    int ah = 0, it = 0;
    for (int64_t k = first; k < P.n; k += stride, ++it) {
      const int s = it & 1;
      mbar_wait(&raw_full[s], (it >> 1) & 1);
      const uint8_t* raw = sRaw + s * RAW_STRIDE;
      for (int t = 0; t < TILES; ++t) {
        const int p = t * TILE_M + r_local;
        const int oy = p / OHW, ox = p - oy * OHW;
        const uint8_t* src_row = raw + (STRIDE * oy) * HW + STRIDE * ox;
        const int sw = r_local & 7;
#pragma unroll
        for (int h = 0; h < 2; ++h, ++ah) {
          const int stage = ah % A_STAGES;
          mbar_wait(&a_empty[stage], ((ah / A_STAGES) & 1) ^ 1);
          if (p < POS) {
            uint8_t* dst_row = sA + stage * HALF_BYTES + (r_local >> 3) * 1024 + sw * 128;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const uint8_t* src_c = src_row + (2 * h + cc) * (HW * HW);
#pragma unroll
              for (int ky = 0; ky < KS; ++ky) {
                const uint32_t* s0 = reinterpret_cast<const uint32_t*>(src_c + ky * HW);
                const uint32_t lo = s0[0], hi = s0[1];
                // byte v -> fp32 via the 2^23 trick (PRMT builds 0x4B0000vv, one FADD removes 2^23: exact);
                // the upper 16 bits of that fp32 are bf16(v); a third PRMT packs two of them.
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t w = (q < 2) ? lo : hi;
                  const uint32_t sel0 = 0x7440u | (uint32_t)((q & 1) * 2), sel1 = sel0 + 1u;
                  const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, sel0)) - 8388608.0f;
                  const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, sel1)) - 8388608.0f;
                  o[q] = __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632u);
                }
                *reinterpret_cast<uint4*>(dst_row + cc * A_CHUNK_BYTES + ((ky ^ sw) << 4)) =
                    make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          }
          fence_async_smem();            // generic-proxy writes -> visible to the tensor core (async proxy)
          mbar_arrive(&a_full[stage]);
        }
      }
      mbar_arrive(&raw_empty[s]);        // this thread is done reading the raw frame
    }
  } else if (warp >= 8) {
    // ------------------------------- epilogue -------------------------------
    const int wq = warp & 3;                     // TMEM lane quarter this warp may access
    const int r_local = wq * 32 + lane;
    int at = 0;
    for (int64_t k = first; k < P.n; k += stride) {
      for (int t = 0; t < TILES; ++t, ++at) {
        const int acc = at & 1;
        mbar_wait(&t_full[acc], (at >> 1) & 1);
        tc_fence_after();
        const int p = t * TILE_M + r_local;
        const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * N_TOTAL);
        uint32_t v[N_TOTAL / 16][16];
#pragma unroll
        for (int g = 0; g < N_TOTAL / 16; ++g) tc_ld16(tbase + g * 16, v[g]);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(&t_empty[acc]);              // accumulator is in registers: release it early
        if (p < POS) {
#pragma unroll
          for (int net = 0; net < N_NETS; ++net) {
            float* o = P.out + (((int64_t)net * P.n + k) * POS + p) * C_OUT;
            const float sc = __ldg(P.scale + net * C_OUT);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                float y[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float x = __uint_as_float(v[net * 2 + g][e4 * 4 + e]) * sc;
                  y[e] = (P.relu && x < 0.0f) ? 0.0f : x;
                }
                *reinterpret_cast<float4*>(o + g * 16 + e4 * 4) = make_float4(y[0], y[1], y[2], y[3]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
  }
}

template <int N_NETS>
constexpr size_t smem_bytes() {
  return (size_t)NSPLIT * N_NETS * C_OUT * K_TOTAL * 2 + (size_t)((N_NETS == 1) ? 3 : 2) * HALF_BYTES +
         2 * (size_t)RAW_STRIDE + 1024;
}

}  // namespace conv1
}  // namespace b2rl

using namespace b2rl;

extern "C" int b2rl_conv1_pack(const float* w_dev, int32_t net, int32_t n_nets, void* bq_out_dev,
                               float* scale_out_dev, void* stream) {
  B2RL_REQUIRE(w_dev && bq_out_dev && scale_out_dev, "null argument");
  B2RL_REQUIRE(n_nets >= 1 && n_nets <= 2 && net >= 0 && net < n_nets, "n_nets must be 1 or 2");
  conv1::k_conv1_pack<<<conv1::C_OUT, conv1::K_TOTAL, 0, (cudaStream_t)stream>>>(w_dev, net, n_nets,
                                                                                (uint16_t*)bq_out_dev, scale_out_dev);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_conv1_fused(const uint8_t* frames_dev, int64_t capacity, const int64_t* idx_dev, int64_t n,
                                const void* bq_dev, const float* scale_dev, int32_t n_nets, float* out_dev,
                                int32_t relu, void* stream) {
  B2RL_REQUIRE(n >= 0, "negative n");
  if (n == 0) return B2RL_OK;
  B2RL_REQUIRE(frames_dev && bq_dev && scale_dev && out_dev, "null argument");
  B2RL_REQUIRE(n_nets == 1 || n_nets == 2, "n_nets must be 1 or 2");
  B2RL_REQUIRE(capacity >= 1, "capacity must be positive");
  B2RL_REQUIRE(((uintptr_t)frames_dev % 16 == 0) && ((uintptr_t)bq_dev % 16 == 0) && ((uintptr_t)out_dev % 16 == 0),
               "frames, packed weights and output must be 16-byte aligned");
  int dev = 0;
  B2RL_CUDA(cudaGetDevice(&dev));
  static int sms[64] = {0};
  static bool init[64] = {false};
  if (!init[dev & 63]) {
    B2RL_CUDA(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    B2RL_CUDA(cudaFuncSetAttribute(conv1::k_conv1_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)conv1::smem_bytes<1>()));
    B2RL_CUDA(cudaFuncSetAttribute(conv1::k_conv1_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)conv1::smem_bytes<2>()));
    init[dev & 63] = true;
  }
  conv1::Params P{frames_dev, idx_dev, n, capacity, (const uint16_t*)bq_dev, scale_dev, out_dev, relu};
  const unsigned grid = (unsigned)((n < sms[dev & 63]) ? n : sms[dev & 63]);
  if (n_nets == 1)
    conv1::k_conv1_fused<1><<<grid, conv1::THREADS, conv1::smem_bytes<1>(), (cudaStream_t)stream>>>(P);
  else
    conv1::k_conv1_fused<2><<<grid, conv1::THREADS, conv1::smem_bytes<2>(), (cudaStream_t)stream>>>(P);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}
