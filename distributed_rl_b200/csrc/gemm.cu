// fp32-accurate dense layers on the tensor cores: C[M][N] (+)= A[M][K] * B[N][K]^T with 3xTF32.
//
// The dense heads of the Q-network (3136 -> 512, twice; cfg/ape_x.json:52-71) run as cuBLAS fp32
// SIMT GEMMs at PyTorch's default precision and are 31 % of the learner step.
// TF32 alone (10-bit mantissa) is not what the reference computes, so every fp32 operand is split
//      x = hi + lo,   hi = rn_tf32(x),  lo = x - hi   (exact in fp32)
// and the product is formed as  hi*hi + hi*lo + lo*hi  on tcgen05 (kind::tf32, fp32 accumulation in
// TMEM); the dropped lo*lo term is 2^-22 relative.  SURVEY.md §8f rank 2 ("TF32x3 policy").
//
//   k_split_pack   fp32 matrix (optionally transposed) -> {hi, lo} operand images in exactly the
//                  128B-swizzled, K-major tile layout the MMA reads, so the GEMM's loader is a
//                  plain cp.async.bulk per tile (no tensor map, no SM-side staging)
//   k_gemm_tf32x3  one CTA per (m-tile 128, n-tile 256, K-split): TMA loader warp, one MMA-issuing
//                  thread (12 x tcgen05.mma per 32-float K chunk), 4 epilogue warps that store the tile
//                  (or, when K is split, this split's partial tile) with coalesced 512-byte stores
//   k_splitk_reduce sums the K-split partials in split order: the result is deterministic (no atomics)
#include "common.cuh"

#include <stdlib.h>

namespace b2rl {
namespace gemm {

constexpr int TM = 128, TN = 256, KC = 32;          // tile rows of A / of B, floats per K chunk (128 B)
constexpr int A_TILE = TM * 128, B_TILE = TN * 128;  // bytes of one {term, k-chunk} tile: 16 KiB / 32 KiB
constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;       // hi+lo of both operands: 96 KiB
constexpr int STAGES = 2;
constexpr int THREADS = 224;                         // warp 0 loader, 1 MMA, 2 TMEM alloc, 3-6 epilogue

__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(b)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sptr(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc),
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
// K-major SW128 descriptor (see csrc/conv1.cu): SBO = 1024 B, LBO = 16 B, version 1, layout SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, N>>3 @17, M>>4 @24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);


// ---- cta_group::2 helpers (CTA pair: one 256-row MMA over two SMs) ---------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same CTA-relative address in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(sptr(b)), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(sptr(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {   // arrives on `bar` of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   sptr(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc),
      "r"(idesc), "r"(accumulate) : "memory");
}
// M = 256 over the pair, N = 256
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
constexpr int STAGE2 = 2 * A_TILE + B_TILE;          // A hi+lo (my 128 rows) + my HALF of B hi+lo: 64 KiB
constexpr int STAGES2 = 3;

// ---- operand packing ---------------------------------------------------------
// Image layout: [term 0=hi,1=lo][k_chunk][row_tile][row_in_tile][128 B, 16-byte units XOR (row & 7)]
// One CTA per (32 operand rows, one K chunk); thread = (row r = tid / 8, 16-byte unit = tid % 8), so
// both the source row segment and the image row are one contiguous 128 B per 8 threads.
// TRANSPOSE: the operand's rows are the source's columns; the 32x32 block goes through SMEM so that
// the source is still read along its contiguous dimension.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  const uint32_t u = __float_as_uint(x);
  uint32_t h = (u + 0x1000u) & 0xFFFFE000u;                        // round to nearest on the 13 dropped bits
  if ((h & 0x7F800000u) == 0x7F800000u) h = u & 0xFFFFE000u;        // rounding reached inf (or x is inf/nan): truncate
  hi = __uint_as_float(h);
  lo = ((u & 0x7F800000u) == 0x7F800000u) ? 0.0f : x - hi;          // exact in fp32
  if ((u & 0x7F800000u) == 0x7F800000u) hi = x;
}

template <bool TRANSPOSE>
__global__ void __launch_bounds__(256)
k_split_pack(const float* __restrict__ src, int src_rows, int src_cols, int64_t src_ld, int tile_rows,
             float* __restrict__ out, int rows_pad, int k_chunks, int row_off, int kc_off) {
  // (row, kc) below are local to this piece; the image position is (row_off + row, kc_off + kc)
  const int kc = blockIdx.y, row0 = blockIdx.x * 32;
  const int r = threadIdx.x >> 3, unit = threadIdx.x & 7;
  const int row = row0 + r;
  float v[4];
  if (TRANSPOSE) {
    __shared__ float tile[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc * KC + w * 4 + j, c = row0 + lane;            // source row k, source column c
      tile[w * 4 + j][lane] = (k < src_rows && c < src_cols) ? src[(int64_t)k * src_ld + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = tile[unit * 4 + e][r];
  } else {
    const int k = kc * KC + unit * 4;
    const float* p = src + (int64_t)row * src_ld + k;
    if (row < src_rows && k + 3 < src_cols && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4 q = *reinterpret_cast<const float4*>(p);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = (row < src_rows && k + e < src_cols) ? p[e] : 0.0f;
    }
  }
  const int irow = row_off + row, ikc = kc_off + kc;
  if (irow >= rows_pad) return;
  float hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_tf32(v[e], hi[e], lo[e]);
  const int rt = irow / tile_rows, rr = irow - rt * tile_rows;
  const int tiles = rows_pad / tile_rows;
  const int64_t off = (((int64_t)ikc * tiles + rt) * tile_rows + rr) * 32 + ((unit ^ (rr & 7)) << 2);
  const int64_t term_stride = (int64_t)k_chunks * rows_pad * 32;
  *reinterpret_cast<float4*>(out + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(out + term_stride + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// ---- activation-side packs that fold act_3 (ReLU) + nn.Flatten into the heads' operand images ----------------
// The conv stack's output is NHWC in memory: y[b][hw][c].  The reference flattens the logical NCHW tensor
// (baseline/baseNetwork.py:204-209), so the heads' weights index features as f = c*HW + hw.  Instead of a ReLU
// kernel plus a permuting copy per pass, these kernels read y coalesced, transpose through shared memory and write
// the operand images in f order directly (the weights' packs stay as they are).
//   k_pack_act_nhwc<false>: A-role image of x = relu(y) viewed [B][K = C*HW]          (forward)
//   k_pack_act_nhwc<true> : B-role image of x^T [K rows][contraction B]               (weight gradient)
//   k_unflatten_relu_mask : dL/dy[b][hw][c] = gx[b][c*HW + hw] * (y[b][hw][c] > 0)    (input gradient)
constexpr int ACT_PAD = 1;      // shared-memory row padding: kills the bank conflicts of the transposed reads

template <bool TRANSPOSE>
__global__ void __launch_bounds__(256)
k_pack_act_nhwc(const float* __restrict__ y, int B, int HW, int C, int relu, float* __restrict__ out,
                int rows_pad, int k_chunks) {
  extern __shared__ float s_act[];
  const int K = C * HW;
  if (!TRANSPOSE) {
    // one CTA per image row b (rows >= B are zero padding): s[hw][c], row stride C + 1
    const int b = blockIdx.x;
    const int ld = C + ACT_PAD;
    const uint32_t magic_c = (uint32_t)((0x100000000ull + C - 1) / C);      // i / C == (i * magic) >> 32 for i < 2^24
    const uint32_t magic_hw = (uint32_t)((0x100000000ull + HW - 1) / HW);
    if (b < B) {
      const float* src = y + (int64_t)b * K;
      for (int i = threadIdx.x; i < K; i += 256) {
        float v = src[i];
        if (relu) v = fmaxf(v, 0.0f);
        const int hw = (int)__umulhi((uint32_t)i, magic_c);
        s_act[hw * ld + (i - hw * C)] = v;
      }
    }
    __syncthreads();
    const int rt = b / TM, rr = b - rt * TM, tiles = rows_pad / TM;
    const int64_t term_stride = (int64_t)k_chunks * rows_pad * 32;
    for (int w = threadIdx.x; w < k_chunks * 8; w += 256) {
      const int kc = w >> 3, unit = w & 7;
      const int f0 = kc * KC + unit * 4;
      int c = (int)__umulhi((uint32_t)f0, magic_hw), hw = f0 - c * HW;      // f = c*HW + hw, walked incrementally
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = (b < B && f0 + e < K) ? s_act[hw * ld + c] : 0.0f;
        if (++hw == HW) { hw = 0; ++c; }
      }
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(v[e], hi[e], lo[e]);
      const int64_t off = (((int64_t)kc * tiles + rt) * TM + rr) * 32 + ((unit ^ (rr & 7)) << 2);
      *reinterpret_cast<float4*>(out + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(out + term_stride + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  } else {
    // one CTA per (chunk of 32 b's, hw): s[b][c], row stride C + 1; image rows f = c*HW + hw, contraction = b
    const int kc = blockIdx.x, hw = blockIdx.y;
    const int ld = C + ACT_PAD;
    for (int i = threadIdx.x; i < 32 * C; i += 256) {
      const int bb = i / C, c = i - bb * C, b = kc * KC + bb;
      float v = (b < B) ? y[((int64_t)b * HW + hw) * C + c] : 0.0f;
      if (relu) v = fmaxf(v, 0.0f);
      s_act[bb * ld + c] = v;
    }
    __syncthreads();
    const int tiles = rows_pad / TN;
    const int64_t term_stride = (int64_t)k_chunks * rows_pad * 32;
    for (int w = threadIdx.x; w < C * 8; w += 256) {
      const int c = w >> 3, unit = w & 7;
      const int f = c * HW + hw;
      float v[4], hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = s_act[(unit * 4 + e) * ld + c];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(v[e], hi[e], lo[e]);
      const int rt = f / TN, rr = f - rt * TN;
      const int64_t off = (((int64_t)kc * tiles + rt) * TN + rr) * 32 + ((unit ^ (rr & 7)) << 2);
      *reinterpret_cast<float4*>(out + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(out + term_stride + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// zero rows [K, rows_pad) of the x^T image (the padding of the last row tile)
__global__ void __launch_bounds__(256)
k_pack_zero_rows(float* __restrict__ out, int row0, int rows_pad, int k_chunks, int tile_rows) {
  const int tiles = rows_pad / tile_rows;
  const int64_t term_stride = (int64_t)k_chunks * rows_pad * 32;
  const int n_rows = rows_pad - row0;
  const int64_t total = (int64_t)n_rows * k_chunks * 8;
  for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < total; w += (int64_t)gridDim.x * 256) {
    const int unit = (int)(w & 7);
    const int64_t q = w >> 3;
    const int kc = (int)(q / n_rows), f = row0 + (int)(q - (int64_t)kc * n_rows);
    const int rt = f / tile_rows, rr = f - rt * tile_rows;
    const int64_t off = (((int64_t)kc * tiles + rt) * tile_rows + rr) * 32 + ((unit ^ (rr & 7)) << 2);
    *reinterpret_cast<float4*>(out + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(out + term_stride + off) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(256)
k_unflatten_relu_mask(const float* __restrict__ gx, int64_t gx_ld, const float* __restrict__ y, int HW, int C,
                      float* __restrict__ out) {
  extern __shared__ float s_act[];          // gx row in f order: s[c*HW + hw]
  const int b = blockIdx.x, K = C * HW;
  for (int i = threadIdx.x; i < K; i += 256) s_act[i] = gx[(int64_t)b * gx_ld + i];
  __syncthreads();
  const float* yr = y + (int64_t)b * K;
  float* o = out + (int64_t)b * K;
  for (int i = threadIdx.x; i < K; i += 256) {      // i = hw*C + c (coalesced reads of y, writes of out)
    const int hw = i / C, c = i - hw * C;
    o[i] = (yr[i] > 0.0f) ? s_act[c * HW + hw] : 0.0f;
  }
}

struct Params {
  const float* a;        // packed A image (tile_rows = 128)
  const float* b;        // packed B image (tile_rows = 256)
  float* c;              // splits == 1: C [M][ldc] fp32 (stored); else the partials [split][M][ldc]
  int64_t M, N, ldc;     // logical sizes (rows beyond M / columns beyond N are dropped)
  int64_t m_tiles, n_tiles, k_chunks;
  int32_t splits;        // K splits (gridDim.z)
};

__global__ void __launch_bounds__(THREADS, 1)
k_gemm_tf32x3(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (sptr(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], acc_full;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t mt = blockIdx.x, nt = blockIdx.y;
  // K range of this split
  const int64_t per = (P.k_chunks + P.splits - 1) / P.splits;
  const int64_t k0 = (int64_t)blockIdx.z * per;
  const int64_t k1 = (k0 + per < P.k_chunks) ? k0 + per : P.k_chunks;
  const int64_t nk = k1 - k0;   // > 0: the host never launches an empty trailing split (gemm_splits)

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sptr(&s_tmem)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (nk > 0) {
    const int64_t a_term = P.k_chunks * P.m_tiles * (TM * 32);   // floats between the hi and lo images
    const int64_t b_term = P.k_chunks * P.n_tiles * (TN * 32);
    if (warp == 0) {
      if (lane == 0) {
        for (int64_t i = 0; i < nk; ++i) {
          const int s = (int)(i % STAGES);
          mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
          const int64_t kc = k0 + i;
          const float* a_hi = P.a + (kc * P.m_tiles + mt) * (TM * 32);
          const float* b_hi = P.b + (kc * P.n_tiles + nt) * (TN * 32);
          uint8_t* st = smem + (size_t)s * STAGE;
          mbar_expect_tx(&full[s], STAGE);
          bulk_g2s(st, a_hi, A_TILE, &full[s]);
          bulk_g2s(st + A_TILE, a_hi + a_term, A_TILE, &full[s]);
          bulk_g2s(st + 2 * A_TILE, b_hi, B_TILE, &full[s]);
          bulk_g2s(st + 2 * A_TILE + B_TILE, b_hi + b_term, B_TILE, &full[s]);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        for (int64_t i = 0; i < nk; ++i) {
          const int s = (int)(i % STAGES);
          mbar_wait(&full[s], (i / STAGES) & 1);
          tc_fence_after();
          const uint32_t base = sptr(smem + (size_t)s * STAGE);
          const uint32_t a_hi = base, a_lo = base + A_TILE, b_hi = base + 2 * A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t o = ks * 32;
            tc_mma_tf32(tmem, make_desc(a_lo + o), make_desc(b_hi + o), IDESC, (i | ks) ? 1u : 0u);   // small terms first
            tc_mma_tf32(tmem, make_desc(a_hi + o), make_desc(b_lo + o), IDESC, 1u);
            tc_mma_tf32(tmem, make_desc(a_hi + o), make_desc(b_hi + o), IDESC, 1u);
          }
          tc_commit(&empty[s]);
        }
        tc_commit(&acc_full);
      }
    } else if (warp >= 3) {
      // ------------------------------- epilogue -------------------------------
      const int wq = warp & 3;                          // TMEM lane quarter (warps 3..6 -> 3,0,1,2)
      mbar_wait(&acc_full, 0);
      tc_fence_after();
      uint8_t* stg = smem + (size_t)wq * 4096;          // pipeline SMEM is idle now: 32 rows x 128 B per warp
      const int64_t row0 = mt * TM + wq * 32;
      const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16);
      for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t v0[16], v1[16];
        tc_ld16(tbase + c0, v0);
        tc_ld16(tbase + c0 + 16, v1);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) =
              make_uint4(v0[4 * g], v0[4 * g + 1], v0[4 * g + 2], v0[4 * g + 3]);
          *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + g) ^ (lane & 7)) << 4)) =
              make_uint4(v1[4 * g], v1[4 * g + 1], v1[4 * g + 2], v1[4 * g + 3]);
        }
        __syncwarp();
        const int64_t col0 = nt * TN + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int off = (i * 32 + lane) * 16;
          const int row = off >> 7, unit = (off >> 4) & 7;
          const int64_t m = row0 + row, n = col0 + unit * 4;
          float* dst = P.c + ((int64_t)blockIdx.z * P.M + m) * P.ldc + n;
          if (m < P.M && n + 3 < P.N) {
            *reinterpret_cast<float4*>(dst) =
                *reinterpret_cast<const float4*>(stg + row * 128 + ((unit ^ (row & 7)) << 4));
          } else if (m < P.M && n < P.N) {
            const float* x = reinterpret_cast<const float*>(stg + row * 128 + ((unit ^ (row & 7)) << 4));
            for (int e = 0; e < 4 && n + e < P.N; ++e) dst[e] = x[e];
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
  }
}


// ---- the same GEMM on CTA pairs (tcgen05 cta_group::2) ------------------------------------------------------------
// Two CTAs of a cluster own m-tiles 2c and 2c+1 of the same n-tile and K split.  The leader (rank 0) issues ONE
// M = 256 MMA per step for both; A comes from each CTA's own shared memory (its 128 rows) and every CTA holds only
// its HALF of the B tile (128 of the 256 rows), which the pair's tensor cores share.  Per SM and K chunk 64 KiB are
// staged instead of 96 KiB, for the same MMA work per SM: the single-CTA kernel is bound by exactly that
// (shared-memory fill + operand reads per SM; tensor pipe 49 % active), and the freed space buys a
// third pipeline stage.  Barriers: every CTA's loader arms its LOCAL full barrier (tx bytes); a relay thread per CTA
// forwards "my stage is full" to the leader's pair barrier (count 2, remote arrive); the leader's tcgen05.commit is
// multicast to both CTAs' empty / accumulator barriers.
__global__ void __launch_bounds__(THREADS, 1)
k_gemm_tf32x3_2cta(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (sptr(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[STAGES2], pair_full[STAGES2], empty[STAGES2], acc_full;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int64_t mt = blockIdx.x, nt = blockIdx.y;     // blockIdx.x = 2 * pair + crank
  const int64_t per = (P.k_chunks + P.splits - 1) / P.splits;
  const int64_t k0 = (int64_t)blockIdx.z * per;
  const int64_t k1 = (k0 + per < P.k_chunks) ? k0 + per : P.k_chunks;
  const int64_t nk = k1 - k0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES2; ++i) { mbar_init(&full[i], 1); mbar_init(&pair_full[i], 2); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  cluster_sync_all();                 // both CTAs' barriers exist before anything can arrive at them
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sptr(&s_tmem)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  cluster_sync_all();                 // both halves of the pair's TMEM are allocated

  if (nk > 0) {
    const int64_t a_term = P.k_chunks * P.m_tiles * (TM * 32);
    const int64_t b_term = P.k_chunks * P.n_tiles * (TN * 32);
    if (warp == 0) {
      if (lane == 0) {                // loader: my A tile + my half of the B tile
        for (int64_t i = 0; i < nk; ++i) {
          const int s = (int)(i % STAGES2);
          mbar_wait(&empty[s], ((i / STAGES2) & 1) ^ 1);
          const int64_t kc = k0 + i;
          const float* a_hi = P.a + (kc * P.m_tiles + mt) * (TM * 32);
          const float* b_hi = P.b + (kc * P.n_tiles + nt) * (TN * 32) + (int64_t)crank * ((TN / 2) * 32);
          uint8_t* st = smem + (size_t)s * STAGE2;
          mbar_expect_tx(&full[s], STAGE2);
          bulk_g2s(st, a_hi, A_TILE, &full[s]);
          bulk_g2s(st + A_TILE, a_hi + a_term, A_TILE, &full[s]);
          bulk_g2s(st + 2 * A_TILE, b_hi, B_TILE / 2, &full[s]);
          bulk_g2s(st + 2 * A_TILE + B_TILE / 2, b_hi + b_term, B_TILE / 2, &full[s]);
        }
      }
    } else if (warp == 2) {
      if (lane == 0) {                // relay: my stage landed -> tell the leader's pair barrier
        for (int64_t i = 0; i < nk; ++i) {
          const int s = (int)(i % STAGES2);
          mbar_wait(&full[s], (i / STAGES2) & 1);
          mbar_arrive_remote(&pair_full[s], 0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && crank == 0) {  // the leader issues the pair's MMAs
        for (int64_t i = 0; i < nk; ++i) {
          const int s = (int)(i % STAGES2);
          mbar_wait_cluster(&pair_full[s], (i / STAGES2) & 1);
          tc_fence_after();
          const uint32_t base = sptr(smem + (size_t)s * STAGE2);
          const uint32_t a_hi = base, a_lo = base + A_TILE, b_hi = base + 2 * A_TILE, b_lo = b_hi + B_TILE / 2;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t o = ks * 32;
            tc_mma_tf32_2cta(tmem, make_desc(a_lo + o), make_desc(b_hi + o), IDESC2, (i | ks) ? 1u : 0u);   // small terms first
            tc_mma_tf32_2cta(tmem, make_desc(a_hi + o), make_desc(b_lo + o), IDESC2, 1u);
            tc_mma_tf32_2cta(tmem, make_desc(a_hi + o), make_desc(b_hi + o), IDESC2, 1u);
          }
          tc_commit2(&empty[s]);      // both CTAs' stage s may be refilled
        }
        tc_commit2(&acc_full);        // both CTAs' halves of the accumulator are complete
      }
    } else if (warp >= 3) {
      // ------------------------------- epilogue (my 128 rows) -------------------------------
      const int wq = warp & 3;
      mbar_wait(&acc_full, 0);
      tc_fence_after();
      uint8_t* stg = smem + (size_t)wq * 4096;
      const int64_t row0 = mt * TM + wq * 32;
      const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16);
      for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t v0[16], v1[16];
        tc_ld16(tbase + c0, v0);
        tc_ld16(tbase + c0 + 16, v1);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) =
              make_uint4(v0[4 * g], v0[4 * g + 1], v0[4 * g + 2], v0[4 * g + 3]);
          *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + g) ^ (lane & 7)) << 4)) =
              make_uint4(v1[4 * g], v1[4 * g + 1], v1[4 * g + 2], v1[4 * g + 3]);
        }
        __syncwarp();
        const int64_t col0 = nt * TN + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int off = (i * 32 + lane) * 16;
          const int row = off >> 7, unit = (off >> 4) & 7;
          const int64_t m = row0 + row, n = col0 + unit * 4;
          float* dst = P.c + ((int64_t)blockIdx.z * P.M + m) * P.ldc + n;
          if (m < P.M && n + 3 < P.N) {
            *reinterpret_cast<float4*>(dst) =
                *reinterpret_cast<const float4*>(stg + row * 128 + ((unit ^ (row & 7)) << 4));
          } else if (m < P.M && n < P.N) {
            const float* x = reinterpret_cast<const float*>(stg + row * 128 + ((unit ^ (row & 7)) << 4));
            for (int e = 0; e < 4 && n + e < P.N; ++e) dst[e] = x[e];
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // neither CTA frees TMEM / exits while the pair may still use it
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
  }
}

// C[i] = sum over splits (in split order: deterministic) of partial[z][i]; one thread per 4 columns
__global__ void __launch_bounds__(256)
k_splitk_reduce(const float* __restrict__ partial, int splits, int64_t M, int64_t N, int64_t ldc, float* __restrict__ c) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_row = ldc >> 2;
  if (q >= M * per_row) return;
  const int64_t m = q / per_row, n = (q - m * per_row) << 2;
  if (n >= N) return;
  const int64_t off = m * ldc + n, stride = M * ldc;
  if (n + 3 < N) {
    float4 a = *reinterpret_cast<const float4*>(partial + off);
    for (int z = 1; z < splits; ++z) {
      const float4 b = *reinterpret_cast<const float4*>(partial + z * stride + off);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(c + off) = a;
  } else {
    for (int e = 0; n + e < N; ++e) {
      float a = partial[off + e];
      for (int z = 1; z < splits; ++z) a += partial[z * stride + off + e];
      c[off + e] = a;
    }
  }
}

}  // namespace gemm
}  // namespace b2rl

using namespace b2rl;

extern "C" int64_t b2rl_gemm_packed_floats(int64_t rows, int64_t k, int32_t b_role) {
  const int64_t tr = b_role ? gemm::TN : gemm::TM;
  const int64_t rows_pad = (rows + tr - 1) / tr * tr, kc = (k + gemm::KC - 1) / gemm::KC;
  return 2 * rows_pad * kc * gemm::KC;
}

// One piece of an operand: the piece's rows go to image rows [row_offset, ...), its contraction index to
// [k_offset, ...) of an operand with total_rows x total_k.  Pieces tile the operand (vertically stacked weight
// matrices, or their transposes side by side); the piece that ends the operand also writes the zero padding.
extern "C" int b2rl_gemm_split_pack_into(const float* src_dev, int64_t src_rows, int64_t src_cols, int64_t src_ld,
                                         int32_t transpose, int32_t b_role, float* out_dev, int64_t total_rows,
                                         int64_t total_k, int64_t row_offset, int64_t k_offset, void* stream) {
  B2RL_REQUIRE(src_dev && out_dev, "null argument");
  B2RL_REQUIRE(src_rows >= 1 && src_cols >= 1 && src_ld >= src_cols, "bad shape");
  B2RL_REQUIRE(((uintptr_t)out_dev % 16) == 0, "packed operand must be 16-byte aligned");
  const int64_t rows = transpose ? src_cols : src_rows, k = transpose ? src_rows : src_cols;
  B2RL_REQUIRE(row_offset >= 0 && k_offset >= 0 && row_offset + rows <= total_rows && k_offset + k <= total_k,
               "piece outside the operand");
  B2RL_REQUIRE(row_offset % 32 == 0 && k_offset % gemm::KC == 0, "piece offsets must be multiples of 32");
  B2RL_REQUIRE((rows % 32 == 0 || row_offset + rows == total_rows) && (k % gemm::KC == 0 || k_offset + k == total_k),
               "an inner piece must be a multiple of 32 rows / 32 contraction elements");
  const int tr = b_role ? gemm::TN : gemm::TM;
  const int64_t rows_pad = (total_rows + tr - 1) / tr * tr, kc_total = (total_k + gemm::KC - 1) / gemm::KC;
  B2RL_REQUIRE(rows_pad < (1 << 30) && kc_total <= 65535, "operand too large");
  const int64_t rows_cover = (row_offset + rows == total_rows) ? rows_pad - row_offset : rows;
  const int64_t kc_cover = (k + gemm::KC - 1) / gemm::KC;
  dim3 grid((unsigned)(rows_cover / 32), (unsigned)kc_cover);
  if (transpose)
    gemm::k_split_pack<true><<<grid, 256, 0, (cudaStream_t)stream>>>(src_dev, (int)src_rows, (int)src_cols, src_ld, tr,
                                                                     out_dev, (int)rows_pad, (int)kc_total,
                                                                     (int)row_offset, (int)(k_offset / gemm::KC));
  else
    gemm::k_split_pack<false><<<grid, 256, 0, (cudaStream_t)stream>>>(src_dev, (int)src_rows, (int)src_cols, src_ld, tr,
                                                                      out_dev, (int)rows_pad, (int)kc_total,
                                                                      (int)row_offset, (int)(k_offset / gemm::KC));
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_gemm_split_pack(const float* src_dev, int64_t src_rows, int64_t src_cols, int64_t src_ld,
                                    int32_t transpose, int32_t b_role, float* out_dev, void* stream) {
  const int64_t rows = transpose ? src_cols : src_rows, k = transpose ? src_rows : src_cols;
  return b2rl_gemm_split_pack_into(src_dev, src_rows, src_cols, src_ld, transpose, b_role, out_dev, rows, k, 0, 0, stream);
}

static int64_t gemm_splits(int64_t M, int64_t N, int64_t K, int sms) {
  const int64_t tiles = ((M + gemm::TM - 1) / gemm::TM) * ((N + gemm::TN - 1) / gemm::TN);
  const int64_t kc = (K + gemm::KC - 1) / gemm::KC;
  int64_t splits = sms / (tiles > 0 ? tiles : 1);      // cover the SMs about once (a function of the shape only)
  if (splits < 1) splits = 1;
  if (splits > kc) splits = kc;
  const int64_t per = (kc + splits - 1) / splits;
  return (kc + per - 1) / per;                          // no empty trailing split
}

static int gemm_sms(int* out) {
  int dev = 0;
  B2RL_CUDA(cudaGetDevice(&dev));
  static int sms[64] = {0};
  if (!sms[dev & 63]) B2RL_CUDA(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  *out = sms[dev & 63];
  return B2RL_OK;
}

extern "C" int64_t b2rl_gemm_workspace_floats(int64_t M, int64_t N, int64_t K, int64_t ldc) {
  int sms = 0;
  if (gemm_sms(&sms) != B2RL_OK) return -1;
  const int64_t splits = gemm_splits(M, N, K, sms);
  return splits > 1 ? splits * M * ldc : 0;
}

extern "C" int b2rl_gemm_tf32x3(const float* a_packed_dev, const float* b_packed_dev, float* c_dev, int64_t M,
                                int64_t N, int64_t K, int64_t ldc, float* workspace_dev, void* stream) {
  B2RL_REQUIRE(a_packed_dev && b_packed_dev && c_dev, "null argument");
  B2RL_REQUIRE(M >= 1 && N >= 1 && K >= 1 && ldc >= N, "bad shape");
  B2RL_REQUIRE(((uintptr_t)c_dev % 16) == 0 && (ldc % 4) == 0, "C must be 16-byte aligned with ldc % 4 == 0");
  int sms = 0;
  if (int rc = gemm_sms(&sms)) return rc;
  int dev = 0;
  B2RL_CUDA(cudaGetDevice(&dev));
  static bool attr[64] = {false};
  const size_t smem_bytes = (size_t)gemm::STAGES * gemm::STAGE + 1024;
  const size_t smem_bytes2 = (size_t)gemm::STAGES2 * gemm::STAGE2 + 1024;
  if (!attr[dev & 63]) {
    B2RL_CUDA(cudaFuncSetAttribute(gemm::k_gemm_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    B2RL_CUDA(cudaFuncSetAttribute(gemm::k_gemm_tf32x3_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes2));
    attr[dev & 63] = true;
  }
  static int use_2cta = -1;
  if (use_2cta < 0) {
    const char* e = getenv("B2RL_GEMM_2CTA");     // opt-in: measured 23.5 us vs 21.5 us for the single-CTA kernel on the
    use_2cta = (e && e[0] == '1') ? 1 : 0;        // heads' shapes (the GEMM is overhead-bound there)
  }
  const int64_t splits = gemm_splits(M, N, K, sms);
  B2RL_REQUIRE(splits == 1 || (workspace_dev && ((uintptr_t)workspace_dev % 16) == 0),
               "this shape splits K: pass b2rl_gemm_workspace_floats() floats of 16-byte aligned workspace");
  gemm::Params P{};
  P.a = a_packed_dev; P.b = b_packed_dev;
  P.c = splits > 1 ? workspace_dev : c_dev;
  P.M = M; P.N = N; P.ldc = ldc;
  P.m_tiles = (M + gemm::TM - 1) / gemm::TM;
  P.n_tiles = (N + gemm::TN - 1) / gemm::TN;
  P.k_chunks = (K + gemm::KC - 1) / gemm::KC;
  P.splits = (int32_t)splits;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)P.m_tiles, (unsigned)P.n_tiles, (unsigned)splits);
  if (use_2cta && (P.m_tiles % 2) == 0) {       // CTA pairs along M (tcgen05 cta_group::2)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(gemm::THREADS); cfg.dynamicSmemBytes = smem_bytes2; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    B2RL_CUDA(cudaLaunchKernelEx(&cfg, gemm::k_gemm_tf32x3_2cta, P));
  } else {
    gemm::k_gemm_tf32x3<<<grid, gemm::THREADS, smem_bytes, st>>>(P);
  }
  count_launch();
  B2RL_CHECK_LAUNCH();
  if (splits > 1) {
    const int64_t quads = M * (ldc >> 2);
    gemm::k_splitk_reduce<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(workspace_dev, (int)splits, M, N, ldc, c_dev);
    count_launch();
    B2RL_CHECK_LAUNCH();
  }
  return B2RL_OK;
}


// flatten_NCHW(relu(y)) packed straight from the conv stack's NHWC output: act_3 + nn.Flatten of cfg/ape_x.json:37-51
// (baseline/baseNetwork.py:204-209) folded into the heads' operand packing.  transpose = 0: A-role image of
// x [B][C*HW] (forward); transpose = 1: B-role image of x^T (weight gradient of the heads).
extern "C" int b2rl_gemm_pack_act_nhwc(const float* y_dev, int64_t B, int64_t HW, int64_t C, int32_t relu, int32_t transpose,
                                       float* out_dev, void* stream) {
  B2RL_REQUIRE(y_dev && out_dev, "null argument");
  B2RL_REQUIRE(B >= 1 && HW >= 1 && C >= 1 && C * HW < (1 << 24) && B < (1 << 24), "bad shape");
  B2RL_REQUIRE(((uintptr_t)out_dev % 16) == 0, "packed operand must be 16-byte aligned");
  const int64_t K = C * HW;
  cudaStream_t st = (cudaStream_t)stream;
  if (!transpose) {
    const int64_t rows_pad = (B + gemm::TM - 1) / gemm::TM * gemm::TM, kc = (K + gemm::KC - 1) / gemm::KC;
    const size_t smem = (size_t)HW * (C + gemm::ACT_PAD) * sizeof(float);
    B2RL_REQUIRE(smem <= 48 * 1024, "activation row too large for the staging tile");
    gemm::k_pack_act_nhwc<false><<<(unsigned)rows_pad, 256, smem, st>>>(y_dev, (int)B, (int)HW, (int)C, relu, out_dev,
                                                                       (int)rows_pad, (int)kc);
    count_launch();
  } else {
    const int64_t rows_pad = (K + gemm::TN - 1) / gemm::TN * gemm::TN, kc = (B + gemm::KC - 1) / gemm::KC;
    const size_t smem = (size_t)32 * (C + gemm::ACT_PAD) * sizeof(float);
    B2RL_REQUIRE(smem <= 48 * 1024 && HW <= 65535, "activation tile too large");
    gemm::k_pack_act_nhwc<true><<<dim3((unsigned)kc, (unsigned)HW), 256, smem, st>>>(y_dev, (int)B, (int)HW, (int)C, relu,
                                                                                    out_dev, (int)rows_pad, (int)kc);
    count_launch();
    if (rows_pad > K) {
      gemm::k_pack_zero_rows<<<64, 256, 0, st>>>(out_dev, (int)K, (int)rows_pad, (int)kc, gemm::TN);
      count_launch();
    }
  }
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

// The backward counterpart: dL/dy (NHWC, [B][HW][C]) = dL/dx (NCHW-flatten order, [B][gx_ld]) permuted back and
// masked by the ReLU (y > 0) — nn.Flatten's and act_3's backward in one launch.
extern "C" int b2rl_unflatten_relu_mask(const float* gx_dev, int64_t gx_ld, const float* y_dev, int64_t B, int64_t HW,
                                        int64_t C, float* out_dev, void* stream) {
  B2RL_REQUIRE(gx_dev && y_dev && out_dev, "null argument");
  B2RL_REQUIRE(B >= 1 && HW >= 1 && C >= 1 && gx_ld >= C * HW, "bad shape");
  const size_t smem = (size_t)C * HW * sizeof(float);
  B2RL_REQUIRE(smem <= 48 * 1024, "activation row too large for the staging tile");
  gemm::k_unflatten_relu_mask<<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(gx_dev, gx_ld, y_dev, (int)HW, (int)C, out_dev);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}
