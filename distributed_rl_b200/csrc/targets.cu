// Target / TD-error / priority kernels for the three learners.  Each one fuses
// what the reference spreads over ATen elementwise kernels, two device->host
// syncs and NumPy (APE_X/Learner.py:90,108; R2D2/Learner.py:146,175;
// IMPALA/Learner.py:176-200) into a single launch, and emits dLoss/dQ so the
// network backward can start without touching the host.
//
// fp32 everywhere the reference is fp32, each op individually rounded
// (-fmad=false); fp64 where the reference drops to NumPy fp64 (R2D2 n-step sums).
#include "common.cuh"

#include <math.h>

namespace b2rl {

// Deterministic block sum of doubles (fixed tree order); result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* s_buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) s_buf[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < THREADS / 32; ++w) r += s_buf[w];
  }
  __syncthreads();
  return r;
}

// ----------------------------------------------------------------------------
// Ape-X  (APE_X/Learner.py:85-121)
// ----------------------------------------------------------------------------
constexpr int APEX_THREADS = 256;

__global__ void __launch_bounds__(APEX_THREADS)
k_apex_target(const float* __restrict__ q_s, const float* __restrict__ qn_on,
              const float* __restrict__ qn_tg, const int64_t* __restrict__ action,
              const float* __restrict__ reward, const float* __restrict__ notdone,
              const float* __restrict__ weight, int B, int A, float gamma_n, float alpha,
              float* __restrict__ target_out, float* __restrict__ td_out, float* __restrict__ prio_out,
              float* __restrict__ grad_q, float* __restrict__ scalars) {
  __shared__ double s_buf[APEX_THREADS / 32];
  double acc_loss = 0.0, acc_y = 0.0, acc_w = 0.0;
  for (int b = threadIdx.x; b < B; b += APEX_THREADS) {
    const float* qo = qn_on + (int64_t)b * A;
    int a_star = 0;
    float best = qo[0];
    for (int a = 1; a < A; ++a) {          // first maximum, like torch.argmax
      const float v = qo[a];
      if (v > best) { best = v; a_star = a; }
    }
    const float nxt = __fmul_rn(qn_tg[(int64_t)b * A + a_star], notdone[b]);        // :93-95
    const float y = __fadd_rn(reward[b], __fmul_rn(gamma_n, nxt));                    // :103
    int a_b = (int)action[b];
    a_b = a_b < 0 ? 0 : (a_b >= A ? A - 1 : a_b);
    const float td_raw = __fsub_rn(y, q_s[(int64_t)b * A + a_b]);                     // :105
    const float td = fminf(fmaxf(td_raw, -1.0f), 1.0f);                               // :106
    const float w = weight[b];
    if (target_out) target_out[b] = y;
    if (td_out) td_out[b] = td;
    if (prio_out) prio_out[b] = powcr(__fadd_rn(fabsf(td), 1e-7f), alpha);            // :108-110
    if (grad_q) {
      // d/dq_sa [0.5 * mean(w * clamp(y - q_sa)^2)] = -w * td / B inside [-1, 1]
      const bool inside = (td_raw >= -1.0f) && (td_raw <= 1.0f);
      const float g = inside ? __fdiv_rn(-__fmul_rn(w, td), (float)B) : 0.0f;
      for (int a = 0; a < A; ++a) grad_q[(int64_t)b * A + a] = (a == a_b) ? g : 0.0f;
    }
    acc_loss += (double)__fmul_rn(w, __fmul_rn(td, td));
    acc_y += (double)y;
    acc_w += (double)w;
  }
  const double sl = block_sum<APEX_THREADS>(acc_loss, s_buf);
  const double sy = block_sum<APEX_THREADS>(acc_y, s_buf);
  const double sw = block_sum<APEX_THREADS>(acc_w, s_buf);
  if (threadIdx.x == 0 && scalars) {
    scalars[0] = __fmul_rn((float)(sl / (double)B), 0.5f);   // :112-114
    scalars[1] = (float)(sy / (double)B);                    // info['mean_value'] :119
    scalars[2] = (float)(sw / (double)B);                    // mean weight :120
  }
}

// ----------------------------------------------------------------------------
// R2D2  (R2D2/Learner.py:22-35, 110-198)
// ----------------------------------------------------------------------------
__device__ __forceinline__ float sgnf(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

__device__ __forceinline__ float value_transform(float x) {           // h(x), :22-27
  const float s = __fsub_rn(__fsqrt_rn(__fadd_rn(fabsf(x), 1.0f)), 1.0f);
  return __fadd_rn(__fmul_rn(sgnf(x), s), __fmul_rn(1e-3f, x));
}
__device__ __forceinline__ float value_inv_transform(float x) {       // h^-1(x), :30-35
  const float inner = __fadd_rn(__fadd_rn(fabsf(x), 1.0f), 1e-3f);
  const float s = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn((float)(4 * 1e-3), inner)));
  const float t = __fdiv_rn(__fsub_rn(s, 1.0f), (float)(2 * 1e-3));
  return __fmul_rn(sgnf(x), __fsub_rn(__fmul_rn(t, t), 1.0f));
}

constexpr int R2D2_THREADS = 128;
constexpr int R2D2_MAX_NSTEP = 32;
struct GammaPow { float v[R2D2_MAX_NSTEP]; };  // fp32(gamma^i), passed by value (graph-capturable)

// One CTA per sequence b; thread t handles time step t of the L-1 trained steps.
__global__ void __launch_bounds__(R2D2_THREADS)
k_r2d2_target(const float* __restrict__ q, const float* __restrict__ qt,
              const int64_t* __restrict__ action, const float* __restrict__ reward,
              const float* __restrict__ notdone, const float* __restrict__ weight, int L, int B, int A,
              int n, double gamma, float alpha, int rescale, const GammaPow gp,
              float* __restrict__ target_out, float* __restrict__ td_out, float* __restrict__ prio_out,
              float* __restrict__ grad_q, double* __restrict__ partial /*[B][2]*/) {
  extern __shared__ float s_nmv[];  // [L] Qbar(s_t, argmax_a Q(s_t, a))
  __shared__ double s_buf[R2D2_THREADS / 32];
  __shared__ float s_max[R2D2_THREADS / 32];
  const int b = blockIdx.x;
  const int Lm1 = L - 1;
  for (int t = threadIdx.x; t < L; t += R2D2_THREADS) {
    const float* qo = q + ((int64_t)t * B + b) * A;
    int am = 0; float best = qo[0];
    for (int a = 1; a < A; ++a) { const float v = qo[a]; if (v > best) { best = v; am = a; } }   // :134
    s_nmv[t] = qt[((int64_t)t * B + b) * A + am];                                                // :137
  }
  __syncthreads();
  const float nd = notdone[b];
  const float w = weight[b];
  const float denom = (float)((int64_t)B * Lm1);
  double acc_loss = 0.0, acc_abs = 0.0, acc_sel = 0.0;
  float mx = 0.0f;
  for (int t = threadIdx.x; t < Lm1; t += R2D2_THREADS) {
    float y;
    if (t < Lm1 - n) {
      // rewards[t] = sum_i fp32(fp32(gamma^i) * r[t+i]) accumulated in fp64, cast to fp32  (:145-154)
      double acc = 0.0;
      for (int i = 0; i < n; ++i) acc += (double)__fmul_rn(gp.v[i], reward[(int64_t)(t + i) * B + b]);
      float tv = s_nmv[t + n];                                                                    // :142
      if (rescale) tv = value_inv_transform(tv);                                                  // :143-144
      y = __fadd_rn((float)acc, __fmul_rn(gp.v[n], tv));                                          // :161
    } else {
      // tail rows come from the `remainder` recursion (:146-157): row t = L-1-n+k holds rem[n-k],
      // rem[0] = bootstrap*notdone (fp64), rem[j+1] = reward[L-3-j] + gamma*rem[j]
      const int k = t - (Lm1 - n);
      const int m = n - k;  // 1..n
      double rem = (double)s_nmv[L - 1] * (double)nd;
      for (int j = 0; j < m; ++j) rem = (double)reward[(int64_t)(Lm1 - 2 - j) * B + b] + gamma * rem;
      y = (float)rem;
    }
    if (rescale) y = value_transform(y);                                                          // :165-166
    int a_b = (int)action[(int64_t)t * B + b];
    a_b = a_b < 0 ? 0 : (a_b >= A ? A - 1 : a_b);
    const int64_t qoff = ((int64_t)t * B + b) * A;
    const float sel = q[qoff + a_b];                                                              // :123
    const float td = __fsub_rn(y, sel);                                                           // :172
    if (target_out) target_out[(int64_t)t * B + b] = y;
    if (td_out) td_out[(int64_t)t * B + b] = td;
    if (grad_q) {
      const float g = __fdiv_rn(-__fmul_rn(w, td), denom);
      for (int a = 0; a < A; ++a) grad_q[qoff + a] = (a == a_b) ? g : 0.0f;
    }
    const float atd = fabsf(td);
    mx = fmaxf(mx, atd);
    acc_abs += (double)atd;
    acc_loss += (double)__fmul_rn(w, __fmul_rn(td, td));
    acc_sel += (double)sel;
  }
  if (grad_q) {  // last window row receives no gradient
    for (int a = threadIdx.x; a < A; a += R2D2_THREADS) grad_q[((int64_t)Lm1 * B + b) * A + a] = 0.0f;
  }
  // block max
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  const double sum_abs = block_sum<R2D2_THREADS>(acc_abs, s_buf);   // contains a __syncthreads
  const double sum_loss = block_sum<R2D2_THREADS>(acc_loss, s_buf);
  const double sum_sel = block_sum<R2D2_THREADS>(acc_sel, s_buf);
  if (threadIdx.x == 0) {
    float m = s_max[0];
    for (int wq = 1; wq < R2D2_THREADS / 32; ++wq) m = fmaxf(m, s_max[wq]);
    const float mean = (float)(sum_abs / (double)Lm1);
    const float mixed = __fadd_rn(__fmul_rn(m, 0.9f), __fmul_rn(0.1f, mean));                    // :180
    if (prio_out) prio_out[b] = powcr(mixed, alpha);                                              // :181
    partial[2 * b + 0] = sum_loss;
    partial[2 * b + 1] = sum_sel;
  }
}

__global__ void k_r2d2_finish(const double* __restrict__ partial, int B, int Lm1, float* __restrict__ scalars) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sl = 0.0, ss = 0.0;
  for (int b = 0; b < B; ++b) { sl += partial[2 * b]; ss += partial[2 * b + 1]; }
  const double cnt = (double)B * (double)Lm1;
  scalars[0] = __fmul_rn((float)(sl / cnt), 0.5f);   // :189-191
  scalars[1] = (float)(ss / cnt);                    // info['mean_value'] :195
}

// ----------------------------------------------------------------------------
// IMPALA V-trace  (IMPALA/Learner.py:141-215).  Thread per rollout, time-major
// (T, B) so a warp's loads at step t are one coalesced 128-byte line.
// ----------------------------------------------------------------------------
constexpr int VTRACE_THREADS = 128;

__global__ void __launch_bounds__(VTRACE_THREADS)
k_vtrace(const float* __restrict__ pi_a, const float* __restrict__ mu_a, const float* __restrict__ value,
         const float* __restrict__ boot, const float* __restrict__ reward, int T, int B, float gamma,
         float c_lambda, float c_bar, float p_bar, float* __restrict__ vtarget, float* __restrict__ adv) {
  const int b = blockIdx.x * VTRACE_THREADS + threadIdx.x;
  if (b >= B) return;
  const float bs = boot[b];
  float vmt_next = 0.0f;      // value_minus_target[i+1]
  float vt_next = bs;         // Vtarget[i+1]; for i = T-1 it is the bootstrap (:203-206)
  float v_next = 0.0f;        // V[i+1]
  for (int i = T - 1; i >= 0; --i) {
    const int64_t o = (int64_t)i * B + b;
    const float r = reward[o], v = value[o];
    const float ratio = expf(__fsub_rn(logf(pi_a[o]), logf(mu_a[o])));                  // :151-174
    float vmt;
    if (i == T - 1) {
      vmt = __fsub_rn(__fadd_rn(r, __fmul_rn(gamma, bs)), v);                           // :178-182
    } else {
      const float td = __fsub_rn(__fadd_rn(r, __fmul_rn(gamma, v_next)), v);            // :187-191
      const float cr = fminf(c_bar, ratio);                                             // :192
      const float cs = __fmul_rn(c_lambda, cr);                                         // :193
      vmt = __fadd_rn(__fmul_rn(td, cr), __fmul_rn(__fmul_rn(gamma, cs), vmt_next));    // :194-197
    }
    const float vt = __fadd_rn(v, vmt);                                                 // :202
    const float at = __fadd_rn(r, __fmul_rn(gamma, vt_next));                           // :207
    const float pt = fminf(p_bar, ratio);                                               // :209
    if (vtarget) vtarget[o] = vt;
    if (adv) adv[o] = __fmul_rn(__fsub_rn(at, v), pt);                                  // :212
    vmt_next = vmt; vt_next = vt; v_next = v;
  }
}

}  // namespace b2rl

using namespace b2rl;

extern "C" int b2rl_apex_target(const float* q_s, const float* qn_on, const float* qn_tg,
                                const int64_t* action, const float* reward, const float* notdone,
                                const float* weight, int32_t B, int32_t A, float gamma_n, float alpha,
                                float* target_out, float* td_out, float* prio_out, float* grad_q_out,
                                float* scalars_out, void* stream) {
  B2RL_REQUIRE(B >= 1 && A >= 1, "B and A must be positive");
  B2RL_REQUIRE(q_s && qn_on && qn_tg && action && reward && notdone && weight, "null input");
  k_apex_target<<<1, APEX_THREADS, 0, (cudaStream_t)stream>>>(q_s, qn_on, qn_tg, action, reward, notdone,
                                                             weight, B, A, gamma_n, alpha, target_out,
                                                             td_out, prio_out, grad_q_out, scalars_out);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

// per-device scratch for the R2D2 partial sums and gamma powers (grown on demand)
struct R2d2Scratch { double* partial = nullptr; int cap = 0; };
static R2d2Scratch g_r2d2[64];

extern "C" int b2rl_r2d2_target(const float* q, const float* qt, const int64_t* action,
                                const float* reward, const float* notdone, const float* weight,
                                int32_t L, int32_t B, int32_t A, int32_t n_step, double gamma, float alpha,
                                int32_t use_rescaling, float* target_out, float* td_out, float* prio_out,
                                float* grad_q_out, float* scalars_out, void* stream) {
  B2RL_REQUIRE(L >= 2 && B >= 1 && A >= 1, "bad shape");
  B2RL_REQUIRE(n_step >= 1 && n_step < R2D2_MAX_NSTEP && n_step <= L - 2, "need 1 <= n_step <= L-2");
  B2RL_REQUIRE(q && qt && action && reward && notdone && weight, "null input");
  int dev = 0;
  B2RL_CUDA(cudaGetDevice(&dev));
  R2d2Scratch& S = g_r2d2[dev & 63];
  cudaStream_t st = (cudaStream_t)stream;
  if (S.cap < B) {
    if (S.partial) B2RL_CUDA(cudaFree(S.partial));
    B2RL_CUDA(cudaMalloc(&S.partial, sizeof(double) * 2 * (size_t)B));
    S.cap = B;
  }
  GammaPow gp{};
  for (int i = 0; i <= n_step; ++i) gp.v[i] = (float)pow(gamma, (double)i);  // python: GAMMA ** i, cast to fp32
  k_r2d2_target<<<B, R2D2_THREADS, sizeof(float) * L, st>>>(q, qt, action, reward, notdone, weight, L, B, A,
                                                            n_step, gamma, alpha, use_rescaling, gp,
                                                            target_out, td_out, prio_out, grad_q_out,
                                                            S.partial);
  count_launch();
  if (scalars_out) {
    k_r2d2_finish<<<1, 32, 0, st>>>(S.partial, B, L - 1, scalars_out);
    count_launch();
  }
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_vtrace(const float* pi_a, const float* mu_a, const float* value, const float* boot,
                           const float* reward, int32_t T, int32_t B, float gamma, float c_lambda,
                           float c_bar, float p_bar, float* vtarget_out, float* advantage_out,
                           void* stream) {
  B2RL_REQUIRE(T >= 1 && B >= 1, "bad shape");
  B2RL_REQUIRE(pi_a && mu_a && value && boot && reward, "null input");
  k_vtrace<<<(B + VTRACE_THREADS - 1) / VTRACE_THREADS, VTRACE_THREADS, 0, (cudaStream_t)stream>>>(
      pi_a, mu_a, value, boot, reward, T, B, gamma, c_lambda, c_bar, p_bar, vtarget_out, advantage_out);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}
