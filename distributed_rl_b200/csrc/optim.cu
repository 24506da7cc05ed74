// Fused optimizer step for the learners' RMSprop (cfg/ape_x.json:27-35 centered, cfg/impala.json:19-23
// plain): ONE pass over (param, grad, square_avg, grad_avg) of every tensor that
//   * applies torch.optim.RMSprop's update (baseline/utils.py getOptim :124-130),
//   * zeroes the gradient (optim.zero_grad),
//   * accumulates per-tensor sum(g^2) for the reference's "norm" = sqrt(sum_i ||g_i||_2)
//     (APE_X/Learner.py:123-138, a per-tensor .norm() kernel + sync each in the reference).
// Replaces ~14 foreach / elementwise launches per step.  SURVEY.md §8f rank 2.
#include "common.cuh"

namespace b2rl {

constexpr int OPT_MAX_TENSORS = 24;
constexpr int OPT_THREADS = 256;
constexpr int OPT_PER_THREAD = 4;
constexpr int64_t OPT_CHUNK = (int64_t)OPT_THREADS * OPT_PER_THREAD;

struct OptTable {
  float* p[OPT_MAX_TENSORS];
  float* g[OPT_MAX_TENSORS];
  float* sq[OPT_MAX_TENSORS];
  float* ga[OPT_MAX_TENSORS];
  int64_t numel[OPT_MAX_TENSORS];
  int32_t block_start[OPT_MAX_TENSORS + 1];   // first block of each tensor (a block never straddles tensors)
  int32_t n_tensors;
};

__global__ void __launch_bounds__(OPT_THREADS)
k_rmsprop(const __grid_constant__ OptTable T, float lr, float alpha, float one_m_alpha, float eps, int centered,
          double* __restrict__ sumsq /*[n_tensors] or nullptr*/) {
  __shared__ double s_part[OPT_THREADS / 32];
  int t = 0;
  while (t + 1 < T.n_tensors && (int)blockIdx.x >= T.block_start[t + 1]) ++t;
  const int64_t base = (int64_t)(blockIdx.x - T.block_start[t]) * OPT_CHUNK;
  const int64_t n = T.numel[t];
  float* __restrict__ P = T.p[t];
  float* __restrict__ G = T.g[t];
  float* __restrict__ SQ = T.sq[t];
  float* __restrict__ GA = T.ga[t];
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < OPT_PER_THREAD; ++u) {
    const int64_t i = base + (int64_t)u * OPT_THREADS + threadIdx.x;
    if (i < n) {
      const float g = G[i];
      const float sq = __fmaf_rn(one_m_alpha * g, g, SQ[i] * alpha);   // square_avg.mul_(alpha).addcmul_(g, g, 1-alpha)
      float avg;
      if (centered) {
        float ga = GA[i];
        ga = __fmaf_rn(one_m_alpha, g - ga, ga);                       // grad_avg.lerp_(g, 1-alpha)
        GA[i] = ga;
        avg = __fsqrt_rn(__fmaf_rn(-ga, ga, sq));                      // addcmul(ga, ga, -1).sqrt_()
      } else {
        avg = __fsqrt_rn(sq);
      }
      SQ[i] = sq;
      P[i] = P[i] - lr * (g / (avg + eps));                            // addcdiv_(g, avg + eps, -lr)
      G[i] = 0.0f;                                                     // zero_grad
      acc += (double)g * (double)g;
    }
  }
  if (sumsq) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < OPT_THREADS / 32; ++w) s += s_part[w];
      atomicAdd(sumsq + t, s);     // diagnostic value only: fp64 accumulation, order matters at the 1e-16 level
    }
  }
}

__global__ void k_grad_norm_finish(double* __restrict__ sumsq, int n, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < n; ++i) { s += sqrt(sumsq[i]); sumsq[i] = 0.0; }   // self-clean for the next step
  *out = (float)sqrt(s);                                      // "p_norm ** .5" of APE_X/Learner.py:130
}

}  // namespace b2rl

using namespace b2rl;

extern "C" int b2rl_rmsprop_step(float* const* params, float* const* grads, float* const* square_avg,
                                 float* const* grad_avg, const int64_t* numel, int32_t n_tensors, double lr,
                                 double alpha, double eps, int32_t centered, double* sumsq_scratch_dev,
                                 float* grad_norm_out_dev, void* stream) {
  B2RL_REQUIRE(n_tensors >= 1 && n_tensors <= OPT_MAX_TENSORS, "1..24 tensors");
  B2RL_REQUIRE(params && grads && square_avg && numel, "null argument");
  B2RL_REQUIRE(!centered || grad_avg, "centered RMSprop needs grad_avg");
  B2RL_REQUIRE(!grad_norm_out_dev || sumsq_scratch_dev,
               "the gradient norm needs a zero-initialised scratch of n_tensors doubles");
  OptTable T{};
  T.n_tensors = n_tensors;
  int64_t blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    B2RL_REQUIRE(numel[i] >= 1 && params[i] && grads[i] && square_avg[i] && (!centered || grad_avg[i]),
                 "bad tensor entry");
    T.p[i] = params[i]; T.g[i] = grads[i]; T.sq[i] = square_avg[i]; T.ga[i] = centered ? grad_avg[i] : nullptr;
    T.numel[i] = numel[i];
    T.block_start[i] = (int32_t)blocks;
    blocks += (numel[i] + OPT_CHUNK - 1) / OPT_CHUNK;
  }
  T.block_start[n_tensors] = (int32_t)blocks;
  B2RL_REQUIRE(blocks < (1LL << 31), "too many elements");
  cudaStream_t st = (cudaStream_t)stream;
  // python-double hyper-parameters, cast once like torch's foreach kernels do (1 - alpha formed in double)
  k_rmsprop<<<(unsigned)blocks, OPT_THREADS, 0, st>>>(T, (float)lr, (float)alpha, (float)(1.0 - alpha), (float)eps,
                                                      centered, sumsq_scratch_dev);
  count_launch();
  if (grad_norm_out_dev) {
    k_grad_norm_finish<<<1, 32, 0, st>>>(sumsq_scratch_dev, n_tensors, grad_norm_out_dev);
    count_launch();
  }
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}

extern "C" int b2rl_rmsprop_norm_finish(double* sumsq_scratch_dev, int32_t n_tensors, float* grad_norm_out_dev,
                                        void* stream) {
  B2RL_REQUIRE(sumsq_scratch_dev && grad_norm_out_dev, "null argument");
  B2RL_REQUIRE(n_tensors >= 1 && n_tensors <= OPT_MAX_TENSORS, "1..24 tensors");
  k_grad_norm_finish<<<1, 32, 0, (cudaStream_t)stream>>>(sumsq_scratch_dev, n_tensors, grad_norm_out_dev);
  count_launch();
  B2RL_CHECK_LAUNCH();
  return B2RL_OK;
}
