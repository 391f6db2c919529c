// Optimiser step of the training path (a17) on a flat parameter arena.
//
// train.py:166-183: grads = pmean(grads); clip by value (grad_max_val); clip by global norm (grad_max_norm);
// state.apply_gradients -> optax.adam(lr) (train.py:312-317, optax defaults b1 = 0.9, b2 = 0.999, eps = 1e-8):
//     mu = b1 mu + (1-b1) g ; nu = b2 nu + (1-b2) g^2 ; theta -= lr * (mu / (1-b1^t)) / (sqrt(nu / (1-b2^t)) + eps)
// The weight-decay term of the loss, weight_decay_mult * mean(theta^2) over ALL leaves (train.py:146-150), has the
// closed-form gradient (2 weight_decay_mult / numel) * theta, added here instead of being differentiated.
//
// All per-step scalars live in a small device array so that a captured CUDA graph of the whole training step can be
// replayed with new values:  hyper[0] lr, [1] b1, [2] b2, [3] eps, [4] 1-b1^t, [5] 1-b2^t, [6] gradient scale,
// [7] weight-decay coefficient, [8] grad_max_val (0 = off), [9] grad_max_norm (0 = off).
// Memory-bound: 16 B read + 12 B written per parameter, one pass.
#include "common.cuh"

namespace rnerf {

__device__ __forceinline__ float final_grad(float g, float th, float gscale, float wd, float clampv) {
  g = fmaf(wd, th, g * gscale);
  if (clampv > 0.f) g = fminf(fmaxf(g, -clampv), clampv);
  return g;
}

// out[0] += sum over i of final_grad(i)^2   (the squared global norm the norm clip needs)
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, const float* __restrict__ theta,
                                                         int64_t n, const float* __restrict__ hyper,
                                                         float* __restrict__ out) {
  const float gscale = hyper[6], wd = hyper[7], clampv = hyper[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = final_grad(g[i], theta ? theta[i] : 0.f, gscale, wd, clampv);
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// plain sum of squares (weight_l2 statistic): out[0] += sum x^2
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ theta, const float* __restrict__ g,
                                                   float* __restrict__ mu, float* __restrict__ nu, int64_t n,
                                                   const float* __restrict__ hyper, const float* __restrict__ norm_sq) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], bc1 = hyper[4], bc2 = hyper[5];
  const float gscale = hyper[6], wd = hyper[7], clampv = hyper[8], max_norm = hyper[9];
  float mult = 1.f;
  if (max_norm > 0.f && norm_sq != nullptr) mult = fminf(1.f, max_norm / (1e-7f + sqrtf(*norm_sq)));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float th = theta[i];
    const float gi = final_grad(g[i], th, gscale, wd, clampv) * mult;
    const float m = b1 * mu[i] + (1.f - b1) * gi;
    const float v = b2 * nu[i] + (1.f - b2) * gi * gi;
    mu[i] = m; nu[i] = v;
    theta[i] = th - lr * ((m / bc1) / (sqrtf(v / bc2) + eps));
  }
}

static unsigned grid_for(int64_t n) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)n_sm * 8;
  return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace rnerf

using namespace rnerf;

extern "C" int rnerf_sumsq(const float* x, int64_t n, float* out_accum, void* stream) {
  if (n <= 0) return 0;
  RNERF_REQUIRE_PTR(x); RNERF_REQUIRE_PTR(out_accum);
  sumsq_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, n, out_accum);
  count_launch();
  return check_launch("rnerf_sumsq");
}

extern "C" int rnerf_grad_sumsq(const float* grad, const float* theta, int64_t n, const float* hyper, float* out_accum,
                                void* stream) {
  if (n <= 0) return 0;
  RNERF_REQUIRE_PTR(grad); RNERF_REQUIRE_PTR(hyper); RNERF_REQUIRE_PTR(out_accum);
  grad_sumsq_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(grad, theta, n, hyper, out_accum);
  count_launch();
  return check_launch("rnerf_grad_sumsq");
}

extern "C" int rnerf_adam_step(float* theta, const float* grad, float* mu, float* nu, int64_t n, const float* hyper,
                               const float* norm_sq, void* stream) {
  if (n <= 0) return 0;
  RNERF_REQUIRE_PTR(theta); RNERF_REQUIRE_PTR(grad); RNERF_REQUIRE_PTR(mu); RNERF_REQUIRE_PTR(nu); RNERF_REQUIRE_PTR(hyper);
  adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(theta, grad, mu, nu, n, hyper, norm_sq);
  count_launch();
  return check_launch("rnerf_adam_step");
}
