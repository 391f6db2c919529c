// The radiance-stage training loss and its gradient in two launches (train.py:75-162 of the reference):
//   loss      = mean((rgb   - px)^2)                                  (fine pass,   train.py:86)
//   loss_c    = mean((rgb_c - px)^2)                                  (coarse pass, train.py:107)
//   loss_bg   = gate * sum(mask |trans_rgb_bkgd - px|) / (sum(mask) + 1),  mask = trans > 0.5       (train.py:89-95)
//   loss_bg_smooth = gate * mean(0.5 dv^2 + 0.5 dh^2) over the env patch's vertical / horizontal differences (train.py:110-118)
//   total     = loss + loss_c + bg_weight loss_bg + bg_smooth_weight loss_bg_smooth
// Written as tensor expressions this is ~70 elementwise / reduction launches of a few microseconds each in forward and
// backward -- 0.3 ms of an 7 ms step on operands of 48 KB.  Here: one reduction kernel (per-block partial sums in a FIXED
// order, finished by the last block to arrive: deterministic, no float atomics) and one gradient kernel that takes the
// upstream gradient of `total` from device memory.
#include "common.cuh"

namespace rnerf {

constexpr int RL_THREADS = 256;
constexpr int RL_MAX_BLOCKS = 64;
constexpr int RL_NSUM = 6;     // sum (rgb-px)^2, sum (rgb_c-px)^2, sum mask |trb-px|, sum mask, sum dv^2, sum dh^2

struct RadianceLossArgs {
  const float* rgb;      // [B][3] fine-pass colour
  const float* rgb_c;    // [B][3] coarse-pass colour
  const float* trb;      // [B][3] trans * rgb_bkgd of the fine pass, or null (bg_weight = 0)
  const float* trans;    // [B]    fine-pass transmittance (with trb)
  const float* px;       // [B][3] target pixels
  const float* env;      // [P][P][C] environment patch, or null (bg_smooth_weight = 0)
  int64_t n_rays;
  int patch, env_c;      // C = 3 for a whole patch; the reference reshapes a device's [P/N][P][3] shard to (P/N, P/N, -1) (train.py:127-128)
  float bg_weight, bg_smooth_weight, gate;
};

__device__ __forceinline__ float block_sum(float v, float* scratch /* [8] */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < RL_THREADS / 32; ++w) r += scratch[w];
  return r;
}

// ws: [0] arrival counter (as uint32, zero on entry, zero again on exit), [1 + b * RL_NSUM + i] partial sum i of block b.
// out: [0] total, [1] loss, [2] loss_c, [3] loss_bg (gated, unweighted), [4] loss_bg_smooth (gated), [5] psnr, [6] psnr_c,
//      [7] sum(mask)
__global__ void __launch_bounds__(RL_THREADS) radiance_loss_fwd_kernel(const RadianceLossArgs a, float* __restrict__ ws,
                                                                        float* __restrict__ out) {
  __shared__ float scratch[8];
  __shared__ bool is_last;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  float s[RL_NSUM] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t r = tid; r < a.n_rays; r += nthr) {
    const float p0 = a.px[3 * r], p1 = a.px[3 * r + 1], p2 = a.px[3 * r + 2];
    float d0 = a.rgb[3 * r] - p0, d1 = a.rgb[3 * r + 1] - p1, d2 = a.rgb[3 * r + 2] - p2;
    s[0] += d0 * d0 + d1 * d1 + d2 * d2;
    d0 = a.rgb_c[3 * r] - p0; d1 = a.rgb_c[3 * r + 1] - p1; d2 = a.rgb_c[3 * r + 2] - p2;
    s[1] += d0 * d0 + d1 * d1 + d2 * d2;
    if (a.trb != nullptr && a.trans[r] > 0.5f) {
      s[2] += fabsf(a.trb[3 * r] - p0) + fabsf(a.trb[3 * r + 1] - p1) + fabsf(a.trb[3 * r + 2] - p2);
      s[3] += 1.f;
    }
  }
  if (a.env != nullptr) {
    const int P = a.patch, C = a.env_c;
    const int64_t n_el = (int64_t)P * P * C;
    for (int64_t e = tid; e < n_el; e += nthr) {
      const int i = (int)(e / ((int64_t)C * P)), j = (int)((e / C) % P);
      const float v = a.env[e];
      if (i + 1 < P) { const float d = a.env[e + (int64_t)C * P] - v; s[4] += d * d; }
      if (j + 1 < P) { const float d = a.env[e + C] - v; s[5] += d * d; }
    }
  }
#pragma unroll
  for (int i = 0; i < RL_NSUM; ++i) {
    const float t = block_sum(s[i], scratch);
    if (threadIdx.x == 0) ws[1 + blockIdx.x * RL_NSUM + i] = t;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u);
    is_last = prev == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    float t[RL_NSUM] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (unsigned b = 0; b < gridDim.x; ++b)
#pragma unroll
      for (int i = 0; i < RL_NSUM; ++i) t[i] += __ldcg(ws + 1 + b * RL_NSUM + i);
    const float inv_n = 1.f / (float)(3 * a.n_rays);
    const float loss = t[0] * inv_n, loss_c = t[1] * inv_n;
    const float loss_bg = a.trb != nullptr ? a.gate * t[2] / (t[3] + 1.f) : 0.f;
    float loss_sm = 0.f;
    if (a.env != nullptr) loss_sm = a.gate * 0.5f * (t[4] + t[5]) / (float)((int64_t)(a.patch - 1) * a.patch * a.env_c);
    out[0] = loss + loss_c + a.bg_weight * loss_bg + a.bg_smooth_weight * loss_sm;
    out[1] = loss; out[2] = loss_c; out[3] = loss_bg; out[4] = loss_sm;
    out[5] = -10.f * logf(loss) / 2.302585092994046f;      // utils.compute_psnr
    out[6] = -10.f * logf(loss_c) / 2.302585092994046f;
    out[7] = t[3];
    *reinterpret_cast<unsigned*>(ws) = 0u;                  // the workspace can be reused without a memset
  }
}

// d total / d (rgb, rgb_c, trans_rgb_bkgd, env), times the upstream gradient g[0] of `total`
__global__ void __launch_bounds__(RL_THREADS) radiance_loss_bwd_kernel(const RadianceLossArgs a, const float* __restrict__ out,
                                                                        const float* __restrict__ g, float* __restrict__ d_rgb,
                                                                        float* __restrict__ d_rgb_c, float* __restrict__ d_trb,
                                                                        float* __restrict__ d_env) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  const float up = g[0];
  const float k_mse = up * 2.f / (float)(3 * a.n_rays);
  const float k_bg = a.trb != nullptr ? up * a.bg_weight * a.gate / (out[7] + 1.f) : 0.f;
  for (int64_t e = tid; e < 3 * a.n_rays; e += nthr) {
    const float p = a.px[e];
    d_rgb[e] = k_mse * (a.rgb[e] - p);
    d_rgb_c[e] = k_mse * (a.rgb_c[e] - p);
    if (d_trb != nullptr) {
      const float d = a.trb[e] - p;
      const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      d_trb[e] = a.trans[e / 3] > 0.5f ? k_bg * sgn : 0.f;
    }
  }
  if (d_env != nullptr) {
    const int P = a.patch, C = a.env_c;
    const int64_t n_el = (int64_t)P * P * C, row = (int64_t)C * P;
    const float k = up * a.bg_smooth_weight * a.gate / (float)((int64_t)(P - 1) * P * C);   // d/dx of 0.5 d^2 = d
    for (int64_t e = tid; e < n_el; e += nthr) {
      const int i = (int)(e / row), j = (int)((e / C) % P);
      const float v = a.env[e];
      float acc = 0.f;
      if (i > 0) acc += v - a.env[e - row];
      if (i + 1 < P) acc -= a.env[e + row] - v;
      if (j > 0) acc += v - a.env[e - C];
      if (j + 1 < P) acc -= a.env[e + C] - v;
      d_env[e] = k * acc;
    }
  }
}

static int loss_grid(const RadianceLossArgs& a) {
  int64_t work = a.n_rays;
  if (a.env != nullptr) work = work > (int64_t)a.patch * a.patch * a.env_c ? work : (int64_t)a.patch * a.patch * a.env_c;
  int64_t b = (work + RL_THREADS - 1) / RL_THREADS;
  return (int)(b < 1 ? 1 : (b > RL_MAX_BLOCKS ? RL_MAX_BLOCKS : b));
}

}  // namespace rnerf

using namespace rnerf;

static int fill_args(RadianceLossArgs& a, const float* rgb, const float* rgb_c, const float* trb, const float* trans, const float* px,
                     int64_t n_rays, const float* env, int patch, int env_c, double bg_weight, double bg_smooth_weight, double gate) {
  RNERF_REQUIRE(n_rays > 0, RNERF_E_SHAPE, "rnerf_radiance_loss: n_rays must be positive");
  RNERF_REQUIRE_PTR(rgb); RNERF_REQUIRE_PTR(rgb_c); RNERF_REQUIRE_PTR(px);
  RNERF_REQUIRE(trb == nullptr || trans != nullptr, RNERF_E_NULL, "rnerf_radiance_loss: trans_rgb_bkgd given without trans");
  RNERF_REQUIRE(env == nullptr || (patch >= 2 && env_c >= 1), RNERF_E_SHAPE, "rnerf_radiance_loss: the env patch must be at least 2 x 2 x 1");
  a.rgb = rgb; a.rgb_c = rgb_c; a.trb = trb; a.trans = trans; a.px = px; a.env = env; a.n_rays = n_rays; a.patch = patch; a.env_c = env_c;
  a.bg_weight = (float)bg_weight; a.bg_smooth_weight = (float)bg_smooth_weight; a.gate = (float)gate;
  return 0;
}

extern "C" size_t rnerf_radiance_loss_ws_floats(void) { return 1 + (size_t)RL_MAX_BLOCKS * RL_NSUM; }

extern "C" int rnerf_radiance_loss_fwd(const float* rgb, const float* rgb_c, const float* trans_rgb_bkgd, const float* trans,
                                       const float* pixels, int64_t n_rays, const float* env, int patch, int env_channels, double bg_weight,
                                       double bg_smooth_weight, double gate, float* ws, float* out, void* stream) {
  RadianceLossArgs a;
  int rc = fill_args(a, rgb, rgb_c, trans_rgb_bkgd, trans, pixels, n_rays, env, patch, env_channels, bg_weight, bg_smooth_weight, gate);
  if (rc) return rc;
  RNERF_REQUIRE_PTR(ws); RNERF_REQUIRE_PTR(out);
  radiance_loss_fwd_kernel<<<loss_grid(a), RL_THREADS, 0, (cudaStream_t)stream>>>(a, ws, out);
  count_launch();
  return check_launch("rnerf_radiance_loss_fwd");
}

extern "C" int rnerf_radiance_loss_bwd(const float* rgb, const float* rgb_c, const float* trans_rgb_bkgd, const float* trans,
                                       const float* pixels, int64_t n_rays, const float* env, int patch, int env_channels, double bg_weight,
                                       double bg_smooth_weight, double gate, const float* out, const float* g_total, float* d_rgb,
                                       float* d_rgb_c, float* d_trans_rgb_bkgd, float* d_env, void* stream) {
  RadianceLossArgs a;
  int rc = fill_args(a, rgb, rgb_c, trans_rgb_bkgd, trans, pixels, n_rays, env, patch, env_channels, bg_weight, bg_smooth_weight, gate);
  if (rc) return rc;
  RNERF_REQUIRE_PTR(out); RNERF_REQUIRE_PTR(g_total); RNERF_REQUIRE_PTR(d_rgb); RNERF_REQUIRE_PTR(d_rgb_c);
  RNERF_REQUIRE((trans_rgb_bkgd == nullptr) == (d_trans_rgb_bkgd == nullptr) && (env == nullptr) == (d_env == nullptr), RNERF_E_NULL,
                "rnerf_radiance_loss_bwd: a gradient buffer must be given exactly for the inputs that are");
  radiance_loss_bwd_kernel<<<loss_grid(a), RL_THREADS, 0, (cudaStream_t)stream>>>(a, out, g_total, d_rgb, d_rgb_c, d_trans_rgb_bkgd, d_env);
  count_launch();
  return check_launch("rnerf_radiance_loss_bwd");
}
