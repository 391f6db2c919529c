// Activations (a11) + volumetric compositing (a12), forward and backward, and the bd_cut_dist mask (a15).
//
// rnerf/models.py:334-338     rgb = sigmoid(raw)*(1+2*pad) - pad ; sigma = softplus(raw_sigma + sigma_bias)
// rnerf/model_utils.py:247-309 volumetric_rendering:
//     delta_i = (t_{i+1}-t_i)*|dir_i| (last: 1e-3*|dir|); dd = sigma*delta (*mask); alpha = 1-exp(-dd)
//     T_i = exp(-sum_{j<i} dd_j); w = alpha*T; rgb = sum w c + T_N*bkgd; acc = sum w
//     dist = clip(nan_to_num(sum w t / acc), t_0, t_last)          (NaN -> 0)
//
// One warp per ray; the ray's samples are processed 32 at a time with a shuffle inclusive scan of dd
// and a running carry, so Ns is arbitrary (64 and 192 on the shipped configs).
#include <float.h>
#include "common.cuh"

namespace rnerf {

// The kernels were issue-bound (88 % issue-slot utilisation at 35 % of HBM peak, profiles/r1a_composite_*): the
// activations dominated the instruction count.  They now use the SFU forms (ex2.approx / rcp.approx / lg2.approx, each
// <= 2^-21 relative): their errors enter the outputs multiplied by weights that sum to <= 1, so the composited values
// move by < 1e-6 absolute.  alpha = 1 - exp(-sigma delta) keeps the accurate expf: there an error of the exponential is
// NOT scaled down (alpha ~ 0 in empty space) and would accumulate over the samples of a ray.
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// jax.nn.softplus(x) = logaddexp(x, 0) = max(x, 0) + log(1 + exp(-|x|))
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + __logf(1.f + __expf(-fabsf(x))); }
__device__ __forceinline__ float trans_exp(float x) { return __expf(x); }   // transmittance: relative error only

struct CompositeArgs {
  const float4* raw;   // [B][Ns] (r,g,b,sigma) raw
  const float* t;      // [B][Ns]
  const float* dirs;   // [B][Ns][3]
  const float* bkgd_raw;  // [B][3] or null
  const float* mask;   // [B][Ns] or null
  int64_t n_rays;
  int n_samples;
  int white_bkgd;
  float rgb_scale, rgb_pad, sigma_bias;
};

// per-sample quantities shared by fwd and bwd
struct SampleEval {
  float r, g, b, sigma, delta, dd, tval;
};

__device__ __forceinline__ SampleEval eval_sample(const CompositeArgs& a, int64_t base, int i, bool valid) {
  SampleEval s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (!valid) return s;
  float4 raw = __ldg(a.raw + base + i);
  s.r = sigmoidf_(raw.x) * a.rgb_scale - a.rgb_pad;
  s.g = sigmoidf_(raw.y) * a.rgb_scale - a.rgb_pad;
  s.b = sigmoidf_(raw.z) * a.rgb_scale - a.rgb_pad;
  s.sigma = softplusf_(raw.w + a.sigma_bias);
  s.tval = __ldg(a.t + base + i);
  float tn = (i + 1 < a.n_samples) ? __ldg(a.t + base + i + 1) : 0.f;
  float tdist = (i + 1 < a.n_samples) ? (tn - s.tval) : 1e-3f;
  const float* d = a.dirs + (base + i) * 3;
  float dn = sqrtf(sumsq3(__ldg(d), __ldg(d + 1), __ldg(d + 2)));
  s.delta = tdist * dn;
  s.dd = s.sigma * s.delta;
  if (a.mask) s.dd *= __ldg(a.mask + base + i);
  return s;
}

// MAXC > 0: the ray's samples (<= 32*MAXC) are evaluated up front so that all their loads are in flight together,
// then the dependent scan runs over registers; MAXC == 0 streams chunks of 32 (any Ns).
template <int MAXC>
__global__ void __launch_bounds__(128) composite_fwd_kernel(CompositeArgs a, float* __restrict__ comp_rgb,
                                                            float* __restrict__ distance, float* __restrict__ acc_out,
                                                            float* __restrict__ weights, float* __restrict__ alpha_out,
                                                            float* __restrict__ trans, float* __restrict__ trb) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.n_rays) return;
  const int64_t base = ray * a.n_samples;
  float carry = 0.f;  // sum of dd over previous chunks
  float sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f, swt = 0.f;
  auto chunk = [&](int i0, const SampleEval& s) {
    const int i = i0 + lane;
    const bool valid = i < a.n_samples;
    float incl = warp_incl_scan(s.dd, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.f;
    float T = trans_exp(-(carry + excl));
    float al = 1.f - expf(-s.dd);
    float w = al * T;
    if (valid) {
      if (weights) weights[base + i] = w;
      if (alpha_out) alpha_out[base + i] = al;
      sr += w * s.r; sg += w * s.g; sb += w * s.b; sw += w; swt += w * s.tval;
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  };
  if (MAXC > 0) {
    SampleEval ev[MAXC > 0 ? MAXC : 1];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) ev[c] = eval_sample(a, base, c * 32 + lane, c * 32 + lane < a.n_samples);
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c * 32 < a.n_samples) chunk(c * 32, ev[c]);
  } else {
    for (int i0 = 0; i0 < a.n_samples; i0 += 32) chunk(i0, eval_sample(a, base, i0 + lane, i0 + lane < a.n_samples));
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sw = warp_sum(sw); swt = warp_sum(swt);
  if (lane == 0) {
    const float Tend = trans_exp(-carry);
    float br = 1.f, bg = 1.f, bb = 1.f;  // rgb_bkgd=None -> ones (model_utils.py:301)
    if (a.bkgd_raw) {
      br = sigmoidf_(a.bkgd_raw[3 * ray]) * a.rgb_scale - a.rgb_pad;
      bg = sigmoidf_(a.bkgd_raw[3 * ray + 1]) * a.rgb_scale - a.rgb_pad;
      bb = sigmoidf_(a.bkgd_raw[3 * ray + 2]) * a.rgb_scale - a.rgb_pad;
      sr += Tend * br; sg += Tend * bg; sb += Tend * bb;
    }
    if (a.white_bkgd) { float e = 1.f - sw; sr += e; sg += e; sb += e; }
    comp_rgb[3 * ray] = sr; comp_rgb[3 * ray + 1] = sg; comp_rgb[3 * ray + 2] = sb;
    if (acc_out) acc_out[ray] = sw;
    if (distance) {
      float d = swt / sw;
      if (isnan(d)) d = 0.f;                           // jnp.nan_to_num(x, copy=inf): NaN -> 0
      else if (isinf(d)) d = d > 0.f ? FLT_MAX : -FLT_MAX;
      float t0 = a.t[base], t1 = a.t[base + a.n_samples - 1];
      distance[ray] = fminf(fmaxf(d, t0), t1);
    }
    if (trans) trans[ray] = Tend;
    if (trb) { trb[3 * ray] = Tend * br; trb[3 * ray + 1] = Tend * bg; trb[3 * ray + 2] = Tend * bb; }
  }
}

// Backward.  With C = sum_i w_i c_i + T_N b,  w_i = T_i - T_{i+1}:
//   dC/dc_k = w_k ;  dC/db = T_N ;  dC/d(dd_k) = c_k T_{k+1} - (sum_{i>k} w_i c_i + T_N b)
//   dT_N/d(dd_k) = -T_N ;  trans_rgb_bkgd = T_N * stopgrad(b)
// Suffix sums come from "total - inclusive prefix" with the totals accumulated in a first pass.
__global__ void __launch_bounds__(128) composite_bwd_kernel(CompositeArgs a, const float* __restrict__ d_rgb,
                                                            const float* __restrict__ d_trans,
                                                            const float* __restrict__ d_trb, float4* __restrict__ d_raw,
                                                            float* __restrict__ d_bkgd_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.n_rays) return;
  const int64_t base = ray * a.n_samples;
  float gr = 0.f, gg = 0.f, gb = 0.f;
  if (d_rgb) { gr = d_rgb[3 * ray]; gg = d_rgb[3 * ray + 1]; gb = d_rgb[3 * ray + 2]; }
  float br = 1.f, bg = 1.f, bb = 1.f, sbr = 0.f, sbg = 0.f, sbb = 0.f;
  if (a.bkgd_raw) {
    sbr = sigmoidf_(a.bkgd_raw[3 * ray]); sbg = sigmoidf_(a.bkgd_raw[3 * ray + 1]); sbb = sigmoidf_(a.bkgd_raw[3 * ray + 2]);
    br = sbr * a.rgb_scale - a.rgb_pad; bg = sbg * a.rgb_scale - a.rgb_pad; bb = sbb * a.rgb_scale - a.rgb_pad;
  }
  // pass 1: total of gC . (w c) and total dd
  float carry = 0.f, tot = 0.f, totw = 0.f;
  for (int i0 = 0; i0 < a.n_samples; i0 += 32) {
    const int i = i0 + lane;
    SampleEval s = eval_sample(a, base, i, i < a.n_samples);
    float incl = warp_incl_scan(s.dd, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.f;
    float T = trans_exp(-(carry + excl));
    float w = (1.f - expf(-s.dd)) * T;
    tot += w * (gr * s.r + gg * s.g + gb * s.b);
    totw += w;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  tot = warp_sum(tot);
  totw = warp_sum(totw);
  const float Tend = trans_exp(-carry);
  // gradient flowing into T_N: comp_rgb bkgd term, trans, trans_rgb_bkgd (bkgd stop-grad)
  float gT = 0.f;
  if (a.bkgd_raw) gT += gr * br + gg * bg + gb * bb;
  if (d_trans) gT += d_trans[ray];
  if (d_trb) gT += d_trb[3 * ray] * br + d_trb[3 * ray + 1] * bg + d_trb[3 * ray + 2] * bb;
  // white_bkgd adds (1 - acc) to every channel: d/d(dd_k) of -acc = -T_N  (acc = 1 - T_N)
  const float gsum = gr + gg + gb;
  if (a.white_bkgd) gT += gsum;
  // pass 2
  carry = 0.f;
  float run = 0.f;  // inclusive prefix of gC.(w c) over previous chunks
  for (int i0 = 0; i0 < a.n_samples; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < a.n_samples;
    SampleEval s = eval_sample(a, base, i, valid);
    float incl = warp_incl_scan(s.dd, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.f;
    float T = trans_exp(-(carry + excl));
    float Tn = trans_exp(-(carry + incl));
    float w = (1.f - expf(-s.dd)) * T;
    float gc = gr * s.r + gg * s.g + gb * s.b;
    float pre = warp_incl_scan(w * gc, lane);
    float suffix = tot - (run + pre);  // sum_{i>k} w_i (gC.c_i)
    float g_dd = gc * Tn - suffix - gT * Tend;
    if (valid) {
      float g_sigma = g_dd * s.delta;
      if (a.mask) g_sigma *= __ldg(a.mask + base + i);
      float4 raw = __ldg(a.raw + base + i);
      float ss = sigmoidf_(raw.w + a.sigma_bias);  // d softplus / dx
      float sr_ = sigmoidf_(raw.x), sg_ = sigmoidf_(raw.y), sb_ = sigmoidf_(raw.z);
      float4 o;
      o.x = gr * w * a.rgb_scale * sr_ * (1.f - sr_);
      o.y = gg * w * a.rgb_scale * sg_ * (1.f - sg_);
      o.z = gb * w * a.rgb_scale * sb_ * (1.f - sb_);
      o.w = g_sigma * ss;
      d_raw[base + i] = o;
    }
    run += __shfl_sync(0xffffffffu, pre, 31);
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0 && d_bkgd_raw) {
    if (a.bkgd_raw) {
      d_bkgd_raw[3 * ray] = gr * Tend * a.rgb_scale * sbr * (1.f - sbr);
      d_bkgd_raw[3 * ray + 1] = gg * Tend * a.rgb_scale * sbg * (1.f - sbg);
      d_bkgd_raw[3 * ray + 2] = gb * Tend * a.rgb_scale * sbb * (1.f - sbb);
    } else {
      d_bkgd_raw[3 * ray] = d_bkgd_raw[3 * ray + 1] = d_bkgd_raw[3 * ray + 2] = 0.f;
    }
  }
}

// rnerf/models.py:498-503: inside-bbox test, then mask = reverse-cumsum(inside) > 0 (everything up to the
// last in-box sample).  inv_mask = 1 - mask.
__global__ void __launch_bounds__(128) bbox_tail_mask_kernel(const float* __restrict__ pos, int64_t n_rays, int n_samples,
                                                             float lox, float loy, float loz, float hix, float hiy,
                                                             float hiz, float* __restrict__ mask,
                                                             float* __restrict__ inv_mask) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int64_t base = ray * n_samples;
  int last = -1;
  for (int i = lane; i < n_samples; i += 32) {
    const float* p = pos + (base + i) * 3;
    float x = p[0], y = p[1], z = p[2];
    bool in = (x >= lox) && (x <= hix) && (y >= loy) && (y <= hiy) && (z >= loz) && (z <= hiz);
    if (in) last = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  for (int i = lane; i < n_samples; i += 32) {
    float m = (i <= last) ? 1.f : 0.f;
    if (mask) mask[base + i] = m;
    if (inv_mask) inv_mask[base + i] = 1.f - m;
  }
}

static int fill_args(CompositeArgs& a, const float* raw, const float* t, const float* dirs, const float* bkgd_raw,
                     const float* mask, int64_t n_rays, int n_samples, int white_bkgd, double rgb_padding,
                     double sigma_bias, const char* who) {
  if (!raw || !t || !dirs) { set_error("%s: null input", who); return RNERF_E_NULL; }
  if (n_rays < 0 || n_samples < 1) { set_error("%s: bad sizes (%lld rays, %d samples)", who, (long long)n_rays, n_samples); return RNERF_E_SHAPE; }
  if (!aligned16(raw)) { set_error("%s: raw must be 16-byte aligned", who); return RNERF_E_ALIGN; }
  a.raw = (const float4*)raw; a.t = t; a.dirs = dirs; a.bkgd_raw = bkgd_raw; a.mask = mask;
  a.n_rays = n_rays; a.n_samples = n_samples; a.white_bkgd = white_bkgd;
  a.rgb_scale = (float)(1.0 + 2.0 * rgb_padding); a.rgb_pad = (float)rgb_padding; a.sigma_bias = (float)sigma_bias;
  return 0;
}

}  // namespace rnerf

using namespace rnerf;

extern "C" int rnerf_composite_fwd(const float* raw, const float* t, const float* dirs, const float* bkgd_raw,
                                   const float* mask, int64_t n_rays, int n_samples, int white_bkgd, double rgb_padding,
                                   double sigma_bias, float* comp_rgb, float* distance, float* acc, float* weights,
                                   float* alpha, float* trans, float* trans_rgb_bkgd, void* stream) {
  if (n_rays == 0) return 0;
  CompositeArgs a;
  int rc = fill_args(a, raw, t, dirs, bkgd_raw, mask, n_rays, n_samples, white_bkgd, rgb_padding, sigma_bias,
                     "rnerf_composite_fwd");
  if (rc) return rc;
  RNERF_REQUIRE_PTR(comp_rgb);
  const int wpb = 4;
  const unsigned grid = (unsigned)((n_rays + wpb - 1) / wpb);
  cudaStream_t st = (cudaStream_t)stream;
  // measured on B200: up-front evaluation helps the 64-sample pass slightly and hurts the 192-sample one (register
  // pressure lowers occupancy), so longer rays stream 32 samples at a time
  if (n_samples <= 64) composite_fwd_kernel<2><<<grid, wpb * 32, 0, st>>>(a, comp_rgb, distance, acc, weights, alpha, trans, trans_rgb_bkgd);
  else                 composite_fwd_kernel<0><<<grid, wpb * 32, 0, st>>>(a, comp_rgb, distance, acc, weights, alpha, trans, trans_rgb_bkgd);
  count_launch();
  return check_launch("rnerf_composite_fwd");
}

extern "C" int rnerf_composite_bwd(const float* raw, const float* t, const float* dirs, const float* bkgd_raw,
                                   const float* mask, int64_t n_rays, int n_samples, int white_bkgd, double rgb_padding,
                                   double sigma_bias, const float* d_comp_rgb, const float* d_trans,
                                   const float* d_trans_rgb_bkgd, float* d_raw, float* d_bkgd_raw, void* stream) {
  if (n_rays == 0) return 0;
  CompositeArgs a;
  int rc = fill_args(a, raw, t, dirs, bkgd_raw, mask, n_rays, n_samples, white_bkgd, rgb_padding, sigma_bias,
                     "rnerf_composite_bwd");
  if (rc) return rc;
  RNERF_REQUIRE_PTR(d_raw);
  RNERF_REQUIRE(aligned16(d_raw), RNERF_E_ALIGN, "rnerf_composite_bwd: d_raw must be 16-byte aligned");
  const int wpb = 4;
  composite_bwd_kernel<<<(unsigned)((n_rays + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      a, d_comp_rgb, d_trans, d_trans_rgb_bkgd, (float4*)d_raw, d_bkgd_raw);
  count_launch();
  return check_launch("rnerf_composite_bwd");
}

extern "C" int rnerf_bbox_tail_mask(const float* pos, int64_t n_rays, int n_samples, const double lo[3],
                                    const double hi[3], float* mask, float* inv_mask, void* stream) {
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(pos); RNERF_REQUIRE_PTR(lo); RNERF_REQUIRE_PTR(hi);
  RNERF_REQUIRE(n_rays > 0 && n_samples > 0, RNERF_E_SHAPE, "rnerf_bbox_tail_mask: bad sizes");
  const int wpb = 4;
  bbox_tail_mask_kernel<<<(unsigned)((n_rays + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      pos, n_rays, n_samples, (float)lo[0], (float)lo[1], (float)lo[2], (float)hi[0], (float)hi[1], (float)hi[2], mask,
      inv_mask);
  count_launch();
  return check_launch("rnerf_bbox_tail_mask");
}
