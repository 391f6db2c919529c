// Hierarchical resampling along the marched path (a13 + a14).  One warp per ray.
//
// sorted_piecewise_constant_pdf (rnerf/model_utils.py:312-374):
//     bins = mid-points of the coarse t (Nc-1), weights = w[1:-1] (Nc-2), padded to sum >= 1e-5;
//     cdf = [0, min(1, cumsum(pdf[:-1])), 1];  for each u: k = last i with u >= cdf[i];
//     z = bins[k] + clip(nan->0((u-cdf[k])/(cdf[k+1]-cdf[k])), 0, 1) * (bins[k+1]-bins[k])
// sample_pdf (rnerf/model_utils.py:377-435):
//     z_all = sort(concat(t_coarse, z));  idx = max(#{ray_dist < z} - 1, 0);
//     pos = ray_pos[idx] + ray_dir[idx]*(z - ray_dist[idx]);  dir = ray_dir[idx];  grad = idx_grad[idx]
// The reference runs the second part as a sequential fori_loop over rays; here every ray is a warp.
#include "common.cuh"

namespace rnerf {

constexpr int RS_MAX_COARSE = 128;
constexpr int RS_MAX_FINE = 256;
constexpr int RS_WARPS = 4;
constexpr int RS_L1_STRIDE = 32;
constexpr int RS_MAX_L1 = 64;            // first-level table covers n_steps <= 2048; longer paths use the plain search
constexpr int RS_MAX_TROW = 4096;        // longest path whose t column is staged in shared memory (dynamic smem)

struct ResampleSmem {
  float tc[RS_MAX_COARSE];
  float bins[RS_MAX_COARSE];
  float cdf[RS_MAX_COARSE];
  float z[RS_MAX_FINE];
  float merged[RS_MAX_COARSE + RS_MAX_FINE];
  float tk[RS_MAX_L1];   // ray_dist at every RS_L1_STRIDE-th march step (first level of the step search)
};

// `t_col` (optional): the dense ray_dist column [B][S] written by the march.  When present (and S <= RS_MAX_TROW) the
// ray's column is staged in shared memory with coalesced loads and the step search never touches global memory;
// otherwise the search probes the strided t field of the path records.
__global__ void __launch_bounds__(RS_WARPS * 32) resample_kernel(const float4* __restrict__ path, int recf4,
                                                                 const float* __restrict__ t_col, int64_t n_rays,
                                                                 int n_steps, const float* __restrict__ t_c,
                                                                 const float* __restrict__ w_c, int nc,
                                                                 const float* __restrict__ u, int u_per_ray, int nf,
                                                                 float* __restrict__ t_f, float* __restrict__ pos_f,
                                                                 float* __restrict__ dir_f, float* __restrict__ grad_f) {
  __shared__ ResampleSmem sm[RS_WARPS];
  extern __shared__ __align__(16) float trow_all[];      // [RS_WARPS][n_steps] when t_col is staged
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ray = blockIdx.x * (int64_t)RS_WARPS + warp;
  if (ray >= n_rays) return;
  ResampleSmem& s = sm[warp];
  const int nb = nc - 1;  // number of bin edges
  const int nw = nc - 2;  // number of pdf bins
  const float* tc = t_c + ray * nc;
  const float* wc = w_c + ray * nc;

  for (int i = lane; i < nc; i += 32) s.tc[i] = __ldg(tc + i);
  __syncwarp();
  for (int i = lane; i < nb; i += 32) s.bins[i] = 0.5f * (s.tc[i + 1] + s.tc[i]);
  // weight sum (w[1:-1]) and padding
  float ws = 0.f;
  for (int i = lane; i < nw; i += 32) ws += __ldg(wc + 1 + i);
  ws = warp_sum(ws);
  const float padding = fmaxf(0.f, 1e-5f - ws);
  const float wsum = ws + padding;
  const float padw = padding / (float)nw;
  // pdf -> cdf.  The cumsum is done sequentially by one lane (61 adds) so that it is monotone and has the
  // same association as the reference's cumsum.
  for (int i = lane; i < nw; i += 32) s.cdf[i + 1] = (__ldg(wc + 1 + i) + padw) / wsum;  // pdf[i] parked at cdf[i+1]
  __syncwarp();
  if (lane == 0) {
    float c = 0.f;
    s.cdf[0] = 0.f;
    for (int i = 0; i < nw - 1; ++i) {
      c += s.cdf[i + 1];
      s.cdf[i + 1] = fminf(1.f, c);
    }
    s.cdf[nb - 1] = 1.f;
  }
  __syncwarp();
  // inverse-CDF samples
  const float* uu = u + (u_per_ray ? ray * nf : 0);
  for (int j = lane; j < nf; j += 32) {
    const float uj = __ldg(uu + j);
    // k = (number of i with cdf[i] <= u) - 1 ; cdf[0] = 0 <= u always
    int lo = 0, hi = nb;  // invariant: cdf[lo] <= u ; (hi == nb or cdf[hi] > u)
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (s.cdf[mid] <= uj) lo = mid; else hi = mid;
    }
    const int k = lo, k1 = min(lo + 1, nb - 1);
    const float c0 = s.cdf[k], c1 = s.cdf[k1], b0 = s.bins[k], b1 = s.bins[k1];
    float tt = (uj - c0) / (c1 - c0);
    if (isnan(tt)) tt = 0.f;
    tt = fminf(fmaxf(tt, 0.f), 1.f);
    s.z[j] = b0 + tt * (b1 - b0);
  }
  __syncwarp();
  // merge the two sorted lists (coarse t, fine z) by rank
  const int nt = nc + nf;
  for (int i = lane; i < nc; i += 32) {
    const float v = s.tc[i];
    int lo = 0, hi = nf;  // count of z < v
    while (lo < hi) { int mid = (lo + hi) >> 1; if (s.z[mid] < v) lo = mid + 1; else hi = mid; }
    s.merged[i + lo] = v;
  }
  for (int j = lane; j < nf; j += 32) {
    const float v = s.z[j];
    int lo = 0, hi = nc;  // count of tc <= v  (ties: coarse first)
    while (lo < hi) { int mid = (lo + hi) >> 1; if (s.tc[mid] <= v) lo = mid + 1; else hi = mid; }
    s.merged[j + lo] = v;
  }
  __syncwarp();
  // nearest march step strictly below, then linear extrapolation.  ray_dist is increasing along the path, so the
  // count of steps with ray_dist < z is found in two levels: a 32-stride table staged in shared memory (one strided
  // load per lane), then 5 probes inside the 32-step segment instead of 10-11 over the whole path.
  const float4* rec0 = path + ray * (int64_t)n_steps * recf4;
  const int n1 = (n_steps + RS_L1_STRIDE - 1) / RS_L1_STRIDE;
  const bool staged = t_col != nullptr;          // host passes t_col only when the dynamic smem was sized for it
  const bool two_level = !staged && n1 <= RS_MAX_L1;
  float* trow = trow_all + warp * n_steps;
  if (staged) {
    const float* tr = t_col + ray * (int64_t)n_steps;
    if ((n_steps & 3) == 0 && (reinterpret_cast<uintptr_t>(t_col) & 15u) == 0) {
      for (int i = lane; i < (n_steps >> 2); i += 32)
        reinterpret_cast<float4*>(trow)[i] = __ldcs(reinterpret_cast<const float4*>(tr) + i);
    } else {
      for (int i = lane; i < n_steps; i += 32) trow[i] = __ldcs(tr + i);
    }
    __syncwarp();
  } else if (two_level) {
    for (int i = lane; i < n1; i += 32) s.tk[i] = __ldg(&rec0[(i * RS_L1_STRIDE) * recf4].w);
    __syncwarp();
  }
  for (int m = lane; m < nt; m += 32) {
    const float zv = s.merged[m];
    int lo = 0, hi = n_steps;  // count of ray_dist < z
    if (staged) {
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (trow[mid] < zv) lo = mid + 1; else hi = mid;
      }
    } else if (two_level) {
      int a = 0, b = n1;       // number of table entries < z
      while (a < b) { int mid = (a + b) >> 1; if (s.tk[mid] < zv) a = mid + 1; else b = mid; }
      // entries [0, a) are < z, entry a (if any) is >= z: the count lies in ((a-1)*32, a*32]
      lo = a == 0 ? 0 : (a - 1) * RS_L1_STRIDE + 1;
      hi = a == 0 ? 0 : min(a * RS_L1_STRIDE, n_steps);
    }
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(&rec0[mid * recf4].w) < zv) lo = mid + 1; else hi = mid;
    }
    const int idx = max(lo - 1, 0);
    const float4 a = __ldg(rec0 + idx * recf4);
    const float3 b = path_dir(__ldg(rec0 + idx * recf4 + 1));
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grad_f) c = __ldg(rec0 + idx * recf4 + 2);
    const float dz = sub(zv, a.w);
    const int64_t o = ray * nt + m;
    t_f[o] = zv;
    pos_f[3 * o] = add(a.x, mul(b.x, dz)); pos_f[3 * o + 1] = add(a.y, mul(b.y, dz)); pos_f[3 * o + 2] = add(a.z, mul(b.z, dz));
    dir_f[3 * o] = b.x; dir_f[3 * o + 1] = b.y; dir_f[3 * o + 2] = b.z;
    if (grad_f) { grad_f[3 * o] = c.x; grad_f[3 * o + 1] = c.y; grad_f[3 * o + 2] = c.z; }
  }
}

}  // namespace rnerf

using namespace rnerf;

extern "C" int rnerf_resample(const float* path, int rec_floats, const float* t_col, int64_t n_rays, int n_steps,
                              const float* t_c, const float* weights_c, int n_coarse, const float* u, int u_per_ray,
                              int n_fine, float* t_f, float* pos_f, float* dir_f, float* grad_f, void* stream) {
  RNERF_REQUIRE(n_rays >= 0 && n_steps > 0, RNERF_E_SHAPE, "rnerf_resample: bad sizes");
  RNERF_REQUIRE(rec_floats == 8 || rec_floats == 12, RNERF_E_SHAPE, "rnerf_resample: rec_floats must be 8 or 12");
  RNERF_REQUIRE(grad_f == nullptr || rec_floats == 12, RNERF_E_SHAPE, "rnerf_resample: grad_f needs the full (12-float) path records");
  RNERF_REQUIRE(n_coarse >= 3 && n_coarse <= RS_MAX_COARSE, RNERF_E_SHAPE, "rnerf_resample: n_coarse=%d unsupported (3..%d)",
                n_coarse, RS_MAX_COARSE);
  RNERF_REQUIRE(n_fine >= 1 && n_fine <= RS_MAX_FINE, RNERF_E_SHAPE, "rnerf_resample: n_fine=%d unsupported (1..%d)", n_fine,
                RS_MAX_FINE);
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(path); RNERF_REQUIRE_PTR(t_c); RNERF_REQUIRE_PTR(weights_c); RNERF_REQUIRE_PTR(u);
  RNERF_REQUIRE_PTR(t_f); RNERF_REQUIRE_PTR(pos_f); RNERF_REQUIRE_PTR(dir_f);
  RNERF_REQUIRE(aligned16(path), RNERF_E_ALIGN, "rnerf_resample: path must be 16-byte aligned");
  const bool staged = t_col != nullptr && n_steps <= RS_MAX_TROW;
  const size_t dyn = staged ? (size_t)RS_WARPS * n_steps * sizeof(float) : 0;
  if (dyn > 24 * 1024) {   // static 17 KB + dynamic must be opted into above 48 KB
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_WARPS * RS_MAX_TROW * 4);
      if (e != cudaSuccess) { set_error("rnerf_resample: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
      attr_set[dev] = true;
    }
  }
  resample_kernel<<<(unsigned)((n_rays + RS_WARPS - 1) / RS_WARPS), RS_WARPS * 32, dyn, (cudaStream_t)stream>>>(
      (const float4*)path, rec_floats / 4, staged ? t_col : nullptr, n_rays, n_steps, t_c, weights_c, n_coarse, u, u_per_ray,
      n_fine, t_f, pos_f, dir_f, grad_f);
  count_launch();
  return check_launch("rnerf_resample");
}
