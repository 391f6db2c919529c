// Shared helpers for the rnerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/rnerf_b200.h"

namespace rnerf {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// returns 0 / records the CUDA error of the launch just issued
int check_launch(const char* what);

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define RNERF_REQUIRE_PTR(p)                                   \
  do {                                                         \
    if ((p) == nullptr) {                                      \
      rnerf::set_error("%s: null pointer '%s'", __func__, #p); \
      return RNERF_E_NULL;                                     \
    }                                                          \
  } while (0)

#define RNERF_REQUIRE(cond, code, ...)  \
  do {                                  \
    if (!(cond)) {                      \
      rnerf::set_error(__VA_ARGS__);    \
      return (code);                    \
    }                                   \
  } while (0)

struct GridGeom {
  int gx, gy, gz;
  float nmin[3];
  float ndelta[3];      // fp32(ndelta_double)      (rnerf/ior_utils.py:140-144)
  float two_ndelta[3];  // fp32(2 * ndelta_double)  (rnerf/ior_utils.py:169-171)
};

static inline bool grid_fits_int32(const int ndim[3]) {
  return ndim[0] > 0 && ndim[1] > 0 && ndim[2] > 0 && (double)ndim[0] * ndim[1] * ndim[2] < 2147483648.0;
}

static inline GridGeom make_geom(const int ndim[3], const double nmin[3], const double nmax[3]) {
  GridGeom g;
  g.gx = ndim[0]; g.gy = ndim[1]; g.gz = ndim[2];
  for (int i = 0; i < 3; ++i) {
    double nd = (nmax[i] - nmin[i]) / (ndim[i] - 1.0);
    g.nmin[i] = (float)nmin[i];
    g.ndelta[i] = (float)nd;
    g.two_ndelta[i] = (float)(2.0 * nd);
  }
  return g;
}

// ---- exact (never FMA-contracted) fp32 arithmetic, same association as the reference ----
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float divf(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sumsq3(float x, float y, float z) { return add(add(mul(x, x), mul(y, y)), mul(z, z)); }
// a*(1-w) + b*w with separate roundings (rnerf/ior_utils.py:214-222)
__device__ __forceinline__ float lerp_ref(float a, float b, float omw, float w) { return add(mul(a, omw), mul(b, w)); }
__device__ __forceinline__ float4 lerp4_ref(float4 a, float4 b, float omw, float w) {
  return make_float4(lerp_ref(a.x, b.x, omw, w), lerp_ref(a.y, b.y, omw, w), lerp_ref(a.z, b.z, omw, w),
                     lerp_ref(a.w, b.w, omw, w));
}

// Brick map: the grid is cut into BRICK^3-voxel bricks; bricks[b] holds the common value c when every corner a
// lookup inside the brick can touch (voxels [BRICK*b, BRICK*b + BRICK] per axis, clamped) has n == c bit-for-bit
// and grad n == 0, and NaN otherwise.  In such a brick the 8 gathers are skipped; the lerp arithmetic is kept.
constexpr int BRICK_LOG2 = 3;
constexpr int BRICK = 1 << BRICK_LOG2;

// VoxMLP._linear3 (rnerf/ior_utils.py:188-223): unclamped floor/frac, clamp-to-edge indices, x->y->z lerps.
// `bricks` may be null (no skipping).  Results are bit-identical with and without the brick map: with all eight
// corners equal to c the seven lerps collapse to three lerps of identical operands, and 0*(1-w) + 0*w == 0.
__device__ __forceinline__ float4 trilinear(const float4* __restrict__ table, const GridGeom& g, float px, float py,
                                            float pz, const float* __restrict__ bricks = nullptr) {
  float x = divf(sub(px, g.nmin[0]), g.ndelta[0]);
  float y = divf(sub(py, g.nmin[1]), g.ndelta[1]);
  float z = divf(sub(pz, g.nmin[2]), g.ndelta[2]);
  float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  float xd = sub(x, xf), yd = sub(y, yf), zd = sub(z, zf);
  // clamp in float first so the int conversion cannot overflow; result identical after the clip
  int x0 = (int)fminf(fmaxf(xf, -2.f), (float)g.gx), y0 = (int)fminf(fmaxf(yf, -2.f), (float)g.gy),
      z0 = (int)fminf(fmaxf(zf, -2.f), (float)g.gz);
  int x1 = min(max(x0 + 1, 0), g.gx - 1), y1 = min(max(y0 + 1, 0), g.gy - 1), z1 = min(max(z0 + 1, 0), g.gz - 1);
  x0 = min(max(x0, 0), g.gx - 1); y0 = min(max(y0, 0), g.gy - 1); z0 = min(max(z0, 0), g.gz - 1);
  if (bricks != nullptr) {
    const int nby = (g.gy + BRICK - 1) >> BRICK_LOG2, nbz = (g.gz + BRICK - 1) >> BRICK_LOG2;
    const float c = __ldg(bricks + ((x0 >> BRICK_LOG2) * nby + (y0 >> BRICK_LOG2)) * nbz + (z0 >> BRICK_LOG2));
    if (c == c) {  // not NaN: homogeneous brick
      const float oxd = sub(1.f, xd), oyd = sub(1.f, yd), ozd = sub(1.f, zd);
      const float c00 = lerp_ref(c, c, oxd, xd);
      const float c0 = lerp_ref(c00, c00, oyd, yd);
      return make_float4(lerp_ref(c0, c0, ozd, zd), 0.f, 0.f, 0.f);
    }
  }
  // 32-bit voxel indices: the host rejects grids with more than 2^31 voxels
  const int sx = g.gy * g.gz, sy = g.gz;
  const int b00 = sx * x0 + sy * y0, b10 = sx * x1 + sy * y0, b01 = sx * x0 + sy * y1, b11 = sx * x1 + sy * y1;
  float4 d000 = __ldg(table + b00 + z0), d100 = __ldg(table + b10 + z0);
  float4 d001 = __ldg(table + b00 + z1), d101 = __ldg(table + b10 + z1);
  float4 d010 = __ldg(table + b01 + z0), d110 = __ldg(table + b11 + z0);
  float4 d011 = __ldg(table + b01 + z1), d111 = __ldg(table + b11 + z1);
  float oxd = sub(1.f, xd), oyd = sub(1.f, yd), ozd = sub(1.f, zd);
  float4 c00 = lerp4_ref(d000, d100, oxd, xd);
  float4 c01 = lerp4_ref(d001, d101, oxd, xd);
  float4 c10 = lerp4_ref(d010, d110, oxd, xd);
  float4 c11 = lerp4_ref(d011, d111, oxd, xd);
  float4 c0 = lerp4_ref(c00, c10, oyd, yd);
  float4 c1 = lerp4_ref(c01, c11, oyd, yd);
  return lerp4_ref(c0, c1, ozd, zd);
}

// safe_l2_normalize (rnerf/math_utils.py:6-12) of the direction state stored in a path record's 2nd float4
__device__ __forceinline__ float3 path_dir(float4 rec1) {
  const float vn = sqrtf(fmaxf(sumsq3(rec1.x, rec1.y, rec1.z), 1e-6f));
  return make_float3(divf(rec1.x, vn), divf(rec1.y, vn), divf(rec1.z, vn));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

}  // namespace rnerf
