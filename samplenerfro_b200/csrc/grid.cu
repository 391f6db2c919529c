// IoR grid preparation and lookup: Gaussian prefilter (a1), (n, grad n) table (a2), trilinear lookup (a3).
#include <math.h>
#include "march_common.cuh"

namespace rnerf {

#define RNERF_MAX_BLUR_WS 11
__constant__ float c_blur[RNERF_MAX_BLUR_WS * RNERF_MAX_BLUR_WS * RNERF_MAX_BLUR_WS];

// conv3d_normal (rnerf/ior_utils.py:327-363): edge padding + 'VALID' correlation == clamp-to-edge taps.
__global__ void __launch_bounds__(256) blur_kernel(const float* __restrict__ in, float* __restrict__ out, int gx, int gy,
                                                   int gz, int ws) {
  const int64_t total = (int64_t)gx * gy * gz;
  const int hws = ws / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int z = (int)(i % gz);
    int y = (int)((i / gz) % gy);
    int x = (int)(i / ((int64_t)gz * gy));
    float acc = 0.f;
    int t = 0;
    for (int a = 0; a < ws; ++a) {
      int xx = min(max(x + a - hws, 0), gx - 1);
      for (int b = 0; b < ws; ++b) {
        int yy = min(max(y + b - hws, 0), gy - 1);
        const float* row = in + ((int64_t)xx * gy + yy) * gz;
        for (int c = 0; c < ws; ++c, ++t) {
          int zz = min(max(z + c - hws, 0), gz - 1);
          acc = fmaf(c_blur[t], __ldg(row + zz), acc);
        }
      }
    }
    out[i] = acc;
  }
}

// a / d for a divisor d that is constant over the launch: the exhaustively verified 3-instruction sequence of the march
// (march_common.cuh, recip = recip_for(d) or 0) where it applies -- same bits as the IEEE division -- else the division.
__device__ __forceinline__ float div_const(float a, float d, float recip) {
  if (recip != 0.f) {
    const float aa = fabsf(a);
    if (aa >= 0x1p-60f && aa <= 0x1p60f) return div_by_const(a, d, recip);
    if (a == 0.f) return d > 0.f ? a : -a;          // (most voxels: a homogeneous neighbourhood)
  }
  return divf(a, d);
}
struct TableRecips { float r[3]; };

// VoxMLP._compute_grad (rnerf/ior_utils.py:165-172): central differences on the edge-padded grid / (2*ndelta).
// One block per (x, y) row, threads along z: 32-bit index arithmetic (the first version derived (x, y, z) from a 64-bit
// flat index with three divisions per voxel and ran at 1.5 TB/s on a 2.5 GB pass: instruction-bound).
__global__ void __launch_bounds__(256) table_kernel(const float* __restrict__ n, float4* __restrict__ table, GridGeom g,
                                                    TableRecips rc) {
  const int rows = g.gx * g.gy;
  const int64_t sx = (int64_t)g.gy * g.gz, sy = g.gz;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int x = row / g.gy, y = row - x * g.gy;
    const int xm = max(x - 1, 0), xp = min(x + 1, g.gx - 1), ym = max(y - 1, 0), yp = min(y + 1, g.gy - 1);
    const float* c = n + sx * x + sy * y;
    const float* cxp = n + sx * xp + sy * y; const float* cxm = n + sx * xm + sy * y;
    const float* cyp = n + sx * x + sy * yp; const float* cym = n + sx * x + sy * ym;
    float4* out = table + (int64_t)row * g.gz;
    for (int z = threadIdx.x; z < g.gz; z += blockDim.x) {
      const int zm = max(z - 1, 0), zp = min(z + 1, g.gz - 1);
      const float dx = div_const(sub(__ldg(cxp + z), __ldg(cxm + z)), g.two_ndelta[0], rc.r[0]);
      const float dy = div_const(sub(__ldg(cyp + z), __ldg(cym + z)), g.two_ndelta[1], rc.r[1]);
      const float dz = div_const(sub(__ldg(c + zp), __ldg(c + zm)), g.two_ndelta[2], rc.r[2]);
      out[z] = make_float4(__ldg(c + z), dx, dy, dz);
    }
  }
}

// Adjoint of table_kernel (extension, no reference counterpart: a learned IoR grid): d_n[i] = d_table[i].n + the transposed
// central differences, edge clamping included (n[0] and n[G-1] each enter their own boundary difference twice).
__global__ void __launch_bounds__(256) table_bwd_kernel(const float4* __restrict__ dt, float* __restrict__ dn, GridGeom g) {
  const int rows = g.gx * g.gy;
  const int64_t sx = (int64_t)g.gy * g.gz, sy = g.gz;
  const float ix = 1.f / g.two_ndelta[0], iy = 1.f / g.two_ndelta[1], iz = 1.f / g.two_ndelta[2];
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int x = row / g.gy, y = row - x * g.gy;
    const float4* c = dt + (int64_t)row * g.gz;
    float* o = dn + (int64_t)row * g.gz;
    for (int z = threadIdx.x; z < g.gz; z += blockDim.x) {
      const float4 here = __ldg(c + z);
      float acc = here.x;
      float s = 0.f;
      if (x >= 1) s += __ldg(c + z - sx).y;
      if (x == g.gx - 1) s += here.y;
      if (x <= g.gx - 2) s -= __ldg(c + z + sx).y;
      if (x == 0) s -= here.y;
      acc += s * ix;
      s = 0.f;
      if (y >= 1) s += __ldg(c + z - sy).z;
      if (y == g.gy - 1) s += here.z;
      if (y <= g.gy - 2) s -= __ldg(c + z + sy).z;
      if (y == 0) s -= here.z;
      acc += s * iy;
      s = 0.f;
      if (z >= 1) s += __ldg(c + z - 1).w;
      if (z == g.gz - 1) s += here.w;
      if (z <= g.gz - 2) s -= __ldg(c + z + 1).w;
      if (z == 0) s -= here.w;
      acc += s * iz;
      o[z] = acc;
    }
  }
}

__global__ void __launch_bounds__(256) lookup_kernel(const float4* __restrict__ table, GridGeom g,
                                                     const float* __restrict__ pts, int64_t n_pts,
                                                     float4* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  out[i] = trilinear(table, g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
}

// One thread per brick: homogeneous iff every voxel of [B*b, B*b + B] (clamped) equals the first bit-for-bit and
// has a zero gradient.  The +1 halo covers the x1/y1/z1 corners of the last cell row of the brick.
__global__ void __launch_bounds__(128) brick_kernel(const float4* __restrict__ table, GridGeom g, float* __restrict__ bricks,
                                                    int nbx, int nby, int nbz) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= (int64_t)nbx * nby * nbz) return;
  const int bz = (int)(b % nbz), by = (int)((b / nbz) % nby), bx = (int)(b / ((int64_t)nbz * nby));
  const int x0 = bx * BRICK, y0 = by * BRICK, z0 = bz * BRICK;
  const int x1 = min(x0 + BRICK, g.gx - 1), y1 = min(y0 + BRICK, g.gy - 1), z1 = min(z0 + BRICK, g.gz - 1);
  const float4 first = __ldg(table + ((int64_t)x0 * g.gy + y0) * g.gz + z0);
  bool ok = true;
  for (int x = x0; x <= x1 && ok; ++x)
    for (int y = y0; y <= y1 && ok; ++y)
      for (int z = z0; z <= z1; ++z) {
        const float4 v = __ldg(table + ((int64_t)x * g.gy + y) * g.gz + z);
        if (__float_as_uint(v.x) != __float_as_uint(first.x) || v.y != 0.f || v.z != 0.f || v.w != 0.f) { ok = false; break; }
      }
  ok = ok && (first.x == first.x) && !isinf(first.x);
  bricks[b] = ok ? first.x : __int_as_float(0x7fc00000);
}

static int grid_blocks(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148 * 32;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace rnerf

using namespace rnerf;

extern "C" int rnerf_grid_blur(const float* n_in, float* n_out, const int ndim[3], int ws, double sigma, void* stream) {
  RNERF_REQUIRE_PTR(n_in); RNERF_REQUIRE_PTR(n_out); RNERF_REQUIRE_PTR(ndim);
  RNERF_REQUIRE(ws >= 1 && ws <= RNERF_MAX_BLUR_WS && (ws & 1), RNERF_E_SHAPE, "rnerf_grid_blur: ws=%d unsupported (odd, <=%d)", ws,
                RNERF_MAX_BLUR_WS);
  RNERF_REQUIRE(n_in != n_out, RNERF_E_SHAPE, "rnerf_grid_blur: in-place blur is not supported");
  // kernel weights in fp32 like the reference: linspace(-hws,hws,ws), exp(-(x^2+y^2+z^2)/(2 s^2)) / sum
  static thread_local float k[RNERF_MAX_BLUR_WS * RNERF_MAX_BLUR_WS * RNERF_MAX_BLUR_WS];
  const int hws = ws / 2;
  const float two_s2 = (float)(2.0 * sigma * sigma);
  float sum = 0.f;
  int t = 0;
  for (int a = 0; a < ws; ++a)
    for (int b = 0; b < ws; ++b)
      for (int c = 0; c < ws; ++c, ++t) {
        float xa = (float)(a - hws), xb = (float)(b - hws), xc = (float)(c - hws);
        k[t] = expf(-((xa * xa + xb * xb) + xc * xc) / two_s2);
        sum += k[t];
      }
  for (int i = 0; i < t; ++i) k[i] = k[i] / sum;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyToSymbolAsync(c_blur, k, sizeof(float) * t, 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { set_error("rnerf_grid_blur: %s", cudaGetErrorString(e)); return (int)e; }
  const int64_t total = (int64_t)ndim[0] * ndim[1] * ndim[2];
  blur_kernel<<<grid_blocks(total, 256), 256, 0, st>>>(n_in, n_out, ndim[0], ndim[1], ndim[2], ws);
  count_launch();
  return check_launch("rnerf_grid_blur");
}

extern "C" int rnerf_grid_table(const float* n, const int ndim[3], const double nmin[3], const double nmax[3],
                                float* table, void* stream) {
  RNERF_REQUIRE_PTR(n); RNERF_REQUIRE_PTR(table); RNERF_REQUIRE_PTR(ndim); RNERF_REQUIRE_PTR(nmin); RNERF_REQUIRE_PTR(nmax);
  RNERF_REQUIRE(ndim[0] >= 2 && ndim[1] >= 2 && ndim[2] >= 2, RNERF_E_SHAPE, "rnerf_grid_table: ndim must be >= 2");
  RNERF_REQUIRE(aligned16(table), RNERF_E_ALIGN, "rnerf_grid_table: table must be 16-byte aligned");
  GridGeom g = make_geom(ndim, nmin, nmax);
  RNERF_REQUIRE((int64_t)ndim[0] * ndim[1] < 2147483647LL, RNERF_E_SHAPE, "rnerf_grid_table: ndim[0] * ndim[1] must fit 31 bits");
  TableRecips rc;
  for (int i = 0; i < 3; ++i) rc.r[i] = fast_div_enabled() ? recip_for(g.two_ndelta[i], (cudaStream_t)stream) : 0.f;
  const int rows = ndim[0] * ndim[1];
  table_kernel<<<rows < 1048576 ? rows : 1048576, 256, 0, (cudaStream_t)stream>>>(n, (float4*)table, g, rc);
  count_launch();
  return check_launch("rnerf_grid_table");
}

extern "C" int rnerf_grid_table_bwd(const float* d_table, const int ndim[3], const double nmin[3], const double nmax[3],
                                    float* d_n, void* stream) {
  RNERF_REQUIRE_PTR(d_table); RNERF_REQUIRE_PTR(ndim); RNERF_REQUIRE_PTR(nmin); RNERF_REQUIRE_PTR(nmax); RNERF_REQUIRE_PTR(d_n);
  RNERF_REQUIRE(ndim[0] > 1 && ndim[1] > 1 && ndim[2] > 1, RNERF_E_SHAPE, "rnerf_grid_table_bwd: every grid side must be >= 2");
  RNERF_REQUIRE(aligned16(d_table), RNERF_E_ALIGN, "rnerf_grid_table_bwd: d_table must be 16-byte aligned");
  GridGeom g = make_geom(ndim, nmin, nmax);
  RNERF_REQUIRE((int64_t)ndim[0] * ndim[1] < 2147483647LL, RNERF_E_SHAPE, "rnerf_grid_table_bwd: ndim[0] * ndim[1] must fit 31 bits");
  const int rows = ndim[0] * ndim[1];
  table_bwd_kernel<<<rows < 1048576 ? rows : 1048576, 256, 0, (cudaStream_t)stream>>>((const float4*)d_table, d_n, g);
  count_launch();
  return check_launch("rnerf_grid_table_bwd");
}

extern "C" int rnerf_grid_lookup(const float* table, const int ndim[3], const double nmin[3], const double nmax[3],
                                 const float* pts, int64_t n_pts, float* out, void* stream) {
  RNERF_REQUIRE_PTR(table); RNERF_REQUIRE_PTR(ndim); RNERF_REQUIRE_PTR(nmin); RNERF_REQUIRE_PTR(nmax);
  if (n_pts == 0) return 0;
  RNERF_REQUIRE_PTR(pts); RNERF_REQUIRE_PTR(out);
  RNERF_REQUIRE(n_pts > 0, RNERF_E_SHAPE, "rnerf_grid_lookup: n_pts < 0");
  RNERF_REQUIRE(aligned16(table) && aligned16(out), RNERF_E_ALIGN, "rnerf_grid_lookup: table/out must be 16-byte aligned");
  RNERF_REQUIRE(grid_fits_int32(ndim), RNERF_E_SHAPE, "rnerf_grid_lookup: grids with >= 2^31 voxels are not supported");
  GridGeom g = make_geom(ndim, nmin, nmax);
  lookup_kernel<<<(unsigned)((n_pts + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)table, g, pts, n_pts,
                                                                                    (float4*)out);
  count_launch();
  return check_launch("rnerf_grid_lookup");
}

extern "C" int64_t rnerf_grid_brick_count(const int ndim[3]) {
  if (!ndim) return 0;
  return (int64_t)((ndim[0] + BRICK - 1) / BRICK) * ((ndim[1] + BRICK - 1) / BRICK) * ((ndim[2] + BRICK - 1) / BRICK);
}

extern "C" int rnerf_grid_bricks(const float* table, const int ndim[3], float* bricks, void* stream) {
  RNERF_REQUIRE_PTR(table); RNERF_REQUIRE_PTR(ndim); RNERF_REQUIRE_PTR(bricks);
  RNERF_REQUIRE(aligned16(table), RNERF_E_ALIGN, "rnerf_grid_bricks: table must be 16-byte aligned");
  const double zero[3] = {0, 0, 0}, one[3] = {1, 1, 1};
  GridGeom g = make_geom(ndim, zero, one);
  const int nbx = (ndim[0] + BRICK - 1) / BRICK, nby = (ndim[1] + BRICK - 1) / BRICK, nbz = (ndim[2] + BRICK - 1) / BRICK;
  const int64_t total = (int64_t)nbx * nby * nbz;
  brick_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const float4*)table, g, bricks, nbx, nby, nbz);
  count_launch();
  return check_launch("rnerf_grid_bricks");
}
