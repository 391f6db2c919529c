// Ray generation and image error on the device: the steps either side of model.apply.
//
// rnerf/datasets.py:216-242 (Blender cameras) and :486-518 (OpenCV cameras), per pixel (x = column, y = row):
//     blender: cam = ((x + pc - W/2) / focal, -(y + pc - H/2) / focal, -1)
//     opencv : cam = ((x - cx + pc) / fx,      (y - cy + pc) / fy,      1)
//     directions = R cam (sum over the camera axis, left to right); origins = t; viewdirs = directions / |directions|
//     radii = |directions[y] - directions[y+1]| * 2 / sqrt(12)   (row neighbour; the last row repeats dx[H-3], sic)
// fp32 throughout, same operation order as the numpy code.  One thread per pixel; 52 B written per ray, nothing read:
// bound by the HBM write (the arrays exist only because the reference's render API takes them as inputs).
#include <math.h>
#include "common.cuh"

namespace rnerf {

struct CamArgs {
  float r[9];         // camtoworld[:3,:3] row-major (fp32 like the reference's arrays)
  float t[3];         // camtoworld[:3,3]
  float fx, fy, cx, cy, pc;
  int opencv;
  int H, W;
};

__device__ __forceinline__ void pixel_dir(const CamArgs& c, int x, int y, float& dx, float& dy, float& dz) {
  float cxv, cyv, czv;
  if (c.opencv) {
    cxv = divf(add(sub((float)x, c.cx), c.pc), c.fx);
    cyv = divf(add(sub((float)y, c.cy), c.pc), c.fy);
    czv = 1.f;
  } else {
    cxv = divf(sub(add((float)x, c.pc), mul((float)c.W, 0.5f)), c.fx);
    cyv = -divf(sub(add((float)y, c.pc), mul((float)c.H, 0.5f)), c.fx);
    czv = -1.f;
  }
  dx = add(add(mul(cxv, c.r[0]), mul(cyv, c.r[1])), mul(czv, c.r[2]));
  dy = add(add(mul(cxv, c.r[3]), mul(cyv, c.r[4])), mul(czv, c.r[5]));
  dz = add(add(mul(cxv, c.r[6]), mul(cyv, c.r[7])), mul(czv, c.r[8]));
}

__global__ void __launch_bounds__(256) generate_rays_kernel(const CamArgs c, int row0, int n_rows, float* __restrict__ origins,
                                                            float* __restrict__ directions, float* __restrict__ viewdirs,
                                                            float* __restrict__ radii) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_rows * c.W) return;
  const int y = row0 + (int)(i / c.W), x = (int)(i % c.W);
  float dx, dy, dz;
  pixel_dir(c, x, y, dx, dy, dz);
  if (origins) { origins[3 * i] = c.t[0]; origins[3 * i + 1] = c.t[1]; origins[3 * i + 2] = c.t[2]; }
  if (directions) { directions[3 * i] = dx; directions[3 * i + 1] = dy; directions[3 * i + 2] = dz; }
  if (viewdirs) {
    const float n = sqrtf(sumsq3(dx, dy, dz));
    viewdirs[3 * i] = divf(dx, n); viewdirs[3 * i + 1] = divf(dy, n); viewdirs[3 * i + 2] = divf(dz, n);
  }
  if (radii) {
    // dx[y] = |dir[y] - dir[y+1]| for y < H-1; the last row is filled with dx[-2:-1], i.e. dx[H-3] (the row before the
    // last difference), exactly as the reference concatenates it
    const int ya = (y < c.H - 1) ? y : max(c.H - 3, 0);
    float ax, ay, az, bx, by, bz;
    pixel_dir(c, x, ya, ax, ay, az);
    pixel_dir(c, x, min(ya + 1, c.H - 1), bx, by, bz);
    const float d = sqrtf(sumsq3(sub(ax, bx), sub(ay, by), sub(az, bz)));
    radii[i] = divf(mul(d, 2.f), 3.46410161513775458705f);     // fp32(sqrt(12))
  }
}

// out[0] += sum (a - b)^2   (mse = out / n; psnr = -10 log10(mse), rnerf/utils.py:392-401)
__global__ void __launch_bounds__(256) sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                     float* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc = fmaf(d, d, acc);
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

}  // namespace rnerf

using namespace rnerf;

extern "C" int rnerf_generate_rays(const double camtoworld_host[12], int height, int width, int opencv, double fx, double fy,
                                   double cx, double cy, int use_pixel_centers, int row0, int n_rows, float* origins,
                                   float* directions, float* viewdirs, float* radii, void* stream) {
  RNERF_REQUIRE_PTR(camtoworld_host);
  RNERF_REQUIRE(height > 0 && width > 0 && row0 >= 0 && n_rows >= 0 && row0 + n_rows <= height, RNERF_E_SHAPE,
                "rnerf_generate_rays: bad image geometry (H %d, W %d, rows %d..%d)", height, width, row0, row0 + n_rows);
  RNERF_REQUIRE(fx != 0.0 && (!opencv || fy != 0.0), RNERF_E_SHAPE, "rnerf_generate_rays: zero focal length");
  if (n_rows == 0) return 0;
  CamArgs c;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) c.r[3 * i + j] = (float)camtoworld_host[4 * i + j];
    c.t[i] = (float)camtoworld_host[4 * i + 3];
  }
  c.fx = (float)fx; c.fy = (float)fy; c.cx = (float)cx; c.cy = (float)cy; c.pc = use_pixel_centers ? 0.5f : 0.f;
  c.opencv = opencv; c.H = height; c.W = width;
  const int64_t n = (int64_t)n_rows * width;
  generate_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, row0, n_rows, origins, directions,
                                                                                     viewdirs, radii);
  count_launch();
  return check_launch("rnerf_generate_rays");
}

extern "C" int rnerf_sq_err(const float* a, const float* b, int64_t n, float* out_accum, void* stream) {
  if (n <= 0) return 0;
  RNERF_REQUIRE_PTR(a); RNERF_REQUIRE_PTR(b); RNERF_REQUIRE_PTR(out_accum);
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)n_sm * 8) blocks = (int64_t)n_sm * 8;
  sq_err_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, n, out_accum);
  count_launch();
  return check_launch("rnerf_sq_err");
}
