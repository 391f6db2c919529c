// Bent-ray eikonal march (a5/a6) and coarse-sample selection (a7).
//
// OneEikonalStep (rnerf/eikonal_utils.py:30-49), radiance stage:
//     (n, g) = linear3(p);  p' = p + step/n * v;  v' = v + step * g;  t' = t + |p - p'|
// PathSampler.__call__ (rnerf/eikonal_utils.py:101-124) returns the state BEFORE each step and the
// lookup made at that state; ray_dir is safe_l2_normalize(v) (rnerf/math_utils.py:6-12).
//
// All state arithmetic uses non-contracted fp32 ops in the reference's association order, so the
// emitted path is bit-identical to the fp32 oracle.
#include "common.cuh"

namespace rnerf {

// One thread per ray; a warp is a bundle of 32 consecutive rays (adjacent pixels -> adjacent voxels, so
// the 8 float4 gathers of a warp land in few cache lines and the grid stays L1/L2 resident).  Records are
// staged through shared memory so that the global stores of a warp are sector-complete and contiguous:
// STEPS_PER_FLUSH steps x 48 B = 192 B contiguous per ray, all 32 lanes active in every store.
constexpr int MARCH_THREADS = 128;
constexpr int STEPS_PER_FLUSH = 4;                     // 4 records = 12 float4 per ray per flush
constexpr int F4_PER_FLUSH = STEPS_PER_FLUSH * 3;      // 12
constexpr int STAGE_PITCH = F4_PER_FLUSH + 1;          // +1 float4 pad: conflict-free column writes

__global__ void __launch_bounds__(MARCH_THREADS, 8) march_kernel(const float4* __restrict__ table, GridGeom g,
                                                              const float* __restrict__ origins,
                                                              const float* __restrict__ viewdirs, int64_t n_rays,
                                                              float near, float step, int n_steps,
                                                              float4* __restrict__ path,
                                                              const float* __restrict__ bricks) {
  __shared__ float4 stage[MARCH_THREADS / 32][32 * STAGE_PITCH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_ray0 = (blockIdx.x * (int64_t)MARCH_THREADS) + warp * 32;
  if (warp_ray0 >= n_rays) return;
  const int64_t ray = warp_ray0 + lane;
  const bool live = ray < n_rays;
  const int64_t rr = live ? ray : (n_rays - 1);
  float ox = origins[3 * rr], oy = origins[3 * rr + 1], oz = origins[3 * rr + 2];
  float vx = viewdirs[3 * rr], vy = viewdirs[3 * rr + 1], vz = viewdirs[3 * rr + 2];
  float px = add(ox, mul(near, vx)), py = add(oy, mul(near, vy)), pz = add(oz, mul(near, vz));
  float t = near;
  float4* my_stage = &stage[warp][lane * STAGE_PITCH];
  const int rays_here = (int)min((int64_t)32, n_rays - warp_ray0);

  // one eikonal step: emit the record of the current state, then advance it
  auto one_step = [&](int kk) {
    float4 c = trilinear(table, g, px, py, pz, bricks);  // (n, gx, gy, gz) at the pre-update position
    // the direction is stored un-normalised; readers apply safe_l2_normalize (path_dir()) to the few records they
    // use, which keeps 3 IEEE divides + 1 sqrt per step out of the march loop
    my_stage[kk * 3 + 0] = make_float4(px, py, pz, t);
    my_stage[kk * 3 + 1] = make_float4(vx, vy, vz, c.x);
    my_stage[kk * 3 + 2] = make_float4(c.y, c.z, c.w, 0.f);
    float s = divf(step, c.x);
    float nx = add(px, mul(s, vx)), ny = add(py, mul(s, vy)), nz = add(pz, mul(s, vz));
    vx = add(vx, mul(step, c.y)); vy = add(vy, mul(step, c.z)); vz = add(vz, mul(step, c.w));
    t = add(t, sqrtf(sumsq3(sub(px, nx), sub(py, ny), sub(pz, nz))));
    px = nx; py = ny; pz = nz;
  };

  for (int k0 = 0; k0 < n_steps; k0 += STEPS_PER_FLUSH) {
    const int nk = min(STEPS_PER_FLUSH, n_steps - k0);
    if (nk == STEPS_PER_FLUSH) {
#pragma unroll
      for (int kk = 0; kk < STEPS_PER_FLUSH; ++kk) one_step(kk);
    } else {
      for (int kk = 0; kk < nk; ++kk) one_step(kk);
    }
    __syncwarp();
    // cooperative flush: element e -> (ray e / n4, float4 e % n4); consecutive lanes write consecutive bytes
    float4* wbase = path + (warp_ray0 * (int64_t)n_steps + k0) * 3;   // one 64-bit base per flush, 32-bit offsets below
    const int ray_stride4 = n_steps * 3;                               // float4 units between consecutive rays
    if (nk == STEPS_PER_FLUSH && rays_here == 32) {
      // common case: compile-time trip count and divisor (e / 12 is a multiply-shift)
#pragma unroll
      for (int it = 0; it < F4_PER_FLUSH; ++it) {
        const int e = it * 32 + lane;
        const int r = e / F4_PER_FLUSH, j = e - r * F4_PER_FLUSH;
        __stcs(wbase + r * ray_stride4 + j, stage[warp][r * STAGE_PITCH + j]);
      }
    } else {
      const int n4 = nk * 3;
      const int total4 = rays_here * n4;
      for (int e = lane; e < total4; e += 32) {
        const int r = e / n4, j = e - r * n4;
        __stcs(wbase + r * ray_stride4 + j, stage[warp][r * STAGE_PITCH + j]);
      }
    }
    __syncwarp();
  }
}

// ray_dir of every record, normalised: the array PathSampler returns (rnerf/eikonal_utils.py:113)
__global__ void __launch_bounds__(256) path_dirs_kernel(const float4* __restrict__ path, int64_t n_rec,
                                                        float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rec) return;
  float3 d = path_dir(__ldg(path + i * 3 + 1));
  out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
}

// rnerf/models.py:243-247: ray_pos[:, jitter] etc.
__global__ void __launch_bounds__(256) select_kernel(const float4* __restrict__ path, int64_t n_rays, int n_steps,
                                                     const int32_t* __restrict__ jitter, int n_coarse,
                                                     float* __restrict__ pos_c, float* __restrict__ dir_c,
                                                     float* __restrict__ t_c, float* __restrict__ grad_c) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rays * n_coarse) return;
  int64_t r = i / n_coarse;
  int j = (int)(i % n_coarse);
  int k = min(max(__ldg(jitter + j), 0), n_steps - 1);
  const float4* rec = path + (r * n_steps + k) * 3;
  float4 a = __ldg(rec), b = __ldg(rec + 1);
  pos_c[3 * i] = a.x; pos_c[3 * i + 1] = a.y; pos_c[3 * i + 2] = a.z;
  t_c[i] = a.w;
  const float3 dn = path_dir(b);
  dir_c[3 * i] = dn.x; dir_c[3 * i + 1] = dn.y; dir_c[3 * i + 2] = dn.z;
  if (grad_c) {
    float4 c = __ldg(rec + 2);
    grad_c[3 * i] = c.x; grad_c[3 * i + 1] = c.y; grad_c[3 * i + 2] = c.z;
  }
}

}  // namespace rnerf

using namespace rnerf;

extern "C" int rnerf_march_fwd(const float* table, const float* bricks, const int ndim[3], const double nmin[3],
                               const double nmax[3], const float* origins, const float* viewdirs, int64_t n_rays,
                               double near, double far, int n_steps, float* path, void* stream) {
  RNERF_REQUIRE_PTR(table); RNERF_REQUIRE_PTR(ndim); RNERF_REQUIRE_PTR(nmin); RNERF_REQUIRE_PTR(nmax);
  RNERF_REQUIRE(n_rays >= 0, RNERF_E_SHAPE, "rnerf_march_fwd: n_rays < 0");
  RNERF_REQUIRE(n_steps >= 2, RNERF_E_SHAPE, "rnerf_march_fwd: n_steps must be >= 2 (step = (far-near)/(S-1))");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(origins); RNERF_REQUIRE_PTR(viewdirs); RNERF_REQUIRE_PTR(path);
  RNERF_REQUIRE(aligned16(table) && aligned16(path), RNERF_E_ALIGN, "rnerf_march_fwd: table/path must be 16-byte aligned");
  RNERF_REQUIRE(grid_fits_int32(ndim), RNERF_E_SHAPE, "rnerf_march_fwd: grids with >= 2^31 voxels are not supported");
  GridGeom g = make_geom(ndim, nmin, nmax);
  const float step = (float)((far - near) / (n_steps - 1));
  const unsigned blocks = (unsigned)((n_rays + MARCH_THREADS - 1) / MARCH_THREADS);
  march_kernel<<<blocks, MARCH_THREADS, 0, (cudaStream_t)stream>>>((const float4*)table, g, origins, viewdirs, n_rays,
                                                                   (float)near, step, n_steps, (float4*)path, bricks);
  count_launch();
  return check_launch("rnerf_march_fwd");
}

extern "C" int rnerf_select(const float* path, int64_t n_rays, int n_steps, const int32_t* jitter, int n_coarse,
                            float* pos_c, float* dir_c, float* t_c, float* grad_c, void* stream) {
  RNERF_REQUIRE(n_rays >= 0 && n_steps > 0 && n_coarse > 0, RNERF_E_SHAPE, "rnerf_select: bad sizes");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(path); RNERF_REQUIRE_PTR(jitter); RNERF_REQUIRE_PTR(pos_c); RNERF_REQUIRE_PTR(dir_c); RNERF_REQUIRE_PTR(t_c);
  RNERF_REQUIRE(aligned16(path), RNERF_E_ALIGN, "rnerf_select: path must be 16-byte aligned");
  int64_t total = n_rays * n_coarse;
  select_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)path, n_rays, n_steps,
                                                                                  jitter, n_coarse, pos_c, dir_c, t_c,
                                                                                  grad_c);
  count_launch();
  return check_launch("rnerf_select");
}

extern "C" int rnerf_path_dirs(const float* path, int64_t n_rays, int n_steps, float* ray_dir, void* stream) {
  RNERF_REQUIRE(n_rays >= 0 && n_steps > 0, RNERF_E_SHAPE, "rnerf_path_dirs: bad sizes");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(path); RNERF_REQUIRE_PTR(ray_dir);
  RNERF_REQUIRE(aligned16(path), RNERF_E_ALIGN, "rnerf_path_dirs: path must be 16-byte aligned");
  const int64_t n = n_rays * n_steps;
  path_dirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)path, n, ray_dir);
  count_launch();
  return check_launch("rnerf_path_dirs");
}
