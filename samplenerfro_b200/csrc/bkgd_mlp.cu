// Background MLP (a10): model_utils.MLP(net_width=128, net_depth=4, skip_layer=2) on pos_enc(dir, 0, 4)
// (rnerf/model_utils.py:93-140, rnerf/models.py:116-118,303).  fp32 on CUDA cores: one evaluation per RAY
// (56 448 MAC) against 256 radiance-MLP evaluations per ray (152 M MAC), so this stage is <0.1 % of the work
// and stays in full precision.
//   Dense_0: 27->128  Dense_1: 128->128  Dense_2: 128->128, then concat(x, inputs) -> 155
//   Dense_3: 155->128  Dense_4: 128->3 (no activation)
#include "common.cuh"

namespace rnerf {

constexpr int BK_R = 32;        // rays per block
constexpr int BK_T = 2;         // threads per neuron: thread (j, h) = (tid >> 1, tid & 1) carries rays 16 h .. 16 h + 15 of neuron j
constexpr int BK_RT = BK_R / BK_T;
constexpr int BK_THREADS = 128 * BK_T;
constexpr int BK_PITCH = 36;    // floats per feature row (R + 4: conflict-free float4 column stores)
constexpr int BK_W = 128;
constexpr int BK_IN = 27;
constexpr int BK_K0 = 0;
constexpr int BK_K1 = BK_K0 + 27 * 128;
constexpr int BK_K2 = BK_K1 + 128 * 128;
constexpr int BK_K3 = BK_K2 + 128 * 128;
constexpr int BK_K4 = BK_K3 + 155 * 128;
constexpr int BK_B0 = BK_K4 + 128 * 3;
constexpr int BK_B1 = BK_B0 + 128;
constexpr int BK_B2 = BK_B1 + 128;
constexpr int BK_B3 = BK_B2 + 128;
constexpr int BK_B4 = BK_B3 + 128;
constexpr int BK_TOTAL = BK_B4 + 3;

// One block = 32 rays, 256 threads: two threads per neuron, 16 rays each.  (The first version ran one thread per neuron with
// all 32 rays in registers: 153 registers, 4 warps per block, and a block's serial chain -- 165 us for the backward -- was the
// whole launch time of a training batch, which fills less than one wave: profiles/r5d_train_backward_ncu_summary.txt.)
// acc[r] += sum_k W[k][j] * in[k][r0 + r]   (`in` already points at this thread's first ray)
__device__ __forceinline__ void dense_accum(float (&acc)[BK_RT], const float* __restrict__ W, int K, int j,
                                            const float* __restrict__ in) {
  for (int k = 0; k < K; ++k) {
    const float w = __ldg(W + k * BK_W + j);
    const float4* row = reinterpret_cast<const float4*>(in + k * BK_PITCH);
#pragma unroll
    for (int r4 = 0; r4 < BK_RT / 4; ++r4) {
      float4 a = row[r4];
      acc[4 * r4 + 0] = fmaf(a.x, w, acc[4 * r4 + 0]);
      acc[4 * r4 + 1] = fmaf(a.y, w, acc[4 * r4 + 1]);
      acc[4 * r4 + 2] = fmaf(a.z, w, acc[4 * r4 + 2]);
      acc[4 * r4 + 3] = fmaf(a.w, w, acc[4 * r4 + 3]);
    }
  }
}

__device__ __forceinline__ void store_relu(const float (&acc)[BK_RT], float bias, float* __restrict__ out_row) {
  float4* o = reinterpret_cast<float4*>(out_row);
#pragma unroll
  for (int r4 = 0; r4 < BK_RT / 4; ++r4)
    o[r4] = make_float4(fmaxf(acc[4 * r4] + bias, 0.f), fmaxf(acc[4 * r4 + 1] + bias, 0.f),
                        fmaxf(acc[4 * r4 + 2] + bias, 0.f), fmaxf(acc[4 * r4 + 3] + bias, 0.f));
}

__global__ void __launch_bounds__(BK_THREADS) bkgd_mlp_kernel(const float* __restrict__ w, const float* __restrict__ dirs,
                                                        int64_t n_rays, int64_t dir_stride, float* __restrict__ raw_out) {
  __shared__ __align__(16) float X[BK_W * BK_PITCH];
  __shared__ __align__(16) float Y[BK_W * BK_PITCH];
  __shared__ __align__(16) float E[BK_IN * BK_PITCH];
  const int tid = threadIdx.x, j = tid >> 1, r0 = (tid & 1) * BK_RT;
  const int64_t ray0 = blockIdx.x * (int64_t)BK_R;
  // pos_enc(dir, 0, 4): [x(3), sin(2^k x) k-major (12), sin(2^k x + pi/2) k-major (12)]  (model_utils.py:204-214)
  for (int e = tid; e < BK_IN * BK_R; e += BK_THREADS) {
    const int r = e % BK_R, f = e / BK_R;
    const int64_t ray = min(ray0 + r, n_rays - 1);
    const float* d = dirs + ray * dir_stride;
    float v;
    if (f < 3) {
      v = __ldg(d + f);
    } else {
      const int q = (f - 3) % 12, k = q / 3, c = q % 3;
      float xb = mul(__ldg(d + c), (float)(1 << k));
      if (f >= 15) xb = add(xb, 1.57079632679489661923f);
      v = sinf(xb);
    }
    E[f * BK_PITCH + r] = v;
  }
  __syncthreads();
  float acc[BK_RT];
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
  dense_accum(acc, w + BK_K0, BK_IN, j, E + r0);
  store_relu(acc, __ldg(w + BK_B0 + j), X + j * BK_PITCH + r0);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
  dense_accum(acc, w + BK_K1, BK_W, j, X + r0);
  store_relu(acc, __ldg(w + BK_B1 + j), Y + j * BK_PITCH + r0);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
  dense_accum(acc, w + BK_K2, BK_W, j, Y + r0);
  __syncthreads();
  store_relu(acc, __ldg(w + BK_B2 + j), X + j * BK_PITCH + r0);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
  dense_accum(acc, w + BK_K3, BK_W, j, X + r0);                       // rows 0..127: layer-2 output
  dense_accum(acc, w + BK_K3 + BK_W * BK_W, BK_IN, j, E + r0);        // rows 128..154: the skip-concatenated inputs
  store_relu(acc, __ldg(w + BK_B3 + j), Y + j * BK_PITCH + r0);
  __syncthreads();
  if (tid < 3 * BK_R) {
    const int r = tid % BK_R, c = tid / BK_R;
    float o = 0.f;
    for (int k = 0; k < BK_W; ++k) o = fmaf(Y[k * BK_PITCH + r], __ldg(w + BK_K4 + k * 3 + c), o);
    o += __ldg(w + BK_B4 + c);
    if (ray0 + r < n_rays) raw_out[(ray0 + r) * 3 + c] = o;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the background MLP (train.py differentiates bkgd_mlp through comp_rgb and the env-map smoothness term):
// the forward is recomputed per 32-ray tile with every layer's activations kept in shared memory, then the chain
//   dz_l = (K_{l+1} dz_{l+1}) * relu'(h_l),  gK_l += X_l^T dz_l,  gb_l += sum_r dz_l
// runs layer by layer; weight gradients are accumulated into `gw` (same flat layout as `w`) with red.global.add.
// ---------------------------------------------------------------------------------------------------------
// acc[r] += sum_j W[i][j] * in[j][r0 + r]   (thread (i, h) reads row i of a row-major [rows][ld] matrix: transposed use of K;
// `in` already points at this thread's first ray)
__device__ __forceinline__ void dense_accum_t(float (&acc)[BK_RT], const float* __restrict__ W, int ld, int ncols, int i,
                                              const float* __restrict__ in) {
  const float* wrow = W + (size_t)i * ld;
  auto step = [&](int j, float w) {
    const float4* row = reinterpret_cast<const float4*>(in + j * BK_PITCH);
#pragma unroll
    for (int r4 = 0; r4 < BK_RT / 4; ++r4) {
      float4 a = row[r4];
      acc[4 * r4 + 0] = fmaf(a.x, w, acc[4 * r4 + 0]);
      acc[4 * r4 + 1] = fmaf(a.y, w, acc[4 * r4 + 1]);
      acc[4 * r4 + 2] = fmaf(a.z, w, acc[4 * r4 + 2]);
      acc[4 * r4 + 3] = fmaf(a.w, w, acc[4 * r4 + 3]);
    }
  };
  // every lane walks its own weight row: 16-byte loads (4 columns per instruction; the rows are 512 B apart, so a scalar
  // walk costs one L1 request of 16 sectors per column) whenever the row is aligned; same summation order either way
  if ((reinterpret_cast<uintptr_t>(wrow) & 15u) == 0 && (ncols & 3) == 0) {
    for (int j = 0; j < ncols; j += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + j));
      step(j, w4.x); step(j + 1, w4.y); step(j + 2, w4.z); step(j + 3, w4.w);
    }
  } else {
    for (int j = 0; j < ncols; ++j) step(j, __ldg(wrow + j));
  }
}

// the two threads of a neuron sit in adjacent lanes: sum their halves, the even lane owns the result
__device__ __forceinline__ float pair_sum(float v) { return v + __shfl_xor_sync(0xffffffffu, v, 1); }

// gK[k][j] += sum_r X[k][r] * dz[j][r] for k < K (thread (j, h) holds dz[j][16 h ..] in registers; X points at its first ray);
// called by all threads of the block
__device__ __forceinline__ void wgrad_rows(const float (&dz)[BK_RT], const float* __restrict__ X, int K, int j, bool owner,
                                           int ld_out, float* __restrict__ gK) {
  for (int k = 0; k < K; ++k) {
    const float4* row = reinterpret_cast<const float4*>(X + k * BK_PITCH);
    float s = 0.f;
#pragma unroll
    for (int r4 = 0; r4 < BK_RT / 4; ++r4) {
      float4 a = row[r4];
      s = fmaf(a.x, dz[4 * r4 + 0], s); s = fmaf(a.y, dz[4 * r4 + 1], s);
      s = fmaf(a.z, dz[4 * r4 + 2], s); s = fmaf(a.w, dz[4 * r4 + 3], s);
    }
    s = pair_sum(s);
    if (owner) atomicAdd(gK + (size_t)k * ld_out + j, s);
  }
}

__global__ void __launch_bounds__(BK_THREADS) bkgd_mlp_bwd_kernel(const float* __restrict__ w, const float* __restrict__ dirs,
                                                                  int64_t n_rays, int64_t dir_stride, const float* __restrict__ d_raw,
                                                                  float* __restrict__ gw, float* __restrict__ d_dirs) {
  extern __shared__ __align__(16) float sm[];
  float* E = sm;                               // [27][36]
  float* Hs = E + BK_IN * BK_PITCH;            // [4][128][36]  post-activation outputs of Dense_0..3
  float* D = Hs + 4 * BK_W * BK_PITCH;         // [128][36]     current dz
  const int tid = threadIdx.x, j = tid >> 1, r0 = (tid & 1) * BK_RT;
  const bool owner = (tid & 1) == 0;
  const int64_t ray0 = blockIdx.x * (int64_t)BK_R;
  for (int e = tid; e < BK_IN * BK_R; e += BK_THREADS) {
    const int r = e % BK_R, f = e / BK_R;
    const int64_t ray = min(ray0 + r, n_rays - 1);
    const float* d = dirs + ray * dir_stride;
    float v;
    if (f < 3) {
      v = __ldg(d + f);
    } else {
      const int q = (f - 3) % 12, k = q / 3, c = q % 3;
      float xb = mul(__ldg(d + c), (float)(1 << k));
      if (f >= 15) xb = add(xb, 1.57079632679489661923f);
      v = sinf(xb);
    }
    E[f * BK_PITCH + r] = v;
  }
  __syncthreads();
  float acc[BK_RT];
  // ---- forward recompute, keeping h0..h3
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
  dense_accum(acc, w + BK_K0, BK_IN, j, E + r0);
  store_relu(acc, __ldg(w + BK_B0 + j), Hs + (0 * BK_W + j) * BK_PITCH + r0);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
  dense_accum(acc, w + BK_K1, BK_W, j, Hs + r0);
  store_relu(acc, __ldg(w + BK_B1 + j), Hs + (1 * BK_W + j) * BK_PITCH + r0);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
  dense_accum(acc, w + BK_K2, BK_W, j, Hs + 1 * BK_W * BK_PITCH + r0);
  store_relu(acc, __ldg(w + BK_B2 + j), Hs + (2 * BK_W + j) * BK_PITCH + r0);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
  dense_accum(acc, w + BK_K3, BK_W, j, Hs + 2 * BK_W * BK_PITCH + r0);
  dense_accum(acc, w + BK_K3 + BK_W * BK_W, BK_IN, j, E + r0);
  store_relu(acc, __ldg(w + BK_B3 + j), Hs + (3 * BK_W + j) * BK_PITCH + r0);
  __syncthreads();
  // ---- output layer Dense_4 (128 -> 3): thread (j, h) = hidden unit j, rays 16 h ..
  {
    float dy[3][BK_RT];
#pragma unroll
    for (int r = 0; r < BK_RT; ++r) {
      const bool live = ray0 + r0 + r < n_rays;
#pragma unroll
      for (int c = 0; c < 3; ++c) dy[c][r] = live ? __ldg(d_raw + (ray0 + r0 + r) * 3 + c) : 0.f;
    }
    const float* h3 = Hs + (3 * BK_W + j) * BK_PITCH + r0;
    const float k0 = __ldg(w + BK_K4 + j * 3), k1 = __ldg(w + BK_K4 + j * 3 + 1), k2 = __ldg(w + BK_K4 + j * 3 + 2);
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
    for (int r = 0; r < BK_RT; ++r) {
      const float h = h3[r];
      g0 = fmaf(h, dy[0][r], g0); g1 = fmaf(h, dy[1][r], g1); g2 = fmaf(h, dy[2][r], g2);
      acc[r] = h > 0.f ? (k0 * dy[0][r] + k1 * dy[1][r] + k2 * dy[2][r]) : 0.f;     // dz3[j][r]
    }
    g0 = pair_sum(g0); g1 = pair_sum(g1); g2 = pair_sum(g2);
    if (owner) { atomicAdd(gw + BK_K4 + j * 3, g0); atomicAdd(gw + BK_K4 + j * 3 + 1, g1); atomicAdd(gw + BK_K4 + j * 3 + 2, g2); }
    if (j == 0) {        // bias of the output layer: both lanes of neuron 0 hold the block's 32 rays between them
      float sb0 = 0.f, sb1 = 0.f, sb2 = 0.f;
#pragma unroll
      for (int r = 0; r < BK_RT; ++r) { sb0 += dy[0][r]; sb1 += dy[1][r]; sb2 += dy[2][r]; }
      sb0 += __shfl_xor_sync(0x3u, sb0, 1); sb1 += __shfl_xor_sync(0x3u, sb1, 1); sb2 += __shfl_xor_sync(0x3u, sb2, 1);
      if (owner) { atomicAdd(gw + BK_B4, sb0); atomicAdd(gw + BK_B4 + 1, sb1); atomicAdd(gw + BK_B4 + 2, sb2); }
    }
  }
  // gradient wrt the encoded direction (threads j < 27, "all" stage only): dE = K_3[128:]^T-part dz_3 + K_0 dz_0
  float accE[BK_RT];
#pragma unroll
  for (int r = 0; r < BK_RT; ++r) accE[r] = 0.f;
  // ---- hidden layers 3 -> 0: acc[] holds dz_l[j][16 h ..]
  for (int l = 3; l >= 0; --l) {
    float sb = 0.f;
#pragma unroll
    for (int r = 0; r < BK_RT; ++r) sb += acc[r];
    sb = pair_sum(sb);
    if (owner) atomicAdd(gw + (l == 3 ? BK_B3 : (l == 2 ? BK_B2 : (l == 1 ? BK_B1 : BK_B0))) + j, sb);
    const float* Kl = w + (l == 3 ? BK_K3 : (l == 2 ? BK_K2 : (l == 1 ? BK_K1 : BK_K0)));
    float* gKl = gw + (l == 3 ? BK_K3 : (l == 2 ? BK_K2 : (l == 1 ? BK_K1 : BK_K0)));
    if (l == 0) {
      wgrad_rows(acc, E + r0, BK_IN, j, owner, BK_W, gKl);                        // X_0 = encoding
      if (d_dirs == nullptr) break;
      float4* drow0 = reinterpret_cast<float4*>(D + j * BK_PITCH + r0);
#pragma unroll
      for (int r4 = 0; r4 < BK_RT / 4; ++r4) drow0[r4] = make_float4(acc[4 * r4], acc[4 * r4 + 1], acc[4 * r4 + 2], acc[4 * r4 + 3]);
      __syncthreads();
      if (j < BK_IN) {
        dense_accum_t(accE, Kl, BK_W, BK_W, j, D + r0);
        float* erow = Hs + j * BK_PITCH + r0;                                    // h0 is no longer needed
#pragma unroll
        for (int r = 0; r < BK_RT; ++r) erow[r] = accE[r];
      }
      __syncthreads();
      // chain rule through pos_enc(dir, 0, 4): d dir_c = dE[c] + sum_k 2^k (cos(2^k d_c) dE[3+3k+c] + cos(2^k d_c + pi/2) dE[15+3k+c])
      if (tid < BK_R * 3) {
        const int r = tid / 3, c = tid % 3;
        if (ray0 + r < n_rays) {
          const float x = __ldg(dirs + (ray0 + r) * dir_stride + c);
          float g = Hs[c * BK_PITCH + r], sc = 1.f;
          for (int k = 0; k < 4; ++k) {
            const float xb = mul(x, sc);
            g += sc * (cosf(xb) * Hs[(3 + 3 * k + c) * BK_PITCH + r] +
                       cosf(add(xb, 1.57079632679489661923f)) * Hs[(15 + 3 * k + c) * BK_PITCH + r]);
            sc *= 2.f;
          }
          d_dirs[(ray0 + r) * 3 + c] = g;
        }
      }
      break;
    }
    wgrad_rows(acc, Hs + (l - 1) * BK_W * BK_PITCH + r0, BK_W, j, owner, BK_W, gKl);          // X_l = h_{l-1} ...
    if (l == 3) wgrad_rows(acc, E + r0, BK_IN, j, owner, BK_W, gKl + BK_W * BK_W);            // ... and the skip-concatenated encoding
    // publish dz_l, then dh_{l-1}[i][r] = sum_j K_l[i][j] dz_l[j][r]
    float4* drow = reinterpret_cast<float4*>(D + j * BK_PITCH + r0);
#pragma unroll
    for (int r4 = 0; r4 < BK_RT / 4; ++r4) drow[r4] = make_float4(acc[4 * r4], acc[4 * r4 + 1], acc[4 * r4 + 2], acc[4 * r4 + 3]);
    __syncthreads();
    if (l == 3 && d_dirs != nullptr && j < BK_IN) dense_accum_t(accE, Kl + BK_W * BK_W, BK_W, BK_W, j, D + r0);   // skip-concat rows
#pragma unroll
    for (int r = 0; r < BK_RT; ++r) acc[r] = 0.f;
    dense_accum_t(acc, Kl, BK_W, BK_W, j, D + r0);
    const float* hprev = Hs + ((l - 1) * BK_W + j) * BK_PITCH + r0;
#pragma unroll
    for (int r = 0; r < BK_RT; ++r) acc[r] = hprev[r] > 0.f ? acc[r] : 0.f;
    __syncthreads();      // everyone has read D before the next layer overwrites it
  }
}

}  // namespace rnerf

using namespace rnerf;

extern "C" size_t rnerf_bkgd_weight_floats(void) { return (size_t)BK_TOTAL; }

extern "C" int rnerf_bkgd_mlp_fwd(const float* w, const float* dirs, int64_t n_rays, int64_t dir_stride_floats,
                                  float* raw_out, void* stream) {
  RNERF_REQUIRE(n_rays >= 0 && dir_stride_floats >= 3, RNERF_E_SHAPE, "rnerf_bkgd_mlp_fwd: bad sizes");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(w); RNERF_REQUIRE_PTR(dirs); RNERF_REQUIRE_PTR(raw_out);
  bkgd_mlp_kernel<<<(unsigned)((n_rays + BK_R - 1) / BK_R), BK_THREADS, 0, (cudaStream_t)stream>>>(w, dirs, n_rays,
                                                                                            dir_stride_floats, raw_out);
  count_launch();
  return check_launch("rnerf_bkgd_mlp_fwd");
}

static int bkgd_bwd_impl(const float* w, const float* dirs, int64_t n_rays, int64_t dir_stride_floats, const float* d_raw,
                         float* gw, float* d_dirs, void* stream) {
  RNERF_REQUIRE(n_rays >= 0 && dir_stride_floats >= 3, RNERF_E_SHAPE, "rnerf_bkgd_mlp_bwd: bad sizes");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(w); RNERF_REQUIRE_PTR(dirs); RNERF_REQUIRE_PTR(d_raw); RNERF_REQUIRE_PTR(gw);
  const size_t smem = (size_t)(BK_IN + 5 * BK_W) * BK_PITCH * sizeof(float);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(bkgd_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("rnerf_bkgd_mlp_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set[dev] = true;
  }
  bkgd_mlp_bwd_kernel<<<(unsigned)((n_rays + BK_R - 1) / BK_R), BK_THREADS, smem, (cudaStream_t)stream>>>(w, dirs, n_rays,
                                                                                                   dir_stride_floats, d_raw, gw, d_dirs);
  count_launch();
  return check_launch("rnerf_bkgd_mlp_bwd");
}

extern "C" int rnerf_bkgd_mlp_bwd(const float* w, const float* dirs, int64_t n_rays, int64_t dir_stride_floats,
                                  const float* d_raw, float* gw, void* stream) {
  return bkgd_bwd_impl(w, dirs, n_rays, dir_stride_floats, d_raw, gw, nullptr, stream);
}

extern "C" int rnerf_bkgd_mlp_bwd_dirs(const float* w, const float* dirs, int64_t n_rays, int64_t dir_stride_floats,
                                       const float* d_raw, float* gw, float* d_dirs, void* stream) {
  RNERF_REQUIRE_PTR(d_dirs);
  return bkgd_bwd_impl(w, dirs, n_rays, dir_stride_floats, d_raw, gw, d_dirs, stream);
}
