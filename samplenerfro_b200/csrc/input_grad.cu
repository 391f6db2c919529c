// Gradient of pos_enc + NerfMLP with respect to its INPUTS (sample positions and directions): the edge through which the
// "all"-stage loss reaches the bent path (train.py:164 differentiates rnerf/models.py:243-244 -> pos_enc
// rnerf/model_utils.py:187-214 -> NerfMLP rnerf/model_utils.py:30-90).
//
// The encoded position enters Dense_0 and, by the skip concat, rows 256..318 of Dense_5; the encoded direction enters rows
// 256..282 of Dense_10.  With dZ_l (gradient wrt layer l's pre-activation, bf16, written by rnerf_mlp_dgrad):
//     dEnc_pos [M][63] = dZ_0 K_0^T + dZ_5 K_5[256:319]^T          dEnc_dir [M][27] = dZ_10 K_10[256:283]^T
//     d pos_c = dEnc[c] + sum_k 2^k (cos(2^k x_c) dEnc[3 + 3k + c] + cos(2^k x_c + pi/2) dEnc[33 + 3k + c])      (k < 10)
//     d dir_c likewise with 4 octaves (features 3 + 3k + c and 15 + 3k + c).
// wt is the transposed weight image built by rnerf_mlp_input_grad_pack: [640][64] fp32, rows 0..255 = K_0^T (feature
// columns padded to 64 with zeros), 256..511 = K_5[256:319]^T, 512..639 = K_10[256:283]^T.
//
// 64 samples per CTA, 256 threads, a 4 x 4 register tile per thread over k-chunks of 32 staged in shared memory
// (dZ transposed to [k][sample], weights [k][feature]); fp32 on the CUDA cores: 2 * 640 * 64 flop per sample, 7 % of
// the MLP forward, run only in the "all" stage.
#include <cuda_bf16.h>
#include "common.cuh"

namespace rnerf {

constexpr int IG_TS = 64;                  // samples per CTA
constexpr int IG_KC = 32;                  // k per chunk
constexpr int IG_AP = IG_TS + 4;           // pitch of the transposed dZ chunk
constexpr int IG_WT_ROWS = 640, IG_F = 64;

__global__ void __launch_bounds__(256) input_grad_pack_kernel(const float* __restrict__ k0, const float* __restrict__ k5,
                                                              const float* __restrict__ k10, float* __restrict__ wt) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= IG_WT_ROWS * IG_F) return;
  const int row = i / IG_F, f = i % IG_F;
  float v = 0.f;
  if (row < 256)      { if (f < 63) v = __ldg(k0 + f * 256 + row); }                 // K_0 [63][256]
  else if (row < 512) { if (f < 63) v = __ldg(k5 + (256 + f) * 256 + (row - 256)); } // K_5 [319][256]
  else                { if (f < 27) v = __ldg(k10 + (256 + f) * 128 + (row - 512)); } // K_10 [283][128]
  wt[i] = v;
}

__global__ void __launch_bounds__(256) mlp_input_grad_kernel(const __nv_bfloat16* __restrict__ dz, int64_t n_samples,
                                                             const float* __restrict__ wt, const float* __restrict__ pos,
                                                             const float* __restrict__ dirs, float* __restrict__ d_pos,
                                                             float* __restrict__ d_dirs) {
  __shared__ __align__(16) float A[IG_KC * IG_AP];       // dZ chunk, [k][sample]
  __shared__ __align__(16) float Wc[IG_KC * IG_F];       // weight chunk, [k][feature]
  __shared__ float E[IG_TS * (IG_F + 1)];                // dEnc tile, [sample][feature]
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;                // samples 4ty..4ty+3, features 4tx..4tx+3
  const int64_t s0 = blockIdx.x * (int64_t)IG_TS;
  const size_t layer_stride = (size_t)n_samples * 256;
  const int ld_s = tid >> 2, ld_p = tid & 3;             // loader: sample, 8-element piece of the 32-k chunk
  const int64_t ld_row = min(s0 + ld_s, n_samples - 1);
  float acc[4][4];
  auto run_segment = [&](const __nv_bfloat16* src, int k_total, int wt_row0) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < k_total; k0 += IG_KC) {
      __syncthreads();                                   // previous chunk fully consumed
      {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + (size_t)ld_row * 256 + k0 + 8 * ld_p));
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(h2[q]);
          A[(8 * ld_p + 2 * q) * IG_AP + ld_s] = f.x;
          A[(8 * ld_p + 2 * q + 1) * IG_AP + ld_s] = f.y;
        }
        const float4* wsrc = reinterpret_cast<const float4*>(wt + (size_t)(wt_row0 + k0) * IG_F);
        float4* wdst = reinterpret_cast<float4*>(Wc);
        wdst[tid] = __ldg(wsrc + tid);
        wdst[tid + 256] = __ldg(wsrc + tid + 256);
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < IG_KC; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(A + k * IG_AP + 4 * ty);
        const float4 w = *reinterpret_cast<const float4*>(Wc + k * IG_F + 4 * tx);
        acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]); acc[0][2] = fmaf(a.x, w.z, acc[0][2]); acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
        acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]); acc[1][2] = fmaf(a.y, w.z, acc[1][2]); acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
        acc[2][0] = fmaf(a.z, w.x, acc[2][0]); acc[2][1] = fmaf(a.z, w.y, acc[2][1]); acc[2][2] = fmaf(a.z, w.z, acc[2][2]); acc[2][3] = fmaf(a.z, w.w, acc[2][3]);
        acc[3][0] = fmaf(a.w, w.x, acc[3][0]); acc[3][1] = fmaf(a.w, w.y, acc[3][1]); acc[3][2] = fmaf(a.w, w.z, acc[3][2]); acc[3][3] = fmaf(a.w, w.w, acc[3][3]);
      }
    }
  };
  auto publish = [&](bool accumulate) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float* e = E + (4 * ty + a) * (IG_F + 1) + 4 * tx + b;
        *e = accumulate ? *e + acc[a][b] : acc[a][b];
      }
  };
  // chain rule through pos_enc for one (sample, component): n_oct octaves, cos block starts at 3 + 3 n_oct
  auto enc_bwd = [&](const float* x_in, float* out, int n_oct) {
    if (tid < IG_TS * 3) {
      const int s = tid / 3, c = tid % 3;
      if (s0 + s < n_samples) {
        const float x = __ldg(x_in + (s0 + s) * 3 + c);
        const float* e = E + s * (IG_F + 1);
        float g = e[c], sc = 1.f;
        for (int k = 0; k < n_oct; ++k) {
          const float xb = mul(x, sc);
          g += sc * (cosf(xb) * e[3 + 3 * k + c] + cosf(add(xb, 1.57079632679489661923f)) * e[3 + 3 * n_oct + 3 * k + c]);
          sc *= 2.f;
        }
        out[(s0 + s) * 3 + c] = g;
      }
    }
  };
  run_segment(dz, 256, 0);                               // dZ_0 K_0^T
  publish(false);                                        // own elements only: no barrier needed before the next segment
  run_segment(dz + 5 * layer_stride, 256, 256);          // + dZ_5 K_5[256:]^T
  publish(true);
  __syncthreads();
  enc_bwd(pos, d_pos, 10);
  run_segment(dz + 9 * layer_stride, 128, 512);          // dZ_10 K_10[256:]^T   (first barrier inside orders the E reads)
  publish(false);
  __syncthreads();
  enc_bwd(dirs, d_dirs, 4);
}

}  // namespace rnerf

using namespace rnerf;

extern "C" size_t rnerf_mlp_input_grad_packed_floats(void) { return (size_t)IG_WT_ROWS * IG_F; }

extern "C" int rnerf_mlp_input_grad_pack(const float* k0, const float* k5, const float* k10, float* wt, void* stream) {
  RNERF_REQUIRE_PTR(k0); RNERF_REQUIRE_PTR(k5); RNERF_REQUIRE_PTR(k10); RNERF_REQUIRE_PTR(wt);
  input_grad_pack_kernel<<<(IG_WT_ROWS * IG_F + 255) / 256, 256, 0, (cudaStream_t)stream>>>(k0, k5, k10, wt);
  count_launch();
  return check_launch("rnerf_mlp_input_grad_pack");
}

extern "C" int rnerf_mlp_input_grad(const uint16_t* dz, int64_t n_samples, const float* wt, const float* pos, const float* dirs,
                                    float* d_pos, float* d_dirs, void* stream) {
  RNERF_REQUIRE(n_samples >= 0, RNERF_E_SHAPE, "rnerf_mlp_input_grad: n_samples < 0");
  if (n_samples == 0) return 0;
  RNERF_REQUIRE_PTR(dz); RNERF_REQUIRE_PTR(wt); RNERF_REQUIRE_PTR(pos); RNERF_REQUIRE_PTR(dirs); RNERF_REQUIRE_PTR(d_pos); RNERF_REQUIRE_PTR(d_dirs);
  RNERF_REQUIRE(aligned16(dz) && aligned16(wt), RNERF_E_ALIGN, "rnerf_mlp_input_grad: dz/wt must be 16-byte aligned");
  const unsigned blocks = (unsigned)((n_samples + IG_TS - 1) / IG_TS);
  mlp_input_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dz, n_samples, wt, pos, dirs, d_pos, d_dirs);
  count_launch();
  return check_launch("rnerf_mlp_input_grad");
}
