// Geometry of the dgrad chain and the argument block shared by its two kernels (mlp_bwd.cu: one CTA per tile group,
// mlp_dgrad_pair.cu: CTA pairs).  See mlp_bwd.cu for the layer numbering.
#pragma once
#include "umma.cuh"

namespace rnerf {

constexpr int DG_GEMMS = 9;                       // d = 0: Dense_10[:256]^T (K = 128), d = 1: Dense_9^T, d >= 2: Dense_(9-d)^T
__host__ __device__ constexpr int dg_dense(int d) { return d == 0 ? 10 : (d == 1 ? 9 : 9 - d); }
__host__ __device__ constexpr int dg_k(int d) { return d == 0 ? 128 : 256; }           // GEMM K = width of the layer's output
__host__ __device__ constexpr int dg_chunks(int d) { return dg_k(d) / KCH; }           // [256 x 32] SWIZZLE_64B chunks
constexpr int DG_NCHUNK = 4 + 8 * 8;              // 68
constexpr size_t DG_PACKED_BYTES = (size_t)DG_NCHUNK * SLOT_BYTES;
// second image for the CTA-pair kernel: every 64-wide k-block of a GEMM is one chunk of two N-halves ([128 x 64] bf16
// SWIZZLE_128B, 16 KB each, one per CTA of the pair) -- 2 chunks for d = 0, 4 for every other GEMM
constexpr int DGP_NCHUNK = 2 + 8 * 4;             // 34
constexpr size_t DG_PAIR_OFF = DG_PACKED_BYTES;
constexpr size_t DG_TOTAL_BYTES = DG_PAIR_OFF + (size_t)DGP_NCHUNK * PAIR_CHUNK_STRIDE;

struct DgradArgs {
  const uint8_t* packed;        // dgrad weight image (dgrad_pack_kernel)
  const float* head_w;          // w_sigma[256] then w_rgb[3][128], bf16-rounded fp32 (forward image tail)
  const uint32_t* masks;        // [10][M][8] ReLU bit-masks written by the training forward (relu_mask32, umma.cuh)
  const float4* d_raw;          // [M] (d rgb_raw[3], d sigma_raw)
  __nv_bfloat16* dZ;            // [10][M][256] out
  int64_t n_samples;
  int n_groups;
};

// CTA-pair version (mlp_dgrad_pair.cu); same arguments, same outputs
int launch_mlp_dgrad_pair(const DgradArgs& a, const CUtensorMap& tm_dz, cudaStream_t st);

}  // namespace rnerf
