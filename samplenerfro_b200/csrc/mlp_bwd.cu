// Backward of the radiance MLP (a17: train.py:164-165 differentiates NerfMLP, rnerf/model_utils.py:30-90).
//
//   mlp_dgrad_kernel  fused chain of data gradients on tcgen05/TMEM: for a 128-row tile the gradient wrt every
//                     layer's pre-activation (dZ_l) is produced layer by layer, dZ_l -> (x W_l^T) -> ReLU mask ->
//                     dZ_{l-1}, with the bf16 A operand living in shared memory exactly like the forward; each dZ_l
//                     is also streamed to HBM once (bf16) for the weight-gradient pass.  ReLU masks come from the
//                     forward's saved activations, prefetched as 256-bit row masks while the MMAs run.
//   mlp_wgrad_kernel  dW_l = X_l^T dZ_l, reduction over the sample rows: both operands are read in their natural
//                     row-major layout as MN-major UMMA operands; each CTA owns a slab of rows and accumulates the
//                     whole [K_l x N_l] block in TMEM (2 x 256 columns), then adds it to the fp32 gradient with
//                     red.global.add; the bias gradient (column sums of dZ_l) is taken from the staged tiles.
//   mlp_head_grad_kernel  the two skinny heads (Dense_8: 256->1, Dense_11: 128->3) on CUDA cores.
//
// Layer numbering: "MMA layer" l = 0..9 as in the forward (Dense_0..7, Dense_9 bottleneck, Dense_10 condition);
// H[l] = saved post-activation output of MMA layer l ([M][256] bf16), dZ[l] = gradient wrt its pre-activation.
#include <stdlib.h>
#include "mlp_bwd.cuh"

namespace rnerf {

// ------------------------------------------------------------------------------------------------
// dgrad chain
// ------------------------------------------------------------------------------------------------
struct DgradPackArgs { const float* kern[12]; };

// chunk c of GEMM d: rows r = input feature kin (0..255), 32 columns = output features n0..n0+31 of the layer:
// B[kin][n] = W[kin][n], i.e. plain row slices of the Flax [in,out] kernel (only its first 256 rows matter).
__global__ void __launch_bounds__(256) dgrad_pack_kernel(DgradPackArgs a, uint8_t* __restrict__ packed) {
  if (blockIdx.x >= DG_NCHUNK) {
    // pair image: chunk pc = k-block kb of GEMM d, N-half h holds rows kin = 128 h + r
    int kb = (int)blockIdx.x - DG_NCHUNK, d = 0;
    while (kb >= dg_k(d) / KB) { kb -= dg_k(d) / KB; ++d; }
    const int out_dim = dg_k(d);
    const float* W = a.kern[dg_dense(d)];
    for (int h = 0; h < 2; ++h) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(packed + DG_PAIR_OFF + (size_t)(blockIdx.x - DG_NCHUNK) * PAIR_CHUNK_STRIDE +
                                                            h * PAIR_HALF_BYTES);
      for (int e = threadIdx.x; e < 128 * KB; e += blockDim.x) {
        const int r = e / KB, k = e % KB;
        dst[sw128_offset(r, k) / 2] = __float2bfloat16_rn(W[(size_t)(h * 128 + r) * out_dim + kb * KB + k]);
      }
    }
    return;
  }
  int c = blockIdx.x, d = 0;
  while (c >= dg_chunks(d)) { c -= dg_chunks(d); ++d; }
  const int out_dim = dg_k(d);
  const float* W = a.kern[dg_dense(d)];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(packed + (size_t)blockIdx.x * SLOT_BYTES);
  for (int e = threadIdx.x; e < NMAX * KCH; e += blockDim.x) {
    const int r = e / KCH, k = e % KCH;
    dst[sw64_offset(r, k) / 2] = __float2bfloat16_rn(W[(size_t)r * out_dim + c * KCH + k]);
  }
}

template <int NT, int NSTAGE>
struct DgradSmem {
  static constexpr uint32_t A_OFF = 0;                                   // [NT][4][16 KB] dZ k-blocks
  static constexpr uint32_t W_OFF = A_OFF + NT * 4 * ABLK_BYTES;         // [NSTAGE][16 KB] weight ring
  static constexpr uint32_t H_OFF = W_OFF + NSTAGE * SLOT_BYTES;         // w_sigma[256] + w_rgb[3][128] fp32
  static constexpr uint32_t BAR_OFF = H_OFF + (256 + 384) * 4;
  static constexpr uint32_t N_BARS = 2 * NSTAGE + 2;
  static constexpr uint32_t TMEM_SLOT = BAR_OFF + N_BARS * 8;
  static constexpr uint32_t BYTES = TMEM_SLOT + 16;
};

template <int NT, int NSTAGE>
__global__ void __launch_bounds__(64 + 128 * NT, 1) mlp_dgrad_kernel(const DgradArgs args,
                                                                     const __grid_constant__ CUtensorMap tm_dz) {
  using SL = DgradSmem<NT, NSTAGE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar_full = [&](int s) { return sbase + SL::BAR_OFF + 8u * s; };
  auto bar_empty = [&](int s) { return sbase + SL::BAR_OFF + 8u * (NSTAGE + s); };
  const uint32_t bar_acc = sbase + SL::BAR_OFF + 8u * (2 * NSTAGE);
  const uint32_t bar_aready = sbase + SL::BAR_OFF + 8u * (2 * NSTAGE + 1);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SL::TMEM_SLOT);
  float* head_s = reinterpret_cast<float*>(smem + SL::H_OFF);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_aready, 4 * NT);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(sbase + SL::TMEM_SLOT, 256 * NT); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 256 + 384; i += blockDim.x) head_s[i] = __ldg(args.head_w + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_groups = (args.n_groups > (int)blockIdx.x) ? (args.n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const size_t layer_stride = (size_t)args.n_samples * 256;

  if (warp == 0) {
    if (lane == 0) {   // weight producer
      int stage = 0; uint32_t phase = 0;
      for (int g = 0; g < my_groups; ++g)
        for (int c = 0; c < DG_NCHUNK; ++c) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          mbar_arrive_expect_tx(bar_full(stage), SLOT_BYTES);
          tma_bulk_g2s(sbase + SL::W_OFF + stage * SLOT_BYTES, args.packed + (size_t)c * SLOT_BYTES, SLOT_BYTES, bar_full(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {   // MMA issuer
      constexpr uint32_t idesc = make_idesc(TILE_M, 256);
      int stage = 0; uint32_t phase = 0, ar_phase = 0;
      for (int g = 0; g < my_groups; ++g)
        for (int d = 0; d < DG_GEMMS; ++d) {
          mbar_wait(bar_aready, ar_phase); ar_phase ^= 1;
          tc_fence_after();
          const int nkb = dg_k(d) / KB;
          for (int kbi = 0; kbi < nkb; ++kbi)
#pragma unroll
            for (int sub = 0; sub < SUBS; ++sub) {
              mbar_wait(bar_full(stage), phase);
              tc_fence_after();
              const uint32_t b_addr = sbase + SL::W_OFF + stage * SLOT_BYTES;
#pragma unroll
              for (int t = 0; t < NT; ++t) {
                const uint32_t a_addr = sbase + SL::A_OFF + (t * 4 + kbi) * ABLK_BYTES + sub * (KCH * 2);
                const uint64_t ad = make_sw128_desc(a_addr), bd = make_sw64_desc(b_addr);
#pragma unroll
                for (int ks = 0; ks < KCH / 16; ++ks)     // +32 bytes of K per step = +2 in the descriptors' address fields
                  umma_bf16_lohi(tmem_base + (uint32_t)(t * 256), (uint32_t)ad + 2u * ks, (uint32_t)(ad >> 32), (uint32_t)bd + 2u * ks,
                                 (uint32_t)(bd >> 32), idesc, (kbi > 0 || sub > 0 || ks > 0) ? 1u : 0u);
              }
              umma_commit(bar_empty(stage));
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
          umma_commit(bar_acc);
        }
    }
  } else {
    // epilogue warpgroups: thread <-> sample row
    const int t = (warp - 2) >> 2, q = warp & 3, row = q * 32 + lane;
    uint8_t* a_row = smem + SL::A_OFF + t * 4 * ABLK_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
    const uint32_t r7s = (uint32_t)(row & 7) << 4;
    const float* w_sigma = head_s;
    const float* w_rgb = head_s + 256;
    uint32_t acc_phase = 0;
    // Every dZ tile is left in shared memory (it is the next GEMM's A operand) and streamed to HBM from there by the
    // TMA engine: 4 tensor stores per warp and layer (one per 64-column k-block of the warp's 32 rows).
    const uint32_t a_warp = sbase + SL::A_OFF + t * 4 * ABLK_BYTES + q * 4096;
    auto store_tile = [&](int64_t wrow0, int layer, int nkb) {     // lane 0 only, after fence + __syncwarp
      if (wrow0 < args.n_samples) {
        for (int kb = 0; kb < nkb; ++kb) tma_store_3d(&tm_dz, a_warp + kb * ABLK_BYTES, kb * KB, (int)wrow0, layer);
        bulk_commit();
      }
    };
    auto tile_free = [&]() {                                       // previous stores have finished reading the tile
      if (lane == 0) bulk_wait_read();
      __syncwarp();
    };
    for (int g = 0; g < my_groups; ++g) {
      const int64_t group = (int64_t)blockIdx.x + (int64_t)g * gridDim.x;
      const int64_t wrow0 = (group * NT + t) * TILE_M + q * 32;
      const int64_t srow = wrow0 + lane;
      const bool live = srow < args.n_samples;
      const int64_t lrow = live ? srow : (args.n_samples - 1);
      const float4 draw = live ? __ldg(args.d_raw + lrow) : make_float4(0.f, 0.f, 0.f, 0.f);
      // ---- prologue: rgb head (Dense_11) backward + ReLU of the condition layer -> dZ[9] (128 columns)
      {
        uint32_t m9[4];                           // condition layer: 128 columns = words 0..3 (one per 32-column group)
        {
          const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(args.masks + ((size_t)9 * args.n_samples + lrow) * 8));
          m9[0] = a4.x; m9[1] = a4.y; m9[2] = a4.z; m9[3] = a4.w;
        }
        tile_free();
#pragma unroll 1
        for (int c8 = 0; c8 < 16; ++c8) {         // 8 columns per 16-byte unit
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j0 = c8 * 8 + i * 2;
            float g0 = draw.x * w_rgb[j0] + draw.y * w_rgb[128 + j0] + draw.z * w_rgb[256 + j0];
            float g1 = draw.x * w_rgb[j0 + 1] + draw.y * w_rgb[128 + j0 + 1] + draw.z * w_rgb[256 + j0 + 1];
            // columns 8 c8 + 2 i, + 1: pair (c8 & 3) * 4 + i of group c8 >> 2 (relu_mask32)
            const uint32_t word = (c8 >> 2) == 0 ? m9[0] : ((c8 >> 2) == 1 ? m9[1] : ((c8 >> 2) == 2 ? m9[2] : m9[3]));
            if (!relu_mask_even(word, (c8 & 3) * 4 + i)) g0 = 0.f;
            if (!relu_mask_odd(word, (c8 & 3) * 4 + i)) g1 = 0.f;
            pk[i] = pack_bf16(g0, g1);
          }
          *reinterpret_cast<uint4*>(a_row + (c8 >> 3) * ABLK_BYTES + ((uint32_t)((c8 & 7) << 4) ^ r7s)) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_aready); store_tile(wrow0, 9, 2); }
      for (int d = 0; d < DG_GEMMS; ++d) {
        const int lo = 8 - d;                       // MMA layer whose dZ this GEMM produces
        uint32_t mask[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};       // this row's ReLU bits of layer lo: columns 0..127 | 128..255
        if (d >= 1) {                                            // 32 bytes per row, hidden under the MMAs
          const uint4* mp = reinterpret_cast<const uint4*>(args.masks + ((size_t)lo * args.n_samples + lrow) * 8);
          const uint4 a4 = __ldg(mp), b4 = __ldg(mp + 1);
          mask[0][0] = a4.x; mask[0][1] = a4.y; mask[0][2] = a4.z; mask[0][3] = a4.w;
          mask[1][0] = b4.x; mask[1][1] = b4.y; mask[1][2] = b4.z; mask[1][3] = b4.w;
        }
        mbar_wait(bar_acc, acc_phase); acc_phase ^= 1;
        tc_fence_after();
        tile_free();
#pragma unroll 1
        for (int cg = 0; cg < 8; ++cg) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + cg * 32), v);
          tmem_ld_wait();
          uint32_t mk = 0u;                         // this 32-column group's word
#pragma unroll
          for (int p = 0; p < 8; ++p) mk = (p == cg) ? mask[p >> 2][p & 3] : mk;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float f0 = __uint_as_float(v[2 * j]), f1 = __uint_as_float(v[2 * j + 1]);
            if (d == 1) {   // the sigma head (Dense_8) also feeds h7: dh7 += d_sigma * w_sigma
              f0 = fmaf(draw.w, w_sigma[cg * 32 + 2 * j], f0);
              f1 = fmaf(draw.w, w_sigma[cg * 32 + 2 * j + 1], f1);
            }
            if (d >= 1) {                          // columns 32 cg + 2 j, + 1: bits j and 16 + j of the group's word
              if (!relu_mask_even(mk, j)) f0 = 0.f;
              if (!relu_mask_odd(mk, j)) f1 = 0.f;
            }
            pk[j] = pack_bf16(f0, f1);
          }
          uint8_t* blk = a_row + (cg >> 1) * ABLK_BYTES;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(blk + ((uint32_t)(((cg & 1) * 4 + c) << 4) ^ r7s)) =
                make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (d < DG_GEMMS - 1) mbar_arrive(bar_aready);
          store_tile(wrow0, lo, 4);
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256 * NT);
}

// ------------------------------------------------------------------------------------------------
// heads + bias gradients of the skinny layers (CUDA cores)
//   gW11[j][c] = sum_m H9[m][j] d_rgb[m][c]   gb11[c] = sum_m d_rgb[m][c]
//   gW8[j]     = sum_m H7[m][j] d_sig[m]      gb8     = sum_m d_sig[m]
// out_rgb: float[128*3 + 3] = (gW11, gb11), out_sig: float[256 + 1] = (gW8, gb8) -- the Flax (kernel, bias) pairs of
// Dense_11 / Dense_8 as they sit in a flat parameter arena; accumulated with atomics (zeroed by the caller)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mlp_head_grad_kernel(const __nv_bfloat16* __restrict__ H, const float4* __restrict__ d_raw,
                                                            int64_t n_samples, int rows_per_block,
                                                            float* __restrict__ out, float* __restrict__ out_sig) {
  // 256 threads = 4 row streams x 64 column quads: thread (s, c) reads columns 4c .. 4c+3 of H7 (and of H9 when c < 32) with
  // one 8-byte load per row -- a warp covers 256 contiguous bytes of a row -- rows s, s+4, ... of a slab, four rows in flight.
  // The grid is persistent (slabs are taken round-robin): the 644 atomics per block are paid ~600 times per launch, not once
  // per 512 rows, and there is no partial last wave.  (The first version, 4-byte loads and one slab per block, ran at 1.8 TB/s
  // with 40 % of the issue slots busy and 2.08 waves: profiles/r5d_train_backward_ncu_summary.txt.)
  const size_t layer_stride = (size_t)n_samples * 256;
  const int c = threadIdx.x & 63, s = threadIdx.x >> 6;
  const uint2* H7 = reinterpret_cast<const uint2*>(H + 7 * layer_stride) + c;     // + m * 64 per row
  const uint2* H9 = reinterpret_cast<const uint2*>(H + 9 * layer_stride) + c;
  float sg[4] = {0.f, 0.f, 0.f, 0.f};
  float rgb[4][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  float sb[4] = {0.f, 0.f, 0.f, 0.f};          // bias gradients (c == 0 only)
  constexpr int U = 4;
  for (int64_t m0 = (int64_t)blockIdx.x * rows_per_block; m0 < n_samples; m0 += (int64_t)gridDim.x * rows_per_block) {
    const int64_t m1 = min(m0 + rows_per_block, n_samples);
    for (int64_t m = m0 + s; m < m1; m += 4 * U) {
      uint2 h7[U], h9[U];
      float4 d[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t mm = m + 4 * u;
        const bool ok = mm < m1;
        const int64_t mc = ok ? mm : m;
        h7[u] = __ldg(H7 + (size_t)mc * 64);
        h9[u] = c < 32 ? __ldg(H9 + (size_t)mc * 64) : make_uint2(0u, 0u);
        d[u] = ok ? __ldg(d_raw + mc) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        sg[0] = fmaf(bf16_lo(h7[u].x), d[u].w, sg[0]); sg[1] = fmaf(bf16_hi(h7[u].x), d[u].w, sg[1]);
        sg[2] = fmaf(bf16_lo(h7[u].y), d[u].w, sg[2]); sg[3] = fmaf(bf16_hi(h7[u].y), d[u].w, sg[3]);
        const float a[4] = {bf16_lo(h9[u].x), bf16_hi(h9[u].x), bf16_lo(h9[u].y), bf16_hi(h9[u].y)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          rgb[k][0] = fmaf(a[k], d[u].x, rgb[k][0]); rgb[k][1] = fmaf(a[k], d[u].y, rgb[k][1]); rgb[k][2] = fmaf(a[k], d[u].z, rgb[k][2]);
        }
        if (c == 0) { sb[0] += d[u].x; sb[1] += d[u].y; sb[2] += d[u].z; sb[3] += d[u].w; }
      }
    }
  }
  // fold the four row streams in shared memory, then one atomic per output element and block
  __shared__ float red[3][64][16 + 4];
  if (s > 0) {
    float* r = red[s - 1][c];
#pragma unroll
    for (int k = 0; k < 4; ++k) { r[k] = sg[k]; r[4 + 3 * k] = rgb[k][0]; r[5 + 3 * k] = rgb[k][1]; r[6 + 3 * k] = rgb[k][2]; r[16 + k] = sb[k]; }
  }
  __syncthreads();
  if (s == 0) {
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const float* r = red[w][c];
#pragma unroll
      for (int k = 0; k < 4; ++k) { sg[k] += r[k]; rgb[k][0] += r[4 + 3 * k]; rgb[k][1] += r[5 + 3 * k]; rgb[k][2] += r[6 + 3 * k]; sb[k] += r[16 + k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(out_sig + 4 * c + k, sg[k]);
    if (c < 32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(out + (4 * c + k) * 3, rgb[k][0]); atomicAdd(out + (4 * c + k) * 3 + 1, rgb[k][1]); atomicAdd(out + (4 * c + k) * 3 + 2, rgb[k][2]);
      }
    }
    if (c == 0) { atomicAdd(out + 384, sb[0]); atomicAdd(out + 385, sb[1]); atomicAdd(out + 386, sb[2]); atomicAdd(out_sig + 256, sb[3]); }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: G[Kx x N] += X[rows][Kx]^T  dZ[rows][N]   (MN-major operands, reduction over rows), gb[N] += colsum(dZ)
// One CTA = one slab of rows; accumulators: Kx/128 M-blocks x N columns of TMEM.
// ------------------------------------------------------------------------------------------------
// Stage geometry, measured on the 4096-ray step [r6k]: 64 rows x 3 stages, 2 in flight 6.13 ms/step; 32 x 6, 4 in flight (the
// first version) 6.43; 48 x 4 6.28; 96 x 2 6.34; 16 x 12 7.35 -- every stage costs two named barriers and a bias pass over
// the tile, so few large stages win as long as one whole stage is in flight behind the one being published.
#ifndef RNERF_WG_ROWS          // (overridable for A/B builds: RNERF_NVCC_EXTRA="-DRNERF_WG_ROWS=32 -DRNERF_WG_STAGES=6 -DRNERF_WG_DEPTH=4")
#define RNERF_WG_ROWS 64
#define RNERF_WG_STAGES 3
#define RNERF_WG_DEPTH 2
#endif
#ifndef RNERF_WG_BIG_ROWS      // rows per CTA and job above which every job gets all the SMs (see rnerf_mlp_wgrad_batched)
#define RNERF_WG_BIG_ROWS 1024
#endif
constexpr int WG_ROWS = RNERF_WG_ROWS;                // rows (GEMM K) per pipeline stage
constexpr int WG_STAGES = RNERF_WG_STAGES;
constexpr int WG_DEPTH = RNERF_WG_DEPTH;              // stages of loads kept in flight (128 KB per SM: the kernel is HBM-bound)
constexpr int WG_STAGE_BYTES = WG_ROWS * 256 * 2 * 2; // X tile [32 x 256] + dZ tile [32 x 256] bf16 = 32 KB
constexpr int WG_THREADS = 64 + 256;                  // warp 0: MMA, warp 1: spare, warps 2-9: loaders/epilogue

struct WgradSmem {
  static constexpr uint32_t ST_OFF = 0;
  static constexpr uint32_t BAR_OFF = WG_STAGES * WG_STAGE_BYTES;
  static constexpr uint32_t TMEM_SLOT = BAR_OFF + (2 * WG_STAGES + 1) * 8;
  static constexpr uint32_t BYTES = TMEM_SLOT + 16;
};

struct WgradArgs {
  const __nv_bfloat16* X;     // [M][ldx] rows; the first kx columns are used
  const __nv_bfloat16* dZ;    // [M][256]
  int ldx, kx, x_cols;        // kx in {128, 256}: gradient rows (UMMA M blocks of 128); x_cols <= kx real columns of X, rest zero
  int n;                      // 128 or 256 columns of dZ
  int64_t n_samples;
  int rows_per_cta;           // multiple of WG_ROWS
  float* gW;                  // [kx_valid][n] fp32, accumulated
  int kx_valid;               // rows of gW actually written (e.g. 63 of 64)
  float* gb;                  // [n] fp32 or null, accumulated
};

// MN-major SWIZZLE_128B descriptor: 64 MN-elements (128 B) contiguous per K-row, 8 K-rows per 1024-byte atom;
// LBO = byte distance between MN-atoms, SBO = byte distance between groups of 8 K-rows.
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {   // both operands MN-major (bits 15, 16)
  return make_idesc(m, n) | (1u << 15) | (1u << 16);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;     // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Tile layout in a stage: [operand (X, dZ)][k-group kb = 0..7][MN-atom mm = 0..3][8 rows x 128 B], i.e. for a row r
// (= GEMM K index) and column c: atom (kb = r/8, mm = c/64), row-in-atom kk = r%8, 16-byte unit (c%64)/8 ^ kk.
// A launch carries up to WG_MAX_JOBS independent weight-gradient GEMMs (all the layers of one MLP's backward): CTA b works
// on job j with cta0[j] <= b < cta0[j + 1].  One launch per MLP instead of one per layer: no launch gaps or per-launch tails
// between the layers, and for small batches (a 512-ray shard of an 8-GPU step) one wave of CTAs instead of 13 launches that
// each pay the fixed costs (TMEM allocation, the 256 KB accumulator flush) on every SM.
constexpr int WG_MAX_JOBS = 16;
struct WgradBatch {
  WgradArgs job[WG_MAX_JOBS];
  int cta0[WG_MAX_JOBS + 1];
  int n_jobs;
};

__global__ void __launch_bounds__(WG_THREADS, 1) mlp_wgrad_kernel(const __grid_constant__ WgradBatch batch) {
  using SL = WgradSmem;
  int jb = 0;
  while (jb + 1 < batch.n_jobs && (int)blockIdx.x >= batch.cta0[jb + 1]) ++jb;
  const WgradArgs& a = batch.job[jb];
  const int cta = (int)blockIdx.x - batch.cta0[jb];
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar_full = [&](int s) { return sbase + SL::BAR_OFF + 8u * s; };
  auto bar_empty = [&](int s) { return sbase + SL::BAR_OFF + 8u * (WG_STAGES + s); };
  const uint32_t bar_done = sbase + SL::BAR_OFF + 8u * (2 * WG_STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SL::TMEM_SLOT);
  const int n_mblk = a.kx / 128;                    // M-blocks of the gradient (1 or 2)
  const int m_rows = 128;                           // UMMA M
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(sbase + SL::TMEM_SLOT, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t row0 = (int64_t)cta * a.rows_per_cta;
  const int64_t row1 = min(row0 + a.rows_per_cta, a.n_samples);
  const int n_steps = row1 > row0 ? (int)((row1 - row0 + WG_ROWS - 1) / WG_ROWS) : 0;
  const int xatoms = a.kx / 64, zatoms = a.n / 64;  // MN-atoms per K-group

  if (warp == 0) {
    if (n_steps > 0 && elect_one_sync()) {
      const uint32_t idesc = make_idesc_mn(m_rows, a.n);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < n_steps; ++it) {
        mbar_wait(bar_full(stage), phase);
        tc_fence_after();
        const uint32_t xs = sbase + SL::ST_OFF + stage * WG_STAGE_BYTES;
        const uint32_t zs = xs + WG_ROWS * 256 * 2;
        // descriptors of the stage's first atoms; every MMA adds its byte offset >> 4 to the low words (umma_bf16_lohi)
        const uint64_t ad0 = make_mn_sw128_desc(xs, 1024, xatoms * 1024), bd0 = make_mn_sw128_desc(zs, 1024, zatoms * 1024);
        for (int mb = 0; mb < n_mblk; ++mb)
#pragma unroll
          for (int ks = 0; ks < WG_ROWS / 16; ++ks) {
            // K-step = 16 rows = 2 k-groups; A: MN-atoms 2*mb, 2*mb+1 of the X tile; B: all n/64 atoms of the dZ tile
            const uint32_t a_off = (uint32_t)(((2 * ks) * xatoms * 1024 + mb * 2 * 1024) >> 4);
            const uint32_t b_off = (uint32_t)(((2 * ks) * zatoms * 1024) >> 4);
            umma_bf16_lohi(tmem_base + (uint32_t)(mb * 256), (uint32_t)ad0 + a_off, (uint32_t)(ad0 >> 32), (uint32_t)bd0 + b_off,
                           (uint32_t)(bd0 >> 32), idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
        umma_commit(bar_empty(stage));
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(bar_done);
    }
  } else if (warp >= 2) {
    // ---- loaders: 256 threads copy one stage (rows x {X, dZ}) with 16-byte cp.async into the swizzled MN-major tiles.
    // Software pipeline: the copies of stage `it` are issued before the wait on stage `it-1`, so one stage of loads is
    // always in flight behind the one being published.
    const int tid = threadIdx.x - 64;
    float colsum = 0.f;                        // thread tid owns dZ column tid (if < n)
    // 16-byte units per row: 16 or 32 -- powers of two, so the per-unit index math below is shifts and masks
    const int xshift = a.kx == 256 ? 5 : 4, zshift = a.n == 256 ? 5 : 4;
    const int xunits = 1 << xshift, zunits = 1 << zshift;
    // Each loader owns one 16-byte unit column (u) of either operand and walks down the rows in steps of 256/units
    // (8 or 16): r & 7 is then fixed per thread, so the swizzled destination is a constant plus a per-row stride.
    const int xu = tid & (xunits - 1), xr0 = tid >> xshift, xrs = 256 >> xshift;
    const int zu = tid & (zunits - 1), zr0 = tid >> zshift, zrs = 256 >> zshift;
    const uint32_t xoff0 = (uint32_t)(((xr0 >> 3) * xatoms + (xu >> 3)) * 1024 + (xr0 & 7) * 128 + (((xu & 7) ^ (xr0 & 7)) << 4));
    const uint32_t zoff0 = (uint32_t)(((zr0 >> 3) * zatoms + (zu >> 3)) * 1024 + (zr0 & 7) * 128 + (((zu & 7) ^ (zr0 & 7)) << 4));
    const uint32_t xdstep = (uint32_t)((xrs >> 3) * xatoms * 1024), zdstep = (uint32_t)((zrs >> 3) * zatoms * 1024);
    const bool in_x = xu * 8 < a.x_cols;
    const int64_t last = a.n_samples - 1;
    auto issue = [&](int it, int stage) {
      const uint32_t xs = sbase + SL::ST_OFF + stage * WG_STAGE_BYTES;
      const uint32_t zs = xs + WG_ROWS * 256 * 2;
      const int64_t rbase = row0 + (int64_t)it * WG_ROWS;
      {
        uint32_t dst = xs + xoff0;
        const __nv_bfloat16* colp = a.X + (in_x ? xu * 8 : 0);
#pragma unroll 4
        for (int r = xr0; r < WG_ROWS; r += xrs, dst += xdstep) {
          const int64_t gr = rbase + r;
          cp_async16(dst, colp + (size_t)min(gr, last) * a.ldx, gr < row1 && in_x);
        }
      }
      {
        uint32_t dst = zs + zoff0;
        const __nv_bfloat16* colp = a.dZ + zu * 8;
#pragma unroll 4
        for (int r = zr0; r < WG_ROWS; r += zrs, dst += zdstep) {
          const int64_t gr = rbase + r;
          cp_async16(dst, colp + (size_t)min(gr, last) * 256, gr < row1);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // publish a completed stage: all 256 loaders' copies have landed (named barrier), bias gradient from the staged dZ
    // tile (column tid: element (r, tid) sits in atom (r/8, tid/64), row r%8, unit ((tid%64)/8) ^ (r%8)), then one arrive
    auto publish = [&](int stage) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (a.gb != nullptr && tid < a.n) {
        const uint8_t* zs = smem + SL::ST_OFF + stage * WG_STAGE_BYTES + WG_ROWS * 256 * 2;
        const int mm = tid >> 6, u = (tid & 63) >> 3, w = tid & 7;
#pragma unroll 8
        for (int r = 0; r < WG_ROWS; ++r) {
          const uint32_t off = (uint32_t)(((r >> 3) * zatoms + mm) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4) + w * 2);
          colsum += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(zs + off));
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 256;" ::: "memory");     // every loader has fenced its writes (and read its column)
      if (tid == 0) mbar_arrive(bar_full(stage));
    };
    // Software pipeline: stage `it` is issued WG_DEPTH - 1 stages ahead of the one being published, so WG_DEPTH - 1
    // stages of loads are always in flight behind it.
    auto wait_pending = [&](int n) {     // at most n of this thread's cp.async groups still pending
      switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
      }
    };
    static_assert(WG_DEPTH >= 2 && WG_DEPTH <= 4 && WG_DEPTH <= WG_STAGES && WG_ROWS % 16 == 0, "wait_pending covers depths up to 4");
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < n_steps; ++it) {
      mbar_wait(bar_empty(stage), phase ^ 1);
      issue(it, stage);
      if (it >= WG_DEPTH - 1) {
        wait_pending(WG_DEPTH - 1);                          // stage it - (WG_DEPTH-1) has landed
        publish((it - (WG_DEPTH - 1)) % WG_STAGES);
      }
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
    {
      const int pending = n_steps < WG_DEPTH - 1 ? n_steps : WG_DEPTH - 1;
      for (int j = 0; j < pending; ++j) {
        wait_pending(pending - 1 - j);
        publish((n_steps - pending + j) % WG_STAGES);
      }
    }
    if (a.gb != nullptr && tid < a.n && n_steps > 0) atomicAdd(a.gb + tid, colsum);
    // ---- epilogue: accumulators -> red.global.add into gW.  TMEM lane = gradient row (X column) within the M-block.
    if (n_steps > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      const int q = warp & 3, wg = (warp - 2) >> 2;       // two warpgroups split the columns
      const bool vec_ok = (reinterpret_cast<uintptr_t>(a.gW) & 15u) == 0;   // a.n is a multiple of 128
      const int lrow = q * 32 + lane;
      for (int mb = 0; mb < n_mblk; ++mb) {
        const int grow = mb * 128 + lrow;
        const bool ok = lrow < m_rows && grow < a.kx_valid;
        for (int cg = wg; cg < a.n / 32; cg += 2) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * 256 + cg * 32), v);
          tmem_ld_wait();
          if (ok) {
            float* dst = a.gW + (size_t)grow * a.n + cg * 32;
            if (vec_ok) {      // 16-byte aligned gradient rows: 8 vector reductions instead of 32 scalar ones
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace rnerf

using namespace rnerf;

extern "C" size_t rnerf_mlp_dgrad_packed_bytes(void) { return DG_TOTAL_BYTES; }

extern "C" int rnerf_mlp_dgrad_pack(const float* const* kernels, void* packed, void* stream) {
  RNERF_REQUIRE_PTR(kernels); RNERF_REQUIRE_PTR(packed);
  RNERF_REQUIRE(aligned16(packed), RNERF_E_ALIGN, "rnerf_mlp_dgrad_pack: packed must be 16-byte aligned");
  DgradPackArgs a;
  for (int i = 0; i < 12; ++i) {
    if (!kernels[i]) { set_error("rnerf_mlp_dgrad_pack: null kernel %d", i); return RNERF_E_NULL; }
    a.kern[i] = kernels[i];
  }
  dgrad_pack_kernel<<<DG_NCHUNK + DGP_NCHUNK, 256, 0, (cudaStream_t)stream>>>(a, (uint8_t*)packed);
  count_launch();
  return check_launch("rnerf_mlp_dgrad_pack");
}

extern "C" int rnerf_mlp_dgrad(const void* dgrad_packed, const void* fwd_packed, const uint32_t* relu_masks, const float* d_raw,
                               int64_t n_samples, uint16_t* dz_out, void* stream) {
  RNERF_REQUIRE(n_samples >= 0, RNERF_E_SHAPE, "rnerf_mlp_dgrad: n_samples < 0");
  if (n_samples == 0) return 0;
  RNERF_REQUIRE_PTR(dgrad_packed); RNERF_REQUIRE_PTR(fwd_packed); RNERF_REQUIRE_PTR(relu_masks); RNERF_REQUIRE_PTR(d_raw); RNERF_REQUIRE_PTR(dz_out);
  RNERF_REQUIRE(aligned16(dgrad_packed) && aligned16(relu_masks) && aligned16(d_raw) && aligned16(dz_out), RNERF_E_ALIGN,
                "rnerf_mlp_dgrad: buffers must be 16-byte aligned");
  constexpr int NT = 2, NSTAGE = 4;
  using SL = DgradSmem<NT, NSTAGE>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kfn = mlp_dgrad_kernel<NT, NSTAGE>;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::BYTES);
    if (e != cudaSuccess) { set_error("rnerf_mlp_dgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set[dev] = true;
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  DgradArgs a;
  a.packed = (const uint8_t*)dgrad_packed;
  a.head_w = reinterpret_cast<const float*>((const uint8_t*)fwd_packed + PK_WSIGMA);
  a.masks = relu_masks; a.d_raw = (const float4*)d_raw; a.dZ = (__nv_bfloat16*)dz_out;
  a.n_samples = n_samples;
  a.n_groups = (int)((n_samples + TILE_M * NT - 1) / (TILE_M * NT));
  const int grid = a.n_groups < n_sm ? a.n_groups : n_sm;
  CUtensorMap tm;
  int rc = make_rows_tmap(&tm, dz_out, n_samples, N_MMA_LAYERS);
  if (rc) return rc;
  // large batches: CTA-pair chain (epilogue of one tile pair under the MMAs of the other); RNERF_DGRAD_KERNEL=single disables
  const char* kenv = getenv("RNERF_DGRAD_KERNEL");
  if (!(kenv && kenv[0] == 's') && n_samples >= 74 * 512) return launch_mlp_dgrad_pair(a, tm, (cudaStream_t)stream);
  kfn<<<grid, 64 + 128 * NT, SL::BYTES, (cudaStream_t)stream>>>(a, tm);
  count_launch();
  return check_launch("rnerf_mlp_dgrad");
}

extern "C" int rnerf_mlp_head_grad(const uint16_t* saved_h, const float* d_raw, int64_t n_samples, float* out_rgb_head,
                                   float* out_sigma_head, void* stream) {
  if (n_samples <= 0) return 0;
  RNERF_REQUIRE_PTR(saved_h); RNERF_REQUIRE_PTR(d_raw); RNERF_REQUIRE_PTR(out_rgb_head); RNERF_REQUIRE_PTR(out_sigma_head);
  const int rows_per_block = 256;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int64_t slabs = (n_samples + rows_per_block - 1) / rows_per_block;
  const unsigned grid = (unsigned)(slabs < (int64_t)n_sm * 4 ? slabs : (int64_t)n_sm * 4);
  mlp_head_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)saved_h, (const float4*)d_raw, n_samples,
                                                                rows_per_block, out_rgb_head, out_sigma_head);
  count_launch();
  return check_launch("rnerf_mlp_head_grad");
}

extern "C" int rnerf_mlp_wgrad(const uint16_t* x, int ldx, int x_cols, int kx_valid, const uint16_t* dz, int n, int64_t n_samples,
                               float* gw, float* gb, void* stream) {
  if (n_samples <= 0) return 0;
  RNERF_REQUIRE_PTR(x); RNERF_REQUIRE_PTR(dz); RNERF_REQUIRE_PTR(gw);
  RNERF_REQUIRE(x_cols >= 8 && x_cols <= 256 && (x_cols % 8) == 0 && x_cols <= ldx && (ldx % 8) == 0 && kx_valid >= 1 && kx_valid <= x_cols,
                RNERF_E_SHAPE, "rnerf_mlp_wgrad: x_cols must be a multiple of 8 in [8,256] (got %d, ldx %d, valid %d)", x_cols, ldx, kx_valid);
  const int kx = x_cols <= 128 ? 128 : 256;
  RNERF_REQUIRE(n == 128 || n == 256, RNERF_E_SHAPE, "rnerf_mlp_wgrad: n must be 128 or 256");
  RNERF_REQUIRE(aligned16(x) && aligned16(dz), RNERF_E_ALIGN, "rnerf_mlp_wgrad: x/dz must be 16-byte aligned");
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgradSmem::BYTES);
    if (e != cudaSuccess) { set_error("rnerf_mlp_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set[dev] = true;
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  WgradBatch b;
  memset(&b, 0, sizeof(b));
  WgradArgs& a = b.job[0];
  a.X = (const __nv_bfloat16*)x; a.dZ = (const __nv_bfloat16*)dz; a.ldx = ldx; a.kx = kx; a.x_cols = x_cols; a.kx_valid = kx_valid; a.n = n;
  a.n_samples = n_samples; a.gW = gw; a.gb = gb;
  int64_t per = (n_samples + n_sm - 1) / n_sm;
  per = (per + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  a.rows_per_cta = (int)per;
  const unsigned grid = (unsigned)((n_samples + per - 1) / per);
  b.n_jobs = 1; b.cta0[0] = 0; b.cta0[1] = (int)grid;
  mlp_wgrad_kernel<<<grid, WG_THREADS, WgradSmem::BYTES, (cudaStream_t)stream>>>(b);
  count_launch();
  return check_launch("rnerf_mlp_wgrad");
}

// All weight-gradient GEMMs of one MLP's backward in ONE launch (same per-job arguments as rnerf_mlp_wgrad, as arrays).
// The SMs are divided between the jobs in proportion to the bytes each job streams (rows x (x_cols + n)), one CTA per SM.
extern "C" int rnerf_mlp_wgrad_batched(int n_jobs, const uint16_t* const* x, const int* ldx, const int* x_cols, const int* kx_valid,
                                       const uint16_t* const* dz, const int* n, int64_t n_samples, float* const* gw,
                                       float* const* gb, void* stream) {
  if (n_samples <= 0 || n_jobs <= 0) return 0;
  RNERF_REQUIRE(n_jobs <= WG_MAX_JOBS, RNERF_E_SHAPE, "rnerf_mlp_wgrad_batched: at most %d jobs per launch (got %d)", WG_MAX_JOBS, n_jobs);
  RNERF_REQUIRE_PTR(x); RNERF_REQUIRE_PTR(ldx); RNERF_REQUIRE_PTR(x_cols); RNERF_REQUIRE_PTR(kx_valid); RNERF_REQUIRE_PTR(dz);
  RNERF_REQUIRE_PTR(n); RNERF_REQUIRE_PTR(gw); RNERF_REQUIRE_PTR(gb);
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgradSmem::BYTES);
  if (e != cudaSuccess) { set_error("rnerf_mlp_wgrad_batched: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  WgradBatch b;
  memset(&b, 0, sizeof(b));
  double total = 0.0;
  for (int j = 0; j < n_jobs; ++j) {
    RNERF_REQUIRE(x[j] != nullptr && dz[j] != nullptr && gw[j] != nullptr, RNERF_E_NULL, "rnerf_mlp_wgrad_batched: job %d has a null pointer", j);
    RNERF_REQUIRE(x_cols[j] >= 8 && x_cols[j] <= 256 && (x_cols[j] % 8) == 0 && x_cols[j] <= ldx[j] && (ldx[j] % 8) == 0 && kx_valid[j] >= 1 &&
                  kx_valid[j] <= x_cols[j], RNERF_E_SHAPE, "rnerf_mlp_wgrad_batched: job %d: bad x_cols / ldx / kx_valid", j);
    RNERF_REQUIRE(n[j] == 128 || n[j] == 256, RNERF_E_SHAPE, "rnerf_mlp_wgrad_batched: job %d: n must be 128 or 256", j);
    RNERF_REQUIRE(aligned16(x[j]) && aligned16(dz[j]), RNERF_E_ALIGN, "rnerf_mlp_wgrad_batched: job %d: x/dz must be 16-byte aligned", j);
    total += (double)(x_cols[j] + n[j]);
  }
  // CTAs per job.  Large batches (>= 1024 rows per CTA even when a job is cut over every SM): every job gets all the
  // SMs, jobs in sequence -- the schedule of 13 separate launches without their gaps and tails (measured: a single wave with
  // ~60 000 rows per CTA is slower, the per-stage floor of the short-row jobs unbalances it).  Small batches (a 512-ray shard
  // of an 8-GPU step): one wave, the SMs divided in proportion to the bytes each job streams, so the fixed costs (TMEM
  // allocation, the 256 KB accumulator flush) are paid once per SM instead of 13 times.
  int ctas[WG_MAX_JOBS], sum = 0;
  const int64_t max_ctas = (n_samples + WG_ROWS - 1) / WG_ROWS;
  const bool big = n_samples >= (int64_t)n_sm * RNERF_WG_BIG_ROWS;
  for (int j = 0; j < n_jobs; ++j) {
    int c = big ? n_sm : (int)((double)n_sm * (x_cols[j] + n[j]) / total);
    if (c < 1) c = 1;
    if (c > max_ctas) c = (int)max_ctas;
    ctas[j] = c; sum += c;
  }
  for (int j = 0; !big && sum < n_sm && j < n_jobs; ++j)  // hand the rounding remainder to the largest jobs first
    if (n[j] == 256 && x_cols[j] == 256 && ctas[j] < max_ctas) { ++ctas[j]; ++sum; }
  int c0 = 0;
  for (int j = 0; j < n_jobs; ++j) {
    WgradArgs& a = b.job[j];
    a.X = (const __nv_bfloat16*)x[j]; a.dZ = (const __nv_bfloat16*)dz[j]; a.ldx = ldx[j]; a.kx = x_cols[j] <= 128 ? 128 : 256;
    a.x_cols = x_cols[j]; a.kx_valid = kx_valid[j]; a.n = n[j]; a.n_samples = n_samples; a.gW = gw[j]; a.gb = gb[j];
    int64_t per = (n_samples + ctas[j] - 1) / ctas[j];
    per = (per + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
    a.rows_per_cta = (int)per;
    b.cta0[j] = c0;
    c0 += (int)((n_samples + per - 1) / per);
  }
  b.cta0[n_jobs] = c0;
  b.n_jobs = n_jobs;
  mlp_wgrad_kernel<<<(unsigned)c0, WG_THREADS, WgradSmem::BYTES, (cudaStream_t)stream>>>(b);
  count_launch();
  return check_launch("rnerf_mlp_wgrad_batched");
}
