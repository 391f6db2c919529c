// Data-gradient chain of the radiance MLP, CTA-pair version (tcgen05 cta_group::2) -- the training backward for large
// batches.  Same arithmetic and outputs as mlp_dgrad_kernel (mlp_bwd.cu; a17: train.py:164-165 differentiating NerfMLP,
// rnerf/model_utils.py:30-90); the schedule is the forward pair kernel's (encmlp_pair.cu):
//
//   * two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 of M = 256 (128 rows from each CTA), N = 256; each CTA
//     stages its N-half of W_l ([128 x 64] per k-block), so a whole GEMM's weights are resident in 64 KB per CTA and are
//     used by tile pair P0, then by P1, before the slots are refilled;
//   * each CTA owns two 128-row tiles; issue order per GEMM: P0's K-loop -> acc[0], P1's K-loop -> acc[1].  While P1's
//     MMAs run, P0's eight epilogue warps (two warpgroups, 128 columns each) apply the ReLU bit-mask, write dZ_l as the next
//     GEMM's A operand and hand the tile to the TMA engine (tensor stores to dZ[l]) -- the single-CTA kernel ran both
//     tiles' MMAs and then both epilogues, leaving the tensor pipe idle 3/4 of the time.
//
// Cross-CTA protocol: as in encmlp_pair.cu (full / empty / acc / aready; the leader = cluster rank 0 issues every MMA).
#include <stdlib.h>
#include "mlp_bwd.cuh"
#include "pair.cuh"

namespace rnerf {

constexpr int DGP_NSLOT = 4;
constexpr int DGP_THREADS = 64 + 512;    // producer, MMA/relay, 16 epilogue warps

struct DgradPairSmem {
  static constexpr uint32_t A_OFF = 0;                                  // [2 tiles][4][16 KB] dZ k-blocks
  static constexpr uint32_t W_OFF = A_OFF + 2 * 4 * ABLK_BYTES;         // [4 slots][16 KB] this CTA's N-half of a k-block
  static constexpr uint32_t H_OFF = W_OFF + DGP_NSLOT * PAIR_HALF_BYTES;   // w_sigma[256] + w_rgb[3][128] fp32
  static constexpr uint32_t BAR_OFF = H_OFF + (256 + 384) * 4;
  static constexpr uint32_t N_BARS = 2 * DGP_NSLOT + 4;                 // full, empty, acc[2], aready[2]
  static constexpr uint32_t TMEM_SLOT = BAR_OFF + N_BARS * 8;
  static constexpr uint32_t BYTES = TMEM_SLOT + 16;
};
static_assert(DgradPairSmem::BYTES <= 232448, "dgrad pair kernel exceeds the shared-memory budget");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DGP_THREADS, 1) mlp_dgrad_pair_kernel(const DgradArgs args,
                                                                                                  const __grid_constant__ CUtensorMap tm_dz) {
  using SL = DgradPairSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  auto bar_full = [&](int s) { return sbase + SL::BAR_OFF + 8u * s; };
  auto bar_empty = [&](int s) { return sbase + SL::BAR_OFF + 8u * (DGP_NSLOT + s); };
  auto bar_acc = [&](int j) { return sbase + SL::BAR_OFF + 8u * (2 * DGP_NSLOT + j); };
  auto bar_aready = [&](int j) { return sbase + SL::BAR_OFF + 8u * (2 * DGP_NSLOT + 2 + j); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SL::TMEM_SLOT);
  float* head_s = reinterpret_cast<float*>(smem + SL::H_OFF);

  if (threadIdx.x == 0) {
    // the leader's full[s] also takes the peer's relayed "my half has landed" (one wait per chunk for the MMA issuer)
    for (int s = 0; s < DGP_NSLOT; ++s) { mbar_init(bar_full(s), leader ? 2 : 1); mbar_init(bar_empty(s), 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(bar_acc(j), 1); mbar_init(bar_aready(j), 16); }
    fence_barrier_init();
  }
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / multicast commit
  if (warp == 1) {
    tmem_alloc2(sbase + SL::TMEM_SLOT, 512);
    tmem_relinquish2();
  }
  for (int i = threadIdx.x; i < 256 + 384; i += blockDim.x) head_s[i] = __ldg(args.head_w + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_pairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
  const int my_groups = (args.n_groups > pair) ? (args.n_groups - pair + n_pairs - 1) / n_pairs : 0;

  if (warp == 0) {
    // ===================== weight producer: this CTA's N-half of every k-block of the chain =====================
    if (lane == 0) {
      uint32_t c = 0;   // running chunk counter: slot = c % 4, phase = (c / 4) & 1
      for (int g = 0; g < my_groups; ++g)
        for (int i = 0; i < DGP_NCHUNK; ++i, ++c) {
          const int s = c % DGP_NSLOT;
          mbar_wait_cluster(bar_empty(s), ((c / DGP_NSLOT) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full(s), PAIR_HALF_BYTES);
          tma_bulk_g2s(sbase + SL::W_OFF + s * PAIR_HALF_BYTES,
                       args.packed + DG_PAIR_OFF + (size_t)i * PAIR_CHUNK_STRIDE + rank * PAIR_HALF_BYTES, PAIR_HALF_BYTES, bar_full(s));
        }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      if (!leader) {
        // ===================== peer: relay "my half of slot s has landed" to the leader =====================
        uint32_t c = 0;
        for (int g = 0; g < my_groups; ++g)
          for (int i = 0; i < DGP_NCHUNK; ++i, ++c) {
            const int s = c % DGP_NSLOT;
            mbar_wait(bar_full(s), (c / DGP_NSLOT) & 1);
            mbar_arrive_remote(mapa_shared(bar_full(s), 0));
          }
      } else {
        // ===================== leader: MMA issuer for the pair =====================
        constexpr uint32_t idesc = make_idesc(2 * TILE_M, 256);
        uint32_t ar_phase[2] = {0, 0};
        uint32_t cbase = 0;   // chunk counter at the start of the current GEMM
        for (int g = 0; g < my_groups; ++g)
          for (int d = 0; d < DG_GEMMS; ++d) {
            const int nkb = dg_k(d) / KB;
            for (int j = 0; j < 2; ++j) {
              mbar_wait_cluster(bar_aready(j), ar_phase[j]);   // dZ of the layer above ready in both CTAs, acc[j] drained
              ar_phase[j] ^= 1;
              tc_fence_after();
              for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t c = cbase + kb;
                const int s = c % DGP_NSLOT;
                if (j == 0) {
                  mbar_wait_cluster(bar_full(s), (c / DGP_NSLOT) & 1);     // both halves of the slot
                  tc_fence_after();
                }
                const uint64_t a_desc = make_sw128_desc(sbase + SL::A_OFF + (j * 4 + kb) * ABLK_BYTES);
                const uint64_t b_desc = make_sw128_desc(sbase + SL::W_OFF + s * PAIR_HALF_BYTES);
                const uint32_t a_lo = (uint32_t)a_desc, b_lo = (uint32_t)b_desc, desc_hi = (uint32_t)(a_desc >> 32);
#pragma unroll
                for (int ks = 0; ks < KB / 16; ++ks)      // +32 bytes of K per step = +2 in the descriptor's address field
                  umma2_bf16_lohi(tmem_base + (uint32_t)(j * 256), a_lo + 2u * ks, b_lo + 2u * ks, desc_hi, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                if (j == 1) umma2_commit_mc(bar_empty(s));   // both CTAs may refill their half of the slot
              }
              umma2_commit_mc(bar_acc(j));                   // accumulators of tile pair j final in both CTAs
            }
            cbase += nkb;
          }
      }
    }
  } else {
    // ===================== epilogue: 8 warps per tile, 128 columns per warpgroup; thread <-> sample row =====================
    const int ew = warp - 2;
    const int t = ew >> 3;                       // tile (pair) index j
    const int half = (ew >> 2) & 1;              // column half handled by this warpgroup
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    uint8_t* a_row = smem + SL::A_OFF + t * 4 * ABLK_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
    const uint32_t r7s = (uint32_t)(row & 7) << 4;
    const float* w_sigma = head_s;
    const float* w_rgb = head_s + 256;
    const uint32_t aready_remote = leader ? 0u : mapa_shared(bar_aready(t), 0);
    const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + half * 128);
    const uint32_t a_warp = sbase + SL::A_OFF + t * 4 * ABLK_BYTES + q * 4096;     // this warp's 32 rows of k-block 0
    uint32_t acc_phase = 0;

    auto signal_ready = [&]() {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive_cluster_local(bar_aready(t)); else mbar_arrive_remote(aready_remote);
      }
    };
    auto tile_free = [&]() {                     // this warp's previous tensor stores have finished reading its rows
      if (lane == 0) bulk_wait_read();
      __syncwarp();
    };

    for (int g = 0; g < my_groups; ++g) {
      const int64_t group = (int64_t)pair + (int64_t)g * n_pairs;
      const int64_t wrow0 = group * 512 + t * 256 + (int64_t)rank * 128 + q * 32;   // first sample row of this warp
      const int64_t srow = wrow0 + lane;
      const bool live = srow < args.n_samples;
      const int64_t lrow = live ? srow : (args.n_samples - 1);
      const float4 draw = live ? __ldg(args.d_raw + lrow) : make_float4(0.f, 0.f, 0.f, 0.f);
      // ---- prologue: rgb head (Dense_11) backward + ReLU of the condition layer -> dZ[9]; this half's 64 columns = k-block `half`
      {
        uint32_t m9[2];                           // condition layer: this half's 64 columns = words 2 half, 2 half + 1
        {
          const uint2 a2 = __ldg(reinterpret_cast<const uint2*>(args.masks + ((size_t)9 * args.n_samples + lrow) * 8) + half);
          m9[0] = a2.x; m9[1] = a2.y;
        }
        tile_free();
        tile_bar_sync(t);                         // k-block 1 was last stored by the OTHER warpgroup (dZ[0], columns 64..127)
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {             // 8 columns per 16-byte unit
          const int c8 = half * 8 + u;
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j0 = c8 * 8 + i * 2;
            float g0 = draw.x * w_rgb[j0] + draw.y * w_rgb[128 + j0] + draw.z * w_rgb[256 + j0];
            float g1 = draw.x * w_rgb[j0 + 1] + draw.y * w_rgb[128 + j0 + 1] + draw.z * w_rgb[256 + j0 + 1];
            // columns 8 c8 + 2 i, + 1: pair (u & 3) * 4 + i of this half's group u >> 2 (relu_mask32)
            const uint32_t word = (u >> 2) == 0 ? m9[0] : m9[1];
            if (!relu_mask_even(word, (u & 3) * 4 + i)) g0 = 0.f;
            if (!relu_mask_odd(word, (u & 3) * 4 + i)) g1 = 0.f;
            pk[i] = pack_bf16(g0, g1);
          }
          *reinterpret_cast<uint4*>(a_row + half * ABLK_BYTES + ((uint32_t)(u << 4) ^ r7s)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      signal_ready();
      if (lane == 0 && wrow0 < args.n_samples) {
        tma_store_3d(&tm_dz, a_warp + half * ABLK_BYTES, half * KB, (int)wrow0, 9);
        bulk_commit();
      }
      for (int d = 0; d < DG_GEMMS; ++d) {
        const int lo = 8 - d;                       // MMA layer whose dZ this GEMM produces
        uint32_t mask[4] = {0, 0, 0, 0};            // this row's ReLU bits of layer lo, this half's 128 columns
        if (d >= 1) {                               // 16 bytes per row and half, hidden under the MMAs
          const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(args.masks + ((size_t)lo * args.n_samples + lrow) * 8) + half);
          mask[0] = m4.x; mask[1] = m4.y; mask[2] = m4.z; mask[3] = m4.w;
        }
        mbar_wait_cluster(bar_acc(t), acc_phase);   // this tile pair's K-loop is complete
        acc_phase ^= 1;
        tc_fence_after();
        tile_free();
        if (d == 0) tile_bar_sync(t);               // dZ[9]'s k-block 1 was stored by the other warpgroup's warp
#pragma unroll 1
        for (int cg = 0; cg < 4; ++cg) {
          uint32_t v[32];
          tmem_ld32(taddr_row + (uint32_t)(cg * 32), v);
          tmem_ld_wait();
          uint32_t mk = 0u;                         // this 32-column group's word
#pragma unroll
          for (int p = 0; p < 4; ++p) mk = (p == cg) ? mask[p] : mk;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float f0 = __uint_as_float(v[2 * j]), f1 = __uint_as_float(v[2 * j + 1]);
            if (d == 1) {   // the sigma head (Dense_8) also feeds h7: dh7 += d_sigma * w_sigma
              f0 = fmaf(draw.w, w_sigma[half * 128 + cg * 32 + 2 * j], f0);
              f1 = fmaf(draw.w, w_sigma[half * 128 + cg * 32 + 2 * j + 1], f1);
            }
            if (d >= 1) {                          // columns 32 cg + 2 j, + 1 of the half: bits j and 16 + j of the group's word
              if (!relu_mask_even(mk, j)) f0 = 0.f;
              if (!relu_mask_odd(mk, j)) f1 = 0.f;
            }
            pk[j] = pack_bf16(f0, f1);
          }
          uint8_t* blk = a_row + (half * 2 + (cg >> 1)) * ABLK_BYTES;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(blk + ((uint32_t)(((cg & 1) * 4 + c) << 4) ^ r7s)) =
                make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        if (d < DG_GEMMS - 1) {
          signal_ready();
        } else {               // last GEMM of the chain: nothing consumes the tile, the accumulators are drained
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
        }
        if (lane == 0 && wrow0 < args.n_samples) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
            tma_store_3d(&tm_dz, a_warp + (half * 2 + kb) * ABLK_BYTES, (half * 2 + kb) * KB, (int)wrow0, lo);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA exits (or frees TMEM) while its peer may still address it
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

int launch_mlp_dgrad_pair(const DgradArgs& a0, const CUtensorMap& tm_dz, cudaStream_t st) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_dgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DgradPairSmem::BYTES);
    if (e != cudaSuccess) { set_error("rnerf_mlp_dgrad: cudaFuncSetAttribute(pair): %s", cudaGetErrorString(e)); return (int)e; }
    attr_set[dev] = true;
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  DgradArgs a = a0;
  a.n_groups = (int)((a.n_samples + 511) / 512);
  int pairs = n_sm / 2;
  if (a.n_groups < pairs) pairs = a.n_groups;
  mlp_dgrad_pair_kernel<<<2 * pairs, DGP_THREADS, DgradPairSmem::BYTES, st>>>(a, tm_dz);
  count_launch();
  return check_launch("rnerf_mlp_dgrad(pair)");
}

}  // namespace rnerf
