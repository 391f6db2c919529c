// Encoding-fused radiance MLP (a8 + a9) on tcgen05 / TMEM, sm_100a.
//
// pos_enc (rnerf/model_utils.py:187-214) of the bent sample position (63-d) and direction (27-d) is
// computed in-kernel straight into the swizzled bf16 A-operand tile; NerfMLP (rnerf/model_utils.py:30-90:
// 8x256 ReLU trunk with the input re-concatenated after layer 4, sigma head, 256 bottleneck, +27-d
// per-sample condition -> 128 -> rgb) runs as a chain of tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) whose
// B operands (weights) are streamed by the TMA engine (cp.async.bulk, 16 KB pre-swizzled chunks) through an
// mbarrier ring, and whose epilogues (bias, ReLU, bf16 pack, heads) read the accumulators with tcgen05.ld
// and write the next layer's A operand back to shared memory.  Activations never touch HBM.
//
// CTA roles: warp 0 = weight producer (TMA), warp 1 = MMA issuer + TMEM owner, warps 2.. = one epilogue
// warpgroup (128 threads, thread <-> sample row) per 128-row tile.  NT tiles (NT*128 samples) share every
// weight chunk, which divides the L2->SM weight traffic by NT.
//
// Tile geometry: M = 128 rows per tile (UMMA_M = 128, cta_group::1), N = the whole layer width per MMA (256, or
// 128 for the condition layer) so that the A operand is read from shared memory once per K-step; activations are
// [128 x 64] SWIZZLE_128B k-blocks, weights stream as [N x 32] SWIZZLE_64B chunks (2 x UMMA_K=16 each).
#include <stdlib.h>
#include <string.h>
#include "umma.cuh"

namespace rnerf {

// ------------------------------------------------------------------------------------------------
// weight packing: Flax [in,out] fp32 kernels -> pre-swizzled bf16 [N x 32] chunks + fp32 tail
// ------------------------------------------------------------------------------------------------
struct PackArgs {
  const float* kern[12];
  const float* bias[12];
};

__global__ void __launch_bounds__(256) encmlp_pack_kernel(PackArgs a, uint8_t* __restrict__ packed) {
  // chunk enumeration must match the MMA issuer: for layer, for k-block (A blocks then E), for 32-wide sub-chunk
  int c = blockIdx.x;
  if (c < N_CHUNKS) {
    int l = 0, rem = c;
    while (rem >= layer_chunks(l)) { rem -= layer_chunks(l); ++l; }
    const int kbi = rem / SUBS, sub = rem % SUBS;
    const bool is_e = kbi >= layer_akb(l);
    const int dense = layer_dense(l);
    const int in_dim = (l == 0) ? POS_ENC : (l == 5 ? 256 + POS_ENC : (l == 9 ? 256 + DIR_ENC : 256));
    const int out_dim = layer_n(l);
    const float* W = a.kern[dense];
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(packed + PK_CHUNKS + (size_t)c * SLOT_BYTES);
    for (int e = threadIdx.x; e < NMAX * KCH; e += blockDim.x) {
      const int n = e / KCH, k = e % KCH;
      const int kk = (is_e ? layer_akb(l) * KB : kbi * KB) + sub * KCH + k;
      float v = (kk < in_dim && n < out_dim) ? W[(size_t)kk * out_dim + n] : 0.f;
      dst[sw64_offset(n, k) / 2] = __float2bfloat16_rn(v);
    }
  } else if (c > N_CHUNKS) {
    // pair-kernel image: chunk pc of the stream, both N-halves
    static const PairChunk stream[PAIR_NCHUNK] = {RNERF_PAIR_STREAM};
    const int pc = c - N_CHUNKS - 1;
    const int l = stream[pc].layer, src = stream[pc].src;
    const int dense = layer_dense(l);
    const int in_dim = (l == 0) ? POS_ENC : (l == 5 ? 256 + POS_ENC : (l == 9 ? 256 + DIR_ENC : 256));
    const int out_dim = layer_n(l), half_n = out_dim / 2;
    const float* W = a.kern[dense];
    for (int h = 0; h < 2; ++h) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(packed + PK_PAIR + (size_t)pc * PAIR_CHUNK_STRIDE + h * PAIR_HALF_BYTES);
      for (int e = threadIdx.x; e < 128 * KB; e += blockDim.x) {
        const int r = e / KB, k = e % KB;
        const int kk = (src == 4 ? layer_akb(l) * KB : src * KB) + k;
        const int n = h * half_n + r;
        float v = (r < half_n && kk < in_dim) ? W[(size_t)kk * out_dim + n] : 0.f;
        dst[sw128_offset(r, k) / 2] = __float2bfloat16_rn(v);
      }
    }
  } else {
    float* bias = reinterpret_cast<float*>(packed + PK_BIAS);
    for (int e = threadIdx.x; e < 10 * 256; e += blockDim.x) {
      const int l = e / 256, j = e % 256;
      bias[e] = j < layer_n(l) ? a.bias[layer_dense(l)][j] : 0.f;
    }
    float* ws = reinterpret_cast<float*>(packed + PK_WSIGMA);
    for (int e = threadIdx.x; e < 256; e += blockDim.x) ws[e] = bf16_round(a.kern[8][e]);  // Dense_8: [256,1]
    float* wr = reinterpret_cast<float*>(packed + PK_WRGB);
    for (int e = threadIdx.x; e < 3 * 128; e += blockDim.x) {                               // Dense_11: [128,3]
      const int ch = e / 128, j = e % 128;
      wr[e] = bf16_round(a.kern[11][j * 3 + ch]);
    }
    if (threadIdx.x == 0)
      *reinterpret_cast<float4*>(packed + PK_HEADB) = make_float4(a.bias[11][0], a.bias[11][1], a.bias[11][2], a.bias[8][0]);
  }
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
template <int NT, int NSTAGE>
struct SmemLayout {
  static constexpr uint32_t A_OFF = 0;                                   // [NT][4][16 KB] activation k-blocks
  static constexpr uint32_t E_OFF = A_OFF + NT * 4 * ABLK_BYTES;         // [NT][16 KB] encoding k-block
  static constexpr uint32_t W_OFF = E_OFF + NT * ABLK_BYTES;             // [NSTAGE][16 KB] weight ring
  static constexpr uint32_t V_OFF = W_OFF + NSTAGE * SLOT_BYTES;         // 2 x 1 KB fp32: bias / head-weight slots
  static constexpr uint32_t BAR_OFF = V_OFF + 2048;                      // mbarriers (8 B each)
  static constexpr uint32_t N_BARS = 2 * NSTAGE + 2;                     // full[], empty[], acc, a_ready
  static constexpr uint32_t TMEM_SLOT = BAR_OFF + N_BARS * 8;
  static constexpr uint32_t BYTES = TMEM_SLOT + 16;
};

// Epilogue of one MMA layer for one sample row (thread): accumulators (TMEM) + bias -> [ReLU] -> bf16 ->
// next layer's A operand (swizzled smem) and/or the fused heads.
//   KIND 0: ReLU, write A            (Dense_0..6)
//   KIND 1: ReLU, write A, sigma head (Dense_7 -> Dense_8)
//   KIND 2: no activation, write A   (Dense_9 bottleneck)
//   KIND 3: ReLU, rgb head only      (Dense_10 -> Dense_11), 128 columns

template <int KIND, bool DUMP>
__device__ __forceinline__ void epilogue_row(uint32_t taddr, const float* __restrict__ bias, const float* __restrict__ vslot,
                                             uint8_t* __restrict__ a_row, uint32_t r7s, EpiOut& o,
                                             __nv_bfloat16* __restrict__ dump_row, uint32_t* __restrict__ mask_row = nullptr) {
  constexpr int NCG = (KIND == 3) ? 4 : 8;
  uint32_t mw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};             // ReLU bit-mask word of every 32-column group (umma.cuh)
#pragma unroll 1
  for (int cg2 = 0; cg2 < NCG; cg2 += 2) {   // 2 x 32 accumulator columns per iteration, both TMEM loads in flight
    uint32_t v[2][32];
    tmem_ld32(taddr + cg2 * 32, v[0]);
    tmem_ld32(taddr + cg2 * 32 + 32, v[1]);
    tmem_ld_wait();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c0 = cg2 * 32 + half * 32;
      uint32_t pk[16];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + c0 + j4 * 4);
        add2(v[half][4 * j4 + 0], v[half][4 * j4 + 1], b4.x, b4.y);
        add2(v[half][4 * j4 + 2], v[half][4 * j4 + 3], b4.z, b4.w);
        const float f0 = __uint_as_float(v[half][4 * j4 + 0]), f1 = __uint_as_float(v[half][4 * j4 + 1]);
        const float f2 = __uint_as_float(v[half][4 * j4 + 2]), f3 = __uint_as_float(v[half][4 * j4 + 3]);
        if (KIND == 2) { pk[2 * j4] = pack_bf16(f0, f1);      pk[2 * j4 + 1] = pack_bf16(f2, f3); }
        else           { pk[2 * j4] = pack_bf16_relu(f0, f1); pk[2 * j4 + 1] = pack_bf16_relu(f2, f3); }
      }
      if (DUMP && KIND != 2) relu_mask_set(mw, cg2 + half, relu_mask32(pk));
      if (KIND == 1) {  // sigma head (Dense_8) on the bf16-rounded trunk output; weights in vslot[0..255]
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 w4 = *reinterpret_cast<const float4*>(vslot + c0 + j4 * 4);
          o.sigma = fmaf(bf16_lo(pk[2 * j4]), w4.x, o.sigma);
          o.sigma = fmaf(bf16_hi(pk[2 * j4]), w4.y, o.sigma);
          o.sigma = fmaf(bf16_lo(pk[2 * j4 + 1]), w4.z, o.sigma);
          o.sigma = fmaf(bf16_hi(pk[2 * j4 + 1]), w4.w, o.sigma);
        }
      }
      if (KIND == 3) {  // rgb head (Dense_11) on the bf16-rounded condition-layer output
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 wr = *reinterpret_cast<const float4*>(vslot + c0 + j4 * 4);
          const float4 wg = *reinterpret_cast<const float4*>(vslot + 128 + c0 + j4 * 4);
          const float4 wb = *reinterpret_cast<const float4*>(vslot + 384 + c0 + j4 * 4);
          const float h0 = bf16_lo(pk[2 * j4]), h1 = bf16_hi(pk[2 * j4]), h2 = bf16_lo(pk[2 * j4 + 1]), h3 = bf16_hi(pk[2 * j4 + 1]);
          o.r = fmaf(h0, wr.x, o.r); o.r = fmaf(h1, wr.y, o.r); o.r = fmaf(h2, wr.z, o.r); o.r = fmaf(h3, wr.w, o.r);
          o.g = fmaf(h0, wg.x, o.g); o.g = fmaf(h1, wg.y, o.g); o.g = fmaf(h2, wg.z, o.g); o.g = fmaf(h3, wg.w, o.g);
          o.b = fmaf(h0, wb.x, o.b); o.b = fmaf(h1, wb.y, o.b); o.b = fmaf(h2, wb.z, o.b); o.b = fmaf(h3, wb.w, o.b);
        }
      } else {
        // next layer's A operand: these 32 columns live in k-block cg2/2 at 16-byte units half*4 .. half*4+3,
        // xor-swizzled with (row & 7)
        uint8_t* blk = a_row + (cg2 >> 1) * ABLK_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(blk + ((uint32_t)((half * 4 + c) << 4) ^ r7s)) =
              make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      if (DUMP) {
        if (dump_row != nullptr) {
          uint4* dst = reinterpret_cast<uint4*>(dump_row + c0);
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
      }
    }
  }
  if (DUMP && KIND != 2 && mask_row != nullptr) {
    reinterpret_cast<uint4*>(mask_row)[0] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
    if (KIND != 3) reinterpret_cast<uint4*>(mask_row)[1] = make_uint4(mw[4], mw[5], mw[6], mw[7]);
  }
}

__device__ __forceinline__ void epi_bar_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <int NT, int NSTAGE, bool DEBUG>
__global__ void __launch_bounds__(64 + 128 * NT, 1) encmlp_kernel(const EncMlpArgs args,
                                                                  const __grid_constant__ CUtensorMap tm_layers) {
  using SL = SmemLayout<NT, NSTAGE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();  // SWIZZLE_128B atoms need 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  auto bar_full = [&](int s) { return sbase + SL::BAR_OFF + 8u * s; };
  auto bar_empty = [&](int s) { return sbase + SL::BAR_OFF + 8u * (NSTAGE + s); };
  const uint32_t bar_acc = sbase + SL::BAR_OFF + 8u * (2 * NSTAGE);
  const uint32_t bar_aready = sbase + SL::BAR_OFF + 8u * (2 * NSTAGE + 1);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SL::TMEM_SLOT);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_aready, 4 * NT);  // one arrive per epilogue warp
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(sbase + SL::TMEM_SLOT, 256 * NT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_groups = (args.n_groups > (int)blockIdx.x) ? (args.n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ===================== weight producer (TMA bulk copies through an mbarrier ring) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int g = 0; g < my_groups; ++g) {
        int c = 0;
        for (int l = 0; l < N_MMA_LAYERS; ++l) {
          const uint32_t bytes = (uint32_t)layer_n(l) * KCH * 2;
          for (int i = 0; i < layer_chunks(l); ++i, ++c) {
            mbar_wait(bar_empty(stage), phase ^ 1);
            mbar_arrive_expect_tx(bar_full(stage), bytes);
            tma_bulk_g2s(sbase + SL::W_OFF + stage * SLOT_BYTES, args.packed + PK_CHUNKS + (size_t)c * SLOT_BYTES, bytes,
                         bar_full(stage));
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0, ar_phase = 0;
      for (int g = 0; g < my_groups; ++g) {
        for (int l = 0; l < N_MMA_LAYERS; ++l) {
          const bool prof = DEBUG && args.prof != nullptr && blockIdx.x == 0 && g == 0;
          if (prof) args.prof[l * 4 + 0] = clock64();
          mbar_wait(bar_aready, ar_phase);  // A operand of layer l written (and accumulators drained)
          ar_phase ^= 1;
          tc_fence_after();
          if (prof) args.prof[l * 4 + 1] = clock64();
          const int akb = layer_akb(l), nkb = akb + layer_has_e(l);
          const uint32_t idesc = make_idesc(TILE_M, layer_n(l));
          for (int kbi = 0; kbi < nkb; ++kbi) {
#pragma unroll
            for (int sub = 0; sub < SUBS; ++sub) {
              mbar_wait(bar_full(stage), phase);
              tc_fence_after();
              const uint32_t b_addr = sbase + SL::W_OFF + stage * SLOT_BYTES;
#pragma unroll
              for (int t = 0; t < NT; ++t) {
                const uint32_t a_addr = ((kbi < akb) ? (sbase + SL::A_OFF + (t * 4 + kbi) * ABLK_BYTES)
                                                     : (sbase + SL::E_OFF + t * ABLK_BYTES)) + sub * (KCH * 2);
                const uint32_t d_addr = tmem_base + (uint32_t)(t * 256);
                const uint64_t ad = make_sw128_desc(a_addr), bd = make_sw64_desc(b_addr);
#pragma unroll
                for (int ks = 0; ks < KCH / 16; ++ks)     // +32 bytes of K per step = +2 in the descriptors' address fields
                  umma_bf16_lohi(d_addr, (uint32_t)ad + 2u * ks, (uint32_t)(ad >> 32), (uint32_t)bd + 2u * ks, (uint32_t)(bd >> 32),
                                 idesc, (kbi > 0 || sub > 0 || ks > 0) ? 1u : 0u);
              }
              umma_commit(bar_empty(stage));  // frees the weight slot once these MMAs have read it
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
          }
          umma_commit(bar_acc);               // accumulators of layer l complete
          if (prof) args.prof[l * 4 + 2] = clock64();
        }
      }
    }
  } else {
    // ===================== encoder / epilogue warpgroups =====================
    const int t = (warp - 2) >> 2;               // tile handled by this warpgroup
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;               // row within the tile == TMEM lane
    const int etid = threadIdx.x - 64;           // 0 .. 128*NT-1
    uint8_t* a_blk = smem + SL::A_OFF + t * 4 * ABLK_BYTES;
    uint8_t* a_row = a_blk + (row >> 3) * 1024 + (row & 7) * 128;   // this row inside a SWIZZLE_128B k-block
    const uint32_t r7s = (uint32_t)(row & 7) << 4;                   // xor mask of the 16-byte unit index
    uint8_t* e_blk = smem + SL::E_OFF + t * ABLK_BYTES;
    float* vslot = reinterpret_cast<float*>(smem + SL::V_OFF);   // [2][256] fp32
    const float* bias_all = reinterpret_cast<const float*>(args.packed + PK_BIAS);
    const float* wsig_g = reinterpret_cast<const float*>(args.packed + PK_WSIGMA);
    const float* wrgb_g = reinterpret_cast<const float*>(args.packed + PK_WRGB);
    const float4 headb = __ldg(reinterpret_cast<const float4*>(args.packed + PK_HEADB));
    uint32_t acc_phase = 0;

    for (int g = 0; g < my_groups; ++g) {
      const int64_t group = (int64_t)blockIdx.x + (int64_t)g * gridDim.x;
      const int64_t srow = (group * NT + t) * TILE_M + row;
      const bool live = srow < args.n_samples;
      const int64_t lrow = live ? srow : (args.n_samples - 1);
      const float p0 = __ldg(args.pos + 3 * lrow), p1 = __ldg(args.pos + 3 * lrow + 1), p2 = __ldg(args.pos + 3 * lrow + 2);
      // ---- layer-0 A operand: pos_enc(pos, 0, 10) ----
      write_encoding<10>(e_blk, row, p0, p1, p2,
                         (DEBUG && args.enc_out && live) ? reinterpret_cast<uint4*>(args.enc_out + (size_t)srow * 64) : nullptr);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_aready);

      EpiOut eo = {0.f, 0.f, 0.f, 0.f};
      for (int l = 0; l < N_MMA_LAYERS; ++l) {
        const bool prof = DEBUG && args.prof != nullptr && blockIdx.x == 0 && g == 0 && warp == 2 && lane == 0;
        if (prof) args.prof[40 + l * 4 + 0] = clock64();
        // stage this layer's bias (slot l&1) and, where a head follows, the head weights (free space of the slots)
        // while the tensor core is busy.  The leading barrier makes sure no warp is still reading these slots in the
        // previous layer's epilogue (the head weights reuse the other slot).
        {
          epi_bar_sync(128 * NT);
          float* bs = vslot + (l & 1) * 256;
          for (int j = etid; j < layer_n(l); j += 128 * NT) bs[j] = __ldg(bias_all + l * 256 + j);
          if (l == 7) for (int j = etid; j < 256; j += 128 * NT) vslot[j] = __ldg(wsig_g + j);            // slot 0
          if (l == 9) for (int j = etid; j < 384; j += 128 * NT)                                          // slot 0 + top of slot 1
              vslot[j < 256 ? j : j + 128] = __ldg(wrgb_g + j);
          epi_bar_sync(128 * NT);
        }
        const float* bias = vslot + (l & 1) * 256;
        mbar_wait(bar_acc, acc_phase);   // every MMA of the layer done: accumulators final, A operand free
        acc_phase ^= 1;
        tc_fence_after();
        if (prof) args.prof[40 + l * 4 + 1] = clock64();
        // Activation dump (training / parity): layers 0..8 leave this warp's 32 rows of the tile in shared memory as the
        // next layer's A operand, so the dump is 4 TMA tensor stores per warp (one per 64-column k-block, fully
        // coalesced, no LSU traffic) instead of per-thread 16-byte stores that touch 32 lines per instruction.
        // The stores issued after the previous layer must have finished reading the tile before it is overwritten.
        const bool dump = DEBUG && args.layer_out != nullptr;
        if (dump) {
          if (lane == 0) bulk_wait_read();
          __syncwarp();
        }
        {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256);
          __nv_bfloat16* dump_row = nullptr;
          if (DEBUG) dump_row = (args.layer_out && live && l == 9) ? args.layer_out + ((size_t)l * args.n_samples + srow) * 256 : nullptr;
          uint32_t* mrow = (DEBUG && args.mask_out && live) ? args.mask_out + ((size_t)l * args.n_samples + srow) * 8 : nullptr;
          if (l == 7)      epilogue_row<1, DEBUG>(taddr, bias, vslot, a_row, r7s, eo, nullptr, mrow);
          else if (l == 8) epilogue_row<2, false>(taddr, bias, vslot, a_row, r7s, eo, nullptr);
          else if (l == 9) epilogue_row<3, DEBUG>(taddr, bias, vslot, a_row, r7s, eo, dump_row, mrow);   // no smem copy: direct stores
          else             epilogue_row<0, DEBUG>(taddr, bias, vslot, a_row, r7s, eo, nullptr, mrow);
        }
        if (l == 8) {
          // condition input for Dense_10: pos_enc(dir, 0, 4) replaces the position encoding (last used by layer 5)
          const float d0 = __ldg(args.dir + 3 * lrow), d1 = __ldg(args.dir + 3 * lrow + 1), d2 = __ldg(args.dir + 3 * lrow + 2);
          write_encoding<4>(e_blk, row, d0, d1, d2,
                            (DEBUG && args.enc_out && live) ? reinterpret_cast<uint4*>(args.enc_out + ((size_t)args.n_samples + srow) * 64) : nullptr);
        }
        if (prof) args.prof[40 + l * 4 + 2] = clock64();
        if (l == 9) {
          if (live) args.raw_out[srow] = make_float4(eo.r + headb.x, eo.g + headb.y, eo.b + headb.z, eo.sigma + headb.w);
          tc_fence_before();  // accumulators drained; the next group's layer 0 may overwrite them
        } else {
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_aready);
            const int64_t wrow0 = (group * NT + t) * TILE_M + q * 32;     // first sample row of this warp
            if (dump && wrow0 < args.n_samples) {
              const uint32_t src = sbase + SL::A_OFF + t * 4 * ABLK_BYTES + q * 4096;
#pragma unroll
              for (int kb = 0; kb < 4; ++kb) tma_store_3d(&tm_layers, src + kb * ABLK_BYTES, kb * KB, (int)wrow0, l);
              bulk_commit();
            }
          }
        }
      }
    }
    if (DEBUG && args.layer_out != nullptr && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256 * NT);
}

template <int NT, int NSTAGE, bool DEBUG>
static int launch_encmlp(const EncMlpArgs& a0, cudaStream_t st) {
  using SL = SmemLayout<NT, NSTAGE>;
  static_assert(SL::BYTES <= 232448, "shared-memory budget exceeded");
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kfn = encmlp_kernel<NT, NSTAGE, DEBUG>;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::BYTES);
    if (e != cudaSuccess) { set_error("rnerf_encmlp_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set[dev] = true;
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  EncMlpArgs a = a0;
  a.n_groups = (int)((a.n_samples + (int64_t)TILE_M * NT - 1) / ((int64_t)TILE_M * NT));
  const int grid = a.n_groups < n_sm ? a.n_groups : n_sm;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (DEBUG && a.layer_out != nullptr) {
    int rc = make_rows_tmap(&tm, a.layer_out, a.n_samples, N_MMA_LAYERS);
    if (rc) return rc;
  }
  kfn<<<grid, 64 + 128 * NT, SL::BYTES, st>>>(a, tm);
  count_launch();
  return check_launch("rnerf_encmlp_fwd");
}

}  // namespace rnerf

using namespace rnerf;

extern "C" size_t rnerf_encmlp_packed_bytes(void) { return PK_TOTAL; }

extern "C" int rnerf_encmlp_pack(const float* const* kernels, const float* const* biases, void* packed, void* stream) {
  RNERF_REQUIRE_PTR(kernels); RNERF_REQUIRE_PTR(biases); RNERF_REQUIRE_PTR(packed);
  RNERF_REQUIRE(aligned16(packed), RNERF_E_ALIGN, "rnerf_encmlp_pack: packed must be 16-byte aligned");
  PackArgs a;
  for (int i = 0; i < 12; ++i) {
    if (!kernels[i] || !biases[i]) { set_error("rnerf_encmlp_pack: null kernel/bias %d", i); return RNERF_E_NULL; }
    a.kern[i] = kernels[i];
    a.bias[i] = biases[i];
  }
  encmlp_pack_kernel<<<N_CHUNKS + 1 + PAIR_NCHUNK, 256, 0, (cudaStream_t)stream>>>(a, (uint8_t*)packed);
  count_launch();
  return check_launch("rnerf_encmlp_pack");
}

namespace rnerf { int launch_encmlp_pair(const EncMlpArgs& a0, cudaStream_t st); }

// RNERF_MLP_KERNEL=single forces the single-CTA kernel (encmlp.cu); default: CTA-pair kernel for large batches
static bool use_pair_kernel() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("RNERF_MLP_KERNEL");
    cached = (e != nullptr && strcmp(e, "single") == 0) ? 0 : 1;
  }
  return cached == 1;
}

static int encmlp_fwd_impl(const void* packed, const float* pos, const float* dir, int64_t n_samples, float* raw_out,
                           uint16_t* layer_out, void* stream, long long* prof = nullptr, uint16_t* enc_out = nullptr,
                           uint32_t* mask_out = nullptr) {
  RNERF_REQUIRE(n_samples >= 0, RNERF_E_SHAPE, "rnerf_encmlp_fwd: n_samples < 0");
  if (n_samples == 0) return 0;
  RNERF_REQUIRE_PTR(packed); RNERF_REQUIRE_PTR(pos); RNERF_REQUIRE_PTR(dir); RNERF_REQUIRE_PTR(raw_out);
  RNERF_REQUIRE(aligned16(packed) && aligned16(raw_out), RNERF_E_ALIGN, "rnerf_encmlp_fwd: packed/raw_out must be 16-byte aligned");
  RNERF_REQUIRE(layer_out == nullptr || aligned16(layer_out), RNERF_E_ALIGN, "rnerf_encmlp_fwd: layer_out must be 16-byte aligned");
  EncMlpArgs a;
  a.packed = (const uint8_t*)packed; a.pos = pos; a.dir = dir; a.n_samples = n_samples;
  a.raw_out = (float4*)raw_out; a.layer_out = (__nv_bfloat16*)layer_out; a.enc_out = (__nv_bfloat16*)enc_out; a.mask_out = mask_out; a.prof = prof; a.n_groups = 0;
  { const char* d = getenv("RNERF_PAIR_DEBUG"); a.dbg = d ? atoi(d) : 0; }
  if (prof != nullptr && layer_out == nullptr && n_samples >= 74 * 512 && getenv("RNERF_PROFILE_PAIR") != nullptr)
    return launch_encmlp_pair(a, (cudaStream_t)stream);
  const bool dbg = layer_out != nullptr || prof != nullptr;
  if (!dbg && n_samples >= 74 * 512 && use_pair_kernel()) return launch_encmlp_pair(a, (cudaStream_t)stream);
  // training forward (activations + encodings saved): the pair kernel in TRAIN mode for large batches
  if (layer_out != nullptr && enc_out != nullptr && prof == nullptr && n_samples >= 74 * 512 && use_pair_kernel())
    return launch_encmlp_pair(a, (cudaStream_t)stream);
  if (n_samples <= 148 * 128) return dbg ? launch_encmlp<1, 8, true>(a, (cudaStream_t)stream) : launch_encmlp<1, 8, false>(a, (cudaStream_t)stream);
  return dbg ? launch_encmlp<2, 4, true>(a, (cudaStream_t)stream) : launch_encmlp<2, 4, false>(a, (cudaStream_t)stream);
}

extern "C" int rnerf_encmlp_fwd(const void* packed, const float* pos, const float* dir, int64_t n_samples, float* raw_out,
                                void* stream) {
  return encmlp_fwd_impl(packed, pos, dir, n_samples, raw_out, nullptr, stream);
}

extern "C" int rnerf_encmlp_fwd_debug(const void* packed, const float* pos, const float* dir, int64_t n_samples,
                                      float* raw_out, uint16_t* layer_out, void* stream) {
  return encmlp_fwd_impl(packed, pos, dir, n_samples, raw_out, layer_out, stream);
}

// development aid (not part of the reference-facing surface): per-layer clock64 stamps of CTA 0
extern "C" int rnerf_encmlp_fwd_profile(const void* packed, const float* pos, const float* dir, int64_t n_samples,
                                        float* raw_out, long long* prof, void* stream) {
  return encmlp_fwd_impl(packed, pos, dir, n_samples, raw_out, nullptr, stream, prof);
}

// training forward: saves every layer's post-activation output (bf16 [10][M][256]) and the two encodings
// (bf16 [2][M][64]) for rnerf_mlp_dgrad / rnerf_mlp_wgrad
extern "C" int rnerf_encmlp_fwd_train(const void* packed, const float* pos, const float* dir, int64_t n_samples,
                                      float* raw_out, uint16_t* layer_out, uint16_t* enc_out, uint32_t* relu_masks, void* stream) {
  RNERF_REQUIRE_PTR(layer_out); RNERF_REQUIRE_PTR(enc_out); RNERF_REQUIRE_PTR(relu_masks);
  RNERF_REQUIRE(aligned16(enc_out) && aligned16(relu_masks), RNERF_E_ALIGN, "rnerf_encmlp_fwd_train: enc_out / relu_masks must be 16-byte aligned");
  return encmlp_fwd_impl(packed, pos, dir, n_samples, raw_out, layer_out, stream, nullptr, enc_out, relu_masks);
}
