// Encoding-fused radiance MLP, CTA-pair version (tcgen05 cta_group::2) -- the production forward for large batches.
//
// Same arithmetic as encmlp.cu (pos_enc + NerfMLP, rnerf/model_utils.py:187-214, 30-90); what changes is the
// schedule, built so that the epilogue of one tile pair always runs under the MMAs of the other:
//
//   * two CTAs of a cluster (an SM pair) issue ONE tcgen05.mma.cta_group::2 of M = 256 (128 rows from each CTA),
//     N = 256; each CTA stages only its N-half of the weights ([128 x 64] per k-block, 16 KB), so a whole layer
//     (4 k-blocks) is resident in 64 KB per CTA and is used twice -- by tile pair P0, then by tile pair P1 --
//     before the slot is refilled with the next layer's k-block.  L2->SM weight traffic per sample is a quarter
//     of the single-tile design, shared-memory operand traffic per MMA is 8 KB per CTA per 128 cycles.
//   * each CTA owns two 128-row tiles (T0 of P0, T1 of P1); accumulators: 2 x 256 TMEM columns.
//   * issue order per layer: P0's K-loop, commit -> acc[0]; P1's K-loop, commit -> acc[1].  While P1's MMAs run,
//     P0's eight epilogue warps (two warpgroups, 128 columns each) turn acc[0] into the next layer's A operand,
//     and vice versa: the tensor pipe never waits for an epilogue as long as it takes < ~2000 cycles.
//
// Cross-CTA protocol (leader = cluster rank 0 issues every MMA):
//   full[s]   (local)   this CTA's half of weight slot s has landed (TMA complete_tx); on the LEADER the barrier takes a
//                       second arrival, the peer's "my half has landed" relayed by the peer's idle warp 1 (remote arrive),
//                       so the MMA issuer waits once per chunk for both halves
//   empty[s]  (both)    multicast tcgen05.commit after the slot's last consumer
//   acc[j]    (both)    multicast tcgen05.commit after tile pair j's K-loop
//   aready[j] (leader)  16 arrivals: the 8 epilogue warps of tile j in both CTAs (the peer's arrive remotely)
#include <stdlib.h>
#include "pair.cuh"

namespace rnerf {

constexpr int PAIR_NSLOT = 4;
constexpr int PAIR_THREADS = 64 + 512;   // producer, MMA/relay, 16 epilogue warps

struct PairSmem {
  static constexpr uint32_t A_OFF = 0;                               // [2 tiles][4][16 KB]
  static constexpr uint32_t E_OFF = A_OFF + 2 * 4 * ABLK_BYTES;      // [2 tiles][16 KB]
  static constexpr uint32_t W_OFF = E_OFF + 2 * ABLK_BYTES;          // [4 slots][16 KB]
  static constexpr uint32_t V_OFF = W_OFF + PAIR_NSLOT * PAIR_HALF_BYTES;   // [2 tiles][256] fp32 bias slot
  static constexpr uint32_t BAR_OFF = V_OFF + 2 * 1024;
  static constexpr uint32_t N_BARS = 3 * PAIR_NSLOT + 4;             // full, pfull, empty, acc[2], aready[2]
  static constexpr uint32_t TMEM_SLOT = BAR_OFF + N_BARS * 8;
  static constexpr uint32_t BYTES = TMEM_SLOT + 16;
};
static_assert(PairSmem::BYTES <= 232448, "pair kernel exceeds the shared-memory budget");

__constant__ PairChunk c_pair_stream[PAIR_NCHUNK] = {RNERF_PAIR_STREAM};

// Epilogue of one layer for one row and one 128-column half (64 columns for the condition layer).
//   KIND 0: ReLU, write A   1: + sigma partial   2: no activation, write A   3: ReLU, rgb partial only
template <int KIND, bool DUMP = false>
__device__ __forceinline__ void pair_epilogue(uint32_t taddr, const float* __restrict__ bias, const float* __restrict__ hw,
                                              uint8_t* __restrict__ a_row, uint32_t r7s, int col0, EpiOut& o,
                                              __nv_bfloat16* __restrict__ dump_row = nullptr /* KIND 3 in training mode */,
                                              uint32_t* __restrict__ mask_row = nullptr /* training: this row's 4 mask words */) {
  constexpr int NCG = (KIND == 3) ? 2 : 4;
  uint32_t mw[4] = {0u, 0u, 0u, 0u};      // ReLU bit-mask word of each of this thread's 32-column groups (umma.cuh)
#ifdef RNERF_PAIR_PIPELINED_LD
  // two TMEM loads in flight: group cg+1 is fetched while group cg is converted and stored
  uint32_t vv[2][32];
  tmem_ld32(taddr, vv[0]);
#pragma unroll
  for (int cg = 0; cg < NCG; ++cg) {
    const int c0 = col0 + cg * 32;          // first of these 32 columns within the layer
    uint32_t (&v)[32] = vv[cg & 1];
    tmem_ld_wait();
    if (cg + 1 < NCG) tmem_ld32(taddr + (cg + 1) * 32, vv[(cg + 1) & 1]);
#else
#pragma unroll 1
  for (int cg = 0; cg < NCG; ++cg) {
    const int c0 = col0 + cg * 32;          // first of these 32 columns within the layer
    uint32_t v[32];
    tmem_ld32(taddr + cg * 32, v);
    tmem_ld_wait();
#endif
    uint32_t pk[16];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias + c0 + j4 * 4);
      add2(v[4 * j4 + 0], v[4 * j4 + 1], b4.x, b4.y);
      add2(v[4 * j4 + 2], v[4 * j4 + 3], b4.z, b4.w);
      const float f0 = __uint_as_float(v[4 * j4 + 0]), f1 = __uint_as_float(v[4 * j4 + 1]);
      const float f2 = __uint_as_float(v[4 * j4 + 2]), f3 = __uint_as_float(v[4 * j4 + 3]);
      if (KIND == 2) { pk[2 * j4] = pack_bf16(f0, f1);      pk[2 * j4 + 1] = pack_bf16(f2, f3); }
      else           { pk[2 * j4] = pack_bf16_relu(f0, f1); pk[2 * j4 + 1] = pack_bf16_relu(f2, f3); }
    }
    if (DUMP && KIND != 2) relu_mask_set(mw, cg, relu_mask32(pk));
    if (KIND == 1) {  // sigma head (Dense_8): hw = w_sigma[256]
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 w4 = *reinterpret_cast<const float4*>(hw + c0 + j4 * 4);
        o.sigma = fmaf(bf16_lo(pk[2 * j4]), w4.x, o.sigma);
        o.sigma = fmaf(bf16_hi(pk[2 * j4]), w4.y, o.sigma);
        o.sigma = fmaf(bf16_lo(pk[2 * j4 + 1]), w4.z, o.sigma);
        o.sigma = fmaf(bf16_hi(pk[2 * j4 + 1]), w4.w, o.sigma);
      }
    }
    if (KIND == 3 && DUMP) {   // the condition layer leaves no shared-memory copy: its saved activations go out directly
      if (dump_row != nullptr) {
        uint4* dst = reinterpret_cast<uint4*>(dump_row + c0);
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
    }
    if (KIND == 3) {  // rgb head (Dense_11): hw = w_rgb[3][128]
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 wr = *reinterpret_cast<const float4*>(hw + c0 + j4 * 4);
        const float4 wg = *reinterpret_cast<const float4*>(hw + 128 + c0 + j4 * 4);
        const float4 wb = *reinterpret_cast<const float4*>(hw + 256 + c0 + j4 * 4);
        const float h0 = bf16_lo(pk[2 * j4]), h1 = bf16_hi(pk[2 * j4]), h2 = bf16_lo(pk[2 * j4 + 1]), h3 = bf16_hi(pk[2 * j4 + 1]);
        o.r = fmaf(h0, wr.x, o.r); o.r = fmaf(h1, wr.y, o.r); o.r = fmaf(h2, wr.z, o.r); o.r = fmaf(h3, wr.w, o.r);
        o.g = fmaf(h0, wg.x, o.g); o.g = fmaf(h1, wg.y, o.g); o.g = fmaf(h2, wg.z, o.g); o.g = fmaf(h3, wg.w, o.g);
        o.b = fmaf(h0, wb.x, o.b); o.b = fmaf(h1, wb.y, o.b); o.b = fmaf(h2, wb.z, o.b); o.b = fmaf(h3, wb.w, o.b);
      }
    } else {
      // next layer's A operand: columns c0..c0+31 -> k-block c0/64, 16-byte units ((c0/32)&1)*4 .. +3, xor (row&7)
      uint8_t* blk = a_row + (c0 >> 6) * ABLK_BYTES;
      const uint32_t u0 = (uint32_t)((c0 >> 5) & 1) * 4;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(blk + (((u0 + c) << 4) ^ r7s)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
    }
  }
  if (DUMP && KIND != 2 && mask_row != nullptr) {
    // (mask_row points at this warpgroup's first word: 4 groups of 32 columns, or 2 for the condition layer's 64)
    if (KIND == 3) *reinterpret_cast<uint2*>(mask_row) = make_uint2(mw[0], mw[1]);
    else *reinterpret_cast<uint4*>(mask_row) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
  }
}

// TRAIN: additionally saves what the backward kernels need -- every layer's post-activation output (TMA tensor stores
// of the swizzled shared-memory tiles, 2 k-blocks per warp and layer; the condition layer by direct stores) and the two
// encodings -- into args.layer_out [10][M][256] / args.enc_out [2][M][64] (bf16).
template <bool TRAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1) encmlp_pair_kernel(const EncMlpArgs args,
                                                                                                const __grid_constant__ CUtensorMap tm_layers) {
  using SL = PairSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  auto bar_full = [&](int s) { return sbase + SL::BAR_OFF + 8u * s; };
  auto bar_pfull = [&](int s) { return sbase + SL::BAR_OFF + 8u * (PAIR_NSLOT + s); };
  auto bar_empty = [&](int s) { return sbase + SL::BAR_OFF + 8u * (2 * PAIR_NSLOT + s); };
  auto bar_acc = [&](int j) { return sbase + SL::BAR_OFF + 8u * (3 * PAIR_NSLOT + j); };
  auto bar_aready = [&](int j) { return sbase + SL::BAR_OFF + 8u * (3 * PAIR_NSLOT + 2 + j); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SL::TMEM_SLOT);

  if (threadIdx.x == 0) {
    // full[s] of the LEADER completes on two arrivals: its own producer's arrive.expect_tx (+ the bytes) and the peer's
    // relayed "my half has landed" -- one wait per chunk on the MMA issuer's critical path instead of two (~150 cycles each)
    for (int s = 0; s < PAIR_NSLOT; ++s) { mbar_init(bar_full(s), leader ? 2 : 1); mbar_init(bar_pfull(s), 1); mbar_init(bar_empty(s), 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(bar_acc(j), 1); mbar_init(bar_aready(j), 16); }
    fence_barrier_init();
  }
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / multicast commit
  if (warp == 1) {
    tmem_alloc2(sbase + SL::TMEM_SLOT, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_pairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
  const int my_groups = (args.n_groups > pair) ? (args.n_groups - pair + n_pairs - 1) / n_pairs : 0;

  if (warp == 0) {
    // ===================== weight producer: this CTA's N-half of every chunk of the stream =====================
    if (lane == 0) {
      uint32_t c = 0;   // running chunk counter: slot = c % 4, phase = (c / 4) & 1
      for (int g = 0; g < my_groups; ++g) {
        for (int i = 0; i < PAIR_NCHUNK; ++i, ++c) {
          const int s = c % PAIR_NSLOT;
          const uint32_t ph = (c / PAIR_NSLOT) & 1;
          const uint32_t bytes = (uint32_t)(layer_n(c_pair_stream[i].layer) / 2) * KB * 2;
          mbar_wait_cluster(bar_empty(s), ph ^ 1);
          if (args.prof != nullptr && blockIdx.x == 0 && g == 1 && c_pair_stream[i].layer == 3)
            args.prof[256 + (i - pair_layer_first(3))] = clock64();     // slot free -> TMA issued (leader CTA)
          mbar_arrive_expect_tx(bar_full(s), bytes);
          tma_bulk_g2s(sbase + SL::W_OFF + s * PAIR_HALF_BYTES,
                       args.packed + PK_PAIR + (size_t)i * PAIR_CHUNK_STRIDE + rank * PAIR_HALF_BYTES, bytes, bar_full(s));
        }
      }
    }
  } else if (warp == 1) {
    // one elected lane of the (converged) warp: elect.sync lets ptxas emit the uniform-datapath MMA issue without the
    // per-instruction ELECT / BRA.U.ANY retry loop it wraps around UTCHMMA in code it must assume divergent
    if (elect_one_sync()) {
      if (!leader) {
        // ===================== peer: relay "my half of slot s has landed" to the leader =====================
        uint32_t c = 0;
        for (int g = 0; g < my_groups; ++g) {
          for (int i = 0; i < PAIR_NCHUNK; ++i, ++c) {
            const int s = c % PAIR_NSLOT;
            mbar_wait(bar_full(s), (c / PAIR_NSLOT) & 1);
            mbar_arrive_remote(mapa_shared(bar_full(s), 0));
          }
        }
      } else {
        // ===================== leader: MMA issuer for the pair =====================
        uint32_t ar_phase[2] = {0, 0};
        uint32_t cbase = 0;   // chunk counter at the start of the current layer
        for (int g = 0; g < my_groups; ++g) {
          for (int l = 0; l < N_MMA_LAYERS; ++l) {
            const int first = pair_layer_first(l), cnt = pair_layer_count(l);
            const uint32_t idesc = make_idesc(2 * TILE_M, layer_n(l));
            for (int j = 0; j < 2; ++j) {
              const bool prof = args.prof != nullptr && blockIdx.x == 0 && g == 1;
              if (prof) args.prof[(l * 2 + j) * 4 + 0] = clock64();
              mbar_wait_cluster(bar_aready(j), ar_phase[j]);   // A operand of layer l ready in both CTAs, acc[j] drained
              ar_phase[j] ^= 1;
              tc_fence_after();
              if (prof) args.prof[(l * 2 + j) * 4 + 1] = clock64();
              if (prof && (l == 2 || l == 3)) {
                unsigned long long gt;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                args.prof[192 + (l - 2) * 2 + j] = (long long)gt;     // aready[j] observed for layer l
              }
              bool started = false;
              for (int i = 0; i < cnt; ++i) {
                const PairChunk ch = c_pair_stream[first + i];
                if (!(ch.cons & (1 << j))) continue;
                const uint32_t c = cbase + i;
                const int s = c % PAIR_NSLOT;
                const uint32_t ph = (c / PAIR_NSLOT) & 1;
                const bool first_consumer = (ch.cons == 3) ? (j == 0) : true;
                const bool last_consumer = (ch.cons == 3) ? (j == 1) : true;
                if (first_consumer) {
                  const bool pw = prof && l == 3 && j == 0;     // layer 3, P0: stamps for each of the 4 chunks
                  if (pw) args.prof[240 + i * 4 + 0] = clock64();
                  if (!(args.dbg & 4)) mbar_wait_cluster(bar_full(s), ph);     // both halves of the slot (see the barrier init)
                  if (pw) args.prof[240 + i * 4 + 1] = clock64();
                  if (pw) args.prof[240 + i * 4 + 2] = clock64();
                  tc_fence_after();
                }
                const uint32_t a_addr = (ch.src < 4) ? (sbase + SL::A_OFF + (j * 4 + ch.src) * ABLK_BYTES)
                                                     : (sbase + SL::E_OFF + j * ABLK_BYTES);
                const uint32_t b_addr = sbase + SL::W_OFF + s * PAIR_HALF_BYTES;
                const uint32_t d_addr = tmem_base + (uint32_t)(j * 256);
                const uint64_t a_desc = make_sw128_desc(a_addr), b_desc = make_sw128_desc(b_addr);
                const uint32_t a_lo = (uint32_t)a_desc, b_lo = (uint32_t)b_desc, desc_hi = (uint32_t)(a_desc >> 32);
#pragma unroll
                for (int ks = 0; ks < KB / 16; ++ks)      // +32 bytes of K per step = +2 in the descriptor's address field
                  umma2_bf16_lohi(d_addr, a_lo + 2u * ks, b_lo + 2u * ks, desc_hi, idesc, (started || ks > 0) ? 1u : 0u);
                started = true;
                if (last_consumer) umma2_commit_mc(bar_empty(s));   // both CTAs may refill their half of the slot
              }
              umma2_commit_mc(bar_acc(j));                          // accumulators of tile pair j final in both CTAs
              if (prof) args.prof[(l * 2 + j) * 4 + 2] = clock64();
              if (prof && l == 2) {
                unsigned long long gt;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                args.prof[196 + j] = (long long)gt;                   // K-loop of layer 2 issued + committed
              }
            }
            cbase += cnt;
          }
        }
      }
    }
  } else {
    // ===================== encoder / epilogue: 8 warps per tile, 128 columns per warpgroup =====================
    const int ew = warp - 2;
    const int t = ew >> 3;                       // tile (pair) index j
    const int half = (ew >> 2) & 1;              // column half handled by this warpgroup
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int ttid = threadIdx.x - 64 - t * 256; // 0..255 within the tile's epilogue threads
    uint8_t* a_blk = smem + SL::A_OFF + t * 4 * ABLK_BYTES;
    uint8_t* e_blk = smem + SL::E_OFF + t * ABLK_BYTES;
    uint8_t* a_row = a_blk + (row >> 3) * 1024 + (row & 7) * 128;
    const uint32_t r7s = (uint32_t)(row & 7) << 4;
    float* bias_s = reinterpret_cast<float*>(smem + SL::V_OFF) + t * 256;
    float* e_scratch = reinterpret_cast<float*>(e_blk);            // free between layer 5's MMAs and the dir-enc write
    float* a_scratch = reinterpret_cast<float*>(a_blk);            // free once layer 9's MMAs have completed
    const float* bias_all = reinterpret_cast<const float*>(args.packed + PK_BIAS);
    const float* wsig_g = reinterpret_cast<const float*>(args.packed + PK_WSIGMA);
    const float* wrgb_g = reinterpret_cast<const float*>(args.packed + PK_WRGB);
    const float4 headb = __ldg(reinterpret_cast<const float4*>(args.packed + PK_HEADB));
    const uint32_t aready_remote = leader ? 0u : mapa_shared(bar_aready(t), 0);
    uint32_t acc_phase = 0;
    const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256);

    auto signal_ready = [&]() {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive_cluster_local(bar_aready(t)); else mbar_arrive_remote(aready_remote);
      }
    };

    for (int g = 0; g < my_groups; ++g) {
      const int64_t group = (int64_t)pair + (int64_t)g * n_pairs;
      const int64_t srow = group * 512 + t * 256 + (int64_t)rank * 128 + row;
      const bool live = srow < args.n_samples;
      const int64_t lrow = live ? srow : (args.n_samples - 1);
      if (half == 0) {
        const float p0 = __ldg(args.pos + 3 * lrow), p1 = __ldg(args.pos + 3 * lrow + 1), p2 = __ldg(args.pos + 3 * lrow + 2);
        write_encoding<10>(e_blk, row, p0, p1, p2,      // layer-0 A operand: pos_enc(pos, 0, 10)
                           (TRAIN && live) ? reinterpret_cast<uint4*>(args.enc_out + (size_t)srow * 64) : nullptr);
      }
      signal_ready();

      EpiOut eo = {0.f, 0.f, 0.f, 0.f};
      for (int l = 0; l < N_MMA_LAYERS; ++l) {
        // stage this layer's bias (and the sigma-head weights) while this tile's MMAs run
        tile_bar_sync(t);                                 // nobody still reads the slot from the previous layer
        if (ttid < layer_n(l)) bias_s[ttid] = __ldg(bias_all + l * 256 + ttid);
        if (l == 7) e_scratch[ttid] = __ldg(wsig_g + ttid);
        tile_bar_sync(t);
        const bool prof = args.prof != nullptr && blockIdx.x == 0 && g == 1 && half == 0 && q == 0 && lane == 0;
        if (prof) args.prof[80 + (l * 2 + t) * 4 + 0] = clock64();
        mbar_wait_cluster(bar_acc(t), acc_phase);         // this tile pair's K-loop is complete
        acc_phase ^= 1;
        tc_fence_after();
        if (prof) args.prof[80 + (l * 2 + t) * 4 + 1] = clock64();
        if (args.prof != nullptr && (blockIdx.x >> 1) == 0 && g == 1 && l == 2 && lane == 0) {
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          args.prof[200 + rank * 16 + ew] = (long long)gt;            // acc[t] of layer 2 observed by this warp
        }
        if (TRAIN) {           // this warp's activation stores of the previous layer have finished reading the tile
          if (lane == 0) bulk_wait_read();
          __syncwarp();
        }
        if (l == 9) {
          if (TRAIN) tile_bar_sync(t);                    // ... and so have everybody else's: the scratch below overlaps their rows
          // rgb-head weights into the (now free) activation buffer
          for (int i = ttid; i < 384; i += 256) a_scratch[i] = __ldg(wrgb_g + i);
          tile_bar_sync(t);
          pair_epilogue<3, TRAIN>(taddr_row + half * 64, bias_s, a_scratch, a_row, r7s, half * 64, eo,
                                  (TRAIN && live) ? args.layer_out + ((size_t)9 * args.n_samples + srow) * 256 : nullptr,
                                  (TRAIN && live && args.mask_out) ? args.mask_out + ((size_t)9 * args.n_samples + srow) * 8 + half * 2 : nullptr);
          // combine the two column halves: half 1 parks its partial sums, half 0 adds and writes the row
          float4* part = reinterpret_cast<float4*>(a_scratch + 1024);
          if (half == 1) part[row] = make_float4(eo.r, eo.g, eo.b, eo.sigma);
          tile_bar_sync(t);
          if (half == 0 && live) {
            const float4 p = part[row];
            args.raw_out[srow] = make_float4(eo.r + p.x + headb.x, eo.g + p.y + headb.y, eo.b + p.z + headb.z,
                                             eo.sigma + p.w + headb.w);
          }
          tc_fence_before();   // accumulators drained; signalled with the next group's encoding (or never)
        } else {
          const uint32_t ta = taddr_row + half * 128;
          uint32_t* mrow = (TRAIN && live && args.mask_out) ? args.mask_out + ((size_t)l * args.n_samples + srow) * 8 + half * 4 : nullptr;
          if (l == 7)      pair_epilogue<1, TRAIN>(ta, bias_s, e_scratch, a_row, r7s, half * 128, eo, nullptr, mrow);
          else if (l == 8) pair_epilogue<2>(ta, bias_s, nullptr, a_row, r7s, half * 128, eo);
          else             pair_epilogue<0, TRAIN>(ta, bias_s, nullptr, a_row, r7s, half * 128, eo, nullptr, mrow);
          if (l == 8 && half == 0) {
            // condition input for Dense_10: pos_enc(dir, 0, 4) replaces the position encoding
            const float d0 = __ldg(args.dir + 3 * lrow), d1 = __ldg(args.dir + 3 * lrow + 1), d2 = __ldg(args.dir + 3 * lrow + 2);
            write_encoding<4>(e_blk, row, d0, d1, d2,
                              (TRAIN && live) ? reinterpret_cast<uint4*>(args.enc_out + ((size_t)args.n_samples + srow) * 64) : nullptr);
          }
          if (prof) args.prof[80 + (l * 2 + t) * 4 + 2] = clock64();
          if (args.prof != nullptr && (blockIdx.x >> 1) == 0 && g == 1 && l == 2 && lane == 0) {
            // cross-SM view (globaltimer, ns): when does every epilogue warp of both CTAs finish layer 2?
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            args.prof[160 + rank * 16 + ew] = (long long)gt;
          }
          signal_ready();
          if (TRAIN && lane == 0) {
            const int64_t wrow0 = group * 512 + t * 256 + (int64_t)rank * 128 + q * 32;     // first sample row of this warp
            if (wrow0 < args.n_samples) {
              const uint32_t src = sbase + SL::A_OFF + t * 4 * ABLK_BYTES + q * 4096;
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
                tma_store_3d(&tm_layers, src + (half * 2 + kb) * ABLK_BYTES, (half * 2 + kb) * KB, (int)wrow0, l);
              bulk_commit();
            }
          }
        }
      }
    }
    if (TRAIN && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA exits (or frees TMEM) while its peer may still address it
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

int launch_encmlp_pair(const EncMlpArgs& a0, cudaStream_t st) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(encmlp_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PairSmem::BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(encmlp_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PairSmem::BYTES);
    if (e != cudaSuccess) { set_error("rnerf_encmlp_fwd: cudaFuncSetAttribute(pair): %s", cudaGetErrorString(e)); return (int)e; }
    attr_set[dev] = true;
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  EncMlpArgs a = a0;
  a.n_groups = (int)((a.n_samples + 511) / 512);
  int pairs = n_sm / 2;
  if (a.n_groups < pairs) pairs = a.n_groups;
  if (const char* lim = getenv("RNERF_PAIR_LIMIT")) {       // development aid: run on fewer SM pairs
    const int l = atoi(lim);
    if (l > 0 && l < pairs) pairs = l;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (a.layer_out != nullptr) {                        // training forward: activations + encodings are saved
    if (a.enc_out == nullptr) { set_error("rnerf_encmlp_fwd(pair): layer_out given without enc_out"); return RNERF_E_NULL; }
    int rc = make_rows_tmap(&tm, a.layer_out, a.n_samples, N_MMA_LAYERS);
    if (rc) return rc;
    encmlp_pair_kernel<true><<<2 * pairs, PAIR_THREADS, PairSmem::BYTES, st>>>(a, tm);
  } else {
    encmlp_pair_kernel<false><<<2 * pairs, PAIR_THREADS, PairSmem::BYTES, st>>>(a, tm);
  }
  count_launch();
  return check_launch("rnerf_encmlp_fwd(pair)");
}

}  // namespace rnerf
