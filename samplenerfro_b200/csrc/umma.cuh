// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and the network geometry shared by the enc+MLP kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace rnerf {

// ------------------------------------------------------------------------------------------------
// network geometry (flag defaults rnerf/utils.py:138-157; identical in every shipped config)
// ------------------------------------------------------------------------------------------------
constexpr int TILE_M = 128;
constexpr int KB = 64;                       // K elements per activation k-block (one 128-byte swizzle row)
constexpr int KCH = 32;                      // K elements per weight chunk (one 64-byte swizzle row)
constexpr int SUBS = KB / KCH;               // weight chunks per activation k-block
constexpr int NMAX = 256;                    // widest layer = N of one MMA
constexpr int SLOT_BYTES = NMAX * KCH * 2;   // 16384: one ring slot / packed chunk
constexpr int ABLK_BYTES = TILE_M * KB * 2;  // 16384: one [128 x 64] bf16 activation k-block
constexpr int N_MMA_LAYERS = 10;             // Dense_0..7, Dense_9 (bottleneck), Dense_10 (condition)
constexpr int POS_ENC = 63, DIR_ENC = 27;

// per MMA layer: number of A k-blocks taken from the activation buffer, whether the E (encoding) block is
// appended, output width, ReLU, and the Flax Dense index it implements
__host__ __device__ constexpr int layer_akb(int l) { return l == 0 ? 0 : 4; }
__host__ __device__ constexpr int layer_has_e(int l) { return (l == 0 || l == 5 || l == 9) ? 1 : 0; }
__host__ __device__ constexpr int layer_n(int l) { return l == 9 ? 128 : 256; }
__host__ __device__ constexpr int layer_relu(int l) { return l == 8 ? 0 : 1; }
__host__ __device__ constexpr int layer_dense(int l) { return l < 8 ? l : l + 1; }   // 8->Dense_9, 9->Dense_10
__host__ __device__ constexpr int layer_chunks(int l) { return (layer_akb(l) + layer_has_e(l)) * SUBS; }
__host__ __device__ constexpr int total_chunks() {
  int c = 0;
  for (int l = 0; l < N_MMA_LAYERS; ++l) c += layer_chunks(l);
  return c;
}
constexpr int N_CHUNKS = total_chunks();     // 78

// packed image: [chunks, one slot each][fp32 tail]
constexpr size_t PK_CHUNKS = 0;
constexpr size_t PK_BIAS = (size_t)N_CHUNKS * SLOT_BYTES;          // float bias[10][256]
constexpr size_t PK_WSIGMA = PK_BIAS + 10 * 256 * 4;               // float w_sigma[256] (bf16-rounded)
constexpr size_t PK_WRGB = PK_WSIGMA + 256 * 4;                    // float w_rgb[3][128] (bf16-rounded, channel-major)
constexpr size_t PK_HEADB = PK_WRGB + 3 * 128 * 4;                 // float4 (b_r, b_g, b_b, b_sigma)
// second weight image for the CTA-pair kernel (encmlp_pair.cu): 41 chunks of [N x 64] split in two N-halves
// ([N/2 x 64] SWIZZLE_128B, 16 KB each at N = 256), one per CTA of the pair
constexpr int PAIR_NCHUNK = 41;
constexpr int PAIR_HALF_BYTES = 128 * KB * 2;                       // 16384
constexpr int PAIR_CHUNK_STRIDE = 2 * PAIR_HALF_BYTES;              // 32768
constexpr size_t PK_PAIR = (PK_HEADB + 16 + 1023) / 1024 * 1024;
constexpr size_t PK_TOTAL = PK_PAIR + (size_t)PAIR_NCHUNK * PAIR_CHUNK_STRIDE;

// Weight stream of the pair kernel.  Every 64-wide k-block of a layer is one chunk; `cons` says which tile pair
// consumes it (1 = P0, 2 = P1, 3 = both).  Layers 5 and 9 read five k-blocks but only four weight slots exist, so
// their encoding block is streamed twice: first for P0 (consumed first), last for P1 (consumed last).
struct PairChunk { int8_t layer, src /* 0..3 = activation k-block, 4 = encoding block */, cons; };
#define RNERF_PAIR_STREAM                                                                                          \
  {0, 4, 3},                                                                                                       \
  {1, 0, 3}, {1, 1, 3}, {1, 2, 3}, {1, 3, 3}, {2, 0, 3}, {2, 1, 3}, {2, 2, 3}, {2, 3, 3},                          \
  {3, 0, 3}, {3, 1, 3}, {3, 2, 3}, {3, 3, 3}, {4, 0, 3}, {4, 1, 3}, {4, 2, 3}, {4, 3, 3},                          \
  {5, 4, 1}, {5, 0, 3}, {5, 1, 3}, {5, 2, 3}, {5, 3, 3}, {5, 4, 2},                                                \
  {6, 0, 3}, {6, 1, 3}, {6, 2, 3}, {6, 3, 3}, {7, 0, 3}, {7, 1, 3}, {7, 2, 3}, {7, 3, 3},                          \
  {8, 0, 3}, {8, 1, 3}, {8, 2, 3}, {8, 3, 3},                                                                      \
  {9, 4, 1}, {9, 0, 3}, {9, 1, 3}, {9, 2, 3}, {9, 3, 3}, {9, 4, 2}
__host__ __device__ constexpr int pair_layer_first(int l) { return l == 0 ? 0 : (l <= 5 ? 1 + 4 * (l - 1) : (l <= 9 ? 23 + 4 * (l - 6) : 41)); }
__host__ __device__ constexpr int pair_layer_count(int l) { return l == 0 ? 1 : ((l == 5 || l == 9) ? 6 : 4); }

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA engine bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// TMA engine tensor store shared -> global through a CUtensorMap (SASS: UTMASTG); completion tracked by bulk groups.
// The source tile must have been made visible to the async proxy (fence.proxy.async) by its writers.
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk stores have finished READING shared memory (the tile may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have completed entirely
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Host: tensor map of a bf16 activation dump [layers][M][256] (row-major) whose box is one warp's share of a
// SWIZZLE_128B k-block: 64 columns x 32 rows.  Rows beyond M are clipped by the TMA engine.  Returns 0 or a
// negative/CUDA error code (message via set_error).
int make_rows_tmap(CUtensorMap* out, const void* base, int64_t n_rows, int n_layers);

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate (SASS: UTCHMMA)
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same with each shared-memory descriptor given as (low word, high word).  The K loop of every kernel here advances a
// descriptor by adding (byte offset >> 4) to its LOW word (the 14-bit start-address field; it cannot carry for addresses
// inside the 227 KB window): one uniform add per operand and MMA instead of rebuilding the 64-bit descriptor from the
// address (shift, mask, or -- a ~12-deep dependent chain on the uniform datapath that held the single issuing thread to
// one MMA per ~145 cycles, slower than the tensor pipe executes them).
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a CONVERGED warp.  Unlike `if (lane == 0)`, elect.sync tells ptxas the branch holds exactly one thread, so
// the uniform-datapath instructions inside (UTCHMMA, UTCBAR, UBLKCP) are emitted plainly instead of each being wrapped in
// an ELECT / BRA.U.ANY retry loop.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 columns of fp32 accumulator -> 32 registers per thread (SASS: LDTM)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30) (=1, unused for
// swizzled K-major), SBO>>4 [32,46), version=1 [46,48), layout_type [61,64)).
//   A operand: K-major SWIZZLE_128B -- rows of 128 B (64 bf16), 8-row atoms 1024 B apart (layout_type 2)
//   B operand: K-major SWIZZLE_64B  -- rows of  64 B (32 bf16), 8-row atoms  512 B apart (layout_type 4)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::f16: c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// byte offset of element (row, k) inside a [rows x 64] bf16 K-major SWIZZLE_128B block (Swizzle<3,4,3>)
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}
// byte offset of element (row, k) inside a [rows x 32] bf16 K-major SWIZZLE_64B block (Swizzle<2,4,3>:
// address bits [4,6) ^= bits [7,9), i.e. the 16-byte unit index is xored with (row >> 1) & 3)
__host__ __device__ __forceinline__ uint32_t sw64_offset(int row, int k) {
  return (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((((k >> 3) ^ ((row >> 1) & 3)) & 3) << 4) + (k & 7) * 2);
}

// two fp32 -> packed bf16x2 (lo in bits [0,16)), optionally with ReLU fused into the conversion
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }


// ---- ReLU bit-masks (training) ---------------------------------------------------------------------------------------
// The dgrad chain needs, per layer, only WHETHER an activation was positive.  Reading that back from the saved bf16
// activations costs 512 B per sample and layer (3.7 GB of the 7.4 GB the dgrad kernel moved per 4096-ray step); the forward
// epilogue has the values in registers, so it emits the bits: 32 B per sample and layer.
// Format: mask [layer][row][8] uint32, word g = columns 32 g .. 32 g + 31 (layer 9, 128 columns: words 0..3): the even column
// 32 g + 2 i sits at bit i, the odd column 32 g + 2 i + 1 at bit 16 + i -- the epilogue holds a group as 16 packed bf16 pairs,
// and one word per group costs four integer ops per pair: the values are >= 0 after ReLU, (x & 0x7FFF) + 0x7FFF carries into
// bit 15 of its half exactly when the half is non-zero, and `acc = (t & 0x80008000) | (acc >> 1)` walks pair i's two flags
// down to bits i and 16 + i.  (The first format interleaved four words per 128 columns at six ops per pair; the mask work
// was a third of the training forward's instructions, profiles/r6q.)
__device__ __forceinline__ uint32_t relu_mask32(const uint32_t (&pk)[16]) {
  uint32_t acc = 0u;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t t = (pk[i] & 0x7FFF7FFFu) + 0x7FFF7FFFu;
    acc = (t & 0x80008000u) | (acc >> 1);
  }
  return acc;
}
// w[idx] = m for a loop-variant idx without sending the array to local memory (N predicated moves)
template <int N>
__device__ __forceinline__ void relu_mask_set(uint32_t (&w)[N], int idx, uint32_t m) {
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = (i == idx) ? m : w[i];
}
// reader: are columns 2 i / 2 i + 1 of the group positive?
__device__ __forceinline__ bool relu_mask_even(uint32_t word, int i) { return (word >> i) & 1u; }
__device__ __forceinline__ bool relu_mask_odd(uint32_t word, int i) { return (word >> (16 + i)) & 1u; }

struct EncMlpArgs {
  const uint8_t* packed;
  const float* pos;
  const float* dir;
  int64_t n_samples;
  float4* raw_out;
  __nv_bfloat16* layer_out;  // debug / training dump [10][M][256] or null
  __nv_bfloat16* enc_out;    // training dump of the encodings [2][M][64] (pos_enc, dir_enc) or null
  uint32_t* mask_out;        // training dump of the ReLU bit-masks [10][M][8] (relu_mask32) or null
  long long* prof;           // development aid: clock64 stamps of CTA 0, [3][10][4], or null
  int n_groups;              // ceil(n_samples / (128*NT))
  int dbg;                   // development aid (timing experiments only; results are wrong when set)
};

// pos_enc(x, 0, L) of one 3-vector into columns [0, 3+6L) of this thread's row of a swizzled k-block; the
// remaining columns up to 64 are zero.  Feature order (non-legacy): x, sin(2^k x) k-major, sin(2^k x + pi/2).
//
// sin/cos(2^k x) come from one accurate sincosf(x) per channel followed by the double-angle recurrence
// s' = 2 s c, c' = 1 - 2 s^2 (3 flops per octave instead of a full sinf).  The absolute error doubles per octave:
// <= 2^9 * 1.2e-7 = 6e-5 at the top position octave -- below the reference's own error in the same feature
// (it evaluates cos as sin(fl32(2^k x + pi/2)), off by up to ulp(3072)/2 = 1.2e-4) and ~30x below the bf16
// quantisation step (2^-9 relative) the value is rounded to when it becomes an MMA operand.
template <int L>
__device__ __forceinline__ void write_encoding(uint8_t* blk, int row, float x0, float x1, float x2,
                                               uint4* __restrict__ gout = nullptr /* optional [8] copy of the row to global */) {
  constexpr int NF = 3 + 6 * L;
  float feat[64];
  feat[0] = x0; feat[1] = x1; feat[2] = x2;
  const float xs[3] = {x0, x1, x2};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float s, c;
    sincosf(xs[ch], &s, &c);
#pragma unroll
    for (int k = 0; k < L; ++k) {
      feat[3 + 3 * k + ch] = s;
      feat[3 + 3 * L + 3 * k + ch] = c;
      const float s2 = s + s;
      const float ns = s2 * c;
      c = fmaf(-s2, s, 1.f);
      s = ns;
    }
  }
#pragma unroll
  for (int f = NF; f < 64; ++f) feat[f] = 0.f;
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {  // 8 columns (16 bytes) at a time
    uint4 o = make_uint4(pack_bf16(feat[c8 * 8 + 0], feat[c8 * 8 + 1]), pack_bf16(feat[c8 * 8 + 2], feat[c8 * 8 + 3]),
                         pack_bf16(feat[c8 * 8 + 4], feat[c8 * 8 + 5]), pack_bf16(feat[c8 * 8 + 6], feat[c8 * 8 + 7]));
    *reinterpret_cast<uint4*>(blk + sw128_offset(row, c8 * 8)) = o;
    if (gout != nullptr) gout[c8] = o;
  }
}

// 2 x fp32 packed add (SASS: FADD2)
__device__ __forceinline__ void add2(uint32_t& x0, uint32_t& x1, float b0, float b1) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(x0), "r"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(x0), "=r"(x1) : "l"(c));
}


struct EpiOut { float sigma, r, g, b; };

}  // namespace rnerf
