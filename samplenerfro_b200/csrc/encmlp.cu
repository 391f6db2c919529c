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
// Tile geometry: M = 128 rows per tile (UMMA_M = 128, cta_group::1), N = 128 per MMA (one "n-half" of a
// 256-wide layer), K-chunk = 64 bf16 = one 128-byte swizzle row (4 x UMMA_K=16).
#include <cuda_bf16.h>
#include "common.cuh"

namespace rnerf {

// ------------------------------------------------------------------------------------------------
// network geometry (flag defaults rnerf/utils.py:138-157; identical in every shipped config)
// ------------------------------------------------------------------------------------------------
constexpr int TILE_M = 128;
constexpr int KB = 64;                     // K elements per chunk / swizzle row
constexpr int NH = 128;                    // N per MMA
constexpr int CHUNK_BYTES = NH * KB * 2;   // 16384
constexpr int ABLK_BYTES = TILE_M * KB * 2;  // 16384: one [128 x 64] bf16 activation k-block
constexpr int N_MMA_LAYERS = 10;           // Dense_0..7, Dense_9 (bottleneck), Dense_10 (condition)
constexpr int N_CHUNKS = 73;
constexpr int POS_ENC = 63, DIR_ENC = 27;

// per MMA layer: number of A k-blocks taken from the activation buffer, whether the E (encoding) block is
// appended, number of n-halves, ReLU, and the Flax Dense index it implements
__host__ __device__ constexpr int layer_akb(int l) { return l == 0 ? 0 : 4; }
__host__ __device__ constexpr int layer_has_e(int l) { return (l == 0 || l == 5 || l == 9) ? 1 : 0; }
__host__ __device__ constexpr int layer_nh(int l) { return l == 9 ? 1 : 2; }
__host__ __device__ constexpr int layer_relu(int l) { return l == 8 ? 0 : 1; }
__host__ __device__ constexpr int layer_dense(int l) { return l < 8 ? l : l + 1; }   // 8->Dense_9, 9->Dense_10
__host__ __device__ constexpr int layer_chunks(int l) { return (layer_akb(l) + layer_has_e(l)) * layer_nh(l); }

// packed image: [chunks][fp32 tail]
constexpr size_t PK_CHUNKS = 0;
constexpr size_t PK_BIAS = (size_t)N_CHUNKS * CHUNK_BYTES;         // float bias[10][256]
constexpr size_t PK_WSIGMA = PK_BIAS + 10 * 256 * 4;               // float w_sigma[256] (bf16-rounded)
constexpr size_t PK_WRGB = PK_WSIGMA + 256 * 4;                    // float4 w_rgb[128] (bf16-rounded, .w = 0)
constexpr size_t PK_HEADB = PK_WRGB + 128 * 16;                    // float4 (b_r, b_g, b_b, b_sigma)
constexpr size_t PK_TOTAL = PK_HEADB + 16;

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

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 B, 8-row atoms 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30) (=1, unused), SBO>>4 [32,46) (=64),
//  version=1 [46,48), layout_type=2 (SWIZZLE_128B) [61,64))
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::f16: c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// byte offset of element (row, k) inside a [rows x 64] bf16 K-major SWIZZLE_128B block
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------
// weight packing: Flax [in,out] fp32 kernels -> 73 pre-swizzled bf16 [128 n x 64 k] chunks + fp32 tail
// ------------------------------------------------------------------------------------------------
struct PackArgs {
  const float* kern[12];
  const float* bias[12];
};

__global__ void __launch_bounds__(256) encmlp_pack_kernel(PackArgs a, uint8_t* __restrict__ packed) {
  // chunk enumeration must match the MMA issuer: for layer, for n-half, for k-block (A blocks then E)
  int c = blockIdx.x;
  if (c < N_CHUNKS) {
    int l = 0, rem = c;
    while (rem >= layer_chunks(l)) { rem -= layer_chunks(l); ++l; }
    const int nkb = layer_akb(l) + layer_has_e(l);
    const int nh = rem / nkb, kbi = rem % nkb;
    const bool is_e = kbi >= layer_akb(l);
    const int dense = layer_dense(l);
    const int in_dim = (l == 0) ? POS_ENC : (l == 5 ? 256 + POS_ENC : (l == 9 ? 256 + DIR_ENC : 256));
    const int out_dim = (l == 9) ? 128 : 256;
    const float* W = a.kern[dense];
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(packed + PK_CHUNKS + (size_t)c * CHUNK_BYTES);
    for (int e = threadIdx.x; e < NH * KB; e += blockDim.x) {
      const int n = e / KB, k = e % KB;
      const int kk = is_e ? (layer_akb(l) * KB + k) : (kbi * KB + k);
      const int col = nh * NH + n;
      float v = (kk < in_dim && col < out_dim) ? W[(size_t)kk * out_dim + col] : 0.f;
      dst[sw128_offset(n, k) / 2] = __float2bfloat16_rn(v);
    }
  } else {
    float* bias = reinterpret_cast<float*>(packed + PK_BIAS);
    for (int e = threadIdx.x; e < 10 * 256; e += blockDim.x) {
      const int l = e / 256, j = e % 256;
      const int out_dim = (l == 9) ? 128 : 256;
      bias[e] = j < out_dim ? a.bias[layer_dense(l)][j] : 0.f;
    }
    float* ws = reinterpret_cast<float*>(packed + PK_WSIGMA);
    for (int e = threadIdx.x; e < 256; e += blockDim.x) ws[e] = bf16_round(a.kern[8][e]);  // Dense_8: [256,1]
    float4* wr = reinterpret_cast<float4*>(packed + PK_WRGB);
    for (int e = threadIdx.x; e < 128; e += blockDim.x)                                      // Dense_11: [128,3]
      wr[e] = make_float4(bf16_round(a.kern[11][e * 3]), bf16_round(a.kern[11][e * 3 + 1]),
                          bf16_round(a.kern[11][e * 3 + 2]), 0.f);
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
  static constexpr uint32_t BAR_OFF = W_OFF + NSTAGE * CHUNK_BYTES;      // mbarriers (8 B each)
  static constexpr uint32_t N_BARS = 2 * NSTAGE + 3;                     // full[], empty[], acc[2], a_ready
  static constexpr uint32_t TMEM_SLOT = BAR_OFF + N_BARS * 8;
  static constexpr uint32_t BYTES = TMEM_SLOT + 16;
  static constexpr uint32_t ALLOC = BYTES + 1024;                        // slack for manual 1024-B alignment
};

struct EncMlpArgs {
  const uint8_t* packed;
  const float* pos;
  const float* dir;
  int64_t n_samples;
  float4* raw_out;
  __nv_bfloat16* layer_out;  // debug dump [10][M][256] or null
  long long* prof;           // development aid: clock64 stamps of CTA 0, [2 roles][10 layers][4], or null
  int n_groups;              // ceil(n_samples / (128*NT))
};

// pos_enc(x, 0, L) of one 3-vector into columns [0, 3+6L) of this thread's row of a swizzled k-block; the
// remaining columns up to 64 are zero.  Feature order (non-legacy): x, sin(2^k x) k-major, sin(2^k x + pi/2).
template <int L>
__device__ __forceinline__ void write_encoding(uint8_t* blk, int row, float x0, float x1, float x2) {
  constexpr int NF = 3 + 6 * L;
  const float xs[3] = {x0, x1, x2};
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {  // 8 columns (16 bytes) at a time
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int f = c8 * 8 + j;
      if (f < 3) {
        v[j] = xs[f];
      } else if (f < NF) {
        const int q = (f - 3) % (3 * L), k = q / 3, ch = q % 3;
        float xb = mul(xs[ch], (float)(1 << k));
        if (f >= 3 + 3 * L) xb = add(xb, 1.57079632679489661923f);
        v[j] = sinf(xb);
      } else {
        v[j] = 0.f;
      }
    }
    uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(blk + sw128_offset(row, c8 * 8)) = o;
  }
}

template <int NT, int NSTAGE>
__global__ void __launch_bounds__(64 + 128 * NT, 1) encmlp_kernel(const EncMlpArgs args) {
  using SL = SmemLayout<NT, NSTAGE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  auto bar_full = [&](int s) { return sbase + SL::BAR_OFF + 8u * s; };
  auto bar_empty = [&](int s) { return sbase + SL::BAR_OFF + 8u * (NSTAGE + s); };
  auto bar_acc = [&](int h) { return sbase + SL::BAR_OFF + 8u * (2 * NSTAGE + h); };
  const uint32_t bar_aready = sbase + SL::BAR_OFF + 8u * (2 * NSTAGE + 2);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SL::TMEM_SLOT);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_acc(0), 1);
    mbar_init(bar_acc(1), 1);
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
    // ===================== weight producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int g = 0; g < my_groups; ++g) {
        for (int c = 0; c < N_CHUNKS; ++c) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          mbar_arrive_expect_tx(bar_full(stage), CHUNK_BYTES);
          tma_bulk_g2s(sbase + SL::W_OFF + stage * CHUNK_BYTES, args.packed + PK_CHUNKS + (size_t)c * CHUNK_BYTES,
                       CHUNK_BYTES, bar_full(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TILE_M, NH);
      int stage = 0;
      uint32_t phase = 0, ar_phase = 0;
      for (int g = 0; g < my_groups; ++g) {
        for (int l = 0; l < N_MMA_LAYERS; ++l) {
          const bool prof = args.prof != nullptr && blockIdx.x == 0 && g == 0;
          if (prof) args.prof[l * 4 + 0] = clock64();
          mbar_wait(bar_aready, ar_phase);  // A operand of layer l written (and accumulators drained)
          ar_phase ^= 1;
          tc_fence_after();
          if (prof) args.prof[l * 4 + 1] = clock64();
          const int akb = layer_akb(l), nkb = akb + layer_has_e(l), nhn = layer_nh(l);
          for (int h = 0; h < nhn; ++h) {
            for (int kbi = 0; kbi < nkb; ++kbi) {
              mbar_wait(bar_full(stage), phase);
              tc_fence_after();
              const uint32_t b_addr = sbase + SL::W_OFF + stage * CHUNK_BYTES;
#pragma unroll
              for (int t = 0; t < NT; ++t) {
                const uint32_t a_addr = (kbi < akb) ? (sbase + SL::A_OFF + (t * 4 + kbi) * ABLK_BYTES)
                                                    : (sbase + SL::E_OFF + t * ABLK_BYTES);
                const uint32_t d_addr = tmem_base + (uint32_t)(t * 256 + h * NH);
#pragma unroll
                for (int ks = 0; ks < KB / 16; ++ks) {
                  umma_bf16(d_addr, make_sw128_desc(a_addr + ks * 32), make_sw128_desc(b_addr + ks * 32), idesc,
                            (kbi > 0 || ks > 0) ? 1u : 0u);
                }
              }
              umma_commit(bar_empty(stage));  // frees the weight slot once these MMAs have read it
              if (kbi == nkb - 1) umma_commit(bar_acc(h));
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
          }
          if (prof) args.prof[l * 4 + 2] = clock64();
        }
      }
    }
  } else {
    // ===================== encoder / epilogue warpgroups =====================
    const int t = (warp - 2) >> 2;               // tile handled by this warpgroup
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;               // row within the tile == TMEM lane
    uint8_t* a_blk = smem + SL::A_OFF + t * 4 * ABLK_BYTES;
    uint8_t* e_blk = smem + SL::E_OFF + t * ABLK_BYTES;
    const float* bias_all = reinterpret_cast<const float*>(args.packed + PK_BIAS);
    const float4* wsig4 = reinterpret_cast<const float4*>(args.packed + PK_WSIGMA);
    const float4* wrgb = reinterpret_cast<const float4*>(args.packed + PK_WRGB);
    const float4 headb = __ldg(reinterpret_cast<const float4*>(args.packed + PK_HEADB));
    uint32_t acc_phase[2] = {0, 0};

    for (int g = 0; g < my_groups; ++g) {
      const int64_t group = (int64_t)blockIdx.x + (int64_t)g * gridDim.x;
      const int64_t srow = (group * NT + t) * TILE_M + row;
      const bool live = srow < args.n_samples;
      const int64_t lrow = live ? srow : (args.n_samples - 1);
      const float p0 = __ldg(args.pos + 3 * lrow), p1 = __ldg(args.pos + 3 * lrow + 1), p2 = __ldg(args.pos + 3 * lrow + 2);
      // ---- layer-0 A operand: pos_enc(pos, 0, 10) ----
      write_encoding<10>(e_blk, row, p0, p1, p2);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_aready);

      float sigma_raw = 0.f;
      for (int l = 0; l < N_MMA_LAYERS; ++l) {
        const int nhn = layer_nh(l);
        const float* bias = bias_all + l * 256;
        float r_acc = 0.f, g_acc = 0.f, b_acc = 0.f;
        // all n-halves of the layer must be complete before the in-place overwrite of the A operand
        const bool prof = args.prof != nullptr && blockIdx.x == 0 && g == 0 && warp == 2 && lane == 0;
        if (prof) args.prof[40 + l * 4 + 0] = clock64();
        for (int h = 0; h < nhn; ++h) { mbar_wait(bar_acc(h), acc_phase[h]); acc_phase[h] ^= 1; }
        tc_fence_after();
        if (prof) args.prof[40 + l * 4 + 1] = clock64();
        for (int cg = 0; cg < nhn * 4; ++cg) {  // 32 accumulator columns at a time
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + cg * 32), v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + cg * 32) + j4);
            f[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + b4.x;
            f[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + b4.y;
            f[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + b4.z;
            f[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + b4.w;
          }
          if (layer_relu(l)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
          if (l == 7) {  // sigma head (Dense_8) on the bf16-rounded trunk output
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w4 = __ldg(wsig4 + cg * 8 + j4);
              sigma_raw = fmaf(bf16_round(f[4 * j4 + 0]), w4.x, sigma_raw);
              sigma_raw = fmaf(bf16_round(f[4 * j4 + 1]), w4.y, sigma_raw);
              sigma_raw = fmaf(bf16_round(f[4 * j4 + 2]), w4.z, sigma_raw);
              sigma_raw = fmaf(bf16_round(f[4 * j4 + 3]), w4.w, sigma_raw);
            }
          }
          if (l == 9) {  // rgb head (Dense_11) on the bf16-rounded condition-layer output
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float4 w4 = __ldg(wrgb + cg * 32 + j);
              const float hb = bf16_round(f[j]);
              r_acc = fmaf(hb, w4.x, r_acc); g_acc = fmaf(hb, w4.y, g_acc); b_acc = fmaf(hb, w4.z, b_acc);
            }
          } else {
            // next layer's A operand: columns cg*32..+31 -> k-block cg/2, 16-byte chunks (cg&1)*4..+3
            uint8_t* blk = a_blk + (cg >> 1) * ABLK_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(blk + sw128_offset(row, (cg & 1) * 32 + c * 8)) =
                  make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
          if (args.layer_out && live) {
            uint4* dst = reinterpret_cast<uint4*>(args.layer_out + ((size_t)l * args.n_samples + srow) * 256 + cg * 32);
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
        }
        if (l == 8) {
          // condition input for Dense_10: pos_enc(dir, 0, 4) replaces the position encoding (last used by layer 5)
          const float d0 = __ldg(args.dir + 3 * lrow), d1 = __ldg(args.dir + 3 * lrow + 1), d2 = __ldg(args.dir + 3 * lrow + 2);
          write_encoding<4>(e_blk, row, d0, d1, d2);
        }
        if (prof) args.prof[40 + l * 4 + 2] = clock64();
        if (l == 9) {
          if (live) args.raw_out[srow] = make_float4(r_acc + headb.x, g_acc + headb.y, b_acc + headb.z, sigma_raw + headb.w);
          tc_fence_before();  // accumulators drained; the next group's layer 0 may overwrite them
        } else {
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_aready);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256 * NT);
}

template <int NT, int NSTAGE>
static int launch_encmlp(const EncMlpArgs& a0, cudaStream_t st) {
  using SL = SmemLayout<NT, NSTAGE>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kfn = encmlp_kernel<NT, NSTAGE>;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::ALLOC);
    if (e != cudaSuccess) { set_error("rnerf_encmlp_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set[dev] = true;
  }
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  EncMlpArgs a = a0;
  a.n_groups = (int)((a.n_samples + (int64_t)TILE_M * NT - 1) / ((int64_t)TILE_M * NT));
  const int grid = a.n_groups < n_sm ? a.n_groups : n_sm;
  kfn<<<grid, 64 + 128 * NT, SL::ALLOC, st>>>(a);
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
  encmlp_pack_kernel<<<N_CHUNKS + 1, 256, 0, (cudaStream_t)stream>>>(a, (uint8_t*)packed);
  count_launch();
  return check_launch("rnerf_encmlp_pack");
}

static int encmlp_fwd_impl(const void* packed, const float* pos, const float* dir, int64_t n_samples, float* raw_out,
                           uint16_t* layer_out, void* stream, long long* prof = nullptr) {
  RNERF_REQUIRE(n_samples >= 0, RNERF_E_SHAPE, "rnerf_encmlp_fwd: n_samples < 0");
  if (n_samples == 0) return 0;
  RNERF_REQUIRE_PTR(packed); RNERF_REQUIRE_PTR(pos); RNERF_REQUIRE_PTR(dir); RNERF_REQUIRE_PTR(raw_out);
  RNERF_REQUIRE(aligned16(packed) && aligned16(raw_out), RNERF_E_ALIGN, "rnerf_encmlp_fwd: packed/raw_out must be 16-byte aligned");
  RNERF_REQUIRE(layer_out == nullptr || aligned16(layer_out), RNERF_E_ALIGN, "rnerf_encmlp_fwd: layer_out must be 16-byte aligned");
  EncMlpArgs a;
  a.packed = (const uint8_t*)packed; a.pos = pos; a.dir = dir; a.n_samples = n_samples;
  a.raw_out = (float4*)raw_out; a.layer_out = (__nv_bfloat16*)layer_out; a.prof = prof; a.n_groups = 0;
  if (n_samples <= 148 * 128) return launch_encmlp<1, 8>(a, (cudaStream_t)stream);
  return launch_encmlp<2, 4>(a, (cudaStream_t)stream);
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
