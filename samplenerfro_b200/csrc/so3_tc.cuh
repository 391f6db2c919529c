// so3_mlp on the tensor pipe: tcgen05.mma kind::f16 on fp16 hi/lo SPLIT operands (fp32-grade products, fp32 accumulate in TMEM).
//
// The "all"-stage march evaluates so3_mlp (model_utils.MLP 60 -> 128 -> 128 -> 128 (+60) -> 128 -> 3, rnerf/ior_utils.py:
// 147-152,282-312) for the CTA's active rays at every step that needs it.  On the CUDA cores that evaluation is a chain of
// ~36 block barriers around short FFMA bursts (~50 us per evaluation, profiles/r1w_*).  Here each hidden layer is ONE GEMM
//     D[128 neurons x 64 columns] = W_l^T [128 x K_l] * act_l [K_l x 64]
// issued by one elected thread: A = the transposed weights (K-major, SWIZZLE_128B rows of 64 halves), B = the activations
// stored [column][feature] (K-major as well: thread = TMEM lane = neuron writes its value of column c into row c, so a warp
// writes 64 contiguous bytes), D = 64 fp32 columns of tensor memory.
//
// Precision: the result steers the ray (1e-4 relative on the bent positions after 768 steps), so a single reduced-precision
// pass is not enough.  Every operand is split x = hi + lo with hi = fp16(x), lo = fp16(x - hi) -- 22 significant bits -- and a
// product is hi*hi + hi*lo + lo*hi (each exact in the fp32 accumulator; the dropped lo*lo is 2^-22 relative): three MMAs per
// k-step, ~2^-21 relative error per product, the same order as fp32's own rounding (measured against the fp32 CUDA-core chain:
// 1e-6 relative, tests/test_gpu_model.py).  The first version used kind::tf32 (K = 8 per instruction, 4-byte elements): the
// same three products cost twice the instructions, twice the shared-memory operand traffic and twice the weight stream, and
// the evaluation was bound by exactly those (profiles/r3n_*); fp16 halves all three.  Range: fp16 holds |x| < 65504;
// encodings are in [-1, 1], so3 weights are O(0.1) and activations O(1-10); an activation beyond 65504 is clamped (a network
// with such activations is outside what this evaluator reproduces -- the CUDA-core chain, RNERF_SO3_TC=0, has no such limit).
// The weight halves are packed once (so3_tc_pack_kernel); the activation halves are produced by the epilogue that writes the
// next layer's B operand.
#pragma once
#include <cuda_fp16.h>
#include "march_common.cuh"

namespace rnerf {

constexpr int TC_N = 64;                          // columns (active rays) per pass = MMA N
constexpr int TC_KBLK = 64;                       // k per k-block: one 128-byte SWIZZLE_128B row of halves
constexpr int TC_A_BYTES = SO3_W * 128;           // one weight chunk: [128 neurons][64 k] x 2 B = 16 KB (hi or lo)
constexpr int TC_B_BYTES = TC_N * 128;            // one activation k-block, hi or lo: [64 columns][64 k] x 2 B = 8 KB
constexpr int TC_NKB = 1 + 2 + 2 + 3;             // k-blocks of the four hidden layers (K = 64, 128, 128, 128 + 64)
constexpr int TC_NCHUNK = TC_NKB;                 // one 32 KB chunk per k-block (hi half, then lo half), in consumption order
constexpr int TC_CHUNK_BYTES = 2 * TC_A_BYTES;    // -> one TMA copy and one full / empty barrier pair per k-block
constexpr size_t TC_PACKED_BYTES = (size_t)TC_NCHUNK * TC_CHUNK_BYTES;  // 256 KB
__host__ __device__ constexpr int tc_layer_kb(int l) { return l == 0 ? 1 : (l == 3 ? 3 : 2); }
__host__ __device__ constexpr int tc_layer_kb0(int l) { return l == 0 ? 0 : (l == 1 ? 1 : (l == 2 ? 3 : 5)); }

// cute::UMMA::InstrDescriptor for kind::f16 with fp16 inputs: c = f32 (1 << 4), a = b = f16 (format 0), both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate; descriptors as (low word, shared high word)
__device__ __forceinline__ void umma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// x = hi + lo, both fp16 (|x| clamped to the fp16 range)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  x = fminf(fmaxf(x, -65504.f), 65504.f);
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
// byte offset of element (row, kk) of a [rows x 64] block of halves, K-major SWIZZLE_128B (16-byte units ^= row & 7)
__host__ __device__ __forceinline__ uint32_t tc_sw128_off(int row, int kk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 3) ^ (row & 7)) & 7) << 4) + (kk & 7) * 2);
}

// so3 fp32 image (So3Args::w) -> the 16 chunks [k-block][hi | lo][128 neurons x 64 k], swizzled, zero-padded in K
__global__ void __launch_bounds__(256) so3_tc_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ packed) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= TC_NKB * SO3_W * TC_KBLK) return;
  const int kk = e & 63, m = (e >> 6) & 127, gkb = e >> 13;
  const int layer = gkb < 1 ? 0 : (gkb < 3 ? 1 : (gkb < 5 ? 2 : 3));
  const int k = (gkb - tc_layer_kb0(layer)) * TC_KBLK + kk;
  float v = 0.f;
  if (layer == 0) { if (k < SO3_IN) v = w[k * SO3_W + m]; }
  else if (layer == 1) v = w[SO3_OFF_W1 + k * SO3_W + m];
  else if (layer == 2) v = w[SO3_OFF_W2 + k * SO3_W + m];
  else if (k < SO3_W) v = w[SO3_OFF_W3 + k * SO3_W + m];                          // Dense_3 rows [h (128); inputs (60)]
  else if (k - SO3_W < SO3_IN) v = w[SO3_OFF_W3 + k * SO3_W + m];
  __half hi, lo;
  split_f16(v, hi, lo);
  const uint32_t off = tc_sw128_off(m, kk);
  *reinterpret_cast<__half*>(packed + (size_t)(2 * gkb) * TC_A_BYTES + off) = hi;
  *reinterpret_cast<__half*>(packed + (size_t)(2 * gkb + 1) * TC_A_BYTES + off) = lo;
}

// ---- shared-memory plan of one evaluator (1024-byte aligned base) --------------------------------------------------------
struct TcSmem {
  static constexpr uint32_t X_HI = 0;                                   // [1 k-block][8 KB]    encoding, hi
  static constexpr uint32_t X_LO = X_HI + TC_B_BYTES;                   //                      encoding, lo
  static constexpr uint32_t H_HI = X_LO + TC_B_BYTES;                   // [2 k-blocks][8 KB]   hidden activations, hi
  static constexpr uint32_t H_LO = H_HI + 2 * TC_B_BYTES;               //                      hidden activations, lo
  static constexpr uint32_t HS = H_LO + 2 * TC_B_BYTES;                 // [128][68] fp32: the last hidden layer, for the 3-wide head
  static constexpr int HS_PITCH = TC_N + 4;
  static constexpr uint32_t RING = (HS + SO3_W * HS_PITCH * 4 + 1023u) & ~1023u;      // [n_slots][32 KB] weight chunks (hi | lo)
};

// One MMA pass of layer `l`: everything the issuing thread does between "activations ready" and "accumulators committed".
// full / empty: the ring's barriers (n_slots each); c = running chunk counter (updated).
__device__ __forceinline__ void tc_issue_layer(int l, uint32_t sbase, uint32_t tmem_d, uint32_t bar_full0, uint32_t bar_empty0,
                                               int n_slots, uint32_t& c, uint32_t bar_acc, int dbg = 0) {
  constexpr uint32_t idesc = make_idesc_f16(SO3_W, TC_N);
  const int nkb = tc_layer_kb(l);
  bool started = false;
  for (int kb = 0; kb < nkb; ++kb) {
    const uint32_t s = c % (uint32_t)n_slots;
    mbar_wait(bar_full0 + 8 * s, (c / (uint32_t)n_slots) & 1u);
    tc_fence_after();
    const bool from_x = (l == 0) || (l == 3 && kb >= 2);
    const int bkb = (l == 3 && kb >= 2) ? kb - 2 : kb;
    const uint32_t b_hi_addr = sbase + (from_x ? TcSmem::X_HI : TcSmem::H_HI) + bkb * TC_B_BYTES;
    const uint32_t b_lo_addr = sbase + (from_x ? TcSmem::X_LO : TcSmem::H_LO) + bkb * TC_B_BYTES;
    const uint64_t a_hi_d = make_sw128_desc(sbase + TcSmem::RING + s * TC_CHUNK_BYTES);
    const uint64_t a_lo_d = make_sw128_desc(sbase + TcSmem::RING + s * TC_CHUNK_BYTES + TC_A_BYTES);
    const uint64_t b_hi_d = make_sw128_desc(b_hi_addr), b_lo_d = make_sw128_desc(b_lo_addr);
    const uint32_t dh = (uint32_t)(a_hi_d >> 32);
    if (!(dbg & 1))
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {               // 16 halves = 32 bytes of K per step: +2 in the descriptors' address fields
      umma_f16_lohi(tmem_d, (uint32_t)a_lo_d + 2u * ks, (uint32_t)b_hi_d + 2u * ks, dh, idesc, (started || ks > 0) ? 1u : 0u);
      umma_f16_lohi(tmem_d, (uint32_t)a_hi_d + 2u * ks, (uint32_t)b_lo_d + 2u * ks, dh, idesc, 1u);
      umma_f16_lohi(tmem_d, (uint32_t)a_hi_d + 2u * ks, (uint32_t)b_hi_d + 2u * ks, dh, idesc, 1u);
    }
    started = true;
    umma_commit(bar_empty0 + 8 * s);
    c += 1;
  }
  umma_commit(bar_acc);
}

// Epilogue of a hidden layer for one neuron m (thread = TMEM lane) and 32 of the 64 columns (first column c0): bias + ReLU,
// then either the next layer's B operand (hi / lo halves: element (row c, k = m)) or, for the last hidden layer, the plain
// fp32 head input HS[m][c].  `tmem_addr` addresses this thread's lane and column c0.
__device__ __forceinline__ void tc_epilogue32(int l, uint8_t* smem, uint32_t tmem_addr, int m, int c0, float bias,
                                              float* __restrict__ saved = nullptr, const int* __restrict__ slots = nullptr, int n_here = 0) {
  uint32_t v[32];
  tmem_ld32(tmem_addr, v);
  tmem_ld_wait();
  if (saved != nullptr) {
    // training: the post-ReLU activation of neuron m for every live column goes to that column's (ray, step) slot -- a warp
    // writes 128 contiguous bytes per column -- so that the reverse sweep need not recompute the forward
    float* dst = saved + l * 128 + m;
#pragma unroll
    for (int j = 0; j < 32; ++j)            // (fully unrolled: v[] stays in registers)
      if (c0 + j < n_here) dst[(size_t)slots[c0 + j] * (4 * 128)] = fmaxf(__uint_as_float(v[j]) + bias, 0.f);
  }
  if (l < 3) {
    uint8_t* hi = smem + TcSmem::H_HI + (m >> 6) * TC_B_BYTES;
    uint8_t* lo = smem + TcSmem::H_LO + (m >> 6) * TC_B_BYTES;
    const int kk = m & 63;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      __half h, g;
      split_f16(fmaxf(__uint_as_float(v[j]) + bias, 0.f), h, g);
      const uint32_t off = tc_sw128_off(c0 + j, kk);
      *reinterpret_cast<__half*>(hi + off) = h;
      *reinterpret_cast<__half*>(lo + off) = g;
    }
  } else {
    float* o = reinterpret_cast<float*>(smem + TcSmem::HS) + m * TcSmem::HS_PITCH + c0;
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(o + j) = make_float4(fmaxf(__uint_as_float(v[j]) + bias, 0.f), fmaxf(__uint_as_float(v[j + 1]) + bias, 0.f),
                                                      fmaxf(__uint_as_float(v[j + 2]) + bias, 0.f), fmaxf(__uint_as_float(v[j + 3]) + bias, 0.f));
  }
}

// Annealed encoding of column c into the layer-0 B operand, one warp per column: lane holds features f = lane and lane + 32
// (f = k*6 + axis: sin(2^k p_axis) w_k; k*6 + 3 + axis: sin(2^k p_axis + pi/2) w_k; f >= 60: the K padding, zero).
struct TcEncLane {
  int k[2], ax[2];
  float ph[2], w[2];
};
template <typename A>
__device__ __forceinline__ TcEncLane tc_enc_lane(const A& a, int lane) {
  TcEncLane e;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int f = lane + 32 * hh, k = f / 6, qq = f - 6 * k;
    e.k[hh] = k; e.ax[hh] = qq >= 3 ? qq - 3 : qq; e.ph[hh] = qq >= 3 ? 1.57079632679489661923f : 0.f;
    e.w[hh] = f < SO3_IN ? so3_window_at(a, k) : 0.f;
  }
  return e;
}
__device__ __forceinline__ float tc_enc_value(const TcEncLane& e, int hh, int lane, float px, float py, float pz) {
  if (lane + 32 * hh >= SO3_IN) return 0.f;
  const float p = e.ax[hh] == 0 ? px : (e.ax[hh] == 1 ? py : pz);
  const float xb = mul(p, (float)(1 << e.k[hh]));
  return mul(sinf(e.ph[hh] != 0.f ? add(xb, e.ph[hh]) : xb), e.w[hh]);
}
__device__ __forceinline__ void tc_enc_store(uint8_t* smem, int c, int hh, int lane, float val) {
  __half h, g;
  split_f16(val, h, g);
  const uint32_t off = tc_sw128_off(c, lane + 32 * hh);
  *reinterpret_cast<__half*>(smem + TcSmem::X_HI + off) = h;
  *reinterpret_cast<__half*>(smem + TcSmem::X_LO + off) = g;
}

}  // namespace rnerf
