// so3_mlp on the tensor pipe: tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-grade products, fp32 accumulate in TMEM).
//
// The "all"-stage march evaluates so3_mlp (model_utils.MLP 60 -> 128 -> 128 -> 128 (+60) -> 128 -> 3, rnerf/ior_utils.py:
// 147-152,282-312) for the CTA's active rays at every step that needs it.  On the CUDA cores that evaluation is a chain of
// ~36 block barriers around short FFMA bursts (~50 us per evaluation, profiles/r1w_*).  Here each hidden layer is ONE GEMM
//     D[128 neurons x 64 columns] = W_l^T [128 x K_l] * act_l [K_l x 64]
// issued by one elected thread: A = the transposed weights (K-major, SWIZZLE_128B rows of 32 tf32), B = the activations stored
// [column][feature] (K-major as well: thread = TMEM lane = neuron writes its value of column c into row c, so a warp writes 128
// contiguous bytes), D = 64 fp32 columns of tensor memory.
//
// Precision: the result steers the ray (1e-4 relative on the bent positions after 768 steps), so single-pass TF32 (10-bit
// mantissa) is not enough.  Every operand is split x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and a product is
// hi*hi + hi*lo + lo*hi (the dropped lo*lo is 2^-22 relative): three MMAs per k-step, ~2^-21 relative error per product --
// the same order as fp32's own rounding.  The weight halves are packed once (so3_tc_pack_kernel); the activation halves are
// produced by the epilogue that writes the next layer's B operand.
#pragma once
#include "march_common.cuh"

namespace rnerf {

constexpr int TC_N = 64;                          // columns (active rays) per pass = MMA N
constexpr int TC_A_BYTES = SO3_W * 128;           // one weight chunk: [128 neurons][32 k] x 4 B = 16 KB (hi or lo)
constexpr int TC_B_BYTES = TC_N * 128;            // one activation k-block, hi or lo: [64 columns][32 k] x 4 B = 8 KB
constexpr int TC_NKB = 2 + 4 + 4 + 6;             // k-blocks of the four hidden layers (K = 64, 128, 128, 128 + 64)
constexpr int TC_NCHUNK = 2 * TC_NKB;             // (k-block, hi | lo) chunks of 16 KB, in the order the MMA issuer consumes them
constexpr size_t TC_PACKED_BYTES = (size_t)TC_NCHUNK * TC_A_BYTES;      // 512 KB
__host__ __device__ constexpr int tc_layer_kb(int l) { return l == 0 ? 2 : (l == 3 ? 6 : 4); }
__host__ __device__ constexpr int tc_layer_kb0(int l) { return l == 0 ? 0 : (l == 1 ? 2 : (l == 2 ? 6 : 10)); }

// cute::UMMA::InstrDescriptor for kind::tf32: c = f32 (1 << 4), a = b = tf32 (2 << 7, 2 << 10), both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs (the top 19 bits of each 32-bit element), fp32 accumulate
__device__ __forceinline__ void umma_tf32_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// byte offset of element (row, k) of a [rows x 32] block of 32-bit elements, K-major SWIZZLE_128B (16-byte units ^= row & 7)
__host__ __device__ __forceinline__ uint32_t tc_sw128_off(int row, int kk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ (row & 7)) & 7) << 4) + (kk & 3) * 4);
}

// so3 fp32 image (So3Args::w) -> the 32 chunks [k-block][hi | lo][128 neurons x 32 k], swizzled, zero-padded in K
__global__ void __launch_bounds__(256) so3_tc_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ packed) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= TC_NKB * SO3_W * 32) return;
  const int kk = e & 31, m = (e >> 5) & 127, gkb = e >> 12;
  const int layer = gkb < 2 ? 0 : (gkb < 6 ? 1 : (gkb < 10 ? 2 : 3));
  const int k = (gkb - tc_layer_kb0(layer)) * 32 + kk;
  float v = 0.f;
  if (layer == 0) { if (k < SO3_IN) v = w[k * SO3_W + m]; }
  else if (layer == 1) v = w[SO3_OFF_W1 + k * SO3_W + m];
  else if (layer == 2) v = w[SO3_OFF_W2 + k * SO3_W + m];
  else if (k < SO3_W) v = w[SO3_OFF_W3 + k * SO3_W + m];                          // Dense_3 rows [h (128); inputs (60)]
  else if (k - SO3_W < SO3_IN) v = w[SO3_OFF_W3 + k * SO3_W + m];
  const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
  const uint32_t off = tc_sw128_off(m, kk);
  *reinterpret_cast<float*>(packed + (size_t)(2 * gkb) * TC_A_BYTES + off) = hi;
  *reinterpret_cast<float*>(packed + (size_t)(2 * gkb + 1) * TC_A_BYTES + off) = lo;
}

// ---- shared-memory plan of one evaluator (1024-byte aligned base) --------------------------------------------------------
struct TcSmem {
  static constexpr uint32_t X_HI = 0;                                   // [2 k-blocks][8 KB]   encoding, hi
  static constexpr uint32_t X_LO = X_HI + 2 * TC_B_BYTES;               //                      encoding, lo
  static constexpr uint32_t H_HI = X_LO + 2 * TC_B_BYTES;               // [4 k-blocks][8 KB]   hidden activations, hi
  static constexpr uint32_t H_LO = H_HI + 4 * TC_B_BYTES;               //                      hidden activations, lo
  static constexpr uint32_t RING = H_LO + 4 * TC_B_BYTES;               // [n_slots][16 KB]     weight chunks
  // the last hidden layer is handed to the 3-wide head as plain fp32 [neuron][column]; it overlays H_HI / H_LO, which
  // nobody reads once Dense_3's MMAs have completed
  static constexpr uint32_t HS = H_HI;
  static constexpr int HS_PITCH = TC_N + 4;
  static_assert(SO3_W * HS_PITCH * 4 <= 8 * TC_B_BYTES, "head input does not fit over the activation buffers");
};

// One MMA pass of layer `l`: everything the issuing thread does between "activations ready" and "accumulators committed".
// full / empty: the ring's barriers (n_slots each); c = running chunk counter (updated).
__device__ __forceinline__ void tc_issue_layer(int l, uint32_t sbase, uint32_t tmem_d, uint32_t bar_full0, uint32_t bar_empty0,
                                               int n_slots, uint32_t& c, uint32_t bar_acc, int dbg = 0) {
  constexpr uint32_t idesc = make_idesc_tf32(SO3_W, TC_N);
  const int nkb = tc_layer_kb(l);
  bool started = false;
  for (int kb = 0; kb < nkb; ++kb) {
    const uint32_t s_hi = c % (uint32_t)n_slots, s_lo = (c + 1) % (uint32_t)n_slots;
    const uint32_t ph_hi = (c / (uint32_t)n_slots) & 1u, ph_lo = ((c + 1) / (uint32_t)n_slots) & 1u;
    mbar_wait(bar_full0 + 8 * s_hi, ph_hi);
    mbar_wait(bar_full0 + 8 * s_lo, ph_lo);
    tc_fence_after();
    const bool from_x = (l == 0) || (l == 3 && kb >= 4);
    const int bkb = (l == 3 && kb >= 4) ? kb - 4 : kb;
    const uint32_t b_hi_addr = sbase + (from_x ? TcSmem::X_HI : TcSmem::H_HI) + bkb * TC_B_BYTES;
    const uint32_t b_lo_addr = sbase + (from_x ? TcSmem::X_LO : TcSmem::H_LO) + bkb * TC_B_BYTES;
    const uint64_t a_hi_d = make_sw128_desc(sbase + TcSmem::RING + s_hi * TC_A_BYTES);
    const uint64_t a_lo_d = make_sw128_desc(sbase + TcSmem::RING + s_lo * TC_A_BYTES);
    const uint64_t b_hi_d = make_sw128_desc(b_hi_addr), b_lo_d = make_sw128_desc(b_lo_addr);
    const uint32_t dh = (uint32_t)(a_hi_d >> 32);
    if (!(dbg & 1))
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {               // 8 tf32 = 32 bytes of K per step: +2 in the descriptors' address fields
      umma_tf32_lohi(tmem_d, (uint32_t)a_lo_d + 2u * ks, (uint32_t)b_hi_d + 2u * ks, dh, idesc, (started || ks > 0) ? 1u : 0u);
      umma_tf32_lohi(tmem_d, (uint32_t)a_hi_d + 2u * ks, (uint32_t)b_lo_d + 2u * ks, dh, idesc, 1u);
      umma_tf32_lohi(tmem_d, (uint32_t)a_hi_d + 2u * ks, (uint32_t)b_hi_d + 2u * ks, dh, idesc, 1u);
    }
    started = true;
    umma_commit(bar_empty0 + 8 * s_hi);
    umma_commit(bar_empty0 + 8 * s_lo);
    c += 2;
  }
  umma_commit(bar_acc);
}

// Epilogue of a hidden layer for one neuron (thread = TMEM lane m = 32 q + lane): bias + ReLU of its 64 columns, then either
// the next layer's B operand (hi / lo halves, row c of k-block q) or, for the last hidden layer, the plain fp32 head input.
__device__ __forceinline__ void tc_epilogue(int l, uint8_t* smem, uint32_t tmem_lane_addr, int q, int lane, float bias) {
  uint32_t v[32];
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(tmem_lane_addr + half * 32, v);
    tmem_ld_wait();
    if (l < 3) {
      uint8_t* hi = smem + TcSmem::H_HI + q * TC_B_BYTES;
      uint8_t* lo = smem + TcSmem::H_LO + q * TC_B_BYTES;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = half * 32 + j;
        const float h = fmaxf(__uint_as_float(v[j]) + bias, 0.f);
        const float hh = tf32_rn(h);
        const uint32_t off = tc_sw128_off(c, lane);
        *reinterpret_cast<float*>(hi + off) = hh;
        *reinterpret_cast<float*>(lo + off) = h - hh;         // (the MMA reads its top 19 bits)
      }
    } else {
      float* hs = reinterpret_cast<float*>(smem + TcSmem::HS) + (q * 32 + lane) * TcSmem::HS_PITCH + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(hs + j) = make_float4(fmaxf(__uint_as_float(v[j]) + bias, 0.f), fmaxf(__uint_as_float(v[j + 1]) + bias, 0.f),
                                                         fmaxf(__uint_as_float(v[j + 2]) + bias, 0.f), fmaxf(__uint_as_float(v[j + 3]) + bias, 0.f));
    }
  }
}

// Annealed encoding of the pass's columns into the layer-0 B operand: feature f = k*6 + c (sin) / k*6 + 3 + c (cos) of
// octave k, times the window; features 60..63 are the K padding (zero).  `wt` = worker thread index, `nw` = worker count;
// lanes map to features, so a warp writes one 128-byte row.
template <typename A>
__device__ __forceinline__ void tc_write_encoding(const A& a, uint8_t* smem, const float* __restrict__ P, int p_pitch, int wt, int nw) {
  for (int e = wt; e < TC_N * 64; e += nw) {
    const int c = e >> 6, f = e & 63;
    float val = 0.f;
    if (f < SO3_IN) {
      const int k = f / 6, qq = f - 6 * k, ax = qq >= 3 ? qq - 3 : qq;
      const float xb = mul(P[ax * p_pitch + c], (float)(1 << k));
      val = mul(sinf(qq >= 3 ? add(xb, 1.57079632679489661923f) : xb), so3_window_at(a, k));
    }
    const float hh = tf32_rn(val);
    const uint32_t off = (uint32_t)(f >> 5) * TC_B_BYTES + tc_sw128_off(c, f & 31);
    *reinterpret_cast<float*>(smem + TcSmem::X_HI + off) = hh;
    *reinterpret_cast<float*>(smem + TcSmem::X_LO + off) = val - hh;
  }
}

}  // namespace rnerf
