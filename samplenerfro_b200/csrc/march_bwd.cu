// Adjoint of the "all"-stage eikonal march: gradients of a loss on the coarse samples (ray_pos_c, ray_dir_c) with
// respect to so3_mlp.
//
// What the reference differentiates (train.py:164 jax.value_and_grad through nn.scan, rnerf/eikonal_utils.py:30-49,75-82):
//     (n, g) = linear3(p_k);  G = where(|g| > 1e-3, rodrigues(so3_mlp(annealed_pos_enc(p_k)), g), g)
//     p_{k+1} = p_k + step / n * v_k;    v_{k+1} = v_k + step * G
// with ray_pos = p_k, ray_dir = safe_l2_normalize(v_k) read at the 64 coarse indices (rnerf/models.py:243-244); ray_dist is
// stop_gradient (rnerf/eikonal_utils.py:120) and the fine samples are stop_gradient too (rnerf/model_utils.py:406-411), so
// the coarse samples are the only way a loss reaches the scan.  The reverse sweep over k with adjoints (lp, lv) of
// (p_{k+1}, v_{k+1}):
//     lv_k = lv + step/n * lp                       dn = -(step/n^2) (v_k . lp)          dG = step * lv
//     active ray:  (d raw, dg) = rodrigues^T(dG);   so3_mlp backward: parameter gradients += ..., dp_mlp = enc^T(dX)
//     inactive:    dg = dG
//     lp_k = lp + J^T (dn, dg) + dp_mlp             J = d linear3 / d p (the trilinear weights differentiated, floor/clamp
//                                                       constant -- exactly what autodiff of rnerf/ior_utils.py:201-222 gives)
// then the loss gradients of record k are added.  Nothing is saved by the forward march beyond the path records: p_k, v_k
// are read back from them and the lookup / MLP forward are recomputed (bit-identical p_k => the same `active` set).
//
// One thread per ray, 128 rays per CTA, like the forward.  Steps at which a ray of the CTA is active evaluate so3_mlp
// forward + backward cooperatively: the active rays are compacted into <= 32 columns of [feature][ray] buffers in
// shared memory (all four hidden activations are kept for the ReLU masks and the weight gradients); thread t owns
// neurons 2j, 2j+1 (j = t mod 64) for columns 16h .. 16h+15 (h = t / 64) in every GEMM of the chain:
//     forward       acc[neuron][col] = sum_in  W [in][neuron]   A[in][col]
//     input grad    acc[in][col]     = sum_out Wt[out][in]      dZ[out][col]      (Wt: transposed image, rnerf_so3_transpose)
//     weight grad   gW[in][neuron]  += sum_col A[in][col] dZ[neuron][col]          (red.global.add.v2.f32 per row)
// fp32 on the CUDA cores, like the forward (the gradient steers the ray).
#include <stdio.h>
#include <stdlib.h>
#include "march_common.cuh"

namespace rnerf {

constexpr int BW_COLS = 32;
constexpr int BW_RP = BW_COLS + 4;                              // column pitch (rows stay 16-byte aligned)
constexpr int BW_OFF_X = 0;                                     // X  [60][RP]   annealed encoding
constexpr int BW_OFF_H = BW_OFF_X + SO3_IN * BW_RP;             // H  [4][128][RP] outputs of Dense_0..3 (post ReLU)
constexpr int BW_OFF_D = BW_OFF_H + 4 * SO3_W * BW_RP;          // D  [128][RP]  dZ of the layer being processed
constexpr int BW_OFF_DX = BW_OFF_D + SO3_W * BW_RP;             // DX [60][RP]   gradient wrt the encoding
constexpr int BW_OFF_R = BW_OFF_DX + SO3_IN * BW_RP;            // R  [4][RP]    d raw (3 rows used)
constexpr int BW_OFF_P = BW_OFF_R + 4 * BW_RP;              // P  [4][RP]    column positions, then d p through the encoding
constexpr int BW_OFF_RAW = BW_OFF_P + 4 * BW_RP;            // RAW [4][RP]   so3_mlp output
constexpr int BW_ACT_FLOATS = BW_OFF_RAW + 4 * BW_RP;
constexpr int SO3T_OFF_0 = 0, SO3T_OFF_1 = SO3_W * SO3_IN, SO3T_OFF_2 = SO3T_OFF_1 + SO3_W * SO3_W,
              SO3T_OFF_3A = SO3T_OFF_2 + SO3_W * SO3_W, SO3T_OFF_3B = SO3T_OFF_3A + SO3_W * SO3_W,
              SO3T_FLOATS = SO3T_OFF_3B + SO3_W * SO3_IN;

struct So3BwdArgs {
  const float* w;        // forward image (So3Args::w)
  const float* wt;       // transposed image: T0 [128][60], T1, T2, T3a [128][128], T3b [128][60]  (T_l[out][in] = W_l[in][out])
  float* gw;             // gradient image, same layout as w, accumulated into
  float window[10];
  const float* window_dev;   // device copy of the window or NULL (So3Args::window_dev)
  const float* saved;        // hidden activations left by the training forward (So3Args::saved), or NULL: recompute them
  long long* prof;           // development aid (RNERF_SWEEP_PROF): clock64 stamps [CTA][32] of one evaluation per CTA, or NULL
  int prof_eval;             // which evaluation of a CTA is stamped
};

__global__ void __launch_bounds__(256) so3_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= SO3T_FLOATS) return;
  int src;
  if (i < SO3T_OFF_1)       { const int o = i / SO3_IN, k = i % SO3_IN; src = k * SO3_W + o; }
  else if (i < SO3T_OFF_2)  { const int e = i - SO3T_OFF_1, o = e / SO3_W, k = e % SO3_W; src = SO3_OFF_W1 + k * SO3_W + o; }
  else if (i < SO3T_OFF_3A) { const int e = i - SO3T_OFF_2, o = e / SO3_W, k = e % SO3_W; src = SO3_OFF_W2 + k * SO3_W + o; }
  else if (i < SO3T_OFF_3B) { const int e = i - SO3T_OFF_3A, o = e / SO3_W, k = e % SO3_W; src = SO3_OFF_W3 + k * SO3_W + o; }
  else                      { const int e = i - SO3T_OFF_3B, o = e / SO3_IN, k = e % SO3_IN; src = SO3_OFF_W3 + (SO3_W + k) * SO3_W + o; }
  wt[i] = __ldg(w + src);
}

// ---- trilinear lookup with its derivative wrt p --------------------------------------------------------------------
// c is computed with the forward's arithmetic (bit-identical n, grad n => the same active set); jx/jy/jz = dc/dp_x,y,z.
template <bool FAST>
__device__ __forceinline__ void lookup_with_jacobian(const float4* __restrict__ table, const MarchGeom& mg,
                                                     const float* __restrict__ bricks, float px, float py, float pz, float4& c,
                                                     float4& jx, float4& jy, float4& jz) {
  float x, y, z;
  grid_coords<FAST>(mg, px, py, pz, x, y, z);
  const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  const float xd = sub(x, xf), yd = sub(y, yf), zd = sub(z, zf);
  const float oxd = sub(1.f, xd), oyd = sub(1.f, yd), ozd = sub(1.f, zd);
  const float mx = (float)(mg.g.gx - 1), my = (float)(mg.g.gy - 1), mz = (float)(mg.g.gz - 1);
  const int x0 = (int)fminf(fmaxf(xf, 0.f), mx), y0 = (int)fminf(fmaxf(yf, 0.f), my), z0 = (int)fminf(fmaxf(zf, 0.f), mz);
  jx = jy = jz = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bricks != nullptr) {
    const float b = __ldg(bricks + ((x0 >> BRICK_LOG2) * mg.nby + (y0 >> BRICK_LOG2)) * mg.nbz + (z0 >> BRICK_LOG2));
    if (b == b) {            // homogeneous brick: eight equal corners -> zero derivative
      const float c00 = lerp_ref(b, b, oxd, xd);
      const float c0 = lerp_ref(c00, c00, oyd, yd);
      c = make_float4(lerp_ref(c0, c0, ozd, zd), 0.f, 0.f, 0.f);
      return;
    }
  }
  const int x1 = (int)fminf(fmaxf(add(xf, 1.f), 0.f), mx), y1 = (int)fminf(fmaxf(add(yf, 1.f), 0.f), my),
            z1 = (int)fminf(fmaxf(add(zf, 1.f), 0.f), mz);
  const int sx = mg.g.gy * mg.g.gz, sy = mg.g.gz;
  const int b00 = sx * x0 + sy * y0, b10 = sx * x1 + sy * y0, b01 = sx * x0 + sy * y1, b11 = sx * x1 + sy * y1;
  const float4 d000 = __ldg(table + b00 + z0), d100 = __ldg(table + b10 + z0);
  const float4 d001 = __ldg(table + b00 + z1), d101 = __ldg(table + b10 + z1);
  const float4 d010 = __ldg(table + b01 + z0), d110 = __ldg(table + b11 + z0);
  const float4 d011 = __ldg(table + b01 + z1), d111 = __ldg(table + b11 + z1);
  const float4 c00 = lerp4_ref(d000, d100, oxd, xd);
  const float4 c01 = lerp4_ref(d001, d101, oxd, xd);
  const float4 c10 = lerp4_ref(d010, d110, oxd, xd);
  const float4 c11 = lerp4_ref(d011, d111, oxd, xd);
  const float4 c0 = lerp4_ref(c00, c10, oyd, yd);
  const float4 c1 = lerp4_ref(c01, c11, oyd, yd);
  c = lerp4_ref(c0, c1, ozd, zd);
  const float ix = 1.f / mg.g.ndelta[0], iy = 1.f / mg.g.ndelta[1], iz = 1.f / mg.g.ndelta[2];
#define RNERF_JAC(F)                                                                                                    \
  {                                                                                                                     \
    const float e00 = d100.F - d000.F, e01 = d101.F - d001.F, e10 = d110.F - d010.F, e11 = d111.F - d011.F;             \
    jx.F = ((e00 * oyd + e10 * yd) * ozd + (e01 * oyd + e11 * yd) * zd) * ix;                                           \
    jy.F = ((c10.F - c00.F) * ozd + (c11.F - c01.F) * zd) * iy;                                                         \
    jz.F = (c1.F - c0.F) * iz;                                                                                          \
  }
  RNERF_JAC(x) RNERF_JAC(y) RNERF_JAC(z) RNERF_JAC(w)
#undef RNERF_JAC
}

// Gradient wrt the (n, grad n) table itself (extension: the reference keeps the grid constant): the adjoint
// (dn, dg) of a lookup is spread over its eight corners with the trilinear weights, one red.global.add.v4.f32 per corner.
template <bool FAST>
__device__ __forceinline__ void scatter_table_grad(float4* __restrict__ d_table, const MarchGeom& mg, float px, float py, float pz,
                                                   float dn, const float dg[3]) {
  float x, y, z;
  grid_coords<FAST>(mg, px, py, pz, x, y, z);
  const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  const float xd = x - xf, yd = y - yf, zd = z - zf;
  const float mx = (float)(mg.g.gx - 1), my = (float)(mg.g.gy - 1), mz = (float)(mg.g.gz - 1);
  const int xi[2] = {(int)fminf(fmaxf(xf, 0.f), mx), (int)fminf(fmaxf(xf + 1.f, 0.f), mx)};
  const int yi[2] = {(int)fminf(fmaxf(yf, 0.f), my), (int)fminf(fmaxf(yf + 1.f, 0.f), my)};
  const int zi[2] = {(int)fminf(fmaxf(zf, 0.f), mz), (int)fminf(fmaxf(zf + 1.f, 0.f), mz)};
  const float wx[2] = {1.f - xd, xd}, wy[2] = {1.f - yd, yd}, wz[2] = {1.f - zd, zd};
  const int sx = mg.g.gy * mg.g.gz, sy = mg.g.gz;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float w = wx[a] * wy[b] * wz[c];
        float4* dst = d_table + (sx * xi[a] + sy * yi[b] + zi[c]);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(w * dn), "f"(w * dg[0]), "f"(w * dg[1]),
                     "f"(w * dg[2]) : "memory");
      }
}

// ---- Rodrigues rotation (rnerf/ior_utils.py:300-306), reverse mode ------------------------------------------------------
// pred = a (cos(th) v + sin(th) e x v + (1 - cos(th)) (e.v) e), th = |r|_safe, e = r/th, a = |g|_safe, v = g/a.
// Given d pred, returns d r and d g.
__device__ __forceinline__ void rodrigues_bwd(const float r[3], const float g[3], const float dp[3], float dr[3], float dg[3]) {
  const float s_r = r[0] * r[0] + r[1] * r[1] + r[2] * r[2], s_g = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
  const float th = sqrtf(fmaxf(s_r, 1e-6f)), a = sqrtf(fmaxf(s_g, 1e-6f));
  const float ith = 1.f / th, ia = 1.f / a;
  const float e[3] = {r[0] * ith, r[1] * ith, r[2] * ith}, v[3] = {g[0] * ia, g[1] * ia, g[2] * ia};
  const float ct = cosf(th), st = sinf(th);
  const float c[3] = {e[1] * v[2] - e[2] * v[1], e[2] * v[0] - e[0] * v[2], e[0] * v[1] - e[1] * v[0]};
  const float ev = e[0] * v[0] + e[1] * v[1] + e[2] * v[2];
  const float omc = (1.f - ct) * ev;
  float u[3], du[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { u[i] = ct * v[i] + st * c[i] + omc * e[i]; du[i] = a * dp[i]; }
  float da = u[0] * dp[0] + u[1] * dp[1] + u[2] * dp[2];
  const float e_du = e[0] * du[0] + e[1] * du[1] + e[2] * du[2];
  const float d_ct = (v[0] * du[0] + v[1] * du[1] + v[2] * du[2]) - ev * e_du;
  const float d_st = c[0] * du[0] + c[1] * du[1] + c[2] * du[2];
  const float d_ev = (1.f - ct) * e_du;
  const float dc[3] = {st * du[0], st * du[1], st * du[2]};
  float de[3], dv[3];
  // c = e x v:  de += v x dc,  dv += dc x e
  de[0] = omc * du[0] + (v[1] * dc[2] - v[2] * dc[1]) + d_ev * v[0];
  de[1] = omc * du[1] + (v[2] * dc[0] - v[0] * dc[2]) + d_ev * v[1];
  de[2] = omc * du[2] + (v[0] * dc[1] - v[1] * dc[0]) + d_ev * v[2];
  dv[0] = ct * du[0] + (dc[1] * e[2] - dc[2] * e[1]) + d_ev * e[0];
  dv[1] = ct * du[1] + (dc[2] * e[0] - dc[0] * e[2]) + d_ev * e[1];
  dv[2] = ct * du[2] + (dc[0] * e[1] - dc[1] * e[0]) + d_ev * e[2];
  float d_th = -st * d_ct + ct * d_st;
  d_th -= (e[0] * de[0] + e[1] * de[1] + e[2] * de[2]) * ith;
  const float k_r = s_r > 1e-6f ? d_th : 0.f;          // d max(s, eps) / ds
  da -= (v[0] * dv[0] + v[1] * dv[1] + v[2] * dv[2]) * ia;
  const float k_g = s_g > 1e-6f ? da : 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dr[i] = de[i] * ith + k_r * e[i];
    dg[i] = dv[i] * ia + k_g * v[i];
  }
}

// ---- building blocks of the cooperative MLP chain -----------------------------------------------------------------------
// Weight stream of one evaluation through the TMA ring (march_common.cuh), 18 chunks of <= 64 rows (32 KB slots: the CTA
// has the SM to itself, and every chunk costs a block barrier, so the chunks are as large as three slots allow):
//   0..7    the forward image, segments S0 (Dense_0 x X; 1 chunk), S1, S2, S3a (Dense_3[:128] x H3; 2 chunks each),
//           S3b (Dense_3[128:] x X; 1 chunk)
//   8..17   T3b [128][60], T3a [128][128], T2, T1, T0 [128][60]  (order of use; 2 chunks each)
constexpr int BW_CH = 64;
constexpr int BW_SLOT_FLOATS = BW_CH * SO3_W;
constexpr int BW_NCHUNK = 18;
constexpr int BW_RING_SLOTS = 3;
struct So3BwdStream {
  const float* w;
  const float* wt;
  bool saved;            // the forward's activations are read back: only the transposed chunks 8..17 are streamed
  __device__ __forceinline__ void operator()(uint32_t g, const float*& src, uint32_t& bytes) const {
    const int c = saved ? 8 + (int)(g % (uint32_t)(BW_NCHUNK - 8)) : (int)(g % (uint32_t)BW_NCHUNK);
    if (c < 8) {
      // chunk -> (first row, rows) of the [504][128] forward matrix: segments start at rows 0, 60, 188, 316, 444
      const int row0 = c == 0 ? 0 : (c == 7 ? 444 : 60 + (c - 1) * BW_CH);
      const int rows = (c == 0 || c == 7) ? SO3_IN : BW_CH;
      src = w + (size_t)row0 * SO3_W;
      bytes = (uint32_t)rows * SO3_W * 4;
    } else {
      const int b = (c - 8) >> 1, i = (c - 8) & 1;           // block in order of use, 2 chunks of 64 rows each
      const int off = b == 0 ? SO3T_OFF_3B : (b == 1 ? SO3T_OFF_3A : (b == 2 ? SO3T_OFF_2 : (b == 3 ? SO3T_OFF_1 : SO3T_OFF_0)));
      const int wp = (b == 0 || b == 4) ? SO3_IN : SO3_W;
      src = wt + off + (size_t)i * BW_CH * wp;
      bytes = (uint32_t)BW_CH * wp * 4;
    }
  }
};

// Thread layout of the cooperative chain: BW_THREADS = 512 (16 warps: four per scheduler, so LDS / barrier latencies of one
// warp hide under the FFMA2s of the others -- with 4 warps the chain ran at 13 % issue utilisation, profiles/r1w).  Thread t
// owns neurons 2j, 2j+1 (j = t mod 64) for the BW_CPT = 4 columns 4h .. 4h+3 (h = t / 64) in the forward and input-gradient
// GEMMs; in the weight-gradient loops it owns the same neurons for ALL 32 columns on the rows r = h (mod 8), so every
// gW element is reduced by exactly one red per evaluation.
constexpr int BW_THREADS = 512;
constexpr int BW_H = BW_THREADS / 64;          // 8 column groups / row residues
constexpr int BW_CPT = BW_COLS / BW_H;         // 4 columns per thread

// acc[n][c] += sum_{r < K} W[r][2j + n] * In[r][4h + c]: W ([K][wp], K <= 128) arrives as the next ceil(K/64) chunks of the
// ring.  Every thread of the CTA runs the chunk loop (block barrier per chunk); `work` = this thread owns output rows.
__device__ __forceinline__ void gemm_ring(float (&acc)[2][BW_CPT], So3Ring& ring, int tid, const So3BwdStream& stream, int K, int wp,
                                          const float* __restrict__ In, int j, int h, bool work) {
  const float* in = In + BW_CPT * h;
  f32x2 a00 = pack2(acc[0][0], acc[0][1]), a01 = pack2(acc[0][2], acc[0][3]);
  f32x2 a10 = pack2(acc[1][0], acc[1][1]), a11 = pack2(acc[1][2], acc[1][3]);
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += BW_CH) {
    const float* wq = ring_acquire(ring, tid, stream) + 2 * j;
    const int rows = min(BW_CH, K - k0);
    if (work) {
#pragma unroll 8
      for (int r = 0; r < rows; ++r) {
        const float2 w = *reinterpret_cast<const float2*>(wq + r * wp);
        const f32x2 w0 = pack2(w.x, w.x), w1 = pack2(w.y, w.y);
        const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(in + (k0 + r) * BW_RP);
        a00 = fma2(w0, x.x, a00); a01 = fma2(w0, x.y, a01);
        a10 = fma2(w1, x.x, a10); a11 = fma2(w1, x.y, a11);
      }
    }
    ++ring.pos;
  }
  unpack2(a00, acc[0][0], acc[0][1]); unpack2(a01, acc[0][2], acc[0][3]);
  unpack2(a10, acc[1][0], acc[1][1]); unpack2(a11, acc[1][2], acc[1][3]);
}

__device__ __forceinline__ void zero_acc(float (&acc)[2][BW_CPT]) {
#pragma unroll
  for (int c = 0; c < BW_CPT; ++c) { acc[0][c] = 0.f; acc[1][c] = 0.f; }
}

// rows 2j, 2j+1 of a [.][RP] buffer, columns 4h .. 4h+3
__device__ __forceinline__ void store_rows(const float (&acc)[2][BW_CPT], float* buf, int j, int h) {
#pragma unroll
  for (int n = 0; n < 2; ++n)
    *reinterpret_cast<float4*>(buf + (2 * j + n) * BW_RP + BW_CPT * h) = make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]);
}

__device__ __forceinline__ void relu_bias_store(float (&acc)[2][BW_CPT], const float* __restrict__ bias, float* buf, int j, int h) {
  const float b0 = __ldg(bias + 2 * j), b1 = __ldg(bias + 2 * j + 1);
#pragma unroll
  for (int c = 0; c < BW_CPT; ++c) { acc[0][c] = fmaxf(acc[0][c] + b0, 0.f); acc[1][c] = fmaxf(acc[1][c] + b1, 0.f); }
  store_rows(acc, buf, j, h);
}

// dZ = dH where the saved activation is positive
__device__ __forceinline__ void relu_mask(float (&acc)[2][BW_CPT], const float* __restrict__ Hl, int j, int h) {
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const float4 x = *reinterpret_cast<const float4*>(Hl + (2 * j + n) * BW_RP + BW_CPT * h);
    acc[n][0] = x.x > 0.f ? acc[n][0] : 0.f; acc[n][1] = x.y > 0.f ? acc[n][1] : 0.f;
    acc[n][2] = x.z > 0.f ? acc[n][2] : 0.f; acc[n][3] = x.w > 0.f ? acc[n][3] : 0.f;
  }
}

__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// dZ rows 2j, 2j+1 over all 32 columns, as column pairs, read back from D (published by store_rows + a block barrier)
struct DzRows { f32x2 d[2][BW_COLS / 2]; };
__device__ __forceinline__ void load_dz_rows(DzRows& z, const float* __restrict__ D, int j) {
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const ulonglong2* row = reinterpret_cast<const ulonglong2*>(D + (2 * j + n) * BW_RP);
#pragma unroll
    for (int q = 0; q < BW_COLS / 4; ++q) { const ulonglong2 v = row[q]; z.d[n][2 * q] = v.x; z.d[n][2 * q + 1] = v.y; }
  }
}
// gb[2j + n] += sum_c dz[n][c]   (threads with h == 0)
__device__ __forceinline__ void bias_grad(const DzRows& z, float* gb, int j) {
  f32x2 s0 = 0ull, s1 = 0ull;
  const f32x2 one = pack2(1.f, 1.f);
#pragma unroll
  for (int q = 0; q < BW_COLS / 2; ++q) { s0 = fma2(z.d[0][q], one, s0); s1 = fma2(z.d[1][q], one, s1); }
  float a, b, c, d;
  unpack2(s0, a, b); unpack2(s1, c, d);
  red_add2(gb + 2 * j, a + b, c + d);
}
// gW[r][2j + n] += sum_c A[r][c] dz[n][c]  for the rows r = h, h + 8, ... < rows   (gW row pitch 128)
__device__ __forceinline__ void wgrad_rows(const DzRows& z, const float* __restrict__ A, int rows, float* gW, int j, int h) {
  float* g = gW + 2 * j;
#pragma unroll 2
  for (int r = h; r < rows; r += BW_H) {
    const ulonglong2* xr = reinterpret_cast<const ulonglong2*>(A + r * BW_RP);
    f32x2 s0 = 0ull, s1 = 0ull;                  // (even columns, odd columns) partial sums
#pragma unroll
    for (int q = 0; q < BW_COLS / 4; ++q) {
      const ulonglong2 x = xr[q];
      s0 = fma2(x.x, z.d[0][2 * q], s0); s0 = fma2(x.y, z.d[0][2 * q + 1], s0);
      s1 = fma2(x.x, z.d[1][2 * q], s1); s1 = fma2(x.y, z.d[1][2 * q + 1], s1);
    }
    float a, b, c, d;
    unpack2(s0, a, b); unpack2(s1, c, d);
    red_add2(g + r * SO3_W, a + b, c + d);
  }
}

// so3_mlp forward + backward for the CTA's active rays.  EVERY thread of the CTA must call this (block barriers inside).
// In (threads with `act`): position p, lookup gradient g, adjoint dG of the rotated gradient.
// Out (threads with `act`): dg (adjoint of g), dp (adjoint of p through the encoding).  Parameter gradients -> a.gw.
__device__ __forceinline__ void so3_fwd_bwd(const So3BwdArgs& a, float* sm, int* cnt, So3Ring& ring, int warp, int lane, bool act,
                                            const float p[3], const float g[3], const float dG[3], float dg[3], float dp[3],
                                            int slot /* ray * n_steps + step: the evaluation's record in a.saved */,
                                            int eval_idx = 0) {
  int stamp_i = 0;
  auto STAMP = [&]() {       // development aid: phase boundaries of evaluation a.prof_eval (scripts/sweep_phases.py)
    if (a.prof != nullptr && eval_idx == a.prof_eval && warp == 0 && lane == 0 && stamp_i < 29)
      a.prof[blockIdx.x * 32 + 1 + stamp_i] = clock64();
    ++stamp_i;
  };
  STAMP();   // 0: entry
  float* X = sm + BW_OFF_X;
  float* H = sm + BW_OFF_H;
  float* D = sm + BW_OFF_D;
  float* DX = sm + BW_OFF_DX;
  float* R = sm + BW_OFF_R;                      // d raw  [3][RP]
  float* P = sm + BW_OFF_P;                      // positions of the columns [3][RP]; later their gradient through the encoding
  float* RAW = sm + BW_OFF_RAW;                  // so3_mlp output [3][RP]
  const int tid = warp * 32 + lane;
  const int j = tid & 63, h = tid >> 6;
  const float* bias = a.w + SO3_OFF_B;
  float* gbias = a.gw + SO3_OFF_B;
  const So3BwdStream stream{a.w, a.wt, a.saved != nullptr};
  int* SLOT = reinterpret_cast<int*>(R + 3 * BW_RP);     // (row 3 of R is unused)
  ring_prime(ring, tid, stream);
  __syncthreads();                               // the previous evaluation has finished with cnt and the buffers
  const unsigned bal = __ballot_sync(0xffffffffu, act);
  if (lane == 0) cnt[warp] = __popc(bal);
  __syncthreads();
  int base = 0, n_act = 0;
#pragma unroll
  for (int w = 0; w < MARCH_THREADS / 32; ++w) { // only the first MARCH_THREADS threads carry rays
    const int c = cnt[w];
    if (w < warp) base += c;
    n_act += c;
  }
  const int idx = base + __popc(bal & ((1u << lane) - 1u));
  STAMP();   // 1: counted
  if (a.prof != nullptr && eval_idx == a.prof_eval && warp == 0 && lane == 0) a.prof[blockIdx.x * 32] = n_act;
  constexpr int HL = SO3_W * BW_RP;              // floats per saved layer
  const float half_pi = 1.57079632679489661923f;
#pragma unroll 1
  for (int col0 = 0; col0 < n_act; col0 += BW_COLS) {
    const int n_here = min(BW_COLS, n_act - col0);
    const bool mine = act && idx >= col0 && idx < col0 + BW_COLS;
    const int col = idx - col0;
    if (col0 > 0) __syncthreads();               // another pass: everyone is done with the buffers of the previous one
    if (mine) { P[col] = p[0]; P[BW_RP + col] = p[1]; P[2 * BW_RP + col] = p[2]; SLOT[col] = slot; }
    __syncthreads();
    // ---- encoding, all threads: feature k*6 + c = sin(2^k p_c) w_k, k*6 + 3 + c = sin(2^k p_c + pi/2) w_k.  Unused columns
    // are zeroed: their dZ stays exactly 0 through the chain, and 0 * (finite activation) adds nothing
    for (int e = tid; e < SO3_IN * BW_COLS; e += BW_THREADS) {
      const int f = e >> 5, cc = e & 31;
      float v = 0.f;
      if (cc < n_here) {
        const int k = f / 6, q = f - 6 * k, c = q >= 3 ? q - 3 : q;
        const float xb = mul(P[c * BW_RP + cc], (float)(1 << k));
        v = mul(sinf(q >= 3 ? add(xb, half_pi) : xb), so3_window_at(a, k));
      }
      X[f * BW_RP + cc] = v;
    }
    float acc[2][BW_CPT];
    if (a.saved != nullptr) {
      // ---- the forward march left every hidden activation of this evaluation in global memory (2 KB per column): read them
      // back instead of recomputing four layers (8 of the 18 weight chunks and their barrier rounds).  Consecutive threads
      // read consecutive neurons of one column; unused columns are zero (their dZ stays 0 through the chain)
      // thread tid owns row n = tid (= layer * 128 + neuron) of H: its value for every live column, eight loads in flight,
      // then the whole 32-column row with conflict-free 16-byte stores
      static_assert(BW_THREADS == 4 * SO3_W && BW_COLS == 32, "one thread per saved activation row");
      float hv[BW_COLS];
#pragma unroll
      for (int c8 = 0; c8 < BW_COLS; c8 += 8) {
        if (c8 < n_here) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            hv[c8 + u] = c8 + u < n_here ? __ldg(a.saved + (size_t)SLOT[c8 + u] * (4 * SO3_W) + tid) : 0.f;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) hv[c8 + u] = 0.f;
        }
      }
      float4* hrow = reinterpret_cast<float4*>(H + tid * BW_RP);
#pragma unroll
      for (int q = 0; q < BW_COLS / 4; ++q) hrow[q] = make_float4(hv[4 * q], hv[4 * q + 1], hv[4 * q + 2], hv[4 * q + 3]);
    } else {
    // (the first chunk barrier of the GEMM publishes X)
    // ---- forward, keeping every hidden activation
    zero_acc(acc);
    gemm_ring(acc, ring, tid, stream, SO3_IN, SO3_W, X, j, h, true);
    relu_bias_store(acc, bias, H, j, h);
    zero_acc(acc);
    gemm_ring(acc, ring, tid, stream, SO3_W, SO3_W, H, j, h, true);           // (the chunk barrier publishes H first)
    relu_bias_store(acc, bias + SO3_W, H + HL, j, h);
    zero_acc(acc);
    gemm_ring(acc, ring, tid, stream, SO3_W, SO3_W, H + HL, j, h, true);
    relu_bias_store(acc, bias + 2 * SO3_W, H + 2 * HL, j, h);
    zero_acc(acc);
    gemm_ring(acc, ring, tid, stream, SO3_W, SO3_W, H + 2 * HL, j, h, true);  // skip concat [h, inputs]
    gemm_ring(acc, ring, tid, stream, SO3_IN, SO3_W, X, j, h, true);
    relu_bias_store(acc, bias + 3 * SO3_W, H + 3 * HL, j, h);
    }
    __syncthreads();
    STAMP();   // 2: X + H ready
    const float* H4 = H + 3 * HL;
    const float* W4 = a.w + SO3_OFF_W4;
    // ---- Dense_4: raw[m][col]: four threads per output, each a quarter of the 128-long sum (one thread per output was a
    // 128-deep dependent FMA chain: 3.3 us of every 54 us evaluation, scripts/sweep_phases.py)
    if (tid < 4 * 3 * BW_COLS) {
      const int o = tid >> 2, kq = tid & 3, m = o >> 5, cc = o & 31;
      float r = 0.f;
#pragma unroll 8
      for (int k = kq; k < SO3_W; k += 4) r = fmaf(H4[k * BW_RP + cc], __ldg(W4 + 3 * k + m), r);
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
      if (kq == 0) RAW[m * BW_RP + cc] = r + __ldg(bias + 4 * SO3_W + m);
    }
    __syncthreads();
    STAMP();   // 3: raw
    // ---- rotation, forward and reverse, by the ray's own thread (it holds g and dG); unused columns get d raw = 0
    if (mine) {
      const float r[3] = {RAW[col], RAW[BW_RP + col], RAW[2 * BW_RP + col]};
      float dr[3];
      rodrigues_bwd(r, g, dG, dr, dg);
      R[col] = dr[0]; R[BW_RP + col] = dr[1]; R[2 * BW_RP + col] = dr[2];
    }
    if (tid < BW_COLS && tid >= n_here) { R[tid] = 0.f; R[BW_RP + tid] = 0.f; R[2 * BW_RP + tid] = 0.f; }
    __syncthreads();
    STAMP();   // 4: rotation reverse
    // ---- Dense_4 backward: dH4 = W4 d raw, masked; gW4 += H4 (x) d raw; gb4 += sum d raw
    {
      float w4[2][3];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int m = 0; m < 3; ++m) w4[n][m] = __ldg(W4 + (2 * j + n) * 3 + m);
#pragma unroll
      for (int c = 0; c < BW_CPT; ++c) {
        const float r0 = R[BW_CPT * h + c], r1 = R[BW_RP + BW_CPT * h + c], r2 = R[2 * BW_RP + BW_CPT * h + c];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const float hv = H4[(2 * j + n) * BW_RP + BW_CPT * h + c];
          acc[n][c] = hv > 0.f ? (w4[n][0] * r0 + w4[n][1] * r1 + w4[n][2] * r2) : 0.f;
        }
      }
      if (tid < 3 * SO3_W) {                     // gW4[k][m] += sum_c H4[k][c] d raw[m][c], one thread per element
        const int k = tid / 3, m = tid - 3 * k;
        float s = 0.f;
#pragma unroll 8
        for (int c = 0; c < BW_COLS; ++c) s = fmaf(H4[k * BW_RP + c], R[m * BW_RP + c], s);
        atomicAdd(a.gw + SO3_OFF_W4 + tid, s);
      } else if (tid < 3 * SO3_W + 3) {
        const int m = tid - 3 * SO3_W;
        float s = 0.f;
        for (int c = 0; c < BW_COLS; ++c) s += R[m * BW_RP + c];
        atomicAdd(gbias + 4 * SO3_W + m, s);
      }
    }
    DzRows z;
    // ---- Dense_3 (inputs [H3, X]): acc = dZ4
    store_rows(acc, D, j, h);
    __syncthreads();
    STAMP();   // 5: Dense_4 backward, dZ4 published
    load_dz_rows(z, D, j);
    if (h == 0) bias_grad(z, gbias + 3 * SO3_W, j);
    wgrad_rows(z, H + 2 * HL, SO3_W, a.gw + SO3_OFF_W3, j, h);
    wgrad_rows(z, X, SO3_IN, a.gw + SO3_OFF_W3 + SO3_W * SO3_W, j, h);
    STAMP();   // 6: weight gradients of Dense_3 (this thread)
    {                                                        // gradient wrt the skip-concatenated encoding
      float ax[2][BW_CPT];
      zero_acc(ax);
      gemm_ring(ax, ring, tid, stream, SO3_W, SO3_IN, D, j, h, j < SO3_IN / 2);
      if (j < SO3_IN / 2) store_rows(ax, DX, j, h);
    }
    zero_acc(acc);
    gemm_ring(acc, ring, tid, stream, SO3_W, SO3_W, D, j, h, true);
    relu_mask(acc, H + 2 * HL, j, h);                        // dZ3
    __syncthreads();                                         // D has been read by everyone
    STAMP();   // 7: Dense_3 input gradients (4 chunks)
    // ---- Dense_2 (input H2)
    store_rows(acc, D, j, h);
    __syncthreads();
    load_dz_rows(z, D, j);
    if (h == 0) bias_grad(z, gbias + 2 * SO3_W, j);
    wgrad_rows(z, H + HL, SO3_W, a.gw + SO3_OFF_W2, j, h);
    zero_acc(acc);
    gemm_ring(acc, ring, tid, stream, SO3_W, SO3_W, D, j, h, true);
    relu_mask(acc, H + HL, j, h);                            // dZ2
    __syncthreads();
    STAMP();   // 8: Dense_2
    // ---- Dense_1 (input H1)
    store_rows(acc, D, j, h);
    __syncthreads();
    load_dz_rows(z, D, j);
    if (h == 0) bias_grad(z, gbias + SO3_W, j);
    wgrad_rows(z, H, SO3_W, a.gw + SO3_OFF_W1, j, h);
    zero_acc(acc);
    gemm_ring(acc, ring, tid, stream, SO3_W, SO3_W, D, j, h, true);
    relu_mask(acc, H, j, h);                                 // dZ1
    __syncthreads();
    STAMP();   // 9: Dense_1
    // ---- Dense_0 (input X)
    store_rows(acc, D, j, h);
    __syncthreads();
    load_dz_rows(z, D, j);
    if (h == 0) bias_grad(z, gbias, j);
    wgrad_rows(z, X, SO3_IN, a.gw, j, h);
    {
      float ax[2][BW_CPT];
      zero_acc(ax);
      if (j < SO3_IN / 2) {
#pragma unroll
        for (int n = 0; n < 2; ++n) {                        // continue from the skip part (own rows / columns)
          const float4 x = *reinterpret_cast<const float4*>(DX + (2 * j + n) * BW_RP + BW_CPT * h);
          ax[n][0] = x.x; ax[n][1] = x.y; ax[n][2] = x.z; ax[n][3] = x.w;
        }
      }
      gemm_ring(ax, ring, tid, stream, SO3_W, SO3_IN, D, j, h, j < SO3_IN / 2);
      if (j < SO3_IN / 2) store_rows(ax, DX, j, h);
    }
    __syncthreads();
    STAMP();   // 10: Dense_0
    // ---- encoding backward, one thread per (component, column): d/dp_c of sin(2^k p_c [+ pi/2]) w_k
    float gpc = 0.f;
    if (tid < 3 * BW_COLS) {
      const int c = tid >> 5, cc = tid & 31;
      if (cc < n_here) {
        const float x = P[c * BW_RP + cc];
        float sc = 1.f;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const float xb = mul(x, sc);
          gpc += so3_window_at(a, k) * sc * (cosf(xb) * DX[(k * 6 + c) * BW_RP + cc] + cosf(add(xb, half_pi)) * DX[(k * 6 + 3 + c) * BW_RP + cc]);
          sc *= 2.f;
        }
      }
    }
    __syncthreads();                             // every thread has read P
    if (tid < 3 * BW_COLS) P[(tid >> 5) * BW_RP + (tid & 31)] = gpc;
    __syncthreads();
    if (mine) { dp[0] = P[col]; dp[1] = P[BW_RP + col]; dp[2] = P[2 * BW_RP + col]; }
    STAMP();   // 11: encoding backward, done
  }
}

template <bool FAST>
__global__ void __launch_bounds__(BW_THREADS, 1) march_all_bwd_kernel(
    const float4* __restrict__ table, const MarchGeom mg, const float* __restrict__ bricks, const float4* __restrict__ path,
    int recf4, int64_t n_rays, float near, float step, int n_steps, const int32_t* __restrict__ jitter, int n_coarse,
    const float* __restrict__ d_pos_c, const float* __restrict__ d_dir_c, const So3BwdArgs so3, float* __restrict__ d_origins,
    float* __restrict__ d_viewdirs, int rays_per_cta, float4* __restrict__ d_table) {
  extern __shared__ __align__(16) float sm[];
  int* cnt = reinterpret_cast<int*>(sm + BW_ACT_FLOATS);
  float* ring_mem = sm + BW_ACT_FLOATS + 16 + 2 * SO3_MAX_SLOTS;         // after cnt (64 B) and room for 16 mbarriers (128 B)
  static_assert((BW_ACT_FLOATS + 16 + 2 * SO3_MAX_SLOTS) % 4 == 0, "ring slots must be 16-byte aligned");
  int16_t* kmap = reinterpret_cast<int16_t*>(ring_mem + BW_RING_SLOTS * BW_SLOT_FLOATS);   // march step -> coarse sample, or -1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  So3Ring ring;
  ring_init(ring, ring_mem, cnt + 16, BW_RING_SLOTS, tid, BW_SLOT_FLOATS);
  for (int i = tid; i < n_steps; i += BW_THREADS) kmap[i] = -1;
  __syncthreads();
  int k_last = 0;
  for (int i = 0; i < n_coarse; ++i) k_last = max(k_last, min(max(__ldg(jitter + i), 0), n_steps - 1));
  for (int i = tid; i < n_coarse; i += BW_THREADS) kmap[min(max(__ldg(jitter + i), 0), n_steps - 1)] = (int16_t)i;
  __syncthreads();
  const int64_t ray = blockIdx.x * (int64_t)rays_per_cta + tid;      // threads beyond rays_per_cta only help with the MLP
  const bool live = tid < rays_per_cta && ray < n_rays;
  const int64_t rr = live ? ray : (n_rays - 1);
  const float4* rec = path + rr * (int64_t)n_steps * recf4;
  float lp[3] = {0.f, 0.f, 0.f}, lv[3] = {0.f, 0.f, 0.f};
  // Rays of a CTA are NOT kept in lockstep: a ray sweeps its own steps until one needs so3_mlp, then waits; the CTA
  // evaluates the MLP for all waiting rays together -- whatever their step, the MLP does not care -- so the number of
  // evaluations a CTA runs in series tends to the largest active-step count of one of its rays instead of the size of the
  // union of their active steps (353 vs 165 on a sorted batch of 4096 random pixels, scripts/all_stage_batch_probe.py).
  // A ray that needs nothing runs at most SWEEP_QUANTUM steps ahead per round, so waiting rays are served promptly.
  constexpr int SWEEP_QUANTUM = 16;
  int k = k_last;
  bool done = !live;
  int n_eval = 0;
  // loss gradients of coarse sample jc = record k (v = direction state of that record)
  auto inject = [&](int kk, const float (&v)[3]) {
    const int jc = kmap[kk];
    if (jc < 0) return;
    const float* gp = d_pos_c + (ray * n_coarse + jc) * 3;
    const float* gd = d_dir_c + (ray * n_coarse + jc) * 3;
    const float s = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const float im = 1.f / sqrtf(fmaxf(s, 1e-6f));
    const float d0 = v[0] * im, d1 = v[1] * im, d2 = v[2] * im;           // safe_l2_normalize (rnerf/math_utils.py:6-12)
    const float g0 = __ldg(gd), g1 = __ldg(gd + 1), g2 = __ldg(gd + 2);
    const float proj = s > 1e-6f ? (d0 * g0 + d1 * g1 + d2 * g2) : 0.f;
    lv[0] += (g0 - d0 * proj) * im; lv[1] += (g1 - d1 * proj) * im; lv[2] += (g2 - d2 * proj) * im;
    lp[0] += __ldg(gp); lp[1] += __ldg(gp + 1); lp[2] += __ldg(gp + 2);
  };
  // second half of the transition k -> k+1, reverse, once dg (adjoint of grad n) and dpm (through the encoding) are known
  auto finish = [&](const float (&p)[3], float hn, float dn, const float (&dg)[3], const float (&dpm)[3], const float4& jx,
                    const float4& jy, const float4& jz) {
    if (d_table != nullptr) scatter_table_grad<FAST>(d_table, mg, p[0], p[1], p[2], dn, dg);
#pragma unroll
    for (int i = 0; i < 3; ++i) lv[i] = fmaf(hn, lp[i], lv[i]);
    lp[0] += jx.x * dn + jx.y * dg[0] + jx.z * dg[1] + jx.w * dg[2] + dpm[0];
    lp[1] += jy.x * dn + jy.y * dg[0] + jy.z * dg[1] + jy.w * dg[2] + dpm[1];
    lp[2] += jz.x * dn + jz.y * dg[0] + jz.z * dg[1] + jz.w * dg[2] + dpm[2];
  };
#pragma unroll 1
  while (true) {
    // state of a ray waiting for the MLP (valid when `need`)
    float p[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 0.f}, g[3] = {0.f, 0.f, 0.f}, dG[3] = {0.f, 0.f, 0.f};
    float4 jx = make_float4(0.f, 0.f, 0.f, 0.f), jy = jx, jz = jx;
    float hn = 0.f, dn = 0.f;
    bool need = false;
#pragma unroll 1
    for (int it = 0; it < SWEEP_QUANTUM && !done && !need; ++it) {
      const float4 r0 = __ldg(rec + k * recf4), r1 = __ldg(rec + k * recf4 + 1);
      p[0] = r0.x; p[1] = r0.y; p[2] = r0.z; v[0] = r1.x; v[1] = r1.y; v[2] = r1.z;
      if (k < k_last) {                          // transition k -> k+1, reverse
        float4 c;
        lookup_with_jacobian<FAST>(table, mg, bricks, p[0], p[1], p[2], c, jx, jy, jz);
        g[0] = c.y; g[1] = c.z; g[2] = c.w;
        hn = step / c.x;
        dn = -(hn / c.x) * (v[0] * lp[0] + v[1] * lp[1] + v[2] * lp[2]);
        dG[0] = step * lv[0]; dG[1] = step * lv[1]; dG[2] = step * lv[2];
        // the forward's test, on the same bits (so3.w == NULL: radiance stage, the sweep only yields ray / table gradients)
        need = so3.w != nullptr && sqrtf(sumsq3(g[0], g[1], g[2])) > 1e-3f;
        if (!need) {
          const float zero[3] = {0.f, 0.f, 0.f};
          finish(p, hn, dn, dG, zero, jx, jy, jz);
        }
      }
      if (!need) {
        inject(k, v);
        done = --k < 0;
      }
    }
    if (__syncthreads_or(need)) {
      float dg[3] = {0.f, 0.f, 0.f}, dpm[3] = {0.f, 0.f, 0.f};
      if (so3.prof != nullptr && tid == 0 && (n_eval == so3.prof_eval || n_eval == so3.prof_eval + 1))
        so3.prof[blockIdx.x * 32 + 30 + (n_eval - so3.prof_eval)] = clock64();      // start of this and of the next evaluation
      so3_fwd_bwd(so3, sm, cnt, ring, warp, lane, need, p, g, dG, dg, dpm, (int)(rr * n_steps + k), n_eval);
      ++n_eval;
      if (need) {
        finish(p, hn, dn, dg, dpm, jx, jy, jz);
        inject(k, v);
        done = --k < 0;
      }
    } else if (!__syncthreads_or(!done)) {
      break;
    }
  }
  ring_drain(ring);                              // weight chunks fetched ahead for an evaluation that never came
  if (live) {                                    // p_0 = o + near d, v_0 = d
    if (d_origins != nullptr) { d_origins[3 * ray] = lp[0]; d_origins[3 * ray + 1] = lp[1]; d_origins[3 * ray + 2] = lp[2]; }
    if (d_viewdirs != nullptr) {
      d_viewdirs[3 * ray] = fmaf(near, lp[0], lv[0]); d_viewdirs[3 * ray + 1] = fmaf(near, lp[1], lv[1]);
      d_viewdirs[3 * ray + 2] = fmaf(near, lp[2], lv[2]);
    }
  }
}

}  // namespace rnerf

using namespace rnerf;

extern "C" size_t rnerf_so3_transposed_floats(void) { return SO3T_FLOATS; }

extern "C" int rnerf_so3_transpose(const float* so3_w, float* so3_wt, void* stream) {
  RNERF_REQUIRE_PTR(so3_w); RNERF_REQUIRE_PTR(so3_wt);
  so3_transpose_kernel<<<(SO3T_FLOATS + 255) / 256, 256, 0, (cudaStream_t)stream>>>(so3_w, so3_wt);
  count_launch();
  return check_launch("rnerf_so3_transpose");
}

extern "C" int rnerf_march_all_bwd(const float* table, const float* bricks, const int ndim[3], const double nmin[3],
                                   const double nmax[3], const float* path, int rec_floats, int64_t n_rays, double near,
                                   double far, int n_steps, const int32_t* jitter, int n_coarse, const float* d_pos_c,
                                   const float* d_dir_c, const float* so3_w, const float* so3_wt,
                                   const double so3_window[10], const float* so3_window_dev, const float* so3_saved, float* g_so3,
                                   float* d_origins, float* d_viewdirs,
                                   float* d_table, void* stream) {
  RNERF_REQUIRE_PTR(table); RNERF_REQUIRE_PTR(ndim); RNERF_REQUIRE_PTR(nmin); RNERF_REQUIRE_PTR(nmax);
  RNERF_REQUIRE(n_rays >= 0, RNERF_E_SHAPE, "rnerf_march_all_bwd: n_rays < 0");
  RNERF_REQUIRE(n_steps >= 2 && n_steps <= 32767, RNERF_E_SHAPE, "rnerf_march_all_bwd: n_steps must be in [2, 32767]");
  RNERF_REQUIRE(n_coarse >= 1 && n_coarse <= n_steps, RNERF_E_SHAPE, "rnerf_march_all_bwd: n_coarse must be in [1, n_steps]");
  RNERF_REQUIRE(rec_floats == 8 || rec_floats == 12, RNERF_E_SHAPE, "rnerf_march_all_bwd: rec_floats must be 8 or 12");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(path); RNERF_REQUIRE_PTR(jitter); RNERF_REQUIRE_PTR(d_pos_c); RNERF_REQUIRE_PTR(d_dir_c);
  if (so3_w != nullptr) {
    RNERF_REQUIRE_PTR(so3_wt); RNERF_REQUIRE_PTR(g_so3);
    RNERF_REQUIRE(so3_window != nullptr || so3_window_dev != nullptr, RNERF_E_NULL, "rnerf_march_all_bwd: no so3 window given");
  }
  RNERF_REQUIRE(aligned16(table) && aligned16(path) && aligned16(so3_w) && aligned16(so3_wt) && aligned16(g_so3) && aligned16(d_table),
                RNERF_E_ALIGN, "rnerf_march_all_bwd: table/path/so3 images/d_table must be 16-byte aligned");
  RNERF_REQUIRE(grid_fits_int32(ndim), RNERF_E_SHAPE, "rnerf_march_all_bwd: grids with >= 2^31 voxels are not supported");
  cudaStream_t st = (cudaStream_t)stream;
  MarchGeom mg;
  const bool fast = make_march_geom(ndim, nmin, nmax, st, mg);
  const float step = (float)((far - near) / (n_steps - 1));
  So3BwdArgs a;
  a.w = so3_w; a.wt = so3_wt; a.gw = g_so3;
  for (int k = 0; k < 10; ++k) a.window[k] = so3_window != nullptr ? (float)so3_window[k] : 0.f;
  a.window_dev = so3_w != nullptr ? so3_window_dev : nullptr;
  a.saved = so3_w != nullptr ? so3_saved : nullptr;
  a.prof = nullptr; a.prof_eval = 0;
  const char* prof_path = getenv("RNERF_SWEEP_PROF");      // development aid: synchronous, never set in production
  if (prof_path != nullptr && so3_w != nullptr) {
    cudaMalloc(&a.prof, 4096 * 32 * sizeof(long long));
    cudaMemsetAsync(a.prof, 0, 4096 * 32 * sizeof(long long), (cudaStream_t)stream);
    a.prof_eval = getenv("RNERF_SWEEP_PROF_EVAL") ? atoi(getenv("RNERF_SWEEP_PROF_EVAL")) : 40;
  }
  RNERF_REQUIRE(a.saved == nullptr || (double)n_rays * n_steps < 2147483648.0, RNERF_E_SHAPE,
                "rnerf_march_all_bwd: so3_saved needs n_rays * n_steps < 2^31");
  const size_t dyn = (size_t)BW_ACT_FLOATS * 4 + 64 + 8 * SO3_MAX_SLOTS + (size_t)BW_RING_SLOTS * BW_SLOT_FLOATS * 4 +
                     (((size_t)n_steps * 2 + 15) & ~(size_t)15);
  RNERF_REQUIRE(dyn <= 227 * 1024, RNERF_E_SHAPE, "rnerf_march_all_bwd: n_steps too large for the step map in shared memory");
  cudaError_t e = cudaFuncSetAttribute(march_all_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(march_all_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) { set_error("rnerf_march_all_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int rpc = so3_rays_per_cta(n_rays, n_sm);
  const unsigned blocks = (unsigned)((n_rays + rpc - 1) / rpc);
  if (fast)
    march_all_bwd_kernel<true><<<blocks, BW_THREADS, dyn, st>>>((const float4*)table, mg, bricks, (const float4*)path, rec_floats / 4,
                                                                   n_rays, (float)near, step, n_steps, jitter, n_coarse, d_pos_c,
                                                                   d_dir_c, a, d_origins, d_viewdirs, rpc, (float4*)d_table);
  else
    march_all_bwd_kernel<false><<<blocks, BW_THREADS, dyn, st>>>((const float4*)table, mg, bricks, (const float4*)path, rec_floats / 4,
                                                                    n_rays, (float)near, step, n_steps, jitter, n_coarse, d_pos_c,
                                                                    d_dir_c, a, d_origins, d_viewdirs, rpc, (float4*)d_table);
  count_launch();
  if (a.prof != nullptr) {
    cudaStreamSynchronize(st);
    const int nb = blocks < 4096 ? (int)blocks : 4096;
    long long* h = (long long*)malloc((size_t)nb * 32 * sizeof(long long));
    cudaMemcpy(h, a.prof, (size_t)nb * 32 * sizeof(long long), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(prof_path, "w")) {
      for (int b = 0; b < nb; ++b) {
        if (h[b * 32 + 1] == 0) continue;          // this CTA never reached the stamped evaluation
        fprintf(f, "%d %lld", b, h[b * 32]);
        for (int i = 1; i < 32; ++i) fprintf(f, " %lld", h[b * 32 + i]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    free(h);
    cudaFree(a.prof);
  }
  return check_launch("rnerf_march_all_bwd");
}
