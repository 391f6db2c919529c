// Bent-ray eikonal march (a5/a6) and coarse-sample selection (a7).
//
// OneEikonalStep (rnerf/eikonal_utils.py:30-49), radiance stage:
//     (n, g) = linear3(p);  p' = p + step/n * v;  v' = v + step * g;  t' = t + |p - p'|
// PathSampler.__call__ (rnerf/eikonal_utils.py:101-124) returns the state BEFORE each step and the
// lookup made at that state; ray_dir is safe_l2_normalize(v) (rnerf/math_utils.py:6-12).
//
// All state arithmetic uses non-contracted fp32 ops in the reference's association order, so the
// emitted path is bit-identical to the fp32 oracle.
//
// Path layouts (rec_floats):
//   12 "full"     (pos3, t | v3, n | grad3, 0)  every array PathSampler returns (debug / online-sparsity consumers)
//    8 "compact"  (pos3, t | v3, n)             what select + resample read on the render/train path of every
//                                               shipped config (idx_grad is only consumed by the sparsity term)
// plus an optional dense t column [B][S] (ray_dist) that lets the resampler search in shared memory.
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include "common.cuh"
#include "march_common.cuh"
#include "so3_tc.cuh"

namespace rnerf {

// One thread per ray; a warp is a bundle of 32 consecutive rays (adjacent pixels -> adjacent voxels, so
// the gathers of a warp land in few cache lines and the brick map / grid stay L1/L2 resident).  Records are
// staged through shared memory so that the global stores of a warp are sector-complete and contiguous:
// STEPS_PER_FLUSH steps x 32 (48) B = 128 (192) B contiguous per ray, all 32 lanes active in every store.
constexpr int STEPS_PER_FLUSH = 4;
constexpr int T_FLUSH = 16;              // the dense t column is flushed every 16 steps: 64 B (2 full sectors) per ray

__global__ void __launch_bounds__(256) verify_recip_kernel(float d, float y, int* __restrict__ bad) {
  const uint32_t m = blockIdx.x * 256u + threadIdx.x;          // 2^23 significands
  const float a = __uint_as_float(0x3f800000u | m);
  if (__float_as_uint(div_by_const(a, d, y)) != __float_as_uint(__fdiv_rn(a, d))) atomicOr(bad, 1);
}

// returns the verified reciprocal of d, or 0 if the fast sequence must not be used for this divisor
float recip_for(float d, cudaStream_t st) {
  static std::mutex mu;
  static std::map<uint32_t, float> cache;
  uint32_t key;
  memcpy(&key, &d, 4);
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
    return 0.f;                 // the verification synchronises: not inside a graph capture (IEEE divisions this time, not cached)
  float y = 0.f;
  if (d == d && fabsf(d) >= 0x1p-30f && fabsf(d) <= 0x1p30f) {
    const float cand = (float)(1.0 / (double)d);
    int* bad = nullptr;
    int host_bad = 1;
    if (cudaMalloc(&bad, sizeof(int)) == cudaSuccess) {
      cudaMemsetAsync(bad, 0, sizeof(int), st);
      verify_recip_kernel<<<(1u << 23) / 256, 256, 0, st>>>(d, cand, bad);
      count_launch();
      if (cudaMemcpyAsync(&host_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaStreamSynchronize(st) != cudaSuccess)
        host_bad = 1;
      cudaFree(bad);
    }
    if (!host_bad) y = cand;
  }
  cache[key] = y;
  return y;
}

// VoxMLP._linear3 at p (same arithmetic as trilinear() in common.cuh; index clamps done in float, which gives the same
// integers: xf is integer-valued, so clamp(int(xf), 0, G-1) == int(clamp(xf, 0, G-1)), and the +1 corner likewise)
template <bool FAST>
__device__ __forceinline__ float4 march_lookup(const float4* __restrict__ table, const MarchGeom& mg,
                                               const float* __restrict__ bricks, float px, float py, float pz) {
  float x, y, z;
  grid_coords<FAST>(mg, px, py, pz, x, y, z);
  const float xf = floorf(x), yf = floorf(y), zf = floorf(z);
  const float xd = sub(x, xf), yd = sub(y, yf), zd = sub(z, zf);
  const float oxd = sub(1.f, xd), oyd = sub(1.f, yd), ozd = sub(1.f, zd);
  const float mx = (float)(mg.g.gx - 1), my = (float)(mg.g.gy - 1), mz = (float)(mg.g.gz - 1);
  const int x0 = (int)fminf(fmaxf(xf, 0.f), mx), y0 = (int)fminf(fmaxf(yf, 0.f), my), z0 = (int)fminf(fmaxf(zf, 0.f), mz);
  if (bricks != nullptr) {
    const float c = __ldg(bricks + ((x0 >> BRICK_LOG2) * mg.nby + (y0 >> BRICK_LOG2)) * mg.nbz + (z0 >> BRICK_LOG2));
    if (c == c) {  // not NaN: homogeneous brick -- the lerps of eight equal corners, gradients exactly 0
      const float c00 = lerp_ref(c, c, oxd, xd);
      const float c0 = lerp_ref(c00, c00, oyd, yd);
      return make_float4(lerp_ref(c0, c0, ozd, zd), 0.f, 0.f, 0.f);
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
  return lerp4_ref(c0, c1, ozd, zd);
}

// ---- "all" stage (a4): so3_mlp inside every eikonal step -------------------------------------------------------------
// VoxMLP.__call__ (rnerf/ior_utils.py:269-312): raw = so3_mlp(annealed_pos_enc(p, 0, 10, alpha*10)); theta = |raw|_safe,
// e = raw/theta, a = |grad n|_safe, v = grad n / a; pred = a (cos(theta) v + sin(theta) e x v + (1 - cos(theta)) (e.v) e);
// OneEikonalStep (rnerf/eikonal_utils.py:34-35) then uses where(|grad n| > 1e-3, pred, grad n).
// so3_mlp = model_utils.MLP(net_width=128, net_depth=4, skip_layer=2, 3 outputs): 60 -> 128 -> 128 -> 128 (+60) -> 128 -> 3.
//
// The `where` makes the MLP irrelevant wherever |grad n| <= 1e-3 (everywhere but the blurred object boundary): a CTA
// only evaluates it at steps where one of its rays needs it, and only for those rays.  They are compacted (ballot +
// per-warp counts) into columns of CTA-wide activation buffers [feature][ray] in shared memory, at most 64 per pass (two
// groups of 32; more active rays -- rare -- take another pass).  The so3 variant of the kernel runs SO3_THREADS = 256
// threads: the first 128 carry the rays, all eight warps work on an evaluation (with only the four ray warps the chain
// ran at 26 % issue utilisation, one warp per scheduler: profiles/r1w).  Thread t owns neurons 2j, 2j+1 (j = t mod 64)
// for columns 8h .. 8h+7 (h = t / 64) of every group, i.e. 8 packed accumulators per group; the layer input is read with
// broadcast LDS.128, the weights with LDS.64; the encoding and the 3-wide output layer are spread over the CTA too.
// 113 KB of shared memory and <= 128 registers keep two CTAs per SM.
// fp32 on the CUDA cores (the result steers the ray, so no reduced-precision operands).
constexpr int SO3_THREADS = 256;
constexpr int SO3_COLS = 64;                                 // active rays per pass
constexpr int SO3_RP = SO3_COLS + 4;                         // ray pitch of the activation buffers (16-byte aligned rows)
constexpr int SO3_OFF_P = (SO3_IN + SO3_W) * SO3_RP;         // X[60][68] + H[128][68], then P[3][68] positions, RAW[3][68] outputs
constexpr int SO3_OFF_RAW = SO3_OFF_P + 3 * SO3_RP;
constexpr int SO3_ACT_FLOATS = SO3_OFF_RAW + 3 * SO3_RP;

// The four hidden-layer kernels are contiguous in the weight image: one [504][128] fp32 matrix (60 + 128 + 128 + 188
// rows).  It is streamed through the TMA-fed shared-memory ring of march_common.cuh in chunks of <= 16 rows that never
// straddle a segment (a segment = the rows multiplying one input block): S0 = Dense_0 x X, S1 = Dense_1 x H,
// S2 = Dense_2 x H, S3a = Dense_3[:128] x H, S3b = Dense_3[128:] x X (the skip concat [h, inputs]).  The weights cross
// L2 -> SM once per CTA evaluation, n_slots - 1 chunks ahead of the FMA loop (also across evaluations: the stream is
// periodic), and are read with conflict-free LDS.
// Rows per chunk CH: 16 (8 KB slots; full frames: two CTAs per SM share the shared memory) or 64 (32 KB slots; small
// launches own the SM, and every chunk costs a block barrier, so fewer and larger chunks are faster).
template <int CH> struct So3Chunking {
  static constexpr int C_X = (SO3_IN + CH - 1) / CH, C_H = (SO3_W + CH - 1) / CH;     // chunks of a 60-row / 128-row segment
  static constexpr int NCHUNK = 2 * C_X + 3 * C_H;                                      // 32 (CH = 16) or 8 (CH = 64)
  static constexpr int SLOT_FLOATS = CH * SO3_W;
};
// dynamic shared memory: activations | per-warp active-ray counts (32 B) | mbarriers (128 B) | ring slots
constexpr int SO3_OFF_CNT = SO3_ACT_FLOATS * 4, SO3_OFF_BARS = SO3_OFF_CNT + 4 * (SO3_THREADS / 32), SO3_OFF_RING = SO3_OFF_BARS + 8 * SO3_MAX_SLOTS;
static_assert(SO3_OFF_RING % 16 == 0, "ring slots must be 16-byte aligned");
static size_t so3_smem_bytes(int n_slots, int ch = SO3_CH) { return (size_t)SO3_OFF_RING + (size_t)n_slots * ch * SO3_W * 4; }

struct So3Chunk { int row0, rows, in_k0, in_is_x, last_of_layer; };
template <int CH>
__device__ __forceinline__ So3Chunk so3_chunk(int c) {
  // segments S0 (60 rows), S1, S2, S3a (128 rows each), S3b (60 rows); weight-row starts 0, 60, 188, 316, 444
  constexpr int CX = So3Chunking<CH>::C_X, CHH = So3Chunking<CH>::C_H;
  So3Chunk k;
  int seg, j;
  if (c < CX) { seg = 0; j = c; } else if (c < CX + CHH) { seg = 1; j = c - CX; } else if (c < CX + 2 * CHH) { seg = 2; j = c - CX - CHH; }
  else if (c < CX + 3 * CHH) { seg = 3; j = c - CX - 2 * CHH; } else { seg = 4; j = c - CX - 3 * CHH; }
  const int seg_row0 = seg == 0 ? 0 : (seg == 1 ? 60 : (seg == 2 ? 188 : (seg == 3 ? 316 : 444)));
  const int seg_rows = (seg == 0 || seg == 4) ? 60 : 128;
  k.in_k0 = j * CH;
  k.row0 = seg_row0 + k.in_k0;
  k.rows = min(CH, seg_rows - k.in_k0);
  k.in_is_x = (seg == 0 || seg == 4);
  k.last_of_layer = (seg != 3) && (k.in_k0 + k.rows == seg_rows);     // S3a continues into S3b
  return k;
}

template <int CH>
struct So3FwdStream {
  const float* w;
  __device__ __forceinline__ void operator()(uint32_t g, const float*& src, uint32_t& bytes) const {
    const So3Chunk k = so3_chunk<CH>((int)(g % (uint32_t)So3Chunking<CH>::NCHUNK));
    src = w + (size_t)k.row0 * SO3_W;
    bytes = (uint32_t)k.rows * SO3_W * 4;
  }
};

// raw = so3_mlp(annealed_pos_enc(p)) for the CTA's active rays.  EVERY thread of the CTA must call this (block barriers
// inside); only threads with `act` get a result.
template <int CH = SO3_CH>
__device__ __forceinline__ void so3_eval(const So3Args& a, float* dyn_smem, So3Ring& ring, int warp, int lane, bool act, float px,
                                         float py, float pz, float& r0, float& r1, float& r2) {
  float* X = dyn_smem;                         // [60][68]
  float* Hs = dyn_smem + SO3_IN * SO3_RP;      // [128][68]
  int* cnt = reinterpret_cast<int*>(reinterpret_cast<char*>(dyn_smem) + SO3_OFF_CNT);
  const int tid = warp * 32 + lane;
  const So3FwdStream<CH> stream{a.w};
  constexpr int SO3_NCHUNK = So3Chunking<CH>::NCHUNK;
  ring_prime(ring, tid, stream);
  // ---- compaction: active ray -> column idx of the activation buffers (only the first MARCH_THREADS threads carry rays)
  const unsigned bal = __ballot_sync(0xffffffffu, act);
  if (lane == 0) cnt[warp] = __popc(bal);
  __syncthreads();
  int base = 0, n_act = 0;
#pragma unroll
  for (int w = 0; w < MARCH_THREADS / 32; ++w) {
    const int c = cnt[w];
    if (w < warp) base += c;
    n_act += c;
  }
  const int idx = base + __popc(bal & ((1u << lane) - 1u));
  const float* bias = a.w + SO3_OFF_B;
  float* P = dyn_smem + SO3_OFF_P;
  float* RAW = dyn_smem + SO3_OFF_RAW;
  const int j = tid & 63, h = tid >> 6;        // neurons 2j, 2j+1; columns 8h .. 8h+7 of every group
  r0 = r1 = r2 = 0.f;
#pragma unroll 1
  for (int col0 = 0; col0 < n_act; col0 += SO3_COLS) {
    const int n_here = min(SO3_COLS, n_act - col0);
    const int n_groups = (n_here + 31) >> 5;   // 1 or 2 groups of 32 columns (unused columns hold garbage)
    const bool mine = act && idx >= col0 && idx < col0 + SO3_COLS;
    const int col = idx - col0;
    if (col0 > 0) __syncthreads();             // another pass: everyone is done with the buffers of the previous one
    if (mine) { P[col] = px; P[SO3_RP + col] = py; P[2 * SO3_RP + col] = pz; }
    __syncthreads();
    // ---- encoding, all threads: feature k*6 + c = sin(2^k p_c) w_k, k*6 + 3 + c = sin(2^k p_c + pi/2) w_k
    for (int e = tid; e < SO3_IN * n_here; e += SO3_THREADS) {
      const int f = e / n_here, cc = e - f * n_here;
      const int k = f / 6, q = f - 6 * k, c = q >= 3 ? q - 3 : q;
      const float xb = mul(P[c * SO3_RP + cc], (float)(1 << k));
      X[f * SO3_RP + cc] = mul(sinf(q >= 3 ? add(xb, 1.57079632679489661923f) : xb), so3_window_at(a, k));
    }
    int layer = 0;
    if (n_act <= 8) {
      // Few columns (the usual case once rays are out of lockstep, and at the rim of an object): splitting the 128 / 188
      // input rows of a layer over the four thread groups instead of the columns makes the dependent FMA chain of a layer
      // four times shorter.  Thread (j, h) sums rows r = h (mod 4) for neurons 2j, 2j+1 and all 8 columns; the four
      // partial sums meet in columns 8 + 8h .. of the output rows of Hs (free in an 8-column pass), then thread t adds
      // them for neuron t / 2, columns 4 (t mod 2) .. +3.
      f32x2 pa[2][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { pa[0][q] = 0ull; pa[1][q] = 0ull; }
#pragma unroll 1
      for (int c = 0; c < SO3_NCHUNK; ++c) {
        const float* wbuf = ring_acquire(ring, tid, stream) + 2 * j;
        const So3Chunk k = so3_chunk<CH>(c);
        const float* in = (k.in_is_x ? X : Hs) + k.in_k0 * SO3_RP;
#pragma unroll 4
        for (int r = h; r < k.rows; r += SO3_THREADS / 64) {
          const float2 w = *reinterpret_cast<const float2*>(wbuf + r * SO3_W);
          const f32x2 w0 = pack2(w.x, w.x), w1 = pack2(w.y, w.y);
          const ulonglong2* xr = reinterpret_cast<const ulonglong2*>(in + r * SO3_RP);
          const ulonglong2 x0 = xr[0], x1 = xr[1];
          pa[0][0] = fma2(w0, x0.x, pa[0][0]); pa[0][1] = fma2(w0, x0.y, pa[0][1]); pa[0][2] = fma2(w0, x1.x, pa[0][2]); pa[0][3] = fma2(w0, x1.y, pa[0][3]);
          pa[1][0] = fma2(w1, x0.x, pa[1][0]); pa[1][1] = fma2(w1, x0.y, pa[1][1]); pa[1][2] = fma2(w1, x1.x, pa[1][2]); pa[1][3] = fma2(w1, x1.y, pa[1][3]);
        }
        ++ring.pos;
        if (k.last_of_layer) {
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            ulonglong2* o = reinterpret_cast<ulonglong2*>(Hs + (2 * j + n) * SO3_RP + 8 + 8 * h);
            o[0] = make_ulonglong2(pa[n][0], pa[n][1]); o[1] = make_ulonglong2(pa[n][2], pa[n][3]);
          }
          __syncthreads();                     // all partial sums are there, nobody reads the layer input any more
          const int nrn = tid >> 1, c4 = 4 * (tid & 1);
          const float b = __ldg(bias + layer * SO3_W + nrn);
          float4 sum = make_float4(b, b, b, b);
#pragma unroll
          for (int hh = 0; hh < SO3_THREADS / 64; ++hh) {
            const float4 v = *reinterpret_cast<const float4*>(Hs + nrn * SO3_RP + 8 + 8 * hh + c4);
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
          }
          *reinterpret_cast<float4*>(Hs + nrn * SO3_RP + c4) = make_float4(fmaxf(sum.x, 0.f), fmaxf(sum.y, 0.f), fmaxf(sum.z, 0.f), fmaxf(sum.w, 0.f));
#pragma unroll
          for (int q = 0; q < 4; ++q) { pa[0][q] = 0ull; pa[1][q] = 0ull; }
          ++layer;
        }
      }
    } else {
    f32x2 acc[2][2][4];                        // [group][neuron][column pair]
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int r = 0; r < 4; ++r) { acc[g][0][r] = 0ull; acc[g][1][r] = 0ull; }
#pragma unroll 1
    for (int c = 0; c < SO3_NCHUNK; ++c) {
      // chunk c has landed; the barrier inside also makes the activations written before this point (X, or Hs of the
      // last layer) visible
      const float* wbuf = ring_acquire(ring, tid, stream) + 2 * j;
      const So3Chunk k = so3_chunk<CH>(c);
      const float* in = (k.in_is_x ? X : Hs) + k.in_k0 * SO3_RP + 8 * h;
#pragma unroll 4
      for (int r = 0; r < k.rows; ++r) {
        const float2 w = *reinterpret_cast<const float2*>(wbuf + r * SO3_W);
        const f32x2 w0 = pack2(w.x, w.x), w1 = pack2(w.y, w.y);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < n_groups) {
            const ulonglong2* xr = reinterpret_cast<const ulonglong2*>(in + r * SO3_RP + 32 * g);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const ulonglong2 x = xr[q];      // columns 4q, 4q+1 | 4q+2, 4q+3 of this thread's eight
              acc[g][0][2 * q] = fma2(w0, x.x, acc[g][0][2 * q]); acc[g][0][2 * q + 1] = fma2(w0, x.y, acc[g][0][2 * q + 1]);
              acc[g][1][2 * q] = fma2(w1, x.x, acc[g][1][2 * q]); acc[g][1][2 * q + 1] = fma2(w1, x.y, acc[g][1][2 * q + 1]);
            }
          }
        }
      }
      ++ring.pos;
      if (k.last_of_layer) {                   // bias + ReLU, handed to the next layer in place through Hs
        __syncthreads();                       // every thread has finished reading the layer input
        const float b0 = __ldg(bias + layer * SO3_W + 2 * j), b1 = __ldg(bias + layer * SO3_W + 2 * j + 1);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < n_groups) {
            float4* o0 = reinterpret_cast<float4*>(Hs + (2 * j) * SO3_RP + 32 * g + 8 * h);
            float4* o1 = reinterpret_cast<float4*>(Hs + (2 * j + 1) * SO3_RP + 32 * g + 8 * h);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float a0, a1, a2, a3;
              unpack2(acc[g][0][2 * q], a0, a1); unpack2(acc[g][0][2 * q + 1], a2, a3);
              o0[q] = make_float4(fmaxf(a0 + b0, 0.f), fmaxf(a1 + b0, 0.f), fmaxf(a2 + b0, 0.f), fmaxf(a3 + b0, 0.f));
              unpack2(acc[g][1][2 * q], a0, a1); unpack2(acc[g][1][2 * q + 1], a2, a3);
              o1[q] = make_float4(fmaxf(a0 + b1, 0.f), fmaxf(a1 + b1, 0.f), fmaxf(a2 + b1, 0.f), fmaxf(a3 + b1, 0.f));
            }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) { acc[g][0][r] = 0ull; acc[g][1][r] = 0ull; }
        }
        ++layer;
      }
    }
    }
    __syncthreads();                           // Dense_3 output visible
    {                                          // Dense_4: raw[m][column], one thread per output
      const float* W4 = a.w + SO3_OFF_W4;
      for (int e = tid; e < 3 * n_here; e += SO3_THREADS) {
        const int m = e / n_here, cc = e - m * n_here;
        float r = __ldg(bias + 4 * SO3_W + m);
#pragma unroll 8
        for (int k = 0; k < SO3_W; ++k) r = fmaf(Hs[k * SO3_RP + cc], __ldg(W4 + 3 * k + m), r);
        RAW[m * SO3_RP + cc] = r;
      }
    }
    __syncthreads();
    if (mine) { r0 = RAW[col]; r1 = RAW[SO3_RP + col]; r2 = RAW[2 * SO3_RP + col]; }
  }
}

// Rodrigues rotation of grad n by raw (rnerf/ior_utils.py:300-306); safe_l2_norm = sqrt(max(sum sq, 1e-6))
__device__ __forceinline__ void so3_rotate(float r0, float r1, float r2, float& gx, float& gy, float& gz) {
  const float theta = sqrtf(fmaxf(sumsq3(r0, r1, r2), 1e-6f));
  const float ex = divf(r0, theta), ey = divf(r1, theta), ez = divf(r2, theta);
  const float an = sqrtf(fmaxf(sumsq3(gx, gy, gz), 1e-6f));
  const float vx = divf(gx, an), vy = divf(gy, an), vz = divf(gz, an);
  const float ct = cosf(theta), st = sinf(theta);
  const float cx = sub(mul(ey, vz), mul(ez, vy)), cy = sub(mul(ez, vx), mul(ex, vz)), cz = sub(mul(ex, vy), mul(ey, vx));
  const float ev = add(add(mul(ex, vx), mul(ey, vy)), mul(ez, vz));
  const float omc = mul(sub(1.f, ct), ev);
  gx = mul(an, add(add(mul(ct, vx), mul(st, cx)), mul(omc, ex)));
  gy = mul(an, add(add(mul(ct, vy), mul(st, cy)), mul(omc, ey)));
  gz = mul(an, add(add(mul(ct, vz), mul(st, cz)), mul(omc, ez)));
}

template <int RECF4, bool FAST, bool SO3>
__global__ void __launch_bounds__(SO3 ? SO3_THREADS : MARCH_THREADS, SO3 ? 2 : 8) march_kernel(const float4* __restrict__ table, const MarchGeom mg,
                                                              const float* __restrict__ origins,
                                                              const float* __restrict__ viewdirs, int64_t n_rays,
                                                              float near, float step, int n_steps,
                                                              float4* __restrict__ path, float* __restrict__ t_col,
                                                              const float* __restrict__ bricks, int dbg, const So3Args so3,
                                                              int so3_slots, int rays_per_cta) {
  extern __shared__ __align__(16) float so3_scratch[];     // SO3 only: so3_smem_bytes(so3_slots)
  constexpr int F4_PER_FLUSH = STEPS_PER_FLUSH * RECF4;   // float4 per ray per flush: 8 (compact) / 12 (full)
  constexpr int PITCH = F4_PER_FLUSH + 1;                  // +1 float4 pad: conflict-free column writes
  __shared__ float4 stage[MARCH_THREADS / 32][32 * PITCH];
  __shared__ float tstage[MARCH_THREADS / 32][T_FLUSH * 32];        // ray_dist of the last <= 16 steps, [step][lane]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // SO3, small launches: a CTA may carry fewer than 128 rays (rays_per_cta = 32 or 64) so that a training batch spreads
  // over all SMs and a CTA evaluates the MLP only at the steps its own few rays need; the warps without rays still take
  // part in every evaluation (warp_ray0 = n_rays: not live, nothing staged or flushed)
  const bool carrier = !SO3 || warp * 32 < rays_per_cta;        // warp-uniform; the other warps only join the evaluations
  // Radiance stage, small launches: a warp may carry fewer than 32 rays (rpw = rays_per_cta / 4 = 1 .. 32).  A warp steps in
  // lockstep, so it waits for the table gathers at every step where ANY of its rays is inside a non-homogeneous brick; a
  // training batch of random pixels has little in common (the union of 32 rays' active steps is ~2.5x one ray's) and too
  // few warps to hide the latency (4096 rays = 128 warps on 148 SMs).  Fewer rays per warp -> more warps, each waiting only
  // for its own rays; the surplus lanes shadow the warp's first ray (same addresses: no extra traffic, no extra divergence).
  const int rpw = SO3 ? 32 : rays_per_cta / (MARCH_THREADS / 32);
  const int64_t warp_ray0 = !carrier ? n_rays : (blockIdx.x * (int64_t)rays_per_cta) + warp * rpw;
  if (!SO3 && warp_ray0 >= n_rays) return;      // SO3: every warp stays for the block barriers of so3_eval
  So3Ring ring;
  if (SO3) {
    char* base = reinterpret_cast<char*>(so3_scratch);
    ring_init(ring, reinterpret_cast<float*>(base + SO3_OFF_RING), base + SO3_OFF_BARS, so3_slots, threadIdx.x);
    __syncthreads();
  }
  const int64_t ray = warp_ray0 + lane;
  const bool live = lane < rpw && ray < n_rays;
  const int64_t rr = live ? ray : (SO3 ? n_rays - 1 : warp_ray0);
  float ox = origins[3 * rr], oy = origins[3 * rr + 1], oz = origins[3 * rr + 2];
  float vx = viewdirs[3 * rr], vy = viewdirs[3 * rr + 1], vz = viewdirs[3 * rr + 2];
  float px = add(ox, mul(near, vx)), py = add(oy, mul(near, vy)), pz = add(oz, mul(near, vz));
  float t = near;
  const int sw = carrier ? warp : 0;            // (non-carrier warps never touch the staging buffers)
  float4* my_stage = &stage[sw][lane * PITCH];
  const int rays_here = (int)max((int64_t)0, min((int64_t)rpw, n_rays - warp_ray0));
  const int ray_stride4 = n_steps * RECF4;                            // float4 units between consecutive rays
  const bool t_vec = t_col != nullptr && (n_steps & 3) == 0 && (reinterpret_cast<uintptr_t>(t_col) & 15u) == 0;
  float* ts = tstage[sw];
  const bool want_t = t_col != nullptr && !(dbg & 2);

  // cooperative flush mapping, fixed per thread: element e = it*32 + lane -> (ray e / F4, float4 e % F4)
  //   compact (F4 = 8):  ray = it*4 + (lane >> 3), unit = lane & 7                   -> one (smem, global) base pair
  //   full    (F4 = 12): 96 = 8 rays x 12 units, so the pattern repeats every 3 iterations -> three base pairs
  constexpr int NBASE = (RECF4 == 2) ? 1 : 3;
  constexpr int RAYS_PER_ROUND = (RECF4 == 2) ? 4 : 8;                // rays covered by NBASE iterations
  int s_off[NBASE], g_off[NBASE];
#pragma unroll
  for (int m = 0; m < NBASE; ++m) {
    const int e = m * 32 + lane, r = e / F4_PER_FLUSH, j = e - r * F4_PER_FLUSH;
    s_off[m] = r * PITCH + j;
    g_off[m] = r * ray_stride4 + j;
  }
  const float4* s_rd = stage[sw];
  float4* g_wr = path + warp_ray0 * (int64_t)ray_stride4;             // advanced by F4_PER_FLUSH per flush
  const int round_stride4 = RAYS_PER_ROUND * ray_stride4;

  // one eikonal step: emit the record of the current state, then advance it
  auto one_step = [&](int kk, int trow) {
    float4 c = make_float4(1.f, 0.f, 0.f, 0.f);
    if (carrier) {
      c = march_lookup<FAST>(table, mg, bricks, px, py, pz);                // (n, gx, gy, gz) at the pre-update position
      // the direction is stored un-normalised; readers apply safe_l2_normalize (path_dir()) to the few records they
      // use, which keeps 3 IEEE divides + 1 sqrt per step out of the march loop
      my_stage[kk * RECF4 + 0] = make_float4(px, py, pz, t);
      my_stage[kk * RECF4 + 1] = make_float4(vx, vy, vz, c.x);
      if (RECF4 == 3) my_stage[kk * RECF4 + 2] = make_float4(c.y, c.z, c.w, 0.f);
      ts[(trow + kk) * 32 + lane] = t;
    }
    float gx = c.y, gy = c.z, gz = c.w;
    if (SO3) {
      const bool act = live && sqrtf(sumsq3(gx, gy, gz)) > 1e-3f;     // jnp.linalg.norm(idx_grad) > 1e-3
      if (__syncthreads_or(act)) {
        float r0, r1, r2;
        so3_eval(so3, so3_scratch, ring, warp, lane, act, px, py, pz, r0, r1, r2);
        if (act) so3_rotate(r0, r1, r2, gx, gy, gz);
      }
    }
    if (carrier) {
      const float s = divf(step, c.x);
      const float nx = add(px, mul(s, vx)), ny = add(py, mul(s, vy)), nz = add(pz, mul(s, vz));
      vx = add(vx, mul(step, gx)); vy = add(vy, mul(step, gy)); vz = add(vz, mul(step, gz));
      t = add(t, sqrtf(sumsq3(sub(px, nx), sub(py, ny), sub(pz, nz))));
      px = nx; py = ny; pz = nz;
    }
  };

  for (int k0 = 0; k0 < n_steps; k0 += STEPS_PER_FLUSH) {
    const int nk = min(STEPS_PER_FLUSH, n_steps - k0);
    const int trow = k0 & (T_FLUSH - 1);
    if (nk == STEPS_PER_FLUSH) {
#pragma unroll
      for (int kk = 0; kk < STEPS_PER_FLUSH; ++kk) one_step(kk, trow);
    } else {
#pragma unroll
      for (int kk = 0; kk < STEPS_PER_FLUSH; ++kk)
        if (kk < nk) one_step(kk, trow);
    }
    if (!carrier) continue;                      // nothing staged, nothing to flush
    __syncwarp();
    if (want_t && (trow + nk == T_FLUSH || k0 + nk >= n_steps)) {
      // dense t column: lanes 4r..4r+3 write the 16 staged steps of ray r as four float4 (64 B, sector-complete)
      const int filled = trow + nk, kbase = k0 - trow;
      float* tb = t_col + warp_ray0 * (int64_t)n_steps + kbase;
      if (filled == T_FLUSH && t_vec) {
#pragma unroll
        for (int it = 0; it < T_FLUSH / 4; ++it) {
          const int e = it * 32 + lane, r = e >> 2, q = e & 3;
          if (r < rays_here)
            __stcs(reinterpret_cast<float4*>(tb + r * n_steps + 4 * q),
                   make_float4(ts[(4 * q) * 32 + r], ts[(4 * q + 1) * 32 + r], ts[(4 * q + 2) * 32 + r], ts[(4 * q + 3) * 32 + r]));
        }
      } else {
        for (int e = lane; e < rays_here * filled; e += 32) {
          const int r = e / filled, j = e - r * filled;
          tb[r * n_steps + j] = ts[j * 32 + r];
        }
      }
    }
    if (dbg & 1) {
    } else if (nk == STEPS_PER_FLUSH && rays_here == 32) {
      // common case: fixed trip count, per-thread offsets, one 64-bit base advanced per round
      float4* g = g_wr;
      const float4* sp = s_rd;
#pragma unroll
      for (int round = 0; round < 32 / RAYS_PER_ROUND; ++round) {
#pragma unroll
        for (int m = 0; m < NBASE; ++m) {
          if (dbg & 4) g[g_off[m]] = sp[s_off[m]]; else __stcs(g + g_off[m], sp[s_off[m]]);
        }
        g += round_stride4;
        sp += RAYS_PER_ROUND * PITCH;
      }
    } else {
      const int n4 = nk * RECF4;
      const int total4 = rays_here * n4;
      for (int e = lane; e < total4; e += 32) {
        const int r = e / n4, j = e - r * n4;
        __stcs(g_wr + r * ray_stride4 + j, s_rd[r * PITCH + j]);
      }
    }
    g_wr += F4_PER_FLUSH;
    __syncwarp();
  }
  if (SO3) ring_drain(ring);                     // weight chunks fetched ahead for an evaluation that never came
}

// "all"-stage march for SMALL launches (a training batch): the rays of a CTA are not kept in lockstep.  A ray marches on
// its own until a step needs so3_mlp, then waits; the CTA evaluates the MLP for all waiting rays together, whatever their
// step, so the number of evaluations a CTA runs in series tends to the largest active-step count of one of its rays
// instead of the union of their active steps (random pixels have little in common even after sorting,
// scripts/all_stage_batch_probe.py).  Records are stored straight to global memory (32 or 48 B per ray and step, sector
// complete); the coalesced staging of march_kernel needs lockstep and pays off only for full frames.  Per ray the
// arithmetic is that of march_kernel, so the records are bit-identical.
constexpr int RAGGED_CH = 64, RAGGED_SLOTS = 3;         // 8 chunks per evaluation through three 32 KB slots
template <int RECF4, bool FAST>
__global__ void __launch_bounds__(SO3_THREADS, 1) march_all_ragged_kernel(const float4* __restrict__ table, const MarchGeom mg,
                                                                          const float* __restrict__ origins,
                                                                          const float* __restrict__ viewdirs, int64_t n_rays,
                                                                          float near, float step, int n_steps,
                                                                          float4* __restrict__ path, float* __restrict__ t_col,
                                                                          const float* __restrict__ bricks, const So3Args so3,
                                                                          int so3_slots, int rays_per_cta) {
  extern __shared__ __align__(16) float so3_scratch[];
  constexpr int QUANTUM = 16;                    // steps a ray that needs nothing may run ahead per round
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  So3Ring ring;
  char* base = reinterpret_cast<char*>(so3_scratch);
  ring_init(ring, reinterpret_cast<float*>(base + SO3_OFF_RING), base + SO3_OFF_BARS, so3_slots, threadIdx.x,
            So3Chunking<RAGGED_CH>::SLOT_FLOATS);
  __syncthreads();
  const int64_t ray = blockIdx.x * (int64_t)rays_per_cta + threadIdx.x;
  const bool live = (int)threadIdx.x < rays_per_cta && ray < n_rays;
  const int64_t rr = live ? ray : (n_rays - 1);
  float vx = viewdirs[3 * rr], vy = viewdirs[3 * rr + 1], vz = viewdirs[3 * rr + 2];
  float px = add(origins[3 * rr], mul(near, vx)), py = add(origins[3 * rr + 1], mul(near, vy)), pz = add(origins[3 * rr + 2], mul(near, vz));
  float t = near;
  float4* rec = path + rr * (int64_t)n_steps * RECF4;
  float* tc = t_col != nullptr ? t_col + rr * (int64_t)n_steps : nullptr;
  int k = 0;
  bool done = !live;
  float n_here = 1.f;
  auto advance = [&](float gx, float gy, float gz) {
    const float s = divf(step, n_here);
    const float nx = add(px, mul(s, vx)), ny = add(py, mul(s, vy)), nz = add(pz, mul(s, vz));
    vx = add(vx, mul(step, gx)); vy = add(vy, mul(step, gy)); vz = add(vz, mul(step, gz));
    t = add(t, sqrtf(sumsq3(sub(px, nx), sub(py, ny), sub(pz, nz))));
    px = nx; py = ny; pz = nz;
    done = ++k >= n_steps;
  };
#pragma unroll 1
  while (true) {
    float gx = 0.f, gy = 0.f, gz = 0.f;
    bool need = false;
#pragma unroll 1
    for (int it = 0; it < QUANTUM && !done && !need; ++it) {
      const float4 c = march_lookup<FAST>(table, mg, bricks, px, py, pz);
      __stcs(rec + k * RECF4, make_float4(px, py, pz, t));
      __stcs(rec + k * RECF4 + 1, make_float4(vx, vy, vz, c.x));
      if (RECF4 == 3) __stcs(rec + k * RECF4 + 2, make_float4(c.y, c.z, c.w, 0.f));
      if (tc != nullptr) tc[k] = t;
      n_here = c.x; gx = c.y; gy = c.z; gz = c.w;
      need = sqrtf(sumsq3(gx, gy, gz)) > 1e-3f;          // jnp.linalg.norm(idx_grad) > 1e-3
      if (!need) advance(gx, gy, gz);
    }
    if (__syncthreads_or(need)) {
      float r0, r1, r2;
      so3_eval<RAGGED_CH>(so3, so3_scratch, ring, warp, lane, need, px, py, pz, r0, r1, r2);
      if (need) {
        so3_rotate(r0, r1, r2, gx, gy, gz);
        advance(gx, gy, gz);
      }
    } else if (!__syncthreads_or(!done)) {
      break;
    }
  }
  ring_drain(ring);
}

// VoxMLP.wrapper_grad_mlp (rnerf/ior_utils.py:225-267) on free-standing points: pred = rodrigues(so3_mlp(annealed_pos_enc(x)),
// condition), no |grad n| threshold.  What PathSampler.compute_normal_loss_and_smooth (rnerf/eikonal_utils.py:84-98) evaluates.
__global__ void __launch_bounds__(SO3_THREADS, 2) so3_predict_kernel(const float* __restrict__ pts, const float* __restrict__ cond,
                                                                     int64_t n, const So3Args so3, int so3_slots,
                                                                     float* __restrict__ pred) {
  extern __shared__ __align__(16) float so3_scratch[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  So3Ring ring;
  char* base = reinterpret_cast<char*>(so3_scratch);
  ring_init(ring, reinterpret_cast<float*>(base + SO3_OFF_RING), base + SO3_OFF_BARS, so3_slots, threadIdx.x);
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)MARCH_THREADS + threadIdx.x;
  const bool act = threadIdx.x < MARCH_THREADS && i < n;
  const float px = act ? pts[3 * i] : 0.f, py = act ? pts[3 * i + 1] : 0.f, pz = act ? pts[3 * i + 2] : 0.f;
  float r0, r1, r2;
  so3_eval(so3, so3_scratch, ring, warp, lane, act, px, py, pz, r0, r1, r2);
  if (act) {
    float gx = cond[3 * i], gy = cond[3 * i + 1], gz = cond[3 * i + 2];
    so3_rotate(r0, r1, r2, gx, gy, gz);
    pred[3 * i] = gx; pred[3 * i + 1] = gy; pred[3 * i + 2] = gz;
  }
  ring_drain(ring);
}

// The same on the tensor pipe (so3_tc.cuh): persistent CTAs, 64 points per pass.  Warp 0: weight producer (TMA ring), warp 1:
// MMA issuer, warps 2-5: encoding / epilogue / head (thread = TMEM lane = neuron).  Stand-alone form of the evaluator the
// "all"-stage march uses; also what pins its arithmetic against the CUDA-core chain (tests/test_gpu_kernels.py).
constexpr int TCP_THREADS = 192, TCP_SLOTS = 3;
struct TcPredictSmem {
  static constexpr uint32_t P_OFF = TcSmem::RING + TCP_SLOTS * TC_CHUNK_BYTES;  // positions [3][64], then raw [3][64]
  static constexpr uint32_t BAR_OFF = P_OFF + 6 * TC_N * 4;                      // full[6], empty[6], acc, act
  static constexpr uint32_t TMEM_SLOT = BAR_OFF + (2 * TCP_SLOTS + 2) * 8;
  static constexpr uint32_t BYTES = TMEM_SLOT + 16;
};
// pos_enc(dir, 0, 4) of the background MLP (rnerf/model_utils.py:204-214; same arithmetic as bkgd_mlp.cu): feature f < 27
__device__ __forceinline__ float bkgd_enc_value(int f, float d0, float d1, float d2) {
  if (f >= 27) return 0.f;
  if (f < 3) return f == 0 ? d0 : (f == 1 ? d1 : d2);
  const int q = (f - 3) % 12, k = q / 3, c = q - 3 * k;
  float xb = mul(c == 0 ? d0 : (c == 1 ? d1 : d2), (float)(1 << k));
  if (f >= 15) xb = add(xb, 1.57079632679489661923f);
  return sinf(xb);
}

// BKGD = false: VoxMLP.wrapper_grad_mlp on free-standing points (pred = rodrigues(so3_mlp(enc(pts)), cond)).
// BKGD = true: the background MLP (model_utils.MLP 27 -> 128 x4 (+27) -> 3: the same network shape as so3_mlp, so its weights
// are handed over zero-padded in so3 layout) on ray directions read with a stride; pred = the 3 raw outputs.
template <bool BKGD>
__global__ void __launch_bounds__(TCP_THREADS, 1) so3_predict_tc_kernel(const uint8_t* __restrict__ packed, const So3Args so3,
                                                                        const float* __restrict__ pts, const float* __restrict__ cond,
                                                                        int64_t n, float* __restrict__ pred, int64_t pts_stride) {
  using SL = TcPredictSmem;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  const uint32_t sbase = smem_u32(tc_smem);
  if ((sbase & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full0 = sbase + SL::BAR_OFF, bar_empty0 = bar_full0 + 8 * TCP_SLOTS;
  const uint32_t bar_acc = bar_empty0 + 8 * TCP_SLOTS, bar_act = bar_acc + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(tc_smem + SL::TMEM_SLOT);
  if (threadIdx.x == 0) {
    for (int s = 0; s < TCP_SLOTS; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_empty0 + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_act, 4);                       // the four worker warps
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(sbase + SL::TMEM_SLOT, TC_N); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t n_tiles = (n + TC_N - 1) / TC_N;
  const int64_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    if (elect_one_sync()) {                      // weight producer: the 32 chunks of an evaluation, over and over
      uint32_t c = 0;
      for (int64_t t = 0; t < my_tiles; ++t)
        for (int i = 0; i < TC_NCHUNK; ++i, ++c) {
          const uint32_t s = c % TCP_SLOTS;
          mbar_wait(bar_empty0 + 8 * s, ((c / TCP_SLOTS) & 1u) ^ 1u);
          mbar_arrive_expect_tx(bar_full0 + 8 * s, TC_CHUNK_BYTES);
          tma_bulk_g2s(sbase + TcSmem::RING + s * TC_CHUNK_BYTES, packed + (size_t)i * TC_CHUNK_BYTES, TC_CHUNK_BYTES, bar_full0 + 8 * s);
        }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {                      // MMA issuer
      uint32_t c = 0, act_phase = 0;
      for (int64_t t = 0; t < my_tiles; ++t)
        for (int l = 0; l < 4; ++l) {
          mbar_wait(bar_act, act_phase); act_phase ^= 1u;
          tc_fence_after();
          tc_issue_layer(l, sbase, tmem_base, bar_full0, bar_empty0, TCP_SLOTS, c, bar_acc, so3.dbg);
        }
    }
  } else {
    const int q = warp & 3, wt = threadIdx.x - 64, m = q * 32 + lane;
    float* P = reinterpret_cast<float*>(tc_smem + SL::P_OFF);
    float* RAW = P + 3 * TC_N;
    const float* bias = so3.w + SO3_OFF_B;
    const float* W4 = so3.w + SO3_OFF_W4;
    const float* hs = reinterpret_cast<const float*>(tc_smem + TcSmem::HS);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_phase = 0;
    auto workers_sync = [] { asm volatile("bar.sync 1, 128;" ::: "memory"); };
    auto publish = [&] {                         // generic-proxy writes -> visible to the MMA, accumulators drained
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_act);
    };
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t p0 = (blockIdx.x + t * gridDim.x) * TC_N;
      const int n_here = (int)min((int64_t)TC_N, n - p0);
      workers_sync();                            // the previous pass's head has finished with P / RAW / HS
      if (wt < TC_N) {
        const int64_t i = p0 + min(wt, n_here - 1);
        const float* src = pts + i * pts_stride;
        P[wt] = src[0]; P[TC_N + wt] = src[1]; P[2 * TC_N + wt] = src[2];
      }
      workers_sync();
      if (BKGD) {
        for (int c = warp - 2; c < TC_N; c += 4) {               // one warp per column, lane = feature (and feature + 32: zero)
          tc_enc_store(tc_smem, c, 0, lane, bkgd_enc_value(lane, P[c], P[TC_N + c], P[2 * TC_N + c]));
          tc_enc_store(tc_smem, c, 1, lane, 0.f);
        }
      } else if (!(so3.dbg & 4)) {
        const TcEncLane enc = tc_enc_lane(so3, lane);
        for (int c = warp - 2; c < TC_N; c += 4)                 // one warp per column, lane = feature (and feature + 32)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
            tc_enc_store(tc_smem, c, hh, lane, tc_enc_value(enc, hh, lane, P[c], P[TC_N + c], P[2 * TC_N + c]));
      }
      publish();
      for (int l = 0; l < 4; ++l) {
        mbar_wait(bar_acc, acc_phase); acc_phase ^= 1u;
        tc_fence_after();
        if (!(so3.dbg & 2)) {
          const float b = __ldg(bias + l * SO3_W + m);
          tc_epilogue32(l, tc_smem, tmem_lane, m, 0, b);
          tc_epilogue32(l, tc_smem, tmem_lane + 32, m, 32, b);
        }
        if (l < 3) publish();
      }
      tc_fence_before();
      workers_sync();                            // Dense_3 output complete in HS
      if (!(so3.dbg & 8))
      for (int e = wt; e < 3 * TC_N; e += 128) {     // Dense_4: raw[j][column]
        const int j = e / TC_N, cc = e - j * TC_N;
        float r = __ldg(bias + 4 * SO3_W + j);
#pragma unroll 8
        for (int k = 0; k < SO3_W; ++k) r = fmaf(hs[k * TcSmem::HS_PITCH + cc], __ldg(W4 + 3 * k + j), r);
        RAW[j * TC_N + cc] = r;
      }
      workers_sync();
      if (wt < n_here) {
        const int64_t i = p0 + wt;
        if (BKGD) {
          pred[3 * i] = RAW[wt]; pred[3 * i + 1] = RAW[TC_N + wt]; pred[3 * i + 2] = RAW[2 * TC_N + wt];
        } else {
          float gx = cond[3 * i], gy = cond[3 * i + 1], gz = cond[3 * i + 2];
          so3_rotate(RAW[wt], RAW[TC_N + wt], RAW[2 * TC_N + wt], gx, gy, gz);
          pred[3 * i] = gx; pred[3 * i + 1] = gy; pred[3 * i + 2] = gz;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TC_N);
}

// ---- "all"-stage march with so3_mlp on the tensor pipe ------------------------------------------------------------------
// Lockstep march of 256 rays per CTA (one CTA per SM: the evaluator's operands take ~160 KB of shared memory), compact
// records.  Warps 0-7 carry the rays AND are the evaluator's workers (encoding, epilogues, head); warp 8 streams the weight
// chunks (TMA ring, keeps running across evaluations), warp 9 issues the MMAs.  Per ray the march arithmetic is that of
// march_kernel; the evaluation differs from the CUDA-core chain only by the 3xTF32 products (fp32-grade, so3_tc.cuh).
constexpr int MTC_CW = 8;                            // carrier = worker warps
constexpr int MTC_RAYS = MTC_CW * 32;
constexpr int MTC_THREADS = MTC_RAYS + 64;
// Shared memory goes to the evaluator first: a 3-slot ring of 32 KB weight chunks (one k-block, hi and lo halves, per chunk:
// the k-block in use plus two of look-ahead, which is what hides the ~1 us L2 -> SM latency of a chunk behind the MMAs of a
// k-block).  The record staging is therefore half as deep as march_kernel's: flushes of 2 steps (64 contiguous bytes per ray,
// still sector-complete) and a t column flushed every 8 steps (32 bytes per ray).
constexpr int MTC_SLOTS = 3;                         // 32 KB each: the k-block in use + two of look-ahead
constexpr int MTC_SPF = 2;                           // steps per record flush
constexpr int MTC_TF = 8;                            // steps per t-column flush
constexpr int MTC_PITCH = MTC_SPF * 2 + 1;           // float4 per ray in the staging buffer (compact records + 1 pad)
struct MtcSmem {
  static constexpr uint32_t STAGE = TcSmem::RING + MTC_SLOTS * TC_CHUNK_BYTES;              // [8 warps][32 * MTC_PITCH] float4
  static constexpr uint32_t TSTAGE = STAGE + MTC_CW * 32 * MTC_PITCH * 16;                   // [8 warps][T_FLUSH * 32] float
  static constexpr uint32_t P_OFF = TSTAGE + MTC_CW * MTC_TF * 32 * 4;                       // P[3][64], RAW[3][64]
  static constexpr uint32_t PART_OFF = P_OFF + 3 * TC_N * 4;                                 // head partial sums [4][3][64] (over RAW, + 2.25 KB)
  static constexpr uint32_t W4_OFF = PART_OFF + 12 * TC_N * 4;                               // Dense_4 kernel [128][3] + bias[3] (+pad)
  static constexpr uint32_t CNT = W4_OFF + (3 * SO3_W + 4) * 4;                              // counts[8] | exit flag | chunks consumed
  static constexpr uint32_t SLOT_OFF = CNT + 64;                                              // int[64]: (ray, step) slot of each column (training)
  static constexpr uint32_t BAR_OFF = SLOT_OFF + TC_N * 4;                                    // full[4], empty[4], acc, act
  static constexpr uint32_t TMEM_SLOT = BAR_OFF + (2 * MTC_SLOTS + 2) * 8;
  static constexpr uint32_t BYTES = TMEM_SLOT + 16;
};
static_assert(MtcSmem::BYTES <= 232448, "march_tc_kernel exceeds the shared-memory budget");

__device__ __forceinline__ void mtc_workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ bool mtc_workers_or(bool p) {
  uint32_t r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbarrier.cta.red.or.pred q, 1, 256, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
               : "=r"(r) : "r"((uint32_t)p) : "memory");
  return r != 0;
}

struct MtcEval {
  uint8_t* smem;
  uint32_t tmem_base, bar_acc, bar_act, acc_phase, passes;
};

// raw = so3_mlp(annealed_pos_enc(p)) for the CTA's active rays on the tensor pipe.  All 256 worker threads call this.
__device__ __forceinline__ void so3_eval_tc(const So3Args& a, MtcEval& ev, int warp, int lane, bool act, float px, float py, float pz,
                                            float& r0, float& r1, float& r2, int slot = 0 /* ray * n_steps + step, with a.saved */) {
  uint8_t* smem = ev.smem;
  int* cnt = reinterpret_cast<int*>(smem + MtcSmem::CNT);
  int* slots = reinterpret_cast<int*>(smem + MtcSmem::SLOT_OFF);
  float* P = reinterpret_cast<float*>(smem + MtcSmem::P_OFF);
  const float* hs = reinterpret_cast<const float*>(smem + TcSmem::HS);
  const int tid = warp * 32 + lane;
  const unsigned bal = __ballot_sync(0xffffffffu, act);
  if (lane == 0) cnt[warp] = __popc(bal);
  mtc_workers_sync();
  int base = 0, n_act = 0;
#pragma unroll
  for (int w = 0; w < MTC_CW; ++w) {
    const int c = cnt[w];
    if (w < warp) base += c;
    n_act += c;
  }
  const int idx = base + __popc(bal & ((1u << lane) - 1u));
  const float* bias = a.w + SO3_OFF_B;
  const float* W4 = reinterpret_cast<const float*>(smem + MtcSmem::W4_OFF);     // staged once per CTA: [128][3], then bias[3]
  const int q = warp & 3, half = warp >> 2, m = q * 32 + lane;
  const uint32_t tmem_lane = ev.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
  auto publish = [&] {                           // generic-proxy writes -> visible to the MMA; accumulators drained
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(ev.bar_act);
  };
  r0 = r1 = r2 = 0.f;
#pragma unroll 1
  for (int col0 = 0; col0 < n_act; col0 += TC_N) {
    const int n_here = min(TC_N, n_act - col0);
    const bool mine = act && idx >= col0 && idx < col0 + TC_N;
    const int col = idx - col0;
    mtc_workers_sync();                          // counts read; the previous pass's RAW / HS are no longer needed
    if (mine) { P[col] = px; P[TC_N + col] = py; P[2 * TC_N + col] = pz; slots[col] = slot; }
    mtc_workers_sync();
    // ---- encoding: warp w takes columns w, w + 8, ...; lane = feature (two per lane: f and f + 32) -> 128-byte row writes
    if (!(a.dbg & 4)) {
      // one warp per column, lane = feature f (and f + 32); two columns per iteration give four independent sinf chains
      const TcEncLane enc = tc_enc_lane(a, lane);
      for (int c0 = warp; c0 < n_here; c0 += 2 * MTC_CW) {
        float val[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = min(c0 + u * MTC_CW, TC_N - 1);            // (a column past n_here is garbage either way)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) val[u][hh] = tc_enc_value(enc, hh, lane, P[c], P[TC_N + c], P[2 * TC_N + c]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) tc_enc_store(smem, min(c0 + u * MTC_CW, TC_N - 1), hh, lane, val[u][hh]);
      }
    }
    publish();
    const bool has_cols = half * 32 < n_here;    // (warp-uniform) this warp's 32 columns hold at least one active ray
#pragma unroll 1
    for (int l = 0; l < 4; ++l) {
      mbar_wait(ev.bar_acc, ev.acc_phase); ev.acc_phase ^= 1u;
      tc_fence_after();
      if (has_cols && !(a.dbg & 2)) tc_epilogue32(l, smem, tmem_lane, m, half * 32, __ldg(bias + l * SO3_W + m), a.saved, slots, n_here);
      if (l < 3) publish();
    }
    tc_fence_before();
    mtc_workers_sync();                          // Dense_3 output complete in HS
    // ---- Dense_4 (128 -> 3): thread (column cc = tid mod 64, k-quarter kq = tid / 64) sums 32 inputs for the three outputs
    // (a warp reads 32 consecutive columns of a row: conflict-free; the weights are broadcasts), partial sums meet in PART
    float* PART = reinterpret_cast<float*>(smem + MtcSmem::PART_OFF);         // [4 quarters][3][64]
    if (!(a.dbg & 8)) {
      const int cc = tid & 63, kq = tid >> 6;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      if (cc < n_here) {
#pragma unroll 8
        for (int k = kq * 32; k < kq * 32 + 32; ++k) {
          const float h = hs[k * TcSmem::HS_PITCH + cc];
          s0 = fmaf(h, W4[3 * k], s0); s1 = fmaf(h, W4[3 * k + 1], s1); s2 = fmaf(h, W4[3 * k + 2], s2);
        }
      }
      PART[(kq * 3 + 0) * TC_N + cc] = s0; PART[(kq * 3 + 1) * TC_N + cc] = s1; PART[(kq * 3 + 2) * TC_N + cc] = s2;
    }
    mtc_workers_sync();
    if (mine) {
      float r[3];
#pragma unroll
      for (int j = 0; j < 3; ++j)
        r[j] = ((PART[j * TC_N + col] + PART[(3 + j) * TC_N + col]) + (PART[(6 + j) * TC_N + col] + PART[(9 + j) * TC_N + col])) + W4[3 * SO3_W + j];
      r0 = r[0]; r1 = r[1]; r2 = r[2];
    }
    ++ev.passes;
  }
}

template <bool FAST>
__global__ void __launch_bounds__(MTC_THREADS, 1) march_tc_kernel(const float4* __restrict__ table, const MarchGeom mg,
                                                                  const float* __restrict__ origins, const float* __restrict__ viewdirs,
                                                                  int64_t n_rays, float near, float step, int n_steps,
                                                                  float4* __restrict__ path, float* __restrict__ t_col,
                                                                  const float* __restrict__ bricks, const So3Args so3,
                                                                  const uint8_t* __restrict__ packed) {
  using SL = MtcSmem;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  const uint32_t sbase = smem_u32(tc_smem);
  if ((sbase & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full0 = sbase + SL::BAR_OFF, bar_empty0 = bar_full0 + 8 * MTC_SLOTS;
  const uint32_t bar_acc = bar_empty0 + 8 * MTC_SLOTS, bar_act = bar_acc + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(tc_smem + SL::TMEM_SLOT);
  volatile int* flags = reinterpret_cast<volatile int*>(tc_smem + SL::CNT) + 8;      // [0] exit, [1] chunks consumed
  if (threadIdx.x == 0) {
    for (int s = 0; s < MTC_SLOTS; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_empty0 + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_act, MTC_CW);
    flags[0] = 0; flags[1] = 0;
    fence_barrier_init();
  }
  if (warp == MTC_CW + 1) { tmem_alloc(sbase + SL::TMEM_SLOT, TC_N); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == MTC_CW) {
    if (elect_one_sync()) {                      // weight producer: chunk c % 32 of the periodic stream into slot c % 4
      uint32_t c = 0;
      bool stop = false;
      while (!stop) {
        const uint32_t s = c % MTC_SLOTS, par = ((c / MTC_SLOTS) & 1u) ^ 1u;
        while (!mbar_try_wait(bar_empty0 + 8 * s, par))
          if (flags[0]) { stop = true; break; }
        if (stop) break;
        mbar_arrive_expect_tx(bar_full0 + 8 * s, TC_CHUNK_BYTES);
        tma_bulk_g2s(sbase + TcSmem::RING + s * TC_CHUNK_BYTES, packed + (size_t)(c % TC_NCHUNK) * TC_CHUNK_BYTES, TC_CHUNK_BYTES,
                     bar_full0 + 8 * s);
        ++c;
      }
      // chunks fetched ahead for an evaluation that never came must have landed before the CTA may exit
      for (uint32_t g = (uint32_t)flags[1]; g < c; ++g) mbar_wait(bar_full0 + 8 * (g % MTC_SLOTS), (g / MTC_SLOTS) & 1u);
    }
  } else if (warp == MTC_CW + 1) {
    if (elect_one_sync()) {                      // MMA issuer: one layer per "activations ready"
      uint32_t c = 0, act_phase = 0, layer = 0;
      bool stop = false;
      while (!stop) {
        while (!mbar_try_wait(bar_act, act_phase))
          if (flags[0]) { stop = true; break; }
        if (stop) break;
        act_phase ^= 1u;
        tc_fence_after();
        tc_issue_layer((int)layer, sbase, tmem_base, bar_full0, bar_empty0, MTC_SLOTS, c, bar_acc, so3.dbg);
        layer = (layer + 1) & 3u;
      }
    }
  } else {
    // ===================== the rays: 8 warps x 32 consecutive rays, in lockstep =====================
    {                                            // Dense_4 kernel + bias into shared memory (the head reads them per evaluation)
      float* w4s = reinterpret_cast<float*>(tc_smem + SL::W4_OFF);
      for (int i = threadIdx.x; i < 3 * SO3_W + 3; i += MTC_RAYS)
        w4s[i] = i < 3 * SO3_W ? __ldg(so3.w + SO3_OFF_W4 + i) : __ldg(so3.w + SO3_OFF_B + 4 * SO3_W + (i - 3 * SO3_W));
      mtc_workers_sync();
    }
    MtcEval ev;
    ev.smem = tc_smem; ev.tmem_base = tmem_base; ev.bar_acc = bar_acc; ev.bar_act = bar_act; ev.acc_phase = 0; ev.passes = 0;
    const int64_t warp_ray0 = blockIdx.x * (int64_t)MTC_RAYS + warp * 32;
    const int64_t ray = warp_ray0 + lane;
    const bool live = ray < n_rays;
    const int64_t rr = live ? ray : (n_rays - 1);
    float ox = origins[3 * rr], oy = origins[3 * rr + 1], oz = origins[3 * rr + 2];
    float vx = viewdirs[3 * rr], vy = viewdirs[3 * rr + 1], vz = viewdirs[3 * rr + 2];
    float px = add(ox, mul(near, vx)), py = add(oy, mul(near, vy)), pz = add(oz, mul(near, vz));
    float t = near;
    float4* stage_w = reinterpret_cast<float4*>(tc_smem + SL::STAGE) + warp * 32 * MTC_PITCH;
    float4* my_stage = stage_w + lane * MTC_PITCH;
    float* ts = reinterpret_cast<float*>(tc_smem + SL::TSTAGE) + warp * MTC_TF * 32;
    const int rays_here = (int)max((int64_t)0, min((int64_t)32, n_rays - warp_ray0));
    const int ray_stride4 = n_steps * 2;
    const bool t_vec = t_col != nullptr && (n_steps & 3) == 0 && (reinterpret_cast<uintptr_t>(t_col) & 15u) == 0;
    constexpr int F4 = MTC_SPF * 2;              // float4 per ray per flush (64 bytes)
    const int fr = lane >> 2, fu = lane & 3;     // flush mapping: 8 rays x 4 units per iteration
    float4* g_wr = path + warp_ray0 * (int64_t)ray_stride4;

    for (int k0 = 0; k0 < n_steps; k0 += MTC_SPF) {
      const int nk = min(MTC_SPF, n_steps - k0);
      const int trow = k0 & (MTC_TF - 1);
      for (int kk = 0; kk < nk; ++kk) {
        const float4 c = march_lookup<FAST>(table, mg, bricks, px, py, pz);
        my_stage[kk * 2 + 0] = make_float4(px, py, pz, t);
        my_stage[kk * 2 + 1] = make_float4(vx, vy, vz, c.x);
        ts[(trow + kk) * 32 + lane] = t;
        float gx = c.y, gy = c.z, gz = c.w;
        const bool act = live && sqrtf(sumsq3(gx, gy, gz)) > 1e-3f;     // jnp.linalg.norm(idx_grad) > 1e-3
        if (mtc_workers_or(act) && !(so3.dbg & 16)) {
          float r0, r1, r2;
          so3_eval_tc(so3, ev, warp, lane, act, px, py, pz, r0, r1, r2);
          if (act) so3_rotate(r0, r1, r2, gx, gy, gz);
        }
        const float s = divf(step, c.x);
        const float nx = add(px, mul(s, vx)), ny = add(py, mul(s, vy)), nz = add(pz, mul(s, vz));
        vx = add(vx, mul(step, gx)); vy = add(vy, mul(step, gy)); vz = add(vz, mul(step, gz));
        t = add(t, sqrtf(sumsq3(sub(px, nx), sub(py, ny), sub(pz, nz))));
        px = nx; py = ny; pz = nz;
      }
      __syncwarp();
      if (t_col != nullptr && (trow + nk == MTC_TF || k0 + nk >= n_steps)) {
        const int filled = trow + nk, kbase = k0 - trow;
        float* tb = t_col + warp_ray0 * (int64_t)n_steps + kbase;
        if (filled == MTC_TF && t_vec) {
#pragma unroll
          for (int it = 0; it < MTC_TF / 4; ++it) {      // lanes 2r, 2r + 1 write the 8 staged steps of ray r (32 B, one sector)
            const int e = it * 32 + lane, r = e >> 1, qd = e & 1;
            if (r < rays_here)
              __stcs(reinterpret_cast<float4*>(tb + r * n_steps + 4 * qd),
                     make_float4(ts[(4 * qd) * 32 + r], ts[(4 * qd + 1) * 32 + r], ts[(4 * qd + 2) * 32 + r], ts[(4 * qd + 3) * 32 + r]));
          }
        } else {
          for (int e = lane; e < rays_here * filled; e += 32) {
            const int r = e / filled, j = e - r * filled;
            tb[r * n_steps + j] = ts[j * 32 + r];
          }
        }
      }
      if (nk == MTC_SPF && rays_here == 32) {
#pragma unroll
        for (int round = 0; round < 4; ++round) {
          const int r = round * 8 + fr;
          __stcs(g_wr + r * ray_stride4 + fu, stage_w[r * MTC_PITCH + fu]);
        }
      } else {
        const int n4 = nk * 2, total4 = rays_here * n4;
        for (int e = lane; e < total4; e += 32) {
          const int r = e / n4, j = e - r * n4;
          __stcs(g_wr + r * ray_stride4 + j, stage_w[r * MTC_PITCH + j]);
        }
      }
      g_wr += F4;
      __syncwarp();
    }
    mtc_workers_sync();
    if (threadIdx.x == 0) {
      flags[1] = (int)(ev.passes * TC_NCHUNK);
      __threadfence_block();
      flags[0] = 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MTC_CW + 1) tmem_dealloc(tmem_base, TC_N);
}

// The ragged "all"-stage march (march_all_ragged_kernel) with the tensor-pipe evaluator: small launches (a training batch),
// 32 or 64 rays per CTA so that every SM is busy, rays not in lockstep.  Threads 0 .. rays_per_cta-1 carry a ray each, all
// eight worker warps take part in an evaluation, warps 8 / 9 stream the weights / issue the MMAs as in march_tc_kernel.
template <int RECF4, bool FAST>
__global__ void __launch_bounds__(MTC_THREADS, 1) march_ragged_tc_kernel(const float4* __restrict__ table, const MarchGeom mg,
                                                                         const float* __restrict__ origins,
                                                                         const float* __restrict__ viewdirs, int64_t n_rays,
                                                                         float near, float step, int n_steps,
                                                                         float4* __restrict__ path, float* __restrict__ t_col,
                                                                         const float* __restrict__ bricks, const So3Args so3,
                                                                         const uint8_t* __restrict__ packed, int rays_per_cta) {
  using SL = MtcSmem;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  const uint32_t sbase = smem_u32(tc_smem);
  if ((sbase & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full0 = sbase + SL::BAR_OFF, bar_empty0 = bar_full0 + 8 * MTC_SLOTS;
  const uint32_t bar_acc = bar_empty0 + 8 * MTC_SLOTS, bar_act = bar_acc + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(tc_smem + SL::TMEM_SLOT);
  volatile int* flags = reinterpret_cast<volatile int*>(tc_smem + SL::CNT) + 8;      // [0] exit, [1] chunks consumed
  if (threadIdx.x == 0) {
    for (int s = 0; s < MTC_SLOTS; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_empty0 + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_act, MTC_CW);
    flags[0] = 0; flags[1] = 0;
    fence_barrier_init();
  }
  if (warp == MTC_CW + 1) { tmem_alloc(sbase + SL::TMEM_SLOT, TC_N); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == MTC_CW) {
    if (elect_one_sync()) {                      // weight producer (see march_tc_kernel)
      uint32_t c = 0;
      bool stop = false;
      while (!stop) {
        const uint32_t s = c % MTC_SLOTS, par = ((c / MTC_SLOTS) & 1u) ^ 1u;
        while (!mbar_try_wait(bar_empty0 + 8 * s, par))
          if (flags[0]) { stop = true; break; }
        if (stop) break;
        mbar_arrive_expect_tx(bar_full0 + 8 * s, TC_CHUNK_BYTES);
        tma_bulk_g2s(sbase + TcSmem::RING + s * TC_CHUNK_BYTES, packed + (size_t)(c % TC_NCHUNK) * TC_CHUNK_BYTES, TC_CHUNK_BYTES,
                     bar_full0 + 8 * s);
        ++c;
      }
      for (uint32_t g = (uint32_t)flags[1]; g < c; ++g) mbar_wait(bar_full0 + 8 * (g % MTC_SLOTS), (g / MTC_SLOTS) & 1u);
    }
  } else if (warp == MTC_CW + 1) {
    if (elect_one_sync()) {                      // MMA issuer
      uint32_t c = 0, act_phase = 0, layer = 0;
      bool stop = false;
      while (!stop) {
        while (!mbar_try_wait(bar_act, act_phase))
          if (flags[0]) { stop = true; break; }
        if (stop) break;
        act_phase ^= 1u;
        tc_fence_after();
        tc_issue_layer((int)layer, sbase, tmem_base, bar_full0, bar_empty0, MTC_SLOTS, c, bar_acc, so3.dbg);
        layer = (layer + 1) & 3u;
      }
    }
  } else {
    {
      float* w4s = reinterpret_cast<float*>(tc_smem + SL::W4_OFF);
      for (int i = threadIdx.x; i < 3 * SO3_W + 3; i += MTC_RAYS)
        w4s[i] = i < 3 * SO3_W ? __ldg(so3.w + SO3_OFF_W4 + i) : __ldg(so3.w + SO3_OFF_B + 4 * SO3_W + (i - 3 * SO3_W));
      mtc_workers_sync();
    }
    MtcEval ev;
    ev.smem = tc_smem; ev.tmem_base = tmem_base; ev.bar_acc = bar_acc; ev.bar_act = bar_act; ev.acc_phase = 0; ev.passes = 0;
    constexpr int QUANTUM = 16;                  // steps a ray that needs nothing may run ahead per round
    const int64_t ray = blockIdx.x * (int64_t)rays_per_cta + threadIdx.x;
    const bool live = (int)threadIdx.x < rays_per_cta && ray < n_rays;
    const int64_t rr = live ? ray : (n_rays - 1);
    float vx = viewdirs[3 * rr], vy = viewdirs[3 * rr + 1], vz = viewdirs[3 * rr + 2];
    float px = add(origins[3 * rr], mul(near, vx)), py = add(origins[3 * rr + 1], mul(near, vy)), pz = add(origins[3 * rr + 2], mul(near, vz));
    float t = near;
    float4* rec = path + rr * (int64_t)n_steps * RECF4;
    float* tc = t_col != nullptr ? t_col + rr * (int64_t)n_steps : nullptr;
    int k = 0;
    bool done = !live;
    float n_here = 1.f;
    auto advance = [&](float gx, float gy, float gz) {
      const float s = divf(step, n_here);
      const float nx = add(px, mul(s, vx)), ny = add(py, mul(s, vy)), nz = add(pz, mul(s, vz));
      vx = add(vx, mul(step, gx)); vy = add(vy, mul(step, gy)); vz = add(vz, mul(step, gz));
      t = add(t, sqrtf(sumsq3(sub(px, nx), sub(py, ny), sub(pz, nz))));
      px = nx; py = ny; pz = nz;
      done = ++k >= n_steps;
    };
#pragma unroll 1
    while (true) {
      float gx = 0.f, gy = 0.f, gz = 0.f;
      bool need = false;
#pragma unroll 1
      for (int it = 0; it < QUANTUM && !done && !need; ++it) {
        const float4 c = march_lookup<FAST>(table, mg, bricks, px, py, pz);
        __stcs(rec + k * RECF4, make_float4(px, py, pz, t));
        __stcs(rec + k * RECF4 + 1, make_float4(vx, vy, vz, c.x));
        if (RECF4 == 3) __stcs(rec + k * RECF4 + 2, make_float4(c.y, c.z, c.w, 0.f));
        if (tc != nullptr) tc[k] = t;
        n_here = c.x; gx = c.y; gy = c.z; gz = c.w;
        need = sqrtf(sumsq3(gx, gy, gz)) > 1e-3f;          // jnp.linalg.norm(idx_grad) > 1e-3
        if (!need) advance(gx, gy, gz);
      }
      if (mtc_workers_or(need)) {
        float r0, r1, r2;
        so3_eval_tc(so3, ev, warp, lane, need, px, py, pz, r0, r1, r2, (int)(rr * n_steps + k));
        if (need) {
          so3_rotate(r0, r1, r2, gx, gy, gz);
          advance(gx, gy, gz);
        }
      } else if (!mtc_workers_or(!done)) {
        break;
      }
    }
    mtc_workers_sync();
    if (threadIdx.x == 0) {
      flags[1] = (int)(ev.passes * TC_NCHUNK);
      __threadfence_block();
      flags[0] = 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MTC_CW + 1) tmem_dealloc(tmem_base, TC_N);
}

// ray_dir of every record, normalised: the array PathSampler returns (rnerf/eikonal_utils.py:113)
__global__ void __launch_bounds__(256) path_dirs_kernel(const float4* __restrict__ path, int recf4, int64_t n_rec,
                                                        float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rec) return;
  float3 d = path_dir(__ldg(path + i * recf4 + 1));
  out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
}

// rnerf/models.py:243-247: ray_pos[:, jitter] etc.
__global__ void __launch_bounds__(256) select_kernel(const float4* __restrict__ path, int recf4, int64_t n_rays, int n_steps,
                                                     const int32_t* __restrict__ jitter, int n_coarse,
                                                     float* __restrict__ pos_c, float* __restrict__ dir_c,
                                                     float* __restrict__ t_c, float* __restrict__ grad_c) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rays * n_coarse) return;
  int64_t r = i / n_coarse;
  int j = (int)(i % n_coarse);
  int k = min(max(__ldg(jitter + j), 0), n_steps - 1);
  const float4* rec = path + (r * n_steps + k) * recf4;
  float4 a = __ldg(rec), b = __ldg(rec + 1);
  pos_c[3 * i] = a.x; pos_c[3 * i + 1] = a.y; pos_c[3 * i + 2] = a.z;
  t_c[i] = a.w;
  const float3 dn = path_dir(b);
  dir_c[3 * i] = dn.x; dir_c[3 * i + 1] = dn.y; dir_c[3 * i + 2] = dn.z;
  if (grad_c) {
    float4 c = __ldg(rec + 2);
    grad_c[3 * i] = c.x; grad_c[3 * i + 1] = c.y; grad_c[3 * i + 2] = c.z;
  }
}

}  // namespace rnerf

using namespace rnerf;

// RNERF_MARCH_DIV=ieee forces IEEE divisions for the grid coordinates (results are identical either way)
bool rnerf::fast_div_enabled() {
  const char* e = getenv("RNERF_MARCH_DIV");      // read per call so that a test can compare both modes in one process
  return !(e != nullptr && strcmp(e, "ieee") == 0);
}

// rays per CTA of the so3 kernels (forward and reverse sweep agree): 128, or fewer when that still leaves SMs idle
int rnerf::so3_rays_per_cta(int64_t n_rays, int n_sm) {
  int rpc = n_rays <= (int64_t)32 * n_sm ? 32 : (n_rays <= (int64_t)64 * n_sm ? 64 : MARCH_THREADS);
  const char* e = getenv("RNERF_SO3_RPC");      // development aid
  if (e != nullptr && (atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 128)) rpc = atoi(e);
  return rpc;
}

bool rnerf::make_march_geom(const int ndim[3], const double nmin[3], const double nmax[3], cudaStream_t st, MarchGeom& mg) {
  mg.g = make_geom(ndim, nmin, nmax);
  mg.nby = (ndim[1] + BRICK - 1) >> BRICK_LOG2;
  mg.nbz = (ndim[2] + BRICK - 1) >> BRICK_LOG2;
  bool fast = fast_div_enabled();
  for (int i = 0; i < 3; ++i) {
    mg.rdelta[i] = fast ? recip_for(mg.g.ndelta[i], st) : 0.f;
    fast = fast && mg.rdelta[i] != 0.f;
  }
  return fast;
}

static int march_impl(const float* table, const float* bricks, const int ndim[3], const double nmin[3],
                      const double nmax[3], const float* origins, const float* viewdirs, int64_t n_rays,
                      double near, double far, int n_steps, int rec_floats, const float* so3_w,
                      const double* so3_window, const float* so3_window_dev, const void* so3_tc_packed, float* path, float* t_col,
                      void* stream, float* so3_saved = nullptr) {
  RNERF_REQUIRE_PTR(table); RNERF_REQUIRE_PTR(ndim); RNERF_REQUIRE_PTR(nmin); RNERF_REQUIRE_PTR(nmax);
  RNERF_REQUIRE(n_rays >= 0, RNERF_E_SHAPE, "rnerf_march_fwd: n_rays < 0");
  RNERF_REQUIRE(n_steps >= 2, RNERF_E_SHAPE, "rnerf_march_fwd: n_steps must be >= 2 (step = (far-near)/(S-1))");
  RNERF_REQUIRE(rec_floats == 8 || rec_floats == 12, RNERF_E_SHAPE, "rnerf_march_fwd: rec_floats must be 8 (compact) or 12 (full)");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(origins); RNERF_REQUIRE_PTR(viewdirs); RNERF_REQUIRE_PTR(path);
  RNERF_REQUIRE(aligned16(table) && aligned16(path), RNERF_E_ALIGN, "rnerf_march_fwd: table/path must be 16-byte aligned");
  RNERF_REQUIRE(grid_fits_int32(ndim), RNERF_E_SHAPE, "rnerf_march_fwd: grids with >= 2^31 voxels are not supported");
  RNERF_REQUIRE((double)n_steps * 12 * 32 < 2147483648.0, RNERF_E_SHAPE, "rnerf_march_fwd: n_steps too large");
  cudaStream_t st = (cudaStream_t)stream;
  MarchGeom mg;
  const bool fast = make_march_geom(ndim, nmin, nmax, st, mg);
  const float step = (float)((far - near) / (n_steps - 1));
  const char* dbg_env = getenv("RNERF_MARCH_DEBUG");   // development aid: 1 = no record stores, 2 = no t stores, 4 = plain stores
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  unsigned blocks = (unsigned)((n_rays + MARCH_THREADS - 1) / MARCH_THREADS);
  int rpc = MARCH_THREADS;
  if (so3_w == nullptr) {
    // radiance stage: rays per warp (see march_kernel) halved while the launch still has <= 7 warps per SM (measured on random
    // pixels of the ship frame, scripts/march_rpw_probe.py: 512 rays 0.51 -> 0.35 ms at 1, 4096 rays 0.53 -> 0.46 at 4, 16 384
    // unchanged at 16; with more warps the shadow lanes cost more issue slots than the shorter waits save); full frames keep 32
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    int rpw = 32;
    while (rpw > 1 && (n_rays + rpw / 2 - 1) / (rpw / 2) <= (int64_t)n_sm * 7) rpw >>= 1;
    if (const char* e = getenv("RNERF_MARCH_RPW")) {     // development aid: 1, 2, 4, 8, 16 or 32
      const int v = atoi(e);
      if (v >= 1 && v <= 32 && (v & (v - 1)) == 0) rpw = v;
    }
    rpc = rpw * (MARCH_THREADS / 32);
    blocks = (unsigned)((n_rays + rpc - 1) / rpc);
  }
  So3Args so3;
  memset(&so3, 0, sizeof(so3));
  size_t dyn = 0;
  int slots = 0;
  if (so3_w != nullptr) {
    so3.w = so3_w;
    for (int k = 0; k < 10; ++k) so3.window[k] = so3_window != nullptr ? (float)so3_window[k] : 0.f;
    so3.window_dev = so3_window_dev;
    { const char* d = getenv("RNERF_SO3_TC_DEBUG"); so3.dbg = d ? atoi(d) : 0; }
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    rpc = so3_rays_per_cta(n_rays, n_sm);
    blocks = (unsigned)((n_rays + rpc - 1) / rpc);
    // Small launches (a training batch) leave at most one CTA per SM: a deep ring (12 chunks = 96 KB in flight) hides
    // the L2 latency alone.  Full frames keep two CTAs per SM (110 KB each) with a 4-slot ring and hide it with the
    // other CTA.  RNERF_SO3_SLOTS overrides (development aid).
    slots = blocks <= (unsigned)n_sm ? 13 : 4;
    const char* se = getenv("RNERF_SO3_SLOTS");
    if (se != nullptr && atoi(se) >= 2 && atoi(se) <= SO3_MAX_SLOTS) slots = atoi(se);
    dyn = so3_smem_bytes(slots);
    static size_t attr_set[64] = {0};
    if (dev >= 0 && dev < 64 && attr_set[dev] < dyn) {
      cudaError_t e = cudaSuccess;
      e = cudaFuncSetAttribute(march_kernel<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(march_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(march_kernel<3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(march_kernel<3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      if (e != cudaSuccess) { set_error("rnerf_march_all_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
      attr_set[dev] = dyn;
    }
  }
  // Full frames with compact records: so3_mlp on the tensor pipe (march_tc_kernel), when the caller supplies the packed
  // hi/lo weight image.  RNERF_SO3_TC=0 keeps the CUDA-core chain (development aid for A/B runs).
  if (so3_w != nullptr && so3_tc_packed != nullptr && rec_floats == 8 && rpc == MARCH_THREADS &&
      !(getenv("RNERF_SO3_TC") != nullptr && atoi(getenv("RNERF_SO3_TC")) == 0)) {
    RNERF_REQUIRE(aligned16(so3_tc_packed), RNERF_E_ALIGN, "rnerf_march_all_fwd: so3_tc_packed must be 16-byte aligned");
    RNERF_REQUIRE(so3_saved == nullptr, RNERF_E_SHAPE, "rnerf_march_all_fwd: so3_saved given for a full-frame launch (see rnerf_so3_saved_floats)");
    cudaError_t e = cudaFuncSetAttribute(march_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MtcSmem::BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(march_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MtcSmem::BYTES);
    if (e != cudaSuccess) { set_error("rnerf_march_all_fwd: cudaFuncSetAttribute(tc): %s", cudaGetErrorString(e)); return (int)e; }
    const unsigned tb = (unsigned)((n_rays + MTC_RAYS - 1) / MTC_RAYS);
    if (fast)
      march_tc_kernel<true><<<tb, MTC_THREADS, MtcSmem::BYTES, st>>>((const float4*)table, mg, origins, viewdirs, n_rays, (float)near, step,
                                                                    n_steps, (float4*)path, t_col, bricks, so3, (const uint8_t*)so3_tc_packed);
    else
      march_tc_kernel<false><<<tb, MTC_THREADS, MtcSmem::BYTES, st>>>((const float4*)table, mg, origins, viewdirs, n_rays, (float)near, step,
                                                                     n_steps, (float4*)path, t_col, bricks, so3, (const uint8_t*)so3_tc_packed);
    count_launch();
    return check_launch("rnerf_march_all_fwd(tc)");
  }
#define RNERF_MARCH_LAUNCH(R, F, A)                                                                                       \
  march_kernel<R, F, A><<<blocks, (A) ? SO3_THREADS : MARCH_THREADS, dyn, st>>>((const float4*)table, mg, origins, viewdirs, n_rays, (float)near, \
                                                            step, n_steps, (float4*)path, t_col, bricks, dbg, so3, slots, rpc)
  if (so3_w != nullptr && rpc < MARCH_THREADS && so3_tc_packed != nullptr &&
      !(getenv("RNERF_SO3_TC") != nullptr && atoi(getenv("RNERF_SO3_TC")) == 0)) {
    // a small launch with the packed hi/lo image: the ragged march with the tensor-pipe evaluator
    RNERF_REQUIRE(aligned16(so3_tc_packed), RNERF_E_ALIGN, "rnerf_march_all_fwd: so3_tc_packed must be 16-byte aligned");
    RNERF_REQUIRE(so3_saved == nullptr || (double)n_rays * n_steps < 2147483648.0, RNERF_E_SHAPE,
                  "rnerf_march_all_fwd: so3_saved needs n_rays * n_steps < 2^31");
    so3.saved = so3_saved;
    so3_saved = nullptr;        // consumed
    cudaError_t e = cudaSuccess;
#define RNERF_RAGGED_TC_LAUNCH(R, F)                                                                                         \
    e = cudaFuncSetAttribute(march_ragged_tc_kernel<R, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MtcSmem::BYTES);   \
    if (e == cudaSuccess)                                                                                                      \
      march_ragged_tc_kernel<R, F><<<blocks, MTC_THREADS, MtcSmem::BYTES, st>>>((const float4*)table, mg, origins, viewdirs, n_rays, \
                                                                                (float)near, step, n_steps, (float4*)path, t_col,  \
                                                                                bricks, so3, (const uint8_t*)so3_tc_packed, rpc)
    if (rec_floats == 8) { if (fast) { RNERF_RAGGED_TC_LAUNCH(2, true); } else { RNERF_RAGGED_TC_LAUNCH(2, false); } }
    else                 { if (fast) { RNERF_RAGGED_TC_LAUNCH(3, true); } else { RNERF_RAGGED_TC_LAUNCH(3, false); } }
#undef RNERF_RAGGED_TC_LAUNCH
    if (e != cudaSuccess) { set_error("rnerf_march_all_fwd: cudaFuncSetAttribute(ragged tc): %s", cudaGetErrorString(e)); return (int)e; }
    count_launch();
    return check_launch("rnerf_march_all_fwd(ragged tc)");
  }
  RNERF_REQUIRE(so3_saved == nullptr, RNERF_E_SHAPE,
                "rnerf_march_all_fwd: so3_saved is only written by the ragged tensor-pipe march (small launch + so3_tc_packed); "
                "rnerf_so3_saved_floats() says whether a launch qualifies");
  if (so3_w != nullptr && rpc < MARCH_THREADS) {          // a small launch: rays not in lockstep (see march_all_ragged_kernel)
    cudaError_t e = cudaSuccess;
    slots = RAGGED_SLOTS;
    dyn = so3_smem_bytes(slots, RAGGED_CH);
#define RNERF_RAGGED_LAUNCH(R, F)                                                                                           \
    e = cudaFuncSetAttribute(march_all_ragged_kernel<R, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);             \
    if (e == cudaSuccess)                                                                                                      \
      march_all_ragged_kernel<R, F><<<blocks, SO3_THREADS, dyn, st>>>((const float4*)table, mg, origins, viewdirs, n_rays,       \
                                                                      (float)near, step, n_steps, (float4*)path, t_col, bricks, \
                                                                      so3, slots, rpc)
    if (rec_floats == 8) { if (fast) { RNERF_RAGGED_LAUNCH(2, true); } else { RNERF_RAGGED_LAUNCH(2, false); } }
    else                 { if (fast) { RNERF_RAGGED_LAUNCH(3, true); } else { RNERF_RAGGED_LAUNCH(3, false); } }
#undef RNERF_RAGGED_LAUNCH
    if (e != cudaSuccess) { set_error("rnerf_march_all_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  } else if (so3_w == nullptr) {
    if (rec_floats == 8) { if (fast) RNERF_MARCH_LAUNCH(2, true, false); else RNERF_MARCH_LAUNCH(2, false, false); }
    else                 { if (fast) RNERF_MARCH_LAUNCH(3, true, false); else RNERF_MARCH_LAUNCH(3, false, false); }
  } else {
    if (rec_floats == 8) { if (fast) RNERF_MARCH_LAUNCH(2, true, true); else RNERF_MARCH_LAUNCH(2, false, true); }
    else                 { if (fast) RNERF_MARCH_LAUNCH(3, true, true); else RNERF_MARCH_LAUNCH(3, false, true); }
  }
#undef RNERF_MARCH_LAUNCH
  count_launch();
  return check_launch(so3_w ? "rnerf_march_all_fwd" : "rnerf_march_fwd");
}

extern "C" int rnerf_march_fwd(const float* table, const float* bricks, const int ndim[3], const double nmin[3],
                               const double nmax[3], const float* origins, const float* viewdirs, int64_t n_rays,
                               double near, double far, int n_steps, int rec_floats, float* path, float* t_col,
                               void* stream) {
  return march_impl(table, bricks, ndim, nmin, nmax, origins, viewdirs, n_rays, near, far, n_steps, rec_floats, nullptr,
                    nullptr, nullptr, nullptr, path, t_col, stream);
}

extern "C" size_t rnerf_so3_weight_floats(void) { return SO3_FLOATS; }

extern "C" int rnerf_march_all_fwd(const float* table, const float* bricks, const int ndim[3], const double nmin[3],
                                   const double nmax[3], const float* origins, const float* viewdirs, int64_t n_rays,
                                   double near, double far, int n_steps, int rec_floats, const float* so3_w,
                                   const double so3_window[10], const float* so3_window_dev, const void* so3_tc_packed,
                                   float* so3_saved, float* path, float* t_col, void* stream) {
  RNERF_REQUIRE_PTR(so3_w);
  RNERF_REQUIRE(so3_window != nullptr || so3_window_dev != nullptr, RNERF_E_NULL, "rnerf_march_all_fwd: no so3 window given");
  return march_impl(table, bricks, ndim, nmin, nmax, origins, viewdirs, n_rays, near, far, n_steps, rec_floats, so3_w,
                    so3_window, so3_window_dev, so3_tc_packed, path, t_col, stream, so3_saved);
}

// Floats of the `so3_saved` buffer a training forward of n_rays x n_steps may hand to rnerf_march_all_fwd, or 0 when that
// launch would not run the ragged tensor-pipe march (the only kernel that writes it).
extern "C" size_t rnerf_so3_saved_floats(int64_t n_rays, int n_steps) {
  if (n_rays <= 0 || n_steps <= 0 || (double)n_rays * n_steps >= 2147483648.0) return 0;
  if (getenv("RNERF_SO3_TC") != nullptr && atoi(getenv("RNERF_SO3_TC")) == 0) return 0;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (so3_rays_per_cta(n_rays, n_sm) >= MARCH_THREADS) return 0;
  return (size_t)n_rays * n_steps * SO3_SAVED_FLOATS;
}

extern "C" int rnerf_select(const float* path, int rec_floats, int64_t n_rays, int n_steps, const int32_t* jitter,
                            int n_coarse, float* pos_c, float* dir_c, float* t_c, float* grad_c, void* stream) {
  RNERF_REQUIRE(n_rays >= 0 && n_steps > 0 && n_coarse > 0, RNERF_E_SHAPE, "rnerf_select: bad sizes");
  RNERF_REQUIRE(rec_floats == 8 || rec_floats == 12, RNERF_E_SHAPE, "rnerf_select: rec_floats must be 8 or 12");
  RNERF_REQUIRE(grad_c == nullptr || rec_floats == 12, RNERF_E_SHAPE, "rnerf_select: grad_c needs the full (12-float) path records");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(path); RNERF_REQUIRE_PTR(jitter); RNERF_REQUIRE_PTR(pos_c); RNERF_REQUIRE_PTR(dir_c); RNERF_REQUIRE_PTR(t_c);
  RNERF_REQUIRE(aligned16(path), RNERF_E_ALIGN, "rnerf_select: path must be 16-byte aligned");
  int64_t total = n_rays * n_coarse;
  select_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)path, rec_floats / 4, n_rays,
                                                                                  n_steps, jitter, n_coarse, pos_c, dir_c,
                                                                                  t_c, grad_c);
  count_launch();
  return check_launch("rnerf_select");
}

extern "C" int rnerf_path_dirs(const float* path, int rec_floats, int64_t n_rays, int n_steps, float* ray_dir, void* stream) {
  RNERF_REQUIRE(n_rays >= 0 && n_steps > 0, RNERF_E_SHAPE, "rnerf_path_dirs: bad sizes");
  RNERF_REQUIRE(rec_floats == 8 || rec_floats == 12, RNERF_E_SHAPE, "rnerf_path_dirs: rec_floats must be 8 or 12");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(path); RNERF_REQUIRE_PTR(ray_dir);
  RNERF_REQUIRE(aligned16(path), RNERF_E_ALIGN, "rnerf_path_dirs: path must be 16-byte aligned");
  const int64_t n = n_rays * n_steps;
  path_dirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)path, rec_floats / 4, n, ray_dir);
  count_launch();
  return check_launch("rnerf_path_dirs");
}

extern "C" size_t rnerf_so3_tc_packed_bytes(void) { return TC_PACKED_BYTES; }

extern "C" int rnerf_so3_tc_pack(const float* so3_w, void* packed, void* stream) {
  RNERF_REQUIRE_PTR(so3_w); RNERF_REQUIRE_PTR(packed);
  RNERF_REQUIRE(aligned16(packed), RNERF_E_ALIGN, "rnerf_so3_tc_pack: packed must be 16-byte aligned");
  so3_tc_pack_kernel<<<(TC_NKB * SO3_W * TC_KBLK + 255) / 256, 256, 0, (cudaStream_t)stream>>>(so3_w, (uint8_t*)packed);
  count_launch();
  return check_launch("rnerf_so3_tc_pack");
}

extern "C" int rnerf_so3_predict_tc(const void* so3_tc_packed, const float* so3_w, const double so3_window[10],
                                    const float* so3_window_dev, const float* pts, const float* cond, int64_t n, float* pred,
                                    void* stream) {
  RNERF_REQUIRE(n >= 0, RNERF_E_SHAPE, "rnerf_so3_predict_tc: n < 0");
  if (n == 0) return 0;
  RNERF_REQUIRE_PTR(so3_tc_packed); RNERF_REQUIRE_PTR(so3_w); RNERF_REQUIRE_PTR(pts); RNERF_REQUIRE_PTR(cond); RNERF_REQUIRE_PTR(pred);
  RNERF_REQUIRE(so3_window != nullptr || so3_window_dev != nullptr, RNERF_E_NULL, "rnerf_so3_predict_tc: no so3 window given");
  RNERF_REQUIRE(aligned16(so3_tc_packed) && aligned16(so3_w), RNERF_E_ALIGN, "rnerf_so3_predict_tc: weight images must be 16-byte aligned");
  So3Args so3;
  memset(&so3, 0, sizeof(so3));
  so3.w = so3_w;
  for (int k = 0; k < 10; ++k) so3.window[k] = so3_window != nullptr ? (float)so3_window[k] : 0.f;
  so3.window_dev = so3_window_dev;
  { const char* d = getenv("RNERF_SO3_TC_DEBUG"); so3.dbg = d ? atoi(d) : 0; }
  cudaError_t e = cudaFuncSetAttribute(so3_predict_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcPredictSmem::BYTES);
  if (e != cudaSuccess) { set_error("rnerf_so3_predict_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (n + TC_N - 1) / TC_N;
  so3_predict_tc_kernel<false><<<(unsigned)(tiles < n_sm ? tiles : n_sm), TCP_THREADS, TcPredictSmem::BYTES, (cudaStream_t)stream>>>(
      (const uint8_t*)so3_tc_packed, so3, pts, cond, n, pred, 3);
  count_launch();
  return check_launch("rnerf_so3_predict_tc");
}

// The background MLP on the tensor pipe (render path: one evaluation per ray of a frame).  bkgd_so3 = the background weights
// zero-padded into so3_mlp's layout (rnerf_so3_weight_floats() floats: Dense_0 rows 27..59 and Dense_3 rows 155..187 zero),
// bkgd_tc_packed = rnerf_so3_tc_pack(bkgd_so3).  dirs is read with a ray stride like rnerf_bkgd_mlp_fwd.
extern "C" int rnerf_bkgd_mlp_fwd_tc(const void* bkgd_tc_packed, const float* bkgd_so3, const float* dirs, int64_t n_rays,
                                     int64_t dir_stride_floats, float* raw_out, void* stream) {
  RNERF_REQUIRE(n_rays >= 0 && dir_stride_floats >= 3, RNERF_E_SHAPE, "rnerf_bkgd_mlp_fwd_tc: bad sizes");
  if (n_rays == 0) return 0;
  RNERF_REQUIRE_PTR(bkgd_tc_packed); RNERF_REQUIRE_PTR(bkgd_so3); RNERF_REQUIRE_PTR(dirs); RNERF_REQUIRE_PTR(raw_out);
  RNERF_REQUIRE(aligned16(bkgd_tc_packed) && aligned16(bkgd_so3), RNERF_E_ALIGN, "rnerf_bkgd_mlp_fwd_tc: weight images must be 16-byte aligned");
  So3Args so3;
  memset(&so3, 0, sizeof(so3));
  so3.w = bkgd_so3;
  cudaError_t e = cudaFuncSetAttribute(so3_predict_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcPredictSmem::BYTES);
  if (e != cudaSuccess) { set_error("rnerf_bkgd_mlp_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (n_rays + TC_N - 1) / TC_N;
  so3_predict_tc_kernel<true><<<(unsigned)(tiles < n_sm ? tiles : n_sm), TCP_THREADS, TcPredictSmem::BYTES, (cudaStream_t)stream>>>(
      (const uint8_t*)bkgd_tc_packed, so3, dirs, nullptr, n_rays, raw_out, dir_stride_floats);
  count_launch();
  return check_launch("rnerf_bkgd_mlp_fwd_tc");
}

extern "C" int rnerf_so3_predict(const float* so3_w, const double so3_window[10], const float* so3_window_dev, const float* pts,
                                 const float* cond, int64_t n, float* pred, void* stream) {
  RNERF_REQUIRE(n >= 0, RNERF_E_SHAPE, "rnerf_so3_predict: n < 0");
  if (n == 0) return 0;
  RNERF_REQUIRE_PTR(so3_w); RNERF_REQUIRE_PTR(pts); RNERF_REQUIRE_PTR(cond); RNERF_REQUIRE_PTR(pred);
  RNERF_REQUIRE(so3_window != nullptr || so3_window_dev != nullptr, RNERF_E_NULL, "rnerf_so3_predict: no so3 window given");
  RNERF_REQUIRE(aligned16(so3_w), RNERF_E_ALIGN, "rnerf_so3_predict: so3_w must be 16-byte aligned");
  So3Args so3;
  memset(&so3, 0, sizeof(so3));
  so3.w = so3_w;
  for (int k = 0; k < 10; ++k) so3.window[k] = so3_window != nullptr ? (float)so3_window[k] : 0.f;
  so3.window_dev = so3_window_dev;
  const int slots = 4;
  const size_t dyn = so3_smem_bytes(slots);
  cudaError_t e = cudaFuncSetAttribute(so3_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) { set_error("rnerf_so3_predict: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  so3_predict_kernel<<<(unsigned)((n + MARCH_THREADS - 1) / MARCH_THREADS), SO3_THREADS, dyn, (cudaStream_t)stream>>>(pts, cond, n, so3,
                                                                                                                   slots, pred);
  count_launch();
  return check_launch("rnerf_so3_predict");
}
