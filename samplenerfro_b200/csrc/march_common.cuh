// Definitions shared by the forward march (march.cu) and its adjoint (march_bwd.cu).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace rnerf {

constexpr int MARCH_THREADS = 128;

struct MarchGeom {
  GridGeom g;
  float rdelta[3];   // reciprocal of ndelta, exhaustively verified for the 3-instruction exact division (else unused)
  int nby, nbz;      // brick grid (y, z extents)
};

// ---- exact division by a constant ------------------------------------------------------------------------------
// q = a*y; r = fma(-q, d, a); q' = fma(r, y, q) with y ~ 1/d equals the correctly rounded a/d for MOST (d, a) but not
// provably all, so a divisor is only used this way after the sequence has been compared with __fdiv_rn for every one
// of the 2^23 significands of a (verify_recip_kernel).  With no under/overflow in q, r, q' -- guaranteed by the
// exponent guards in fast_coords() and recip_for() -- the sequence commutes with scaling a by powers of two, so one
// binade of a covers all of them.
__device__ __forceinline__ float div_by_const(float a, float d, float y) {
  const float q = __fmul_rn(a, y);
  const float r = __fmaf_rn(-q, d, a);
  return __fmaf_rn(r, y, q);
}

// grid coordinates x = (p - nmin) / ndelta of VoxMLP._linear3 (rnerf/ior_utils.py:201-203)
template <bool FAST>
__device__ __forceinline__ void grid_coords(const MarchGeom& mg, float px, float py, float pz, float& x, float& y, float& z) {
  const float ax = sub(px, mg.g.nmin[0]), ay = sub(py, mg.g.nmin[1]), az = sub(pz, mg.g.nmin[2]);
  if (FAST) {
    x = div_by_const(ax, mg.g.ndelta[0], mg.rdelta[0]);
    y = div_by_const(ay, mg.g.ndelta[1], mg.rdelta[1]);
    z = div_by_const(az, mg.g.ndelta[2], mg.rdelta[2]);
    const float lo = fminf(fminf(fabsf(ax), fabsf(ay)), fabsf(az)), hi = fmaxf(fmaxf(fabsf(ax), fabsf(ay)), fabsf(az));
    if (lo >= 0x1p-60f && hi <= 0x1p60f) return;     // false for zeros, subnormals, huge values and NaN
  }
  x = divf(ax, mg.g.ndelta[0]); y = divf(ay, mg.g.ndelta[1]); z = divf(az, mg.g.ndelta[2]);
}

// ---- so3_mlp of the "all" stage: model_utils.MLP(net_width=128, net_depth=4, skip_layer=2, 3 outputs) -------------------
constexpr int SO3_IN = 60, SO3_W = 128;
constexpr int SO3_OFF_W1 = SO3_IN * SO3_W, SO3_OFF_W2 = SO3_OFF_W1 + SO3_W * SO3_W, SO3_OFF_W3 = SO3_OFF_W2 + SO3_W * SO3_W,
              SO3_OFF_W4 = SO3_OFF_W3 + (SO3_W + SO3_IN) * SO3_W, SO3_OFF_B = SO3_OFF_W4 + SO3_W * 3,
              SO3_FLOATS = SO3_OFF_B + 4 * SO3_W + 3;

struct So3Args {
  const float* w;        // kernels W0..W4 ([in][out] row-major) then biases b0..b4, fp32, SO3_FLOATS
  float window[10];      // cosine-easing window of annealed_pos_enc (rnerf/model_utils.py:236-245) per octave
  const float* window_dev;   // the same 10 values in device memory, or NULL: read at run time, so a captured CUDA graph follows
                             // a changing annealed_alpha (train.py:350-351) instead of freezing the capture-time window
  int dbg;                   // development aid (RNERF_SO3_TC_DEBUG; timing experiments only, results are wrong when set)
  float* saved;              // training forward (ragged tensor-pipe march): the four hidden activations of every evaluation are
                             // left here for the reverse sweep, [(ray * n_steps + step)][4][128] fp32, or NULL
};
constexpr int SO3_SAVED_FLOATS = 4 * SO3_W;    // per (ray, step) slot
template <typename A>
__device__ __forceinline__ float so3_window_at(const A& a, int k) {
  return a.window_dev != nullptr ? __ldg(a.window_dev + k) : a.window[k];
}


// ---- packed fp32 pairs (SASS: FFMA2) -------------------------------------------------------------------------------------
// A three-register FFMA issues every other cycle per scheduler on this part; the packed form does two IEEE fmas per
// issue, so the so3 GEMM loops keep their accumulators as column pairs.  Per-lane results equal fmaf's bit for bit.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// ---- TMA-fed weight ring -------------------------------------------------------------------------------------------
// The so3 weight matrices are streamed L2 -> shared memory in chunks of <= SO3_CH rows by the TMA engine
// (cp.async.bulk, one elected thread, completion on one mbarrier per slot); the FMA loops read them with conflict-free
// LDS.  The chunk sequence of an evaluation is fixed and periodic, so the ring simply keeps running: chunk number g
// (counted over the whole kernel) lives in slot g % n_slots and is chunk g % period of the stream.  While chunk g is
// consumed, chunks g+1 .. g+n_slots-1 are in flight -- across evaluation boundaries too, so the next evaluation finds
// its first chunks already resident.  Slot reuse is ordered by the block barrier every chunk starts with.
constexpr int SO3_CH = 16;                                   // rows per chunk (forward march; the reverse sweep uses 64)
constexpr int SO3_SLOT_FLOATS = SO3_CH * SO3_W;              // 8 KB slots
constexpr int SO3_MAX_SLOTS = 16;

struct So3Ring {
  float* slots;          // [n_slots][SO3_SLOT_FLOATS]
  uint32_t slots_s;      // the same, as a shared-window address
  uint32_t bars_s;       // n_slots mbarriers (8 bytes each)
  int n_slots;
  int slot_floats;       // capacity of one slot
  uint32_t pos;          // chunks consumed so far (uniform over the CTA)
  bool primed;           // chunks pos .. pos + n_slots - 2 have been issued
};

// thread 0 initialises the barriers; the caller must __syncthreads() before first use
__device__ __forceinline__ void ring_init(So3Ring& r, float* slots, void* bars, int n_slots, int tid,
                                          int slot_floats = SO3_SLOT_FLOATS) {
  r.slots = slots; r.slots_s = smem_u32(slots); r.bars_s = smem_u32(bars); r.n_slots = n_slots; r.slot_floats = slot_floats;
  r.pos = 0; r.primed = false;
  if (tid == 0) {
    for (int i = 0; i < n_slots; ++i) mbar_init(r.bars_s + 8 * i, 1);
    fence_barrier_init();
  }
}
// thread 0 only: start the copy of global chunk number g; `stream(g, src, bytes)` names its source
template <class Stream>
__device__ __forceinline__ void ring_issue(const So3Ring& r, uint32_t g, const Stream& stream) {
  const float* src; uint32_t bytes;
  stream(g, src, bytes);
  const uint32_t slot = g % (uint32_t)r.n_slots, bar = r.bars_s + 8 * slot;
  mbar_arrive_expect_tx(bar, bytes);
  tma_bulk_g2s(r.slots_s + slot * (uint32_t)(r.slot_floats * 4), src, bytes, bar);
}
// every thread, at the start of an evaluation
template <class Stream>
__device__ __forceinline__ void ring_prime(So3Ring& r, int tid, const Stream& stream) {
  if (!r.primed) {
    if (tid == 0)
      for (int i = 0; i < r.n_slots - 1; ++i) ring_issue(r, r.pos + i, stream);
    r.primed = true;
  }
}
// every thread, once per chunk: waits for chunk r.pos, frees the slot of chunk r.pos - 1 (block barrier), refills it, and
// returns the chunk's data.  The caller consumes the chunk and then does ++r.pos.
template <class Stream>
__device__ __forceinline__ const float* ring_acquire(So3Ring& r, int tid, const Stream& stream) {
  const uint32_t g = r.pos, slot = g % (uint32_t)r.n_slots;
  mbar_wait(r.bars_s + 8 * slot, (g / (uint32_t)r.n_slots) & 1u);
  __syncthreads();                 // everyone is done with chunk g-1 (and sees the activations written before this point)
  if (tid == 0) ring_issue(r, g + r.n_slots - 1, stream);
  return r.slots + slot * r.slot_floats;
}
// every thread, before the kernel exits: the copies issued ahead must have landed
__device__ __forceinline__ void ring_drain(So3Ring& r) {
  if (r.primed)
    for (int i = 0; i < r.n_slots - 1; ++i) {
      const uint32_t g = r.pos + i;
      mbar_wait(r.bars_s + 8 * (g % (uint32_t)r.n_slots), (g / (uint32_t)r.n_slots) & 1u);
    }
}

// verified reciprocal of a grid pitch for div_by_const, or 0 (march.cu)
float recip_for(float d, cudaStream_t st);
bool fast_div_enabled();     // RNERF_MARCH_DIV=ieee switches the verified sequence off (march.cu)
int so3_rays_per_cta(int64_t n_rays, int n_sm);
// geometry + verified reciprocals; returns true when the 3-instruction division may be used (march.cu)
bool make_march_geom(const int ndim[3], const double nmin[3], const double nmax[3], cudaStream_t st, MarchGeom& mg);

}  // namespace rnerf
