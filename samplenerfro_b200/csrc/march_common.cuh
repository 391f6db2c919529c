// Definitions shared by the forward march (march.cu) and its adjoint (march_bwd.cu).
#pragma once
#include "common.cuh"

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
};


// verified reciprocal of a grid pitch for div_by_const, or 0 (march.cu)
float recip_for(float d, cudaStream_t st);
// geometry + verified reciprocals; returns true when the 3-instruction division may be used (march.cu)
bool make_march_geom(const int ndim[3], const double nmin[3], const double nmax[3], cudaStream_t st, MarchGeom& mg);

}  // namespace rnerf
