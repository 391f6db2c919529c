/*
 * rnerf_b200.h -- C ABI of the B200-native refractive rendering hot path.
 *
 * The reference (alexkeroro86/SampleNeRFRO) has no FFI layer: the path sits behind Python call
 * signatures and is lowered by XLA.  Each entry point below names the reference function
 * (file:line under /root/reference) whose arithmetic it replaces; INTEGRATION.md shows the ctypes
 * binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host; arrays
 *     are row-major, contiguous, fp32 unless stated; indices are int32.
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on that stream, do not
 *     allocate, do not synchronise.
 *   - return value: 0 = ok, <0 = invalid argument (RNERF_E_*), >0 = cudaError_t.  rnerf_last_error()
 *     returns a thread-local message for the last non-zero return.
 *   - python floats of the reference (nmin, nmax, near, far) cross the ABI as double and are rounded
 *     to fp32 at the same point the reference rounds them.
 *
 * Path record ("bent sample"): rec_floats = 12 ("full") or 8 ("compact") floats per (ray, march step),
 * [B][S][rec_floats]:
 *     0..2 ray_pos   3 ray_dist   4..6 direction state v (UN-normalised)   7 idx_data (n)
 *     8..10 idx_grad (grad n)     11 unused (0)                            <- full records only
 *   i.e. the arrays PathSampler.__call__ returns (rnerf/eikonal_utils.py:118-124), interleaved; ray_dir =
 *   safe_l2_normalize(v) (rnerf/eikonal_utils.py:113) is applied by the readers (rnerf_select, rnerf_resample,
 *   rnerf_path_dirs) to the records they use, with the same arithmetic, instead of at every march step.
 *   Compact records drop idx_grad, which on the render/train path is only read by the online-sparsity term
 *   (rnerf/models.py:351-357; off in every shipped config) and by the debug outputs.
 *   t_col [B][S] is an optional dense copy of ray_dist that lets rnerf_resample search in shared memory.
 */
#ifndef RNERF_B200_H_
#define RNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNERF_ABI_VERSION 11
#define RNERF_PATH_STRIDE 12         /* full records */
#define RNERF_PATH_STRIDE_COMPACT 8

#define RNERF_E_NULL   (-1)  /* null pointer */
#define RNERF_E_SHAPE  (-2)  /* bad size / unsupported shape */
#define RNERF_E_ALIGN  (-3)  /* pointer not 16-byte aligned */
#define RNERF_E_ARCH   (-4)  /* device is not sm_100 */

int rnerf_abi_version(void);
const char* rnerf_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t rnerf_launch_count(void);

/* ---- a1: rnerf/ior_utils.py:327-363 conv3d_normal -- normalised ws^3 Gaussian, edge padding ---- */
int rnerf_grid_blur(const float* n_in, float* n_out, const int ndim_host[3], int ws, double sigma, void* stream);

/* ---- a2: rnerf/ior_utils.py:161,165-172 VoxMLP.setup/_compute_grad -- table[G^3][4] = (n, dn/dx, dn/dy, dn/dz) */
int rnerf_grid_table(const float* n, const int ndim_host[3], const double nmin_host[3], const double nmax_host[3],
                     float* table, void* stream);

/* ---- (no reference counterpart) brick map for the march: bricks[nbx*nby*nbz] (8^3-voxel bricks, x slowest) holds
 * the common n of a homogeneous brick (all corners bit-equal, grad n == 0) or NaN.  Purely an access-skipping aid:
 * lookups give bit-identical results with or without it. */
int64_t rnerf_grid_brick_count(const int ndim_host[3]);
int rnerf_grid_bricks(const float* table, const int ndim_host[3], float* bricks, void* stream);

/* ---- a3: rnerf/ior_utils.py:188-223 VoxMLP._linear3 -- out[N][4] */
int rnerf_grid_lookup(const float* table, const int ndim_host[3], const double nmin_host[3],
                      const double nmax_host[3], const float* pts, int64_t n_pts, float* out, void* stream);

/* ---- a5/a6: rnerf/eikonal_utils.py:30-49,101-124 OneEikonalStep + PathSampler (radiance stage) ----
 * step_size = (far - near) / (S - 1) (rnerf/models.py:121-122).  path: [B][S][rec_floats] records; t_col: [B][S]
 * or NULL.  The first call for a new voxel pitch (ndelta) runs a one-off exhaustive check of the constant-divisor
 * division used for the grid coordinates and synchronises `stream` once; later calls are asynchronous. */
int rnerf_march_fwd(const float* table, const float* bricks /* from rnerf_grid_bricks, or NULL */,
                    const int ndim_host[3], const double nmin_host[3], const double nmax_host[3],
                    const float* origins, const float* viewdirs, int64_t n_rays, double near, double far, int n_steps,
                    int rec_floats, float* path, float* t_col, void* stream);

/* ---- a4 + a5/a6, "all" stage: rnerf/ior_utils.py:269-312 VoxMLP.__call__ (so3_mlp on annealed_pos_enc(p, 0, 10,
 * alpha*10), Rodrigues rotation of grad n) inside rnerf/eikonal_utils.py:30-49 (grad = where(|grad n| > 1e-3, pred,
 * grad n)).  so3_w: the 5 Dense kernels of so3_mlp ([in,out] row-major: 60x128, 128x128, 128x128, 188x128, 128x3) then
 * the 5 biases, fp32, rnerf_so3_weight_floats() values.  so3_window_host[10]: cosine-easing window per octave
 * (rnerf/model_utils.py:222-245 at alpha * 10).  so3_window_dev: the same 10 values as fp32 in DEVICE memory, or NULL;
 * when given it wins and is read at run time, so a captured CUDA graph of the training step follows the annealing
 * schedule (train.py:350-351) instead of replaying its capture-time window (so3_window_host may then be NULL).  The same
 * pair of arguments appears in rnerf_so3_predict and rnerf_march_all_bwd.
 * so3_tc_packed: rnerf_so3_tc_pack(so3_w) or NULL.  When given, full frames with compact records and all small launches
 * (training batches) evaluate so3_mlp on the tensor pipe (tcgen05.mma kind::f16 on fp16 hi/lo split operands: fp32-grade
 * products, positions within the same 1e-4 of the reference); otherwise the fp32 CUDA-core chain runs.
 * so3_saved: NULL, or (training) a buffer of rnerf_so3_saved_floats(n_rays, n_steps) floats in which the forward leaves the four
 * hidden activations of every so3_mlp evaluation, [(ray * n_steps + step)][4][128]; only the slots of evaluated (ray, step)
 * pairs are written.  rnerf_march_all_bwd given the same buffer reads them back instead of recomputing the forward.
 * rnerf_so3_saved_floats returns 0 when the launch would not run the kernel that writes it (pass NULL then).
 * Outputs as rnerf_march_fwd (idx_grad is the un-rotated grad n). */
size_t rnerf_so3_weight_floats(void);
size_t rnerf_so3_saved_floats(int64_t n_rays, int n_steps);
int rnerf_march_all_fwd(const float* table, const float* bricks, const int ndim_host[3], const double nmin_host[3],
                        const double nmax_host[3], const float* origins, const float* viewdirs, int64_t n_rays,
                        double near, double far, int n_steps, int rec_floats, const float* so3_w,
                        const double so3_window_host[10], const float* so3_window_dev, const void* so3_tc_packed,
                        float* so3_saved, float* path, float* t_col, void* stream);

/* ray_dir[B][S][3] = safe_l2_normalize(v) of every record: the `ray_dir` array of PathSampler.__call__
 * (rnerf/eikonal_utils.py:113), for callers that want the whole bent path (extract_mesh.py:178). */
int rnerf_path_dirs(const float* path, int rec_floats, int64_t n_rays, int n_steps, float* ray_dir, void* stream);

/* ---- a7: rnerf/models.py:240-247 coarse selection; jitter[Nc] int32 march-step indices ---- */
int rnerf_select(const float* path, int rec_floats, int64_t n_rays, int n_steps, const int32_t* jitter, int n_coarse,
                 float* pos_c, float* dir_c, float* t_c, float* grad_c, void* stream);

/* ---- a8+a9: rnerf/model_utils.py:187-214 pos_enc + :30-90 NerfMLP, fused, bf16 tcgen05 ----
 * rnerf_encmlp_pack converts the 12 Flax Dense layers ([in,out] kernels, [out] biases; creation order
 * Dense_0..11 of model_utils.py:65-89) into the device image the kernel streams with TMA.
 * kernels_host/biases_host: host arrays of 12 device pointers.  packed: rnerf_encmlp_packed_bytes(). */
size_t rnerf_encmlp_packed_bytes(void);
int rnerf_encmlp_pack(const float* const* kernels_host, const float* const* biases_host, void* packed, void* stream);
/* raw_out[M][4] = (raw_rgb[3], raw_sigma).  pos/dir: [M][3].  M arbitrary (tail rows are masked). */
int rnerf_encmlp_fwd(const void* packed, const float* pos, const float* dir, int64_t n_samples, float* raw_out,
                     void* stream);
/* debug/parity variant: also dumps every layer's post-activation output as bf16, layer_out[L][M][256]
 * (L = 10: Dense_0..7, bottleneck, cond layer (first 128 cols)). */
int rnerf_encmlp_fwd_debug(const void* packed, const float* pos, const float* dir, int64_t n_samples,
                           float* raw_out, uint16_t* layer_out, void* stream);

/* ---- a17 (train.py:164-165 differentiates the MLPs): training forward + backward of pos_enc + NerfMLP ----
 * rnerf_encmlp_fwd_train additionally saves layer_out[10][M][256] (bf16 post-activation outputs of Dense_0..7,
 * Dense_9, Dense_10), enc_out[2][M][64] (bf16 pos_enc / dir_enc rows, zero padded) and relu_masks[10][M][8] (uint32: one bit
 * per activation, set where it is positive -- all the dgrad chain needs of the activations; 32 B instead of 512 B per
 * sample and layer; bit layout: csrc/umma.cuh relu_mask32: word g = columns 32g..32g+31, even column 2i at bit i, odd column 2i+1 at bit 16+i).
 * rnerf_mlp_dgrad: fused tcgen05 chain producing dz_out[10][M][256] (bf16 gradient wrt every layer's pre-activation)
 *   from d_raw[M][4] and relu_masks; dgrad_packed comes from rnerf_mlp_dgrad_pack (transposed weight image, rebuilt when
 *   weights change).
 * rnerf_mlp_wgrad: gw[kx_valid][n] += x[:, :x_cols]^T dz (fp32, accumulating), gb[n] += column sums of dz (or NULL).
 * rnerf_mlp_head_grad: out_rgb_head[387] += (gW11[128][3], gb11[3]), out_sigma_head[257] += (gW8[256], gb8) from d_raw
 *   and the saved activations (each pair laid out like the Flax (kernel, bias) of Dense_11 / Dense_8). */
int rnerf_encmlp_fwd_train(const void* packed, const float* pos, const float* dir, int64_t n_samples, float* raw_out,
                           uint16_t* layer_out, uint16_t* enc_out, uint32_t* relu_masks, void* stream);
size_t rnerf_mlp_dgrad_packed_bytes(void);
int rnerf_mlp_dgrad_pack(const float* const* kernels_host, void* dgrad_packed, void* stream);
int rnerf_mlp_dgrad(const void* dgrad_packed, const void* fwd_packed, const uint32_t* relu_masks, const float* d_raw,
                    int64_t n_samples, uint16_t* dz_out, void* stream);
int rnerf_mlp_wgrad(const uint16_t* x, int ldx, int x_cols, int kx_valid, const uint16_t* dz, int n, int64_t n_samples,
                    float* gw, float* gb, void* stream);
/* rnerf_mlp_wgrad_batched: up to 16 rnerf_mlp_wgrad jobs over the same n_samples rows in ONE launch (per-job arguments as
 * arrays of length n_jobs; gb[j] may be NULL): all the layers of one MLP's backward.  The SMs are divided between the jobs in
 * proportion to the bytes each streams. */
int rnerf_mlp_wgrad_batched(int n_jobs, const uint16_t* const* x, const int* ldx, const int* x_cols, const int* kx_valid,
                            const uint16_t* const* dz, const int* n, int64_t n_samples, float* const* gw, float* const* gb,
                            void* stream);
int rnerf_mlp_head_grad(const uint16_t* layer_out, const float* d_raw, int64_t n_samples, float* out_rgb_head,
                        float* out_sigma_head, void* stream);

/* development aid: same as rnerf_encmlp_fwd, plus clock64 stamps of CTA 0: prof[2 roles][10 layers][4] int64
 * (role 0 = MMA issuer: wait-start, A-ready, issued; role 1 = epilogue: wait-start, acc-ready, done). */
int rnerf_encmlp_fwd_profile(const void* packed, const float* pos, const float* dir, int64_t n_samples,
                             float* raw_out, long long* prof, void* stream);

/* ---- a10: rnerf/model_utils.py:93-140 MLP as bkgd_mlp (27->128->128->128(+27)->128->3), fp32 ----
 * w: the 5 Dense kernels then the 5 biases, concatenated fp32 ([in,out] row-major each).
 * dirs: [B][3] unit directions (encoded in-kernel, pos_enc deg 0..4).  raw_out: [B][3]. */
size_t rnerf_bkgd_weight_floats(void);
int rnerf_bkgd_mlp_fwd(const float* w, const float* dirs, int64_t n_rays, int64_t dir_stride_floats,
                       float* raw_out, void* stream);

/* backward of the above wrt the 5 Dense layers: gw (same flat layout as w, fp32) is ACCUMULATED into; d_raw: [B][3]. */
int rnerf_bkgd_mlp_bwd(const float* w, const float* dirs, int64_t n_rays, int64_t dir_stride_floats,
                       const float* d_raw, float* gw, void* stream);

/* ---- a11+a12: rnerf/models.py:334-338 activations + rnerf/model_utils.py:247-309 volumetric_rendering ----
 * raw: [B][Ns][4]; t: [B][Ns]; dirs: [B][Ns][3]; bkgd_raw: [B][3] or NULL (rgb_bkgd=None);
 * mask: [B][Ns] fp32 or NULL.  Outputs (any may be NULL except comp_rgb): comp_rgb[B][3], distance[B],
 * acc[B], weights[B][Ns], alpha[B][Ns], trans[B], trans_rgb_bkgd[B][3]. */
int rnerf_composite_fwd(const float* raw, const float* t, const float* dirs, const float* bkgd_raw,
                        const float* mask, int64_t n_rays, int n_samples, int white_bkgd, double rgb_padding,
                        double sigma_bias, float* comp_rgb, float* distance, float* acc, float* weights,
                        float* alpha, float* trans, float* trans_rgb_bkgd, void* stream);
/* backward of the above wrt raw and bkgd_raw, from d(comp_rgb)[B][3], d(trans)[B], d(trans_rgb_bkgd)[B][3]
 * (any may be NULL = zero).  d_raw: [B][Ns][4]; d_bkgd_raw: [B][3] or NULL. */
int rnerf_composite_bwd(const float* raw, const float* t, const float* dirs, const float* bkgd_raw,
                        const float* mask, int64_t n_rays, int n_samples, int white_bkgd, double rgb_padding,
                        double sigma_bias, const float* d_comp_rgb, const float* d_trans,
                        const float* d_trans_rgb_bkgd, float* d_raw, float* d_bkgd_raw, void* stream);

/* ---- a13+a14: rnerf/model_utils.py:312-374 sorted_piecewise_constant_pdf + :377-435 sample_pdf ----
 * t_c/weights_c: [B][Nc] coarse distances / compositing weights (the kernel forms the mid-point bins and
 * uses weights[1:-1]); u: [Nf] (u_per_ray=0) or [B][Nf] sorted CDF positions in [0,1).
 * path/rec_floats/t_col as written by rnerf_march_fwd (t_col may be NULL: the search then probes the records).
 * Outputs: t_f[B][Nc+Nf], pos_f/dir_f/grad_f [B][Nc+Nf][3] (grad_f needs full records; may be NULL). */
int rnerf_resample(const float* path, int rec_floats, const float* t_col, int64_t n_rays, int n_steps, const float* t_c,
                   const float* weights_c, int n_coarse, const float* u, int u_per_ray, int n_fine, float* t_f,
                   float* pos_f, float* dir_f, float* grad_f, void* stream);

/* ---- a17: train.py:166-183 gradient clipping + optax.adam (train.py:312-317) on a flat fp32 parameter arena ----
 * hyper (device, 10 floats): lr, b1, b2, eps, 1-b1^t, 1-b2^t, gradient scale, weight-decay coefficient
 * (2 weight_decay_mult / numel: the closed-form gradient of train.py:146-150's weight_l2 term), grad_max_val,
 * grad_max_norm (0 = off).  Per-step scalars are read from device memory so a captured CUDA graph of the step can be
 * replayed.  The effective gradient is clamp(gscale * grad + wd * theta, +-grad_max_val) * min(1, grad_max_norm /
 * (1e-7 + sqrt(*norm_sq))); rnerf_grad_sumsq accumulates that squared norm (before the norm factor) into out_accum.
 * rnerf_sumsq: out_accum[0] += sum x^2 (the weight_l2 statistic). */
int rnerf_sumsq(const float* x, int64_t n, float* out_accum, void* stream);
int rnerf_grad_sumsq(const float* grad, const float* theta /* or NULL */, int64_t n, const float* hyper, float* out_accum,
                     void* stream);
int rnerf_adam_step(float* theta, const float* grad, float* mu, float* nu, int64_t n, const float* hyper,
                    const float* norm_sq /* or NULL */, void* stream);

/* ---- the radiance-stage loss of /root/reference/train.py:75-162 and its gradient, two launches instead of ~70 tensor ops:
 *   total = mean((rgb - px)^2) + mean((rgb_c - px)^2)
 *         + bg_weight * gate * sum(mask |trans_rgb_bkgd - px|) / (sum(mask) + 1)         mask = trans > 0.5   (train.py:89-95)
 *         + bg_smooth_weight * gate * mean(0.5 dv^2 + 0.5 dh^2)   over the differences of the env map [patch][patch][env_channels]
 *           along its first two axes (:126-129; 3 channels for a whole patch -- a device's [P/N][P][3] shard is reshaped to
 *           (P/N, P/N, 3N) by the reference, quirk kept)
 * gate = (annealed_alpha > 0).  trans_rgb_bkgd / trans may be NULL (no background term), env may be NULL (no smoothness
 * term).  ws: rnerf_radiance_loss_ws_floats() floats, ZERO before the first call (the kernel leaves it zeroed).
 * out[8]: total, loss, loss_c, loss_bg (gated, unweighted), loss_bg_smooth (gated), psnr, psnr_c, sum(mask).
 * _bwd: gradients of `total` times the device scalar g_total[0]; d_trans_rgb_bkgd / d_env NULL exactly when their input is. */
size_t rnerf_radiance_loss_ws_floats(void);
int rnerf_radiance_loss_fwd(const float* rgb, const float* rgb_c, const float* trans_rgb_bkgd, const float* trans,
                            const float* pixels, int64_t n_rays, const float* env, int patch, int env_channels, double bg_weight,
                            double bg_smooth_weight, double gate, float* ws, float* out, void* stream);
int rnerf_radiance_loss_bwd(const float* rgb, const float* rgb_c, const float* trans_rgb_bkgd, const float* trans,
                            const float* pixels, int64_t n_rays, const float* env, int patch, int env_channels, double bg_weight,
                            double bg_smooth_weight, double gate, const float* out, const float* g_total, float* d_rgb,
                            float* d_rgb_c, float* d_trans_rgb_bkgd, float* d_env, void* stream);

/* ---- rnerf/datasets.py:216-242 (Blender) / :486-518 (OpenCV) _generate_rays for image rows
 * [row0, row0 + n_rows) of one camera.  camtoworld_host: the 3x4 pose, row-major (rounded to fp32 like the reference's
 * arrays).  Blender: focal = fx (fy, cx, cy unused).  Outputs [n_rows*W][3] (radii [n_rows*W]); any may be NULL. */
int rnerf_generate_rays(const double camtoworld_host[12], int height, int width, int opencv, double fx, double fy,
                        double cx, double cy, int use_pixel_centers, int row0, int n_rows, float* origins,
                        float* directions, float* viewdirs, float* radii, void* stream);
/* out_accum[0] += sum (a - b)^2: the image mse behind compute_psnr (rnerf/utils.py:392-401) */
int rnerf_sq_err(const float* a, const float* b, int64_t n, float* out_accum, void* stream);

/* ---- training of the "all" stage -- what train.py:164 (jax.value_and_grad) differentiates when
 * path_sampler is trainable (train.py:302-310).  The loss reaches the scan only through the coarse samples
 * ray_pos[:, jitter] / ray_dir[:, jitter] (rnerf/models.py:243-244; ray_dist and the fine samples are stop_gradient,
 * rnerf/eikonal_utils.py:120, rnerf/model_utils.py:406-411).
 * rnerf_mlp_input_grad: gradient of pos_enc + NerfMLP wrt its inputs, d_pos[M][3] / d_dirs[M][3], from the dz_out of
 *   rnerf_mlp_dgrad; wt from rnerf_mlp_input_grad_pack(Dense_0, Dense_5, Dense_10 kernels), rebuilt when they change.
 * rnerf_bkgd_mlp_bwd_dirs: rnerf_bkgd_mlp_bwd that also writes d_dirs[B][3], the gradient wrt the encoded direction
 *   (ray_dir_c[:, -1], rnerf/models.py:303).
 * rnerf_march_all_bwd: reverse sweep of the scan (rnerf/eikonal_utils.py:30-49,75-82).  path/rec_floats: the records
 *   written by rnerf_march_all_fwd for the same inputs; jitter[Nc]: strictly increasing march-step indices;
 *   d_pos_c/d_dir_c: [B][Nc][3] loss gradients; so3_wt: rnerf_so3_transpose(so3_w).  g_so3 (layout of so3_w) is
 *   ACCUMULATED into; d_origins/d_viewdirs [B][3] (gradients wrt the ray, not used by train.py) may be NULL.
 *   Extension without a reference counterpart (the reference keeps the grid constant; BASELINE.json's north_star asks for
 *   "learned IoR-grid" gradients): d_table [G^3][4] or NULL is ACCUMULATED with the gradient wrt the (n, grad n) table, and
 *   rnerf_grid_table_bwd takes it on to the n-grid (adjoint of rnerf_grid_table).  so3_w may be NULL (radiance stage: the
 *   sweep then only yields the ray / table gradients; so3_wt, so3_window_host, g_so3 are ignored). */
/* VoxMLP.wrapper_grad_mlp (rnerf/ior_utils.py:225-267) on free-standing points: pred[N][3] = rodrigues(so3_mlp(
 * annealed_pos_enc(pts)), cond) -- the evaluation behind PathSampler.compute_normal_loss_and_smooth
 * (rnerf/eikonal_utils.py:84-98). */
int rnerf_so3_predict(const float* so3_w, const double so3_window_host[10], const float* so3_window_dev, const float* pts,
                      const float* cond, int64_t n, float* pred, void* stream);
/* so3_mlp on the tensor pipe (tcgen05.mma kind::tf32, 3xTF32 split for fp32-grade products): rnerf_so3_tc_pack turns the fp32
 * image so3_w into the pre-swizzled hi/lo weight chunks the evaluator streams (rnerf_so3_tc_packed_bytes() bytes; repack
 * whenever so3_w changes); rnerf_so3_predict_tc = rnerf_so3_predict through that evaluator (same arguments + the packed image). */
size_t rnerf_so3_tc_packed_bytes(void);
int rnerf_so3_tc_pack(const float* so3_w, void* so3_tc_packed, void* stream);
int rnerf_so3_predict_tc(const void* so3_tc_packed, const float* so3_w, const double so3_window_host[10],
                         const float* so3_window_dev, const float* pts, const float* cond, int64_t n, float* pred, void* stream);
/* a10 on the tensor pipe (render path): bkgd_mlp(pos_enc(dir, 0, 4)) (rnerf/models.py:303, rnerf/model_utils.py:93-140) with the
 * so3 evaluator -- the two networks have the same shape (27 / 60 inputs -> 128 x 4 with the inputs re-joined before Dense_3 -> 3).
 * bkgd_so3: the background weights zero-padded into so3_mlp's layout (rnerf_so3_weight_floats() floats); bkgd_tc_packed =
 * rnerf_so3_tc_pack(bkgd_so3).  fp16 hi/lo split operands, fp32 accumulation: 1e-6 of rnerf_bkgd_mlp_fwd. */
int rnerf_bkgd_mlp_fwd_tc(const void* bkgd_tc_packed, const float* bkgd_so3, const float* dirs, int64_t n_rays,
                          int64_t dir_stride_floats, float* raw_out, void* stream);
size_t rnerf_mlp_input_grad_packed_floats(void);
int rnerf_mlp_input_grad_pack(const float* dense0_kernel, const float* dense5_kernel, const float* dense10_kernel,
                              float* wt, void* stream);
int rnerf_mlp_input_grad(const uint16_t* dz, int64_t n_samples, const float* wt, const float* pos, const float* dirs,
                         float* d_pos, float* d_dirs, void* stream);
int rnerf_bkgd_mlp_bwd_dirs(const float* w, const float* dirs, int64_t n_rays, int64_t dir_stride_floats,
                            const float* d_raw, float* gw, float* d_dirs, void* stream);
size_t rnerf_so3_transposed_floats(void);
int rnerf_so3_transpose(const float* so3_w, float* so3_wt, void* stream);
int rnerf_march_all_bwd(const float* table, const float* bricks, const int ndim_host[3], const double nmin_host[3],
                        const double nmax_host[3], const float* path, int rec_floats, int64_t n_rays, double near,
                        double far, int n_steps, const int32_t* jitter, int n_coarse, const float* d_pos_c,
                        const float* d_dir_c, const float* so3_w, const float* so3_wt, const double so3_window_host[10],
                        const float* so3_window_dev, const float* so3_saved /* or NULL */, float* g_so3, float* d_origins,
                        float* d_viewdirs, float* d_table, void* stream);
int rnerf_grid_table_bwd(const float* d_table, const int ndim_host[3], const double nmin_host[3], const double nmax_host[3],
                         float* d_n, void* stream);

/* ---- a15: rnerf/models.py:498-503 bd_cut_dist mask: reverse-cumsum(inside bbox) > 0 ---- */
int rnerf_bbox_tail_mask(const float* pos, int64_t n_rays, int n_samples, const double lo_host[3],
                         const double hi_host[3], float* mask, float* inv_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RNERF_B200_H_ */
