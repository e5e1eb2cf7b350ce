/*
 * turbdiff_b200.h - C ABI of libturbdiff_b200.so, the sm_100a kernel library behind the
 * TurbDiff denoising hot path (3-D U-Net denoiser forward/backward + DDPM ancestral
 * sampling).  Reference = martenlienen/generative-turbulence; citations are
 * `turbdiff/...py:line` in that repository.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer borrowed for the
 *     duration of the stream-ordered call; nothing is allocated or synchronised inside
 *     (all entry points are CUDA-graph capturable);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = ok, >0 = cudaError_t of the failed launch, <0 = TDB_E_* argument
 *     error.  tdb_last_error() returns a static message for the calling thread;
 *   - `dtype`: TDB_F32 (0) or TDB_BF16 (1) = storage type of *activation* buffers.
 *     Parameters (weights, biases, norm scales, FiLM) are always fp32 unless noted;
 *   - activation layout ("halo grid"): channels-last with a materialised one-voxel
 *     replicate halo, [B][X+2][Y+2][Z+2][ld] elements, channel c of voxel (b,x,y,z) at
 *       (((b*(X+2) + x+1)*(Y+2) + y+1)*(Z+2) + z+1)*ld + c
 *     `ld` (>= C) is the channel pitch in elements, so a tensor can be a channel slice of
 *     a wider buffer (skip concatenation without a copy).  X,Y,Z are the *unhaloed* dims of
 *     that U-Net level (194x50x50 at level 0 of the shapes config);
 *   - module-boundary tensors (x_t, eps, noise ...) are the reference's: NCDHW fp32
 *     contiguous (B,F,X,Y,Z).
 */
#ifndef TURBDIFF_B200_H
#define TURBDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TDB_API __attribute__((visibility("default")))
#else
#define TDB_API
#endif

#define TDB_F32 0
#define TDB_BF16 1

#define TDB_E_BADARG (-1)      /* inconsistent sizes / null pointer */
#define TDB_E_UNSUPPORTED (-2) /* shape not supported by this kernel (e.g. channel multiple) */
#define TDB_E_NODEVICE (-3)    /* no sm_100 device / driver entry point missing */

/* pointwise flags (tdb_pointwise) */
#define TDB_PW_SILU 1u   /* apply x*sigmoid(x) after the affine part */
#define TDB_PW_NOHALO 2u /* write interior voxels only */

/* convolution flags (tdb_conv3d_bf16, tdb_conv3d_bf16_fold) */
#define TDB_CONV_ALL_ROWS 1u /* also store the halo rows of the output (input-gradient convolutions) */
#define TDB_CONV_CLUSTER_MC 2u /* tdb_conv3d_bf16: share weight tiles across a 2-CTA cluster by TMA multicast (opt-in) */

/* ddpm step flags (tdb_ddpm_step) */
#define TDB_STEP_NOISE_BCS 1u /* GaussianDiffusion(noise_bcs=True)  */
#define TDB_STEP_CLIP 2u      /* clip_denoised: clamp x0 to [-1,1]  */
#define TDB_STEP_FINAL 4u     /* after the update also pin non-inside voxels to x_bcs (ddpm.py:814) */
#define TDB_STEP_LEARNED_VAR 8u /* learned_variances (ddpm.py:732-741): eps is the (B,2F,nvox) model output, per-voxel variance */

/* ---- fused optimiser step (training path; reference: torch.optim.RAdam, turbdiff/models/diffusion.py:216, with
 * Lightning's gradient_clip_val = 0.1 / norm, config/shapes_experiment.yaml:50-51) ---------------------------------
 * All tensors fp32.  *_ptrs are device arrays of device addresses (one per tensor), numel the element counts;
 * (chunk_tensor[i], chunk_off[i]) name the tensor and element offset of work chunk i (`chunk` elements, one block). */

/* out (pre-zeroed double) += sum over all gradient elements of g*g. */
TDB_API int tdb_grad_sqnorm(const int64_t* grad_ptrs, const int64_t* numel, const int* chunk_tensor, const int64_t* chunk_off,
                    int n_chunks, int chunk, double* out, void* stream);

/* One RAdam step on every tensor.  sqnorm != NULL: gradients are scaled by min(1, max_norm / (sqrt(*sqnorm) + 1e-6))
 * on the fly (clip_grad_norm_).  step_size and `rectified` are the step-dependent scalars of torch's RAdam:
 * rectified ? lr*rect*sqrt(1-beta2^t)/(1-beta1^t) : lr/(1-beta1^t). */
TDB_API int tdb_radam_step(const int64_t* param_ptrs, const int64_t* grad_ptrs, const int64_t* exp_avg_ptrs,
                   const int64_t* exp_avg_sq_ptrs, const int64_t* numel, const int* chunk_tensor, const int64_t* chunk_off,
                   int n_chunks, int chunk, const double* sqnorm, float max_norm, float step_size, double beta1, double beta2,
                   float eps, float weight_decay, int rectified, void* stream);

/* ---- TKE-spectrum statistic ---------------------------------------------------------------- */

/* E[b][j] = 4 pi k_j^2 * sum_p w_p exp(trilinear(log |fftshift(fftn(tke_b))|^2, k_j * points_p + centre)) with
 * tke_b = |u_b - u_mean|^2 / 2: TurbulentKineticEnergySpectrum.forward (turbdiff/models/metrics.py:296-320) with the
 * log-domain interp3 (:211-267).  u: (B,3,n0,n1,n2) fp32; u_mean: (3,n0,n1,n2) or NULL (u is already the perturbation);
 * k: (K) radii; points (P,3) / weights (P): sphere quadrature (the reference's Lebedev nodes); work: 4*B*n0*n1*n2
 * floats of scratch; E: (B,K).  Axes <= 64 (the reference evaluates 48^3 cubes). */
TDB_API int tdb_tke_spectrum(const float* u, const float* u_mean, int B, int n0, int n1, int n2, const float* k, int K,
                     const float* points, const float* weights, int P, float* work, float* E, void* stream);

TDB_API const char* tdb_last_error(void);
TDB_API int tdb_version(void);
/* number of kernel launches issued through this library by the calling process */
TDB_API int64_t tdb_launch_count(void);

/* ---- input / output stages -------------------------------------------------------- */

/* encode_x / encode_c_local 1x1x1 convs + channel concat, written into a halo grid
 * (ddpm.py:433,436,495-501).  x: (B,F,X,Y,Z) fp32; c_local: (Fc,X,Y,Z) fp32, unbatched, may
 * be NULL when Fc == 0.  wx: (dim,F), wc: (dim,Fc) fp32.  out channels: [0,dim) = encode_x,
 * [dim,2dim) = encode_c_local.  parts: bit0 write the x half, bit1 write the c half. */
TDB_API int tdb_encode_input(const float* x, const float* c_local, const float* wx, const float* bx,
                     const float* wc, const float* bc, void* out, int ld_out, int B, int F, int Fc,
                     int dim, int X, int Y, int Z, int parts, int dtype, void* stream);

/* decode[1]: 1x1x1 conv dim -> F on a halo grid, written as NCDHW fp32 (ddpm.py:459,505). */
TDB_API int tdb_decode_output(const void* act, int ld, const float* w, const float* b, float* out, int B,
                      int X, int Y, int Z, int dim, int F, int dtype, void* stream);

/* ---- convolution --------------------------------------------------------------------- */

/* 3x3x3 replicate-padded conv (ntaps=27) or 1x1x1 conv (ntaps=1) over a halo grid
 * (ddpm.py:164,188; nn.Conv3d(padding_mode="replicate") == valid conv over the halo).
 * fp32 CUDA-core path: exact fp32 FMA accumulation (the 1e-5 parity path).
 * w: packed [ntaps][Cin][Cout] fp32 (tap = (kx*3+ky)*3+kz).  bias may be NULL.
 * out is a halo grid whose halo rows hold unspecified values. Cin % 8 == 0, Cout % 4 == 0. */
TDB_API int tdb_conv3d_f32(const float* in, int ld_in, const float* w, const float* bias, float* out,
                   int ld_out, int B, int X, int Y, int Z, int Cin, int Cout, int ntaps,
                   void* stream);

/* bf16 tensor-core path: implicit GEMM on tcgen05 with TMEM accumulators and TMA-staged
 * operand tiles; fp32 accumulation.  w: packed [Cout][ntaps*Cin] bf16 (k = tap*Cin + ci).
 * Cin % 16 == 0, Cout % 16 == 0, ld_in % 8 == 0, ld_out % 8 == 0.
 * gn_stats (nullable): double [B][G][2] (sum, sum of squares) accumulated from the fp32
 * accumulators over interior voxels, G groups of Cout/G channels.
 * splitk_scratch (nullable): fp32 [rows][Cout] workspace; when given, layers with few output tiles
 * and a long K loop (the deep U-Net levels) are split along K over up to 16 CTAs per tile. */
TDB_API int tdb_conv3d_bf16(const void* in, int ld_in, const void* w, const float* bias, void* out,
                    int ld_out, int B, int X, int Y, int Z, int Cin, int Cout, int ntaps,
                    double* gn_stats, int G, unsigned flags, float* splitk_scratch, void* stream);

/* Same convolution (3x3x3 only) for narrow layers, Cout in {16,32,64}: the kz filter axis is folded
 * into the GEMM N dimension (9 row-shifted A boxes instead of 27, 3x wider MMAs), persistent CTAs,
 * double-buffered TMEM accumulators, weights resident in shared memory when they fit.
 * w_fold: packed [3*Cout][9*Cin] bf16, row = kz*Cout + co, col = (kx*3+ky)*Cin + ci.
 * The activation tiles are fetched through an overlapping 4-D TMA view that cannot use TMA's
 * out-of-bounds fill: `in` must be preceded AND followed by `pad_rows` >= Yp*Zp + 2*Zp + 256 rows
 * (of ld_in elements) of readable memory; their contents never reach a stored output. */
TDB_API int tdb_conv3d_bf16_fold(const void* in, int ld_in, int pad_rows, const void* w_fold, const float* bias, void* out,
                         int ld_out, int B, int X, int Y, int Z, int Cin, int Cout, double* gn_stats,
                         int G, unsigned flags, void* stream);

/* cta_group::2 variant of tdb_conv3d_bf16_fold (same contract; Cin % 64 == 0; Cout in {32,64} or a multiple of 128
 * <= 512, processed as 128-channel N tiles with w_fold rows ordered [tile][kz][co]): two CTAs of a cluster issue
 * one 256-row MMA, each staging half of the weight rows, which stay resident in shared memory when the half fits.
 * w_proj (nullable, Cout <= 64): bf16 [Cout][Cin] weights of the ResnetBlock's 1x1x1 residual projection of the SAME
 * input (ddpm.py:188,197); it is evaluated on the centre-tap activation tiles at no extra traffic and written
 * (+ bias_proj) to the halo grid out_proj (interior rows). */
TDB_API int tdb_conv3d_bf16_fold2(const void* in, int ld_in, int pad_rows, const void* w_fold, const float* bias,
                          void* out, int ld_out, int B, int X, int Y, int Z, int Cin, int Cout, double* gn_stats,
                          int G, unsigned flags, const void* w_proj, const float* bias_proj, void* out_proj,
                          int ld_outp, void* stream);

/* Row-window CTA-pair convolution (3x3x3 only; replaces nn.Conv3d(3, padding_mode="replicate") + res_conv of reference
 * turbdiff/models/ddpm.py:164,188,197): for one kx the nine (ky, kz) taps are nine row-shifted views of ONE
 * shared-memory window of 128 + 2*(Z+2) + 2 rows, so activations are staged 3x per channel chunk instead of 9x/27x and
 * all 128 rows of a tile are outputs.  w = the per-tap layout of tdb_conv3d_bf16 ([Cout][27*Cin] bf16).  Cin % 32 == 0,
 * Cout in {32, 64, 128}: the weights stay resident in shared memory, split over the pair, when 27*Cin*Cout bytes leave
 * room for two windows (<= ~116 KB); otherwise Cin % 64 == 0, Cout % 128 == 0 (<= 512): N tiles of 128 channels whose
 * nine weight tiles stream with every window.  Z + 2 <= 63.  No padding rows are required (TMA zero-fills outside the
 * grid).  gn_stats / flags / w_proj, bias_proj, out_proj, ld_outp as in tdb_conv3d_bf16_fold2. */
TDB_API int tdb_conv3d_bf16_win(const void* in, int ld_in, const void* w, const float* bias, void* out, int ld_out, int B,
                        int X, int Y, int Z, int Cin, int Cout, double* gn_stats, int G, unsigned flags,
                        const void* w_proj, const float* bias_proj, void* out_proj, int ld_outp, void* stream);

/* tdb_conv3d_bf16_win plus a 1x1 convolution of a second input accumulated in the same output tile:
 *   out = conv3x3x3(in, w) + bias + conv1x1(in2, w2),   in2 [rows][ld_in2] with Cin channels, w2 [Cout][Cin] bf16.
 * The training backward of a ResnetBlock with a residual projection (reference ddpm.py:190-197 through autograd):
 * grad_x = dgrad(block1.conv)(d_raw1) + res_conv^T(grad_out) - one kernel instead of two convolutions and an add pass. */
TDB_API int tdb_conv3d_bf16_win_add1x1(const void* in, int ld_in, const void* w, const float* bias, void* out, int ld_out, int B,
                               int X, int Y, int Z, int Cin, int Cout, unsigned flags, const void* in2, int ld_in2,
                               const void* w2, void* stream);

/* Row-window CTA-pair convolution with kz folded into N (N = 3*Cout; same reference call sites, ddpm.py:164,188,197): one shared-memory window per kx viewed at the
 * three ky offsets, the +-1 row shift of kz applied in the epilogue (tiles of 128 rows advancing by 126).  For the
 * narrow layers: Cout in {32, 64}, Cin % 32 == 0, folded weights (layout of tdb_conv3d_bf16_fold2: [3*Cout][9*Cin])
 * resident in shared memory split over the pair (27*Cin*Cout bytes <= ~116 KB), Z + 2 <= 64.  Other arguments as
 * tdb_conv3d_bf16_win. */
TDB_API int tdb_conv3d_bf16_winz(const void* in, int ld_in, const void* w_fold, const float* bias, void* out, int ld_out,
                         int B, int X, int Y, int Z, int Cin, int Cout, double* gn_stats, int G, unsigned flags,
                         const void* w_proj, const float* bias_proj, void* out_proj, int ld_outp, void* stream);

/* 32 -> 32 channels (reference ddpm.py:164 in the full-resolution blocks), input pitch exactly 32 (64-byte rows): the kz-folded row-window kernel over PAIRED rows - two
 * consecutive grid rows are fetched as one 128-byte line (half the TMA requests) and the K index of the MMA selects the
 * parity.  Needs an even Z + 2 (<= 128) and a 128-byte aligned input; w_fold as tdb_conv3d_bf16_fold ([96][288]). */
TDB_API int tdb_conv3d_bf16_winp(const void* in, int ld_in, const void* w_fold, const float* bias, void* out, int ld_out,
                         int B, int X, int Y, int Z, int Cin, int Cout, double* gn_stats, int G, unsigned flags,
                         void* stream);

/* ---- normalisation / pointwise --------------------------------------------------------- */

/* GroupNorm statistics over the interior voxels of a halo grid (ddpm.py:165,170,472):
 * stats[b][g] = (sum, sum of squares) in double, ACCUMULATED into a pre-zeroed buffer. */
TDB_API int tdb_gn_stats(const void* raw, int ld, double* stats, int B, int X, int Y, int Z, int C, int G,
                 int dtype, void* stream);

/* Fused GroupNorm-apply + FiLM + SiLU + residual + halo materialisation
 * (ddpm.py:170-176,197,48): for every voxel p of the OUTPUT halo grid (halo voxels read
 * their clamped interior source):
 *     v = raw[src][c]
 *     if stats: v = (v - mean)*rstd*gamma[c] + beta[c]          (eps, biased variance)
 *     if film:  v = film[b][C + c] + (film[b][c] + 1) * v        (scale first, shift second)
 *     if flags & TDB_PW_SILU: v = v * sigmoid(v)
 *     if res:   v += res[src][c]
 *     out[p][c] = v
 * stats: double [B][G][2] from tdb_gn_stats; film: fp32 [B][film_ld] with this block's
 * (scale|shift) at film[b][0..2C). */
TDB_API int tdb_pointwise(const void* raw, int ld_raw, const double* stats, const float* gamma,
                  const float* beta, const float* film, int film_ld, const void* res, int ld_res,
                  void* out, int ld_out, int B, int X, int Y, int Z, int C, int G, float eps,
                  unsigned flags, int dtype, void* stream);

/* Trilinear resampling, align_corners=True, halo grid -> halo grid incl. halo
 * (ddpm.py:358-361,367-369): src = i*(n_in-1)/(n_out-1) per axis in fp32. */
/* Up-sampling runs on one of two kernels with bit-identical results: a line walker, and - for outputs of 1.5 M rows and more
 * whose channel vectors divide a warp - a two-stage kernel (x/y blend of the input line in shared memory, then the z blend).
 * dtype | TDB_TRILINEAR_LINE selects the second one regardless of the size (kernel tests). */
#define TDB_TRILINEAR_LINE 0x100
TDB_API int tdb_trilinear(const void* in, int ld_in, int Xi, int Yi, int Zi, void* out, int ld_out, int Xo,
                  int Yo, int Zo, int B, int C, int dtype, void* stream);

/* ---- bottleneck attention ---------------------------------------------------------------- */

/* softmax(q k^T / sqrt(dh)) v per (sample, head) over the S = X*Y*Z interior voxels
 * (ddpm.py:295-308, attention.py:9-15).  qkv: halo grid with 3*heads*dh channels ordered
 * q|k|v, channel = head*dh + d.  out: halo grid, heads*dh channels (interior rows written). */
TDB_API int tdb_attention(const void* qkv, int ld_qkv, void* out, int ld_out, int B, int X, int Y, int Z,
                  int heads, int dh, int dtype, void* stream);

/* bf16 kernel-layout copy of one convolution weight straight from the fp32 parameter (reference layout (Cout, Cin, kD, kH, kW),
 * nn.Conv3d at ddpm.py:164,188), one launch, no intermediates.  taps = 27 or 1.  folded = 0: per-tap layout [O][taps*I]
 * (tdb_conv3d_bf16, _win); folded = 1: kz-folded [3*O][9*I] in N tiles of tile_n rows (_fold, _fold2, _winz, _winp).
 * transpose = 1: the weights of the input-gradient convolution, W'[o = ci][i = co][tap] = W[co][ci][26 - tap]
 * (O = Cin, I = Cout); transpose = 0: the forward weights (O = Cout, I = Cin). */
TDB_API int tdb_pack_conv_weights(const float* w, void* dst, int Cout, int Cin, int taps, int folded, int tile_n,
                          int transpose, void* stream);

/* The same for n weights in one launch per 64 weights (the job table is a kernel parameter).  In training every layout is
 * re-derived after each optimizer step: 62 one-weight launches were ~0.7 ms at the head of the step.  `jobs` is a host
 * array, read during the call only. */
typedef struct TdbPackJob {
    const float* w; /* fp32 (Cout, Cin, taps) */
    void* dst;      /* bf16, Cout*Cin*taps elements */
    int Cout, Cin, taps, folded, tile_n, transpose;
} TdbPackJob;
TDB_API int tdb_pack_conv_weights_batch(const TdbPackJob* jobs, int n, void* stream);

/* ---- timestep conditioning ------------------------------------------------------------------ */

/* Nyquist embedding -> process_c MLP -> all FiLM projections of the network in one call
 * (ddpm.py:147-148,447-452,184,191).  t: int64 (B,).  emb_scale/emb_bias: (dim,).
 * w1 (4dim,dim) b1 (4dim) w2 (dim,4dim) b2 (dim).  film_wt: (dim, film_rows) = all
 * project_onto_scale_shift weights stacked along rows and TRANSPOSED (coalesced reads),
 * film_b (film_rows).  Outputs: c (B,dim), film (B,film_rows). */
TDB_API int tdb_time_film(const int64_t* t, const float* emb_scale, const float* emb_bias, const float* w1,
                  const float* b1, const float* w2, const float* b2, const float* film_wt,
                  const float* film_b, float* c, float* film, int B, int dim, int film_rows,
                  void* stream);

/* Backward of tdb_time_film (autograd of ddpm.py:447-452, 184/191): from d_film (B, film_rows) = the FiLM scale/shift
 * gradients of every block, and c (B, dim) = the forward's conditioning vector, the gradients of all FiLM projections
 * (g_film_w (film_rows, dim), g_film_b (film_rows)) and of the process_c MLP (g_w1 (4dim, dim), g_b1, g_w2 (dim, 4dim),
 * g_b2).  dc_scratch: (B, dim) floats of workspace (zeroed by the call).  Two launches, no library calls. */
TDB_API int tdb_time_film_bwd(const int64_t* t, const float* emb_scale, const float* emb_bias, const float* w1,
                      const float* b1, const float* w2, const float* b2, const float* film_wt, const float* c,
                      const float* d_film, float* g_film_w, float* g_film_b, float* g_w1, float* g_b1, float* g_w2,
                      float* g_b2, float* dc_scratch, int B, int dim, int film_rows, void* stream);

/* ---- diffusion process ------------------------------------------------------------------------ */

/* One ancestral sampling update, masked to inside cells, as a single bandwidth-bound
 * kernel (ddpm.py:711-715,722-728,745-752,797-814).  All tensors (B,F,nvox) fp32 NCDHW.
 *   coef: device table [T][8] fp32 rows = {sqrt_recip_acp, sqrt_recipm1_acp, post_coef1,
 *         post_coef2, exp(0.5*log_betas), sqrt_acp, sqrt_one_minus_acp, 0}
 *   t_ptr: device int32 holding the current step t (so a captured graph can be replayed)
 *   mask: uint8 (nvox), 1 on inside cells (== where_cells' cell_idx set)
 *   z: posterior noise; z_bc: boundary re-noising draw (NULL unless NOISE_BCS). At t==0 the
 *   noise tensors are ignored (x <- mean). x_out may alias x_t.
 *   TDB_STEP_LEARNED_VAR: eps is the model's (B,2F,nvox) output (noise prediction | variance weights v) and the step's
 *   standard deviation is exp(lerp(log_betas[t], posterior_log_var[t], sigmoid(v)) / 2) per voxel; the coefficient row
 *   then holds log_betas[t] in slot 4 and posterior_log_var[t] in slot 7. */
TDB_API int tdb_ddpm_step(const float* x_t, const float* eps, const float* z, const float* z_bc,
                  const float* x_bcs, const uint8_t* mask, const float* coef, const int32_t* t_ptr,
                  float* x_out, int B, int F, int64_t nvox, unsigned flags, void* stream);

/* Fused tail of one sampling step (everything between the denoiser's last convolution and the next step's first):
 * decode.0's second GroupNorm + SiLU + residual (ddpm.py:168-177,197) -> decode.1 1x1x1 conv (ddpm.py:459,505) -> the
 * tdb_ddpm_step update (same flags / coefficient table / bit-exact arithmetic) -> encode_x of the NEXT step written into
 * the level-0 input halo grid `xin0` (ddpm.py:495), halo rows included.  raw / res: halo grids of decode.0 block2's
 * convolution output and of the block input (dim channels), stats: that norm's [B][G][2] double moments; w_dec (Fo,dim),
 * b_dec (Fo); w_enc (dim,F), b_enc (dim); x_in / z / z_bc / x_bcs / x_out: (B,F,nvox) fp32 NCDHW, x_out != x_in (halo rows
 * still read the old state); eps_out: optional (B,Fo,nvox) copy of the model output (NULL = not stored).
 * dim in {8,16,32,64}, F <= 4, Fo <= 8.  Results equal the unfused launches bit for bit given the same moments. */
TDB_API int tdb_step_tail(const void* raw, int ld_raw, const double* stats, const float* gamma, const float* beta, const void* res,
                  int ld_res, const float* w_dec, const float* b_dec, int Fo, const float* x_in, const float* z,
                  const float* z_bc, const float* x_bcs, const uint8_t* mask, const float* coef, const int32_t* t_ptr,
                  float* x_out, float* eps_out, int F, unsigned flags, const float* w_enc, const float* b_enc, void* xin0,
                  int ld_xin0, int B, int X, int Y, int Z, int dim, int G, float eps_gn, int dtype, void* stream);

/* q_sample with optional inside-cell masking (ddpm.py:818-822,837-838):
 * out = sqrt_acp[t_b]*x0 + sqrt(1-acp)[t_b]*noise ; where mask==0 and !noise_bcs: out = x0.
 * t: int64 (B,) device. coef as in tdb_ddpm_step. */
TDB_API int tdb_q_sample(const float* x0, const float* noise, const int64_t* t, const float* coef,
                 const uint8_t* mask, float* out, int B, int F, int64_t nvox, int noise_bcs,
                 void* stream);

/* Masked training loss and its gradient (ddpm.py:845-852):
 * loss = mean_b mean_{f, inside} |eps-noise|^p  (p=2: l2, p=1: l1), accumulated in double into
 * loss_acc[0] (pre-zeroed); grad (nullable) = dloss/deps, zero outside the mask.
 * n_inside = number of inside cells. */
TDB_API int tdb_masked_loss(const float* eps, const float* noise, const uint8_t* mask, double* loss_acc,
                    float* grad, int B, int F, int64_t nvox, int64_t n_inside, int l1, void* stream);

/* ---- backward (training path; the reference gets these from torch.autograd) ---------------------- */

/* Adjoint of halo materialisation: for every border voxel, add the gradient stored on its halo images
 * (in place); the halo rows are then set to zero, so a folded gradient qualifies for TDB_WGRAD_ZERO_HALO. */
TDB_API int tdb_halo_fold(void* g, int ld, int B, int X, int Y, int Z, int C, int dtype, void* stream);

/* Backward of tdb_pointwise, pass 1: red[b][c] = (sum g_u, sum g_u*xhat, sum raw, sum g_out) over interior voxels in
 * double (pre-zeroed, B*C*4), where u is the forward pre-activation, g_u = g_out*silu'(u) (or g_out) and
 * xhat = (raw-mean)*rstd.  g_out must already be folded.  (sum raw gives the host the per-channel sum of d_raw,
 * i.e. the bias gradient of the preceding convolution, without another pass; sum g_out is the bias gradient of a
 * residual 1x1 projection added to the same block output.) */
TDB_API int tdb_pointwise_bwd_reduce(const void* g_out, int ld_g, const void* raw, int ld_raw, const double* stats,
                             const float* gamma, const float* beta, const float* film, int film_ld, double* red,
                             int B, int X, int Y, int Z, int C, int G, float eps, unsigned flags, int dtype,
                             void* stream);

/* Between the two passes (one block): from red and the forward moments `stats`, grp[b][g] = (m1, m2) for pass 2,
 * colsum[c] = sum of d_raw over samples and interior voxels (= bias gradient of the convolution that produced raw),
 * gw[c] / gb[c] = gradients of the GroupNorm weight / bias, gsum[c] (nullable) = sum of g_out, and (dfilm != NULL) the FiLM gradients
 * dfilm[b][c] = d scale, dfilm[b][C + c] = d shift (reference: autograd through ddpm.py:168-177). */
TDB_API int tdb_pointwise_bwd_finalize(const double* red, const double* stats, const float* gamma, const float* beta,
                               const float* film, int film_ld, float* grp, float* colsum, float* gw, float* gb,
                               float* gsum, float* dfilm, int dfilm_ld, int B, int X, int Y, int Z, int C, int G,
                               float eps, void* stream);

/* Pass 2: d_raw = rstd*(k*g_u - m1 - xhat*m2) on interior rows and 0 on halo rows, k = gamma*(scale+1),
 * grp[b][g] = (m1, m2) = group means of k*A1 and k*A2 (fp32).  Without stats: d_raw = g_u. */
TDB_API int tdb_pointwise_bwd_apply(const void* g_out, int ld_g, const void* raw, int ld_raw, const double* stats,
                            const float* gamma, const float* beta, const float* film, int film_ld,
                            const float* grp, void* d_raw, int ld_d, int B, int X, int Y, int Z, int C, int G,
                            float eps, unsigned flags, int dtype, void* stream);

/* tdb_pointwise_bwd_finalize + tdb_pointwise_bwd_apply in one launch: every block derives the group sums from `red`
 * (the reduce kernel's output) in its prologue and one extra block per sample writes the parameter gradients
 * (colsum, gw, gb, gsum [C]; dfilm [B][dfilm_ld] may be null).  Same results as the two-launch sequence. */
TDB_API int tdb_pointwise_bwd_apply_fused(const void* g_out, int ld_g, const void* raw, int ld_raw, const double* stats,
                                  const float* gamma, const float* beta, const float* film, int film_ld,
                                  const double* red, void* d_raw, int ld_d, float* colsum, float* gw, float* gb,
                                  float* gsum, float* dfilm, int dfilm_ld, int B, int X, int Y, int Z, int C, int G,
                                  float eps, unsigned flags, int dtype, void* stream);

/* Weight gradient of tdb_conv3d_*: dw[tap][ci][co] += sum over interior rows p of
 * in[p+delta(tap)][ci]*d_out[p][co]; dw fp32 [ntaps][Cin][Cout], accumulated (pre-zero it).
 * flags & TDB_WGRAD_ZERO_HALO: the caller guarantees that d_out is zero on halo rows and that `in` is readable
 * Yp*Zp+Zp+1 rows before/after the grid; the bf16 path then runs on tensor cores without per-row masking. */
#define TDB_WGRAD_ZERO_HALO 1u
TDB_API int tdb_conv3d_wgrad(const void* in, int ld_in, const void* d_out, int ld_do, float* dw, int B, int X, int Y,
                     int Z, int Cin, int Cout, int ntaps, int dtype, unsigned flags, void* stream);

/* The bf16 tensor-core form of tdb_conv3d_wgrad (reference: torch.autograd's conv weight gradient behind ddpm.py:164,188
 * when GaussianDiffusion.forward's loss is back-propagated, ddpm.py:874-882) (tcgen05.mma with MN-major operands straight from the halo grids);
 * same contract as TDB_WGRAD_ZERO_HALO (d_out zero on halo rows); needs Cin % 32 == 0 and Cout in {32, 64*n <= 256,
 * 256*n}.  tdb_conv3d_wgrad dispatches here when it can.  mode: TDB_WGRAD_SHARE_KZ = load one row window per
 * (kx, ky) and use it for the three kz taps through row-shifted matrix descriptors (0 = one window per tap). */
#define TDB_WGRAD_SHARE_KZ 1u
/* Cout in {32, 64}, 27 taps: the three kz taps are stacked on the GEMM's N side (N = 3*Cout) - one activation window per
 * (kx, ky, channel chunk) and one output-gradient window whose three swizzle atoms start one row apart. */
#define TDB_WGRAD_KZ_ON_N 2u
TDB_API int tdb_conv3d_wgrad_tc(const void* in, int ld_in, const void* d_out, int ld_do, float* dw, int B, int X, int Y,
                        int Z, int Cin, int Cout, int ntaps, unsigned mode, void* stream);

/* dw [taps][Cin][Cout] fp32 (what tdb_conv3d_wgrad accumulates) -> the parameter layout (Cout, Cin, kD, kH, kW) fp32 of
 * nn.Conv3d.weight.grad (ddpm.py:164,188), tiled through shared memory (as a torch permute + contiguous this was a strided
 * copy at ~1 TB/s, 36 launches and 0.43 ms per training step). */
TDB_API int tdb_unpack_wgrad(const float* dw, float* out, int Cout, int Cin, int taps, void* stream);

/* Transpose of tdb_trilinear (reference: autograd of F.interpolate(mode="trilinear", align_corners=True), ddpm.py:358-369):
 * d_in (interior rows; halo rows zero) from the output gradient's interior rows.
 * flags & TDB_TRIBWD_ACCUMULATE: d_in += the transposed gradient on interior rows, halo rows untouched (the skip
 * connection's gradient and the gradient through the down-sampling are summed in one pass). */
#define TDB_TRIBWD_ACCUMULATE 1u
TDB_API int tdb_trilinear_bwd(const void* g_out, int ld_g, int Xo, int Yo, int Zo, void* d_in, int ld_d, int Xi,
                      int Yi, int Zi, int B, int C, int dtype, unsigned flags, void* stream);

/* Backward of tdb_attention: d_qkv (q|k|v gradient, interior rows) from qkv and d_out.  S <= 135: the S x S probability /
 * score-gradient matrices in shared memory; longer sequences (up to ~650 voxels): streaming form, nothing quadratic stored. */
TDB_API int tdb_attention_bwd(const void* qkv, int ld_qkv, const void* d_out, int ld_do, void* d_qkv, int ld_dq, int B,
                      int X, int Y, int Z, int heads, int dh, int dtype, void* stream);

/* out[c][f] += sum_{b, interior v} G[b,v][c] * Q[b*q_bstride + f*nvox + v]: weight gradients of the 1x1x1
 * encoders / decoder (reference ddpm.py:433,436,459 through autograd) between a halo grid G and NCDHW planes Q
 * (q_bstride = 0 for an unbatched Q).  colsum (optional, fp32 [C], accumulated): colsum[c] += sum_{b, interior v} G[b,v][c],
 * the bias gradient of the same 1x1x1 convolution, from the same pass. */
TDB_API int tdb_cl_nc_outer(const void* G, int ld, const float* Q, int64_t q_bstride, float* out, float* colsum, int B, int X,
                    int Y, int Z, int C, int F, int dtype, void* stream);

/* ---- cell indexing (bit-exact) -------------------------------------------------------------------- */

/* out = mask ? a : other (other NULL -> 0): models/utils.py:22-28 where_cells. (R, nvox) rows. */
TDB_API int tdb_where_cells(const float* a, const float* other, const uint8_t* mask, float* out, int64_t rows,
                    int64_t nvox, void* stream);
/* out[r][j] = x[r][cell_idx[j]]: models/utils.py:14-15 select_cells. */
TDB_API int tdb_select_cells(const float* x, const int64_t* cell_idx, float* out, int64_t rows, int64_t nvox,
                     int64_t n_cells, void* stream);
/* grid[b][f][cell_idx[j]] = samples[b][j][f] on a pre-zeroed grid: data/ofles.py:220-232. */
TDB_API int tdb_scatter_cells(const float* samples, const int64_t* cell_idx, float* grid, int B, int F,
                      int64_t nvox, int64_t n_cells, void* stream);
/* Fused scatter + FIXED_VALUE boundary writes + normalise: grid[b][f][v] = fma(scale[f], value, shift[f])
 * with value = the fixed boundary value of channel f where the voxel's boundary class fixes it, else samples[b][j][f] on voxel
 * cell_idx[j], else 0: OpenFOAMData.grid_embedding data/ofles.py:220-240 (cell scatter :231-232, boundary writes :233-238)
 * followed by Normalization.normalize_grid models/normalization.py:20-24; scale = 1/std, shift = -mean/std.
 * code[v]: bit 0 = cell (tdb_build_mask), bits 1..7 = boundary class (0 = none); bc_has / bc_val: [n_classes + 1][F] tables
 * (row 0 unused), both NULL when the case fixes no boundary values.  Bit-exact with the reference's torch op sequence. */
TDB_API int tdb_scatter_normalize(const float* samples, const int64_t* cell_idx, const uint8_t* code, const uint8_t* bc_has,
                          const float* bc_val, const float* scale, const float* shift, float* grid, int B, int F, int64_t nvox,
                          int64_t n_cells, void* stream);
/* Fused de-normalise + gather, channels-last: out[b][j][f] = fma(scale[f], x[b][f][cell_idx[j]], shift[f]):
 * models/normalization.py:26-30 + models/utils.py:14-15 + the "b f c -> b c f" of models/metrics.py:50-57; scale = std, shift = mean. */
TDB_API int tdb_gather_denormalize(const float* x, const int64_t* cell_idx, const float* scale, const float* shift, float* out,
                           int B, int F, int64_t nvox, int64_t n_cells, void* stream);
/* mask[cell_idx[j]] = 1 on a pre-zeroed mask. */
TDB_API int tdb_build_mask(const int64_t* cell_idx, uint8_t* mask, int64_t n_cells, int64_t nvox, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TURBDIFF_B200_H */
