/*
 * pnpadmm.h — C ABI of the B200-native ADMM / PnP-ADMM CS-MRI reconstruction path.
 *
 * This is the drop-in boundary for ONE hot path of zj15001/PNP_ADMM_CNC_MRI: the ADMM
 * iteration loop that is copy-pasted into every entry script of the reference.  The
 * reference has no FFI of its own for this path (it is inline NumPy inside Python
 * functions), so every entry point below cites the reference code block it replaces
 * (file:line, script aliases S1 = "【1】ADMM_L1.py", S3 = "【3】PNP_ADMM_L1_D  .py",
 * S4 = "【4】ADMM_CNC .py", S6 = "【6】PNP_ADMM_CNC_D .py").  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every array pointer is a DEVICE pointer on the current CUDA device unless its name
 *     starts with "h_" (host).  The *_host entry points take host pointers and do their
 *     own staged H2D / D2H copies.
 *   - images are row-major [B][N][N]; complex arrays are interleaved (re, im), i.e.
 *     [B][N][N][2]; masks are uint8 {0,1} with the DC bin at [0][0] (no fftshift),
 *     exactly like CS_MRI/Q_*.mat['Q1'].
 *   - N is a power of two, 16 <= N <= 2048 (f32) / 1024 (f64).  N == 256 in f32 selects the
 *     thread-block-cluster kernel (whole state resident in distributed shared memory);
 *     everything else uses the streaming kernels.
 *   - the library never allocates per call: the caller owns a workspace of
 *     pnpadmm_workspace_bytes() bytes.  Work is stream-ordered on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream); no host sync inside.
 *   - all functions return 0 on success or a negative PNPADMM_ERR_* code and never throw.
 *     pnpadmm_last_error_string() describes the last failure on the calling thread.
 *   - scalar parameters are passed as double in both precisions and rounded once inside.
 */
#ifndef PNPADMM_H_
#define PNPADMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNPADMM_ABI_VERSION 3

#define PNPADMM_OK                 0
#define PNPADMM_ERR_BAD_ARG       -1   /* NULL pointer, B <= 0, iters < 0, unknown enum ...          */
#define PNPADMM_ERR_BAD_SIZE      -2   /* N not a power of two or outside the supported range        */
#define PNPADMM_ERR_WORKSPACE     -3   /* workspace NULL / too small / misaligned (256 B)            */
#define PNPADMM_ERR_CUDA          -4   /* a CUDA runtime call or kernel launch failed                */
#define PNPADMM_ERR_UNSUPPORTED   -5   /* e.g. cluster kernel requested for N != 256 or on a non-sm_100 device */

#define PNPADMM_PROX_L1    0           /* S1:123                                                    */
#define PNPADMM_PROX_CNC   1           /* S4:127-129                                                */

#define PNPADMM_KERNEL_AUTO       0    /* cluster kernel when N == 256 && f32, else streaming        */
#define PNPADMM_KERNEL_CLUSTER    1    /* K1: 8-CTA cluster, state resident in DSMEM (N == 256, f32) */
#define PNPADMM_KERNEL_STREAMING  2    /* K2: two passes per iteration through L2 / HBM              */
#define PNPADMM_KERNEL_ROWSEP     3    /* K3 (pnpadmm_reconstruct_f32 / host entry points, N == 256): the caller asserts a
                                          ROW-SEPARABLE mask, mask[kr][kc] == mask[0][kc] for all kr (full k-space lines, e.g.
                                          CS_MRI/Q_Cartesian30).  The blend then commutes with the column transforms and every
                                          image row is solved on its own: no column FFTs, no transposes.  The assertion is
                                          verified on the device; a mask that violates it yields NaN in x, z, w.  The host
                                          entry points test h_mask themselves and pick this kernel under KERNEL_AUTO.        */

typedef void* pnpadmm_stream_t;        /* cudaStream_t */

int         pnpadmm_abi_version(void);
const char* pnpadmm_last_error_string(void);

/* sm_count, max co-resident 8-CTA clusters of the N=256 kernel (0 if it cannot launch),
 * compute capability.  Any out pointer may be NULL. */
int pnpadmm_device_info(int* sm_count, int* max_clusters_256, int* cc_major, int* cc_minor);

/* How a reconstruction of (B, N, iters) is scheduled on the current device and how many kernels it launches
 * (introspection for benchmarks; nothing is enqueued): packed planes given to the cluster kernel and to the
 * streaming kernels (hybrid schedule at N == 256), chunks of the cluster schedule, kernel launches of one
 * pnpadmm_acquire_f32, one pnpadmm_solve_f32 and one pnpadmm_reconstruct_f32.  Any out pointer may be NULL. */
int pnpadmm_plan_info(int B, int N, int mask_batched, int iters, int kernel, int* planes_cluster,
                      int* planes_streaming, int* chunks, int* launches_acquire, int* launches_solve,
                      int* launches_reconstruct);

/* Bytes of device workspace needed by every call below for (B, N, precision, mask layout). */
size_t pnpadmm_workspace_bytes(int B, int N, int is_f64, int mask_batched);

/* ---------------------------------------------------------------------------------------
 * a2  acquisition:  y = fft2(img) * mask + noises           (S1:99 == S4:103 == S3:242 == S6:251)
 *   img   [B][N][N] real in [0,1];  mask [N][N] or [B][N][N] u8;  noise [N][N][2] or [B][N][N][2]
 *   y     [B][N][N][2] out.  Noise is added on EVERY bin, sampled or not, like the reference.
 *   spectrum_f32 (f64 entry only): the reference's image is float32 (utils_image.uint2single) and
 *   NumPy >= 2 evaluates fft2 of a float32 array in single precision, rounding each 1-D pass to
 *   complex64 before `* mask + noises` promotes to complex128.  Non-zero reproduces that rounding
 *   (the reference as it runs today); zero keeps the full double-precision spectrum (NumPy 1.x, the
 *   author's environment).  The f32 entry computes in float32 throughout.
 * ------------------------------------------------------------------------------------- */
int pnpadmm_acquire_f32(const float* img, const uint8_t* mask, const float* noise, float* y,
                        int B, int N, int mask_batched, int noise_batched,
                        void* ws, size_t ws_bytes, pnpadmm_stream_t stream);
int pnpadmm_acquire_f64(const double* img, const uint8_t* mask, const double* noise, double* y,
                        int B, int N, int mask_batched, int noise_batched, int spectrum_f32,
                        void* ws, size_t ws_bytes, pnpadmm_stream_t stream);

/* zero-filled start  x0 = |ifft2(y)|  (complex magnitude)              (S1:100,104 == S4:104,108) */
int pnpadmm_zero_filled_f32(const float* y, float* x0, int B, int N,
                            void* ws, size_t ws_bytes, pnpadmm_stream_t stream);
int pnpadmm_zero_filled_f64(const double* y, double* x0, int B, int N,
                            void* ws, size_t ws_bytes, pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Data-term preparation for the x-update (S1:97,117-118: index = nonzero(mask), La2 = 1/(2 reo)).
 * Builds, inside the workspace, the pair-packed Hermitian-symmetrised measurement term and the
 * per-bin blend coefficient for `reo`.  Must precede pnpadmm_xupdate_* / pnpadmm_iterate_* and be
 * repeated when y, mask or reo change.  (pnpadmm_solve_* calls it itself.)
 * ------------------------------------------------------------------------------------- */
int pnpadmm_prepare_f32(const float* y, const uint8_t* mask, int B, int N, int mask_batched,
                        double reo, void* ws, size_t ws_bytes, pnpadmm_stream_t stream);
int pnpadmm_prepare_f64(const double* y, const uint8_t* mask, int B, int N, int mask_batched,
                        double reo, void* ws, size_t ws_bytes, pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a3  closed-form x-update                (S1:115-120 == S4:119-124 == S3:259-264 == S6:266-271)
 *     X = fft2(z - w);  X[idx] = (La2 X[idx] + y[idx]) / (1 + La2);  x = |Re(ifft2(X))|
 *   z, w, x : [B][N][N].  xpw (may be NULL) additionally receives x + w (the PnP-L1 denoiser
 *   input, S3:290).  Uses the workspace prepared by pnpadmm_prepare_*.
 * ------------------------------------------------------------------------------------- */
int pnpadmm_xupdate_f32(const float* z, const float* w, float* x, float* xpw,
                        int B, int N, int mask_batched, int kernel,
                        void* ws, size_t ws_bytes, pnpadmm_stream_t stream);
int pnpadmm_xupdate_f64(const double* z, const double* w, double* x, double* xpw,
                        int B, int N, int mask_batched, int kernel,
                        void* ws, size_t ws_bytes, pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * The ADMM loop proper          (S1:111-126 for PROX_L1, S4:115-132 for PROX_CNC), `iters` times:
 *     x-update (a3);  z-update (a4: S1:123 | a5: S4:127-129);  w = w + x - z (a6: S1:126)
 *   z, w : in/out state [B][N][N];  x : out, the last x-update (what the reference returns in
 *   out[n], S1:132).  Needs pnpadmm_prepare_* first.  iters == 0 leaves z, w untouched, x undefined.
 * ------------------------------------------------------------------------------------- */
int pnpadmm_iterate_f32(float* x, float* z, float* w, int B, int N, int mask_batched,
                        int prox, int iters, double lambda1, double reo, double alpha, double b,
                        int kernel, void* ws, size_t ws_bytes, pnpadmm_stream_t stream);
int pnpadmm_iterate_f64(double* x, double* z, double* w, int B, int N, int mask_batched,
                        int prox, int iters, double lambda1, double reo, double alpha, double b,
                        int kernel, void* ws, size_t ws_bytes, pnpadmm_stream_t stream);

/* Whole reconstruction from measurements (S1:100-126 / S4:104-132):
 *     x = |ifft2(y)|; z = x; w = 0; prepare; iterate.       Outputs x, z, w : [B][N][N]. */
int pnpadmm_solve_f32(const float* y, const uint8_t* mask, float* x, float* z, float* w,
                      int B, int N, int mask_batched,
                      int prox, int iters, double lambda1, double reo, double alpha, double b,
                      int kernel, void* ws, size_t ws_bytes, pnpadmm_stream_t stream);
int pnpadmm_solve_f64(const double* y, const uint8_t* mask, double* x, double* z, double* w,
                      int B, int N, int mask_batched,
                      int prox, int iters, double lambda1, double reo, double alpha, double b,
                      int kernel, void* ws, size_t ws_bytes, pnpadmm_stream_t stream);

/* Whole reconstruction from IMAGES on the device: what the body of the reference's `for img in L_paths` loop does
 * (S1:97-132 / S4:101-138), batched:  y = fft2(img) * mask + noises;  x = |ifft2(y)|; z = x; w = 0;  `iters` iterations.
 *   img [B][N][N] real in [0,1], or (img == NULL) img8 [B][N][N] uint8 gray levels, divided by 255 on the device like
 *   utils_image.uint2single (utils_image.py:181);  mask, noise as for pnpadmm_acquire_*;  outputs x, z, w [B][N][N].
 * For N == 256 in f32 with one mask and one noise array for the batch, acquisition, zero-filled start and the data term run
 * INSIDE the cluster kernel (one preparation launch + one kernel for the whole reconstruction; y is never materialised);
 * otherwise this is pnpadmm_acquire_* into the workspace followed by pnpadmm_solve_*.  The f64 entry rounds the spectrum
 * like NumPy >= 2 does for the reference's float32 image (spectrum_f32 = 1 of pnpadmm_acquire_f64). */
int pnpadmm_reconstruct_f32(const float* img, const uint8_t* img8, const uint8_t* mask, const float* noise,
                            float* x, float* z, float* w, int B, int N, int mask_batched, int noise_batched,
                            int prox, int iters, double lambda1, double reo, double alpha, double b,
                            int kernel, void* ws, size_t ws_bytes, pnpadmm_stream_t stream);
int pnpadmm_reconstruct_f64(const double* img, const uint8_t* img8, const uint8_t* mask, const double* noise,
                            double* x, double* z, double* w, int B, int N, int mask_batched, int noise_batched,
                            int prox, int iters, double lambda1, double reo, double alpha, double b,
                            int kernel, void* ws, size_t ws_bytes, pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Reference-facing convenience with HOST buffers (what `ADMM_L1(mask, noises, **opts)` /
 * `ADMM_CNC(...)` do per image, S1:97-132 / S4:101-138, batched):  h_img u8 [B][N][N] gray
 * levels (divided by 255 on the device like utils_image.uint2single, utils_image.py:181),
 * h_mask u8 [N][N], h_noise complex64 [N][N][2] (already x3, S1:186), h_x f32 [B][N][N] out.
 * d_scratch: device buffer of pnpadmm_host_scratch_bytes(B, N) bytes (holds img, y, x, z, w);
 * copies run on `stream`; returns after the D2H copy has been enqueued (caller syncs the
 * stream).  Host buffers should be pinned for the copies to be asynchronous.
 * ------------------------------------------------------------------------------------- */
size_t pnpadmm_host_scratch_bytes(int B, int N);
int pnpadmm_reconstruct_host_f32(const uint8_t* h_img, const uint8_t* h_mask, const float* h_noise,
                                 float* h_x, int B, int N,
                                 int prox, int iters, double lambda1, double reo, double alpha, double b,
                                 int kernel, void* d_scratch, size_t scratch_bytes,
                                 void* ws, size_t ws_bytes, pnpadmm_stream_t stream);

/* Pipelined form of the same call for back-to-back batches: the H2D copies run on `h2d`, the kernels
 * on `compute`, the D2H copy on `d2h` (three distinct streams), with n_slots (2..4) device slots so that the
 * copies of neighbouring batches overlap the kernels of batch i.  Rotate slot = 0, 1, .., n_slots-1, 0, ..
 * between calls.  The ordering events live in a CALLER-OWNED pipeline object (created once, outside the hot
 * loop), so several pipelines - other host threads, other stream / scratch sets - can share a device; calls on
 * one pipeline object are serialised by the library.  From the second batch on the compute section of a slot
 * (uint8 -> unit scale, acquisition, zero-fill, prepare, iterations) is replayed as ONE CUDA-graph launch.
 * Returns after everything is enqueued; pnpadmm_reconstruct_host_wait(pipe, slot) blocks the host until the
 * h_x passed with that slot has been written.  The host buffers of a slot must stay untouched until then.
 * d_scratch: pnpadmm_host_pipeline_scratch_bytes(B, N, n_slots) bytes, one per pipeline. */
#define PNPADMM_PIPELINE_MAX_SLOTS 4
typedef struct pnpadmm_pipeline_s* pnpadmm_pipeline_t;
int pnpadmm_pipeline_create(pnpadmm_pipeline_t* out, int n_slots);     /* on the current device */
int pnpadmm_pipeline_destroy(pnpadmm_pipeline_t pipe);                 /* after the streams have drained */
/* What a pipeline copies back: PNPADMM_OUT_F32 (default) = the reconstructions x as float32 [B][N][N] (what the reference
 * stores in out[n], S1:132); PNPADMM_OUT_U8 = img_E as the reference SAVES it (S1:133-138: uint8(round(255 x)), saturated),
 * 4x fewer bytes over the host link - h_x then points to B*N*N bytes.  Set between batches, not while a call is in flight. */
#define PNPADMM_OUT_F32 0
#define PNPADMM_OUT_U8  1
int pnpadmm_pipeline_set_output(pnpadmm_pipeline_t pipe, int format);
size_t pnpadmm_host_pipeline_scratch_bytes(int B, int N, int n_slots);
int pnpadmm_reconstruct_host_pipelined_f32(pnpadmm_pipeline_t pipe,
                                           const uint8_t* h_img, const uint8_t* h_mask, const float* h_noise,
                                           float* h_x, int B, int N,
                                           int prox, int iters, double lambda1, double reo, double alpha, double b,
                                           int kernel, void* d_scratch, size_t scratch_bytes,
                                           void* ws, size_t ws_bytes, int slot,
                                           pnpadmm_stream_t compute, pnpadmm_stream_t h2d, pnpadmm_stream_t d2h);
int pnpadmm_reconstruct_host_wait(pnpadmm_pipeline_t pipe, int slot);

/* ---------------------------------------------------------------------------------------
 * Pointwise pieces used by the PnP variants (denoiser runs outside this library).
 * ------------------------------------------------------------------------------------- */
/* a1  soft(x, c) = fmax(|x| - c, 0) * sign(x), sign(0) = 0                       (S1:18-19) */
int pnpadmm_soft_f32(const float* x, float* out, double c, size_t n, pnpadmm_stream_t stream);
int pnpadmm_soft_f64(const double* x, double* out, double c, size_t n, pnpadmm_stream_t stream);

/* a8  t = (1 - alpha) z + alpha (x + w) + coef (z - s),  coef = alpha*reo*lambda1*b
 *                                                                   (S6:301 == S6:518) */
int pnpadmm_cnc_combine_f32(const float* z, const float* x, const float* w, const float* s, float* t,
                            double alpha, double coef, size_t n, pnpadmm_stream_t stream);
int pnpadmm_cnc_combine_f64(const double* z, const double* x, const double* w, const double* s, double* t,
                            double alpha, double coef, size_t n, pnpadmm_stream_t stream);

/* a6 + PnP clamps:  w = w + x - z;  if (clamp01) x, z, w = clamp(., 0, 1)
 *                                                   (S3:293-296 == S6:305-308 == S6:522-525) */
int pnpadmm_dual_update_f32(float* x, float* z, float* w, int clamp01, size_t n, pnpadmm_stream_t stream);
int pnpadmm_dual_update_f64(double* x, double* z, double* w, int clamp01, size_t n, pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Per-image quality metrics on the device: out[B][3] = (PSNR dB, SSIM, RE) of
 * reconstructions x[B][N][N] (unit scale) against uint8 ground truth ref[B][N][N], evaluated in
 * double precision like the reference's utils/utils_image.py calculate_psnr :543-556,
 * calculate_ssim :570-615 (11x11 Gaussian window, valid region) and calculate_re :622-636,
 * border = 0.  quantize = 0: img_E = 255 x (S1:133, S4:139); quantize = 1: img_E =
 * uint8(round(255 clip(x, 0, 1))) (util.single2uint, S6:315, S6:531).  `scratch` (device,
 * 16-byte aligned, pnpadmm_metrics_scratch_bytes(B)) holds the per-image accumulators; `out` is
 * device memory.  Stream-ordered, no host synchronisation.
 * ------------------------------------------------------------------------------------- */
size_t pnpadmm_metrics_scratch_bytes(int B);
int pnpadmm_metrics_f32(const float* x, const uint8_t* ref, int B, int N, int quantize, double* out, void* scratch,
                        size_t scratch_bytes, pnpadmm_stream_t stream);
int pnpadmm_metrics_f64(const double* x, const uint8_t* ref, int B, int N, int quantize, double* out, void* scratch,
                        size_t scratch_bytes, pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a9  DnCNN / FDnCNN denoiser forward on the tensor cores (reference
 * models/network_dncnn.py:36-67 DnCNN.forward = x - model(x), :120-141 FDnCNN.forward = model(x);
 * called from denoising_step S6:353-359, denoising_step1 S3:20-35, denoising_step2 S6:19-34):
 *     conv3x3 cin->64 + ReLU,  n_mid x [conv3x3 64->64 + ReLU],  conv3x3 64->1      (nb = n_mid + 2)
 * bf16 operands, fp32 accumulation (tcgen05 / TMEM), bf16 activations between layers,
 * zero padding 1, bias in every layer (act_mode 'R': no batch norm).
 *   x      [B][cin][H][W] f32 (cin = 1 DnCNN, 2 FDnCNN: noise-level map in channel 1)
 *   out    [B][H][W] f32:  residual != 0 ? x[:, 0] - n(x) : n(x)
 *   w_head [64][cin][3][3] f32 (values already rounded to bf16), b_head [64] f32
 *   w_mid  n_mid x [kx][c_in / 8][ky][c_out 64][c_in % 8] bf16 (73 728 B per layer), b_mid [n_mid][64] f32
 *   w_tail [kx][c_in / 8][ky][16][c_in % 8] bf16, rows 1..15 of every ky zero;  b_tail [1] f32
 *   act0, act1: two device buffers of pnpadmm_dncnn_activation_bytes(B, H, W) bytes, 16-byte aligned.
 * pnpadmm_conv64_bf16 runs ONE 64->64 layer (w / bias as one w_mid layer) and exists for the parity tests;
 * in / out are bf16 in the kernels' inter-layer layout [B][H][8][W][8] ("chunk-planar rows": channel c of pixel
 * (y, x) at [y][c / 8][x][c % 8]), which makes every tile transfer a contiguous 1-D bulk copy.
 * ------------------------------------------------------------------------------------- */
size_t pnpadmm_dncnn_activation_bytes(int B, int H, int W);
int pnpadmm_conv64_bf16(const void* in, void* out, const void* w, const float* bias, int B, int H, int W, int relu,
                        pnpadmm_stream_t stream);
int pnpadmm_dncnn_forward_bf16(const float* x, float* out, int B, int cin, int H, int W, int n_mid,
                               const float* w_head, const float* b_head, const void* w_mid, const float* b_mid,
                               const void* w_tail, const float* b_tail, int residual, void* act0, void* act1,
                               pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a9  IRCNN denoiser forward (reference models/network_dncnn.py:70-109 IRCNN: conv3x3 1->64, five DILATED conv3x3 64->64
 * with dilation = padding = 2, 3, 4, 3, 2, conv3x3 64->1, ReLU between, forward = x - model(x); called from S3:52-62,
 * S3:280-288 with one of 25 weight sets) on the same kernels: pnpadmm_dncnn_forward_bf16 with a dilation per middle layer
 * (`dilation_mid`, n_mid host integers in 1..4; the first and last layers are undilated).  pnpadmm_conv64_dilated_bf16 is the
 * one-layer entry of the parity tests.
 * ------------------------------------------------------------------------------------- */
int pnpadmm_conv64_dilated_bf16(const void* in, void* out, const void* w, const float* bias, int B, int H, int W, int relu,
                                int dilation, pnpadmm_stream_t stream);
int pnpadmm_dncnn_forward_dilated_bf16(const float* x, float* out, int B, int cin, int H, int W, int n_mid,
                                       const int* dilation_mid, const float* w_head, const float* b_head, const void* w_mid,
                                       const float* b_mid, const void* w_tail, const float* b_tail, int residual, void* act0,
                                       void* act1, pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a9  FFDNet denoiser forward on the same tensor-core kernels (reference models/network_ffdnet.py:31-73
 * FFDNet.forward: replicate-pad to even size, PixelUnShuffle(2), concatenate the noise-level map, conv3x3 5->64 + ReLU,
 * n_mid x [conv3x3 64->64 + ReLU], conv3x3 64->4, PixelShuffle(2), crop; called from denoising_step1 S3:64-66 with
 * sigma = 15 / 255).  All convolutions run at half resolution; bf16 operands, fp32 accumulation, bias in every layer.
 *   x      [B][H][W] f32 (one channel);  out [B][H][W] f32, 8-byte aligned
 *   sigma  the noise level of the map (rounded to bf16 like the module's input)
 *   w_head a w_mid-format layer whose input channels 5..63 are zero (channel c < 4: sub-pixel (c / 2, c % 2), channel 4: map);
 *          only its first two 8-channel chunks are read (K = 16 per tap);  b_head [64] f32
 *   w_mid, b_mid as for pnpadmm_dncnn_forward_bf16;  w_tail [kx][c_in / 8][ky][16][c_in % 8] bf16 with rows 4..15 zero, b_tail [4] f32
 *   act0, act1: two device buffers of pnpadmm_dncnn_activation_bytes(B, ceil(H / 2), ceil(W / 2)) bytes, 16-byte aligned.
 * ------------------------------------------------------------------------------------- */
int pnpadmm_ffdnet_forward_bf16(const float* x, float* out, int B, int H, int W, float sigma, int n_mid, const void* w_head,
                                const float* b_head, const void* w_mid, const float* b_mid, const void* w_tail,
                                const float* b_tail, void* act0, void* act1, pnpadmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Measurement helper: runs `iters` dependent-FMA loops on every SM and returns the achieved
 * non-tensor FP32 FLOP/s in *flops (device-timed with CUDA events, synchronous).  Used by
 * bench.py as the measured denominator of the FP32 FFT roofline.
 * ------------------------------------------------------------------------------------- */
int pnpadmm_measure_fp32_peak(double* flops, pnpadmm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PNPADMM_H_ */
