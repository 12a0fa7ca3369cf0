/*
 * qiddm.h — C ABI of libqiddm_b200.so: batched state-vector simulation of the QIDDM
 * quantum layers on NVIDIA B200 (sm_100a).
 *
 * The reference (aaai2026/QIDDM) is pure Python and has no FFI: the seam this library
 * replaces is the PennyLane call `self.qnode(inputs[, weights]) -> Tensor`
 *   nn/qdense.py:58, :115, :198, :279, :465, :549, :1633   (dense families)
 *   nn/qconv.py:78-79  (the call that is missing there), :92-126 (eval-mode unitary)
 * i.e. QNode.__call__ + device (`default.qubit.torch` / `lightning.qubit`) + autograd/parameter-shift.
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions (PennyLane 0.29): wire 0 is the most significant bit of the basis
 * index; Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi); StronglyEntanglingLayers ranges
 * r_l = (l mod (n-1)) + 1 restarting at 0 in every block; ring = CNOT/CZ(i, (i+r) mod n), i = 0..n-1.
 *
 * All pointers are DEVICE pointers borrowed for the duration of the call (never freed or
 * retained).  Work is enqueued on the caller's stream; nothing synchronises.  Every entry
 * point returns 0 on success, a negative QIDDM_E* code on a bad argument, or a positive
 * cudaError_t value.  No exceptions cross the ABI.  A plan is immutable after creation and
 * may be used from several host threads as long as each call gets its own workspace.
 */
#ifndef QIDDM_H
#define QIDDM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QIDDM_ABI_VERSION 1
#define QIDDM_MAX_QUBITS 12

typedef struct CUstream_st *qiddm_stream_t; /* == cudaStream_t */

enum { QIDDM_OK = 0, QIDDM_EINVAL = -1, QIDDM_EUNSUPPORTED = -2, QIDDM_ENOMEM = -3, QIDDM_ENODEVICE = -4 };

/* initial state */
enum { QIDDM_INIT_ZERO = 0,      /* |0...0>                                                            */
       QIDDM_INIT_AMPLITUDE = 1, /* AmplitudeEmbedding(features + add_offset, pad_with, normalize)     */
       QIDDM_INIT_BASIS = 2,     /* |basis[c]> (or |c> when basis == NULL): used to build unitaries    */
       QIDDM_INIT_STATE = 3 };   /* internal (density-matrix path): complex state vectors supplied by the library */
/* per-block data gate that precedes each block's first Rot on every wire (re-upload encoding) */
enum { QIDDM_ENC_NONE = 0, QIDDM_ENC_RZ = 1, QIDDM_ENC_RY = 2 };
enum { QIDDM_IMP_CNOT = 0, QIDDM_IMP_CZ = 1 };
/* weight re-mapping applied before the angles are used (nn/qdense.py:45,:97,:171; nn/qconv.py:55) */
enum { QIDDM_REMAP_NONE = 0, QIDDM_REMAP_TANH = 1, QIDDM_REMAP_PI_TANH = 2 };
enum { QIDDM_READ_PROBS = 0,     /* out[m] = clamp(post_scale * |psi[m*read_stride]|^2), m < read_count */
       QIDDM_READ_EXPVAL_Z = 1,  /* out[j] = post_scale * <Z_j>, j < n_qubits                           */
       QIDDM_READ_STATE = 2 };   /* out = (re, im) interleaved, 2 * 2^n floats                          */
enum { QIDDM_DTYPE_F32 = 0, QIDDM_DTYPE_F64 = 1 };
enum { QIDDM_PATH_AUTO = 0, QIDDM_PATH_GATE = 1, QIDDM_PATH_GEMM = 2 };

/* One circuit "stage" = what one QNode call of the reference computes per circuit instance. */
typedef struct qiddm_circuit_desc {
    int32_t n_qubits;          /* 1..QIDDM_MAX_QUBITS                                                  */
    int32_t n_blocks;          /* L: re-upload blocks (1 for plain SEL circuits)                        */
    int32_t layers_per_block;  /* SEL depth inside a block; weights are (L, D, n, 3)                    */
    int32_t init;              /* QIDDM_INIT_*                                                          */
    int32_t n_features;        /* AMPLITUDE: real features per instance, <= 2^n                         */
    float   pad_value;         /* AMPLITUDE: pad_with                                                   */
    float   add_offset;        /* AMPLITUDE: constant added to every feature first (QConv's x + 0.1)    */
    int32_t enc;               /* QIDDM_ENC_*: gate(enc_scale * in[j]) on wire j before every block     */
    float   enc_scale;
    int32_t imprimitive;       /* QIDDM_IMP_*                                                           */
    int32_t remap;             /* QIDDM_REMAP_*                                                         */
    int32_t readout;           /* QIDDM_READ_*                                                          */
    int32_t read_count;        /* PROBS: K                                                              */
    int32_t read_stride;       /* PROBS: stride between retained probabilities                          */
    float   post_scale;
    int32_t clamp;             /* 0/1: clamp outputs to [clamp_lo, clamp_hi] (PROBS only)               */
    float   clamp_lo, clamp_hi;
    int32_t path;              /* QIDDM_PATH_*                                                          */
} qiddm_circuit_desc;

/* Optional fused patch-unfold addressing for QConv (replaces torch.nn.Unfold + einops,
 * nn/qconv.py:76-86): instance c = (b, y, x); feature f = (ch, ky, kx) reads
 * img[b, ch, y+ky-pad_h, x+kx-pad_w] (0 outside); outputs go to out[b, m, y, x]. */
typedef struct qiddm_unfold_desc {
    int32_t channels, height, width;   /* input image (NCHW)        */
    int32_t kernel_h, kernel_w, pad_h, pad_w;
} qiddm_unfold_desc;

typedef struct qiddm_plan qiddm_plan;

int qiddm_abi_version(void);
const char *qiddm_error_string(int code);

/* Number of inputs / outputs per circuit instance for a descriptor (0 on invalid descriptors). */
int qiddm_n_inputs(const qiddm_circuit_desc *desc);
int qiddm_n_outputs(const qiddm_circuit_desc *desc);
int qiddm_n_weights(const qiddm_circuit_desc *desc); /* L*D*n*3 */

int  qiddm_plan_create(const qiddm_circuit_desc *desc, qiddm_plan **plan);
void qiddm_plan_destroy(qiddm_plan *plan);

/* Bytes of scratch a forward/backward call on `batch` instances needs (256-byte aligned). */
size_t qiddm_workspace_bytes(const qiddm_plan *plan, int64_t batch);

/* out[batch, n_outputs] = circuit(in[batch, n_inputs]; weights).  `in` may be NULL when the
 * descriptor has no inputs; `basis` (int32[batch]) only for QIDDM_INIT_BASIS (may be NULL). */
int qiddm_forward(const qiddm_plan *plan, const float *in, const int32_t *basis, const void *weights,
                  int weights_dtype, float *out, void *workspace, int64_t batch, qiddm_stream_t stream);

/* Adjoint-method backward.  grad_in (batch, n_inputs) may be NULL; grad_weights has the shape and
 * dtype of `weights` and is OVERWRITTEN with the batch-summed gradient (may be NULL). */
int qiddm_backward(const qiddm_plan *plan, const float *in, const int32_t *basis, const void *weights,
                   int weights_dtype, const float *grad_out, float *grad_in, void *grad_weights,
                   void *workspace, int64_t batch, qiddm_stream_t stream);

/* The same pair with the final states kept between them: `state` (qiddm_state_bytes(plan, batch) = batch * 2^n complex fp32)
 * is WRITTEN by the forward and READ by the backward, whose adjoint sweep then starts from psi_final instead of recomputing
 * the forward sweep (20-25 % of its time).  `state` may be NULL (= the plain calls). */
size_t qiddm_state_bytes(const qiddm_plan *plan, int64_t batch);
int qiddm_forward_save(const qiddm_plan *plan, const float *in, const int32_t *basis, const void *weights,
                       int weights_dtype, float *out, float *state, void *workspace, int64_t batch, qiddm_stream_t stream);
int qiddm_backward_saved(const qiddm_plan *plan, const float *in, const int32_t *basis, const void *weights,
                         int weights_dtype, const float *grad_out, const float *state, float *grad_in, void *grad_weights,
                         void *workspace, int64_t batch, qiddm_stream_t stream);

/* QConv: same circuit with the patch-unfold fused.  img (n_images, C, H, W) fp32; out
 * (n_images, read_count, H_out, W_out) fp32; grad_img is OVERWRITTEN (col2im accumulated inside). */
int qiddm_qconv_forward(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, const float *img,
                        const void *weights, int weights_dtype, float *out, void *workspace,
                        int64_t n_images, qiddm_stream_t stream);
int qiddm_qconv_backward(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, const float *img,
                         const void *weights, int weights_dtype, const float *grad_out, float *grad_img,
                         void *grad_weights, void *workspace, int64_t n_images, qiddm_stream_t stream);

/* The same with img / out / grad tensors of io_dtype (QIDDM_DTYPE_F32, or QIDDM_DTYPE_F64 as the reference's float64 UNet
 * passes them: read and written in place of cast kernels; the simulation itself is fp32). */
int qiddm_qconv_forward_io(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, int io_dtype, const void *img,
                           const void *weights, int weights_dtype, void *out, void *workspace, int64_t n_images,
                           qiddm_stream_t stream);
int qiddm_qconv_backward_io(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, int io_dtype, const void *img,
                            const void *weights, int weights_dtype, const void *grad_out, void *grad_img, void *grad_weights,
                            void *workspace, int64_t n_images, qiddm_stream_t stream);

/* Collapse the weight-only part of a circuit into its 2^n x 2^n unitary — the eval-mode matrix
 * of nn/qconv.py:92-126.  Stored TRANSPOSED (row c = U|c>, i.e. unitary[c][k] = U[k][c]),
 * interleaved re/im fp32, 2 * 4^n floats. */
int qiddm_build_unitary(const qiddm_plan *plan, const void *weights, int weights_dtype, float *unitary,
                        void *workspace, qiddm_stream_t stream);

/* ---- Unitary-collapse (tensor-core GEMM) path for amplitude-embedding circuits with a probability
 * readout (QDenseUndirected_old[_noise], QConv2d rows): generalises the eval-mode collapse of
 * nn/qconv.py:92-126 to training.  `prepare` collapses the circuit for the current weights into
 * `collapsed` (qiddm_gemm_collapsed_bytes; call again whenever the weights change); forward/backward
 * then run every instance as one row of a tcgen05 GEMM.  precision: 3 = fp32-grade (3-term fp16 split),
 * 1 = single fp16 pass (about 1e-3 relative).  The backward takes its own precision: 1 behind a precision-3 forward runs
 * the dX / dW GEMMs single-pass on the fp32-grade saved state (outputs unchanged, gradients to about 5e-4 of their
 * maximum).  Workspace: qiddm_gemm_workspace_bytes(plan, batch). */
int    qiddm_gemm_supported(const qiddm_plan *plan);
size_t qiddm_gemm_collapsed_bytes(const qiddm_plan *plan);
size_t qiddm_gemm_workspace_bytes(const qiddm_plan *plan, int64_t batch);
int qiddm_gemm_prepare(const qiddm_plan *plan, const void *weights, int weights_dtype, void *collapsed,
                       void *workspace, qiddm_stream_t stream);
/* `saved` (qiddm_gemm_saved_bytes) is optional: when the forward is given a buffer it keeps the fp16 operand
 * splits and Y there and the backward reuses them (no re-materialisation GEMM); pass NULL for inference, and
 * NULL to the backward to have it recompute them. */
size_t qiddm_gemm_saved_bytes(const qiddm_plan *plan, int64_t batch);
/* workspace of an inference forward (saved == NULL): the operand splits and norms live there */
size_t qiddm_gemm_forward_workspace_bytes(const qiddm_plan *plan, int64_t batch);
int qiddm_gemm_forward(const qiddm_plan *plan, const void *collapsed, const float *in, float *out, void *saved,
                       void *workspace, int64_t batch, int precision, qiddm_stream_t stream);
int qiddm_gemm_backward(const qiddm_plan *plan, const void *collapsed, const float *in, const void *weights,
                        int weights_dtype, const float *grad_out, const void *saved, float *grad_in,
                        void *grad_weights, void *workspace, int64_t batch, int precision, qiddm_stream_t stream);

/* Fused diffusion TRAINING STEP of a single amplitude-embedding layer on the unitary-collapse path: what
 * `Diffusion.run_training_step_data / _noise` (src/models.py:44-104) do around `QDenseUndirected_old[_noise].forward`
 * (nn/qdense.py:56-66, :95-111) with `add_normal_noise_multiple` (src/noise.py:105-126) and MSELoss -- noise ladder ->
 * circuit -> loss -> d loss / d weights -- in four launches with no (rows x pixels) fp32 intermediate in HBM:
 *   rows (b, t), t < T:  in = level_{t+1}(b),  out = layer(in),  d = a out + b - (c0 level_t(b) + c1 level_{t+1}(b)),
 *   level_k(b) = clamp(x[b] (1 - w[k]) + eps[b] w[k], 0, 1),  loss = mean(d^2)  (goal "data": a 1, b 0, c0 1, c1 0;
 *   goal "noise": a 0.1, b -0.05, c0 -1, c1 1),  grad_weights = d loss / d weights (written, not accumulated).
 * x: (n_images, n_features) float32 / float64 (io_dtype), eps: (n_images, n_features) float32, w: (T + 1) level weights in
 * io_dtype, loss: one io_dtype scalar on the device.  Needs read_count == n_features (the layer reconstructs its input).
 * `collapsed` from qiddm_gemm_prepare for the same weights.  workspace: qiddm_dense_mse_step_workspace_bytes. */
size_t qiddm_dense_mse_step_workspace_bytes(const qiddm_plan *plan, int64_t n_images, int T);
int qiddm_dense_mse_step(const qiddm_plan *plan, const void *collapsed, const void *x, const float *eps, const void *w,
                         int io_dtype, int64_t n_images, int T, double a, double b, double c0, double c1,
                         const void *weights, int weights_dtype, void *loss, void *grad_weights, void *workspace,
                         int precision, int bwd_precision, qiddm_stream_t stream);

/* QConv on the unitary-collapse path (replaces torch.nn.Unfold + einops + the missing QNode call of
 * nn/qconv.py:76-86 and, in eval mode, the QubitUnitary path of nn/qconv.py:92-126): `collapsed` comes from
 * qiddm_gemm_prepare; img (n_images, C, H, W) and out / grad_out (n_images, read_count, H_out, W_out) are NCHW tensors
 * of io_dtype (QIDDM_DTYPE_F32, or QIDDM_DTYPE_F64 as the reference's float64 UNet passes them — read and written in
 * place of a cast; the simulation itself is fp32); the patch-unfold is fused into the operand preparation and the col2im of the image gradient runs as one
 * gather kernel (grad_img is OVERWRITTEN; may be NULL). */
/* 1 when the layer runs as a direct fp32 convolution behind the qconv_gemm entry points (csrc/qiddm_conv.cu): 1x1 / 3x3 windows
 * with "same" padding and at most 16 output channels (N = 2 out_channels <= 32 rows of U).  The staged image band and the N
 * accumulators of a patch stay on chip: no patch matrix, no fp16 operand splits; results are plain fp32 (QIDDM_QCONV_DIRECT=0
 * turns it off).  The backward then needs the `saved` buffer of its forward. */
int qiddm_qconv_direct_supported(const qiddm_plan *plan, const qiddm_unfold_desc *unfold);
/* `qiddm_gemm_prepare` for a layer whose every call takes the direct convolution: the collapse (U^T) and its fp32 filter rows,
 * without the fp16 GEMM operands.  The buffer has the size and layout of qiddm_gemm_collapsed_bytes; the GEMM entry points must
 * not be used with it. */
int qiddm_gemm_prepare_direct(const qiddm_plan *plan, const void *weights, int weights_dtype, void *collapsed,
                              void *workspace, qiddm_stream_t stream);
size_t qiddm_qconv_gemm_saved_bytes(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, int64_t n_images);
size_t qiddm_qconv_gemm_workspace_bytes(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, int64_t n_images);
int qiddm_qconv_gemm_forward(const qiddm_plan *plan, const void *collapsed, const qiddm_unfold_desc *unfold,
                             int io_dtype, const void *img, void *out, void *saved, void *workspace, int64_t n_images,
                             int precision, qiddm_stream_t stream);
int qiddm_qconv_gemm_backward(const qiddm_plan *plan, const void *collapsed, const qiddm_unfold_desc *unfold,
                              int io_dtype, const void *img, const void *weights, int weights_dtype, const void *grad_out,
                              const void *saved, void *grad_img, void *grad_weights, void *workspace,
                              int64_t n_images, int precision, qiddm_stream_t stream);

/* Bilinear Upsample (align_corners = False) followed by a 1 x 1 QConv2d -- `UpBlock.up_conv` of nn/unet.py:36-41 -- in one pass:
 * `img_src` is the (n_images, C, h_in, w_in) tensor in front of the upsample, `unfold` the geometry of the 1 x 1 convolution on the
 * upsampled (height, width) image, scale_h / scale_w the source-coordinate scales (1 / scale_factor, or in / out for size=).  The
 * interpolation runs inside the staging of the direct-convolution kernels: the upsampled tensor never exists.  Direct-convolution
 * layers only (qiddm_qconv_direct_supported, kernel 1 x 1; else QIDDM_EUNSUPPORTED); `collapsed` from qiddm_gemm_prepare[_direct],
 * `saved` / `workspace` sized by qiddm_qconv_gemm_saved_bytes / _workspace_bytes for `unfold`.  The backward writes grad_up, the
 * gradient w.r.t. the UPSAMPLED image (n_images, C, height, width; may be NULL): qiddm_upsample_bilinear_backward maps it to the
 * source. */
int qiddm_qconv_up_forward(const qiddm_plan *plan, const void *collapsed, const qiddm_unfold_desc *unfold, int io_dtype,
                           const void *img_src, int h_in, int w_in, double scale_h, double scale_w, void *out, void *saved,
                           int64_t n_images, qiddm_stream_t stream);
int qiddm_qconv_up_backward(const qiddm_plan *plan, const void *collapsed, const qiddm_unfold_desc *unfold, int io_dtype,
                            const void *img_src, int h_in, int w_in, double scale_h, double scale_w, const void *weights,
                            int weights_dtype, const void *grad_out, const void *saved, void *grad_up, void *grad_weights,
                            void *workspace, int64_t n_images, qiddm_stream_t stream);

/* Compatibility: the forward `_QConv2d_FAST` LITERALLY executes (nn/qconv.py:71-90 -- the QNode call is missing there, SURVEY.md
 * H1): out[b, j, y, x] = clamp((patch feature 2j + 0.1) * F * 0.5, 0, 1), F = C * kernel_h * kernel_w, feature = (ch, ky, kx)
 * in torch.nn.Unfold order, zero padding; out_channels = min(module out_channels, ceil(F / 2)).  img / out / grad tensors
 * NCHW of `dtype`; the backward OVERWRITES grad_img.  Every checkpoint the reference saved for a QConv network was trained
 * through this map (its circuit weights never received a gradient). */
int qiddm_qconv_reference_map_forward(const qiddm_unfold_desc *unfold, int dtype, const void *img, void *out, int out_channels,
                                      int64_t n_images, qiddm_stream_t stream);
int qiddm_qconv_reference_map_backward(const qiddm_unfold_desc *unfold, int dtype, const void *img, const void *grad_out,
                                       void *grad_img, int out_channels, int64_t n_images, qiddm_stream_t stream);

/* Optional per-kernel timing for roofline reports: when enabled, CUDA events are recorded on the
 * launching stream around each main kernel.  collect() synchronises on them and returns, per kind
 * (0 gate forward, 1 gate adjoint backward, 2 tcgen05 GEMM (all), 3 other, 4/5/6 GEMM forward / dX / dW,
 * 7 prep_x, 8 transpose_x, 9 g_bound, 10 grad_y, 11 finish_dx, 12 assemble, 13 build_w, 14 / 15 direct QConv forward / backward
 * kernels), the summed
 * milliseconds, the summed algorithmic work (flops) and the launch count, then clears the record.
 * Arrays of QIDDM_TIMING_KINDS = 16. */
#define QIDDM_TIMING_KINDS 16
void qiddm_timing_enable(int enable);
int  qiddm_timing_collect(double *ms_by_kind, double *work_by_kind, int64_t *launches_by_kind);

/* UNet glue around QConv2d (reference nn/unet.py:28-116: torch.nn.Upsample(scale_factor=2, mode="bilinear") and
 * torch.nn.BatchNorm2d in float64).  dtype = QIDDM_DTYPE_F32 / F64 for all tensors of a call; NCHW, contiguous.
 * Bilinear resize with align_corners = False: in (planes, h_in, w_in) -> out (planes, h_out, w_out), planes = N * C,
 * scale_* = the source step per output pixel (1 / scale_factor when a scale factor was given, else in / out); backward
 * is the exact transpose in gather form (grad_in OVERWRITTEN).
 * BatchNorm2d with batch statistics: y = (x - mean_c) * rstd_c * gamma_c + beta_c over (n, hw) per channel; save_mean /
 * save_rstd (float64[c]) are kept for the backward; running_* (nullable, tensor dtype) are updated with `momentum` and the
 * unbiased variance as torch does; gamma / beta nullable.  backward: grad_x nullable, grad_gamma / grad_beta nullable,
 * OVERWRITTEN.  Deterministic (two-stage reductions, no atomics).  workspace: qiddm_batchnorm_workspace_bytes(c). */
int qiddm_upsample_bilinear_forward(const void *in, void *out, int dtype, int64_t planes, int h_in, int w_in, int h_out,
                                    int w_out, double scale_h, double scale_w, qiddm_stream_t stream);
int qiddm_upsample_bilinear_backward(const void *grad_out, void *grad_in, int dtype, int64_t planes, int h_in, int w_in,
                                     int h_out, int w_out, double scale_h, double scale_w, qiddm_stream_t stream);
size_t qiddm_batchnorm_workspace_bytes(int channels);
int qiddm_batchnorm_forward(const void *x, void *y, int dtype, int n, int c, int hw, const void *gamma, const void *beta,
                            double *save_mean, double *save_rstd, void *running_mean, void *running_var, double momentum,
                            double eps, void *workspace, qiddm_stream_t stream);
int qiddm_batchnorm_backward(const void *x, const void *grad_y, void *grad_x, int dtype, int n, int c, int hw, const void *gamma,
                             const double *save_mean, const double *save_rstd, void *grad_gamma, void *grad_beta,
                             void *workspace, qiddm_stream_t stream);

/* The same BatchNorm2d with the neighbouring ReLU of the UNet blocks fused in (nn/unet.py:92-108: Conv -> BN -> ReLU; :55-63:
 * Conv -> ReLU -> BN): relu_mode 0 = none, 1 = y = relu(bn(x)) (the backward masks grad_y where bn(x) <= 0, so it takes beta),
 * 2 = y = bn(relu(x)) (statistics of relu(x); the backward zeroes grad_x where x <= 0).  One pass less over the activation in
 * each direction per fused pair.  MaxPool2d(kernel, stride = kernel, no padding) of nn/unet.py:110 on (planes, h, w): the
 * backward recomputes the window's first maximum instead of storing indices (grad_x OVERWRITTEN). */
int qiddm_batchnorm_relu_forward(const void *x, void *y, int dtype, int n, int c, int hw, const void *gamma, const void *beta,
                                 double *save_mean, double *save_rstd, void *running_mean, void *running_var, double momentum,
                                 double eps, int relu_mode, void *workspace, qiddm_stream_t stream);
int qiddm_batchnorm_relu_backward(const void *x, const void *grad_y, void *grad_x, int dtype, int n, int c, int hw,
                                  const void *gamma, const void *beta, const double *save_mean, const double *save_rstd,
                                  void *grad_gamma, void *grad_beta, int relu_mode, void *workspace, qiddm_stream_t stream);
int qiddm_maxpool2d_forward(const void *x, void *y, int dtype, int64_t planes, int h, int w, int kernel, qiddm_stream_t stream);
int qiddm_maxpool2d_backward(const void *x, const void *grad_y, void *grad_x, int dtype, int64_t planes, int h, int w, int kernel,
                             qiddm_stream_t stream);

/* Linear layers with one narrow side (min(in, out) <= 16) next to the circuits: `linear_down` (pixels -> qubits) and
 * `linear_up` (qubits -> pixels) of nn/qdense.py:219-386, :565-670 (torch.nn.Linear there, float64).  y = x W^T + b with
 * x (rows, in), W (out, in) row-major, b (out) or NULL, as streaming kernels in the tensors' dtype; the backward writes
 * grad_x (rows, in), grad_weight (out, in), grad_bias (out) -- each may be NULL (grad_bias needs grad_weight) -- with
 * deterministic two-stage sums.  workspace (backward): qiddm_skinny_linear_workspace_bytes. */
size_t qiddm_skinny_linear_workspace_bytes(int64_t rows, int in_features, int out_features);
int qiddm_skinny_linear_forward(const void *x, const void *weight, const void *bias, void *y, int dtype, int64_t rows,
                                int in_features, int out_features, qiddm_stream_t stream);
int qiddm_skinny_linear_backward(const void *x, const void *weight, const void *grad_y, void *grad_x, void *grad_weight,
                                 void *grad_bias, int dtype, int64_t rows, int in_features, int out_features, void *workspace,
                                 qiddm_stream_t stream);

/* Diffusion-step glue.  qiddm_noise_ladder = src/noise.py:105-126 (`add_normal_noise_multiple`) fused with the slicing of
 * src/models.py:50-63: for x, eps (batch, pixels) (eps float32 as the reference draws it) and the level weights w[tau]
 * (tensor dtype), level_t = clamp(x (1 - w_t) + eps w_t, 0, 1); writes noisy[(b, t)] = level_{t+1} and clean[(b, t)] =
 * level_t for t < tau - 1, both (batch * (tau - 1), pixels); clean may be NULL.  qiddm_mse_loss_grad = MSELoss + `.mean().backward()` seed
 * (src/models.py:65-67, :95-99): d = scale * pred + shift - target (+ target_add when non-NULL); loss[0] = mean(d^2),
 * grad = 2 scale d / n (deterministic two-stage sum).  workspace: qiddm_mse_workspace_bytes(). */
int qiddm_noise_ladder(const void *x, const float *eps, const void *w, int dtype, int64_t batch, int pixels, int tau,
                       void *noisy, void *clean, qiddm_stream_t stream);
size_t qiddm_mse_workspace_bytes(void);
int qiddm_mse_loss_grad(const void *pred, const void *target, const void *target_add, int dtype, double scale, double shift,
                        int64_t n, void *grad, void *loss, void *workspace, qiddm_stream_t stream);
/* The same loss with the target recomputed from the image and its noise draw (`clean` of qiddm_noise_ladder may then be NULL):
 * row (b, t) of pred (batch * (tau - 1), pixels): d = scale * pred + shift - (c0 * level_t + c1 * level_{t+1}); goal "data"
 * (src/models.py:65-67): c0 = 1, c1 = 0; goal "noise" (:95-99): scale 0.1, shift -0.05, c0 = -1, c1 = 1.  The pass reads pred and
 * (x, eps) once per tau - 1 rows instead of pred and one or two ladder tensors. */
int qiddm_mse_ladder_loss_grad(const void *pred, const void *x, const float *eps, const void *w, int dtype, int64_t batch, int pixels,
                               int tau, double scale, double shift, double c0, double c1, void *grad, void *loss, void *workspace,
                               qiddm_stream_t stream);

/* Tail of the re-upload families' training step in one pass: `linear_up` (nn/qdense.py:642, :676: Linear(hidden -> pixels)) +
 * MSELoss + `.mean().backward()` (src/models.py:65-67, :95-99) with the target recomputed from the noise draw as in
 * qiddm_mse_ladder_loss_grad.  h (batch * (tau - 1), hidden <= 16), weight (pixels, hidden), bias (pixels, may be NULL), x / eps
 * (batch, pixels), w[tau], tau <= 33; writes loss[0], grad_weight (pixels, hidden), grad_bias (pixels, may be NULL) and grad_h (like
 * h).  Neither the layer's output nor its gradient (batch * (tau - 1) x pixels each) is materialised; the products are expanded
 * (second moments of weight and h) so that only K + 4 and K + 8 FP64 operations touch every (row, pixel).  Deterministic. */
size_t qiddm_linear_up_mse_workspace_bytes(int pixels, int hidden);
int qiddm_linear_up_mse_step(const void *h, const void *weight, const void *bias, const void *x, const float *eps, const void *w,
                             int dtype, int64_t batch, int pixels, int tau, int hidden, double scale, double shift, double c0,
                             double c1, void *loss, void *grad_weight, void *grad_bias, void *grad_h, void *workspace,
                             qiddm_stream_t stream);

/* Noise channels right before a probability readout (the `add_noise` branch of nn/qdense.py:98-104, :174-180, :431-439,
 * run on `default.mixed` by src/mnist_noise.py:211-229): the same single-qubit channel on every wire immediately before
 * probs() acts on the probability vector as p' = (M x ... x M) p with one 2 x 2 column-stochastic M = [[m00, m01], [m10,
 * m11]] per wire -- AmplitudeDamping(g): [[1, g], [0, 1 - g]]; DepolarizingChannel(q): [[1 - 2q/3, 2q/3], [2q/3, 1 - 2q/3]];
 * PhaseShift / PhaseDamping: identity.  probs_in / probs_out (batch, 2^n), tensor dtype; the backward is the same call
 * with M transposed.  probs_in == probs_out is allowed. */
int qiddm_readout_channel(const void *probs_in, void *probs_out, int dtype, int64_t batch, int n_qubits, double m00, double m01,
                          double m10, double m11, qiddm_stream_t stream);

/* Mid-circuit noise channels of the re-upload classes (the `add_noise` branches of nn/qdense.py:515-527, :1405-1417,
 * :1599-1617, which src/mnist_noise.py:211-229 evaluates on `default.mixed` with the flag flipped on a TRAINED net: inference
 * only, no backward).  Descriptor: QIDDM_INIT_ZERO, QIDDM_ENC_RZ, readout PROBS or EXPVAL_Z.  After the RZ(a_j) of every block
 * each wire passes a single-qubit channel given in its generic form on the wire's 2 x 2 block: the populations mix as
 * (rho00, rho11) -> [[m00, m01], [m10, m11]] (rho00, rho11), the coherences scale by f_off.  PhaseDamping(g): f_off =
 * sqrt(1 - g), M = I; AmplitudeDamping(g): f_off = sqrt(1 - g), M = [[1, g], [0, 1 - g]]; DepolarizingChannel(p): f_off =
 * 1 - 4p/3, M = [[1 - 2p/3, 2p/3], [2p/3, 1 - 2p/3]].  Full density-matrix simulation ((batch, 2^n, 2^n) complex fp32 in the
 * workspace, the unitary layers through the gate kernels on its rows); in (batch, n) angles, out (batch, n_outputs). */
size_t qiddm_noisy_workspace_bytes(const qiddm_plan *plan, int64_t batch);
int qiddm_noisy_forward(const qiddm_plan *plan, const float *in, const void *weights, int weights_dtype, double f_off,
                        double m00, double m01, double m10, double m11, float *out, void *workspace, int64_t batch,
                        qiddm_stream_t stream);

/* On-device PCA support (replaces the sklearn `PCA.fit_transform` host round trip of nn/qdense.py:456, :1429):
 * eigen-decomposition of a symmetric m x m float64 matrix (the Gram matrix of the centred batch rows), one CTA, parallel
 * cyclic Jacobi.  evals[m] in DESCENDING order, evecs (m x m row-major) column j = eigenvector of evals[j].
 * m <= qiddm_sym_eigh_max_dim(); asynchronous on `stream`, no status read-back (CUDA-graph capturable). */
int qiddm_sym_eigh_max_dim(void);
int qiddm_sym_eigh_f64(const double *a, int m, double *evals, double *evecs, qiddm_stream_t stream);
/* `count` independent matrices, contiguous (count, m, m) -> evals (count, m), evecs (count, m, m); one CTA each.  Used for
 * per-group PCA: a batch of N images x tau noise levels keeps the reference's batch-1 semantics (one PCA per image's
 * tau-ladder, nn/qdense.py:1429 with src/mnist_exm.py:144) as N groups of tau rows in one launch. */
int qiddm_sym_eigh_f64_batched(const double *a, int m, int64_t count, double *evals, double *evecs, qiddm_stream_t stream);

/* Id of the CUDA-graph capture `stream` is currently part of, 0 when it is not capturing (lets the host side keep
 * per-capture caches of the collapsed operator). */
int64_t qiddm_stream_capture_id(qiddm_stream_t stream);

/* Measurement support: one launch of a register-only packed-FMA loop (8 independent fma.rn.f32x2 chains per thread,
 * 8 CTAs of 256 threads per SM, `iters` rounds of 64 flop per thread); *flops (host, nullable) = flops the launch executes.
 * The caller times it with events on `stream`: the FP32-pipe roofline denominator of the gate kernels (SURVEY.md 8d:
 * MEASURED_PEAKS.json holds no FP32 figure).  `sink` = any 4-byte device buffer (never written). */
int qiddm_probe_fp32_fma(int iters, float *sink, double *flops, qiddm_stream_t stream);

/* Kernel launches enqueued by this library since load (for bench.py's gpu_launches). */
int64_t qiddm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QIDDM_H */
