// Internal declarations shared by the translation units of libqiddm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "qiddm.h"

namespace qiddm {

// Everything a gate-path kernel needs, passed by value.
struct GateParams {
    // circuit
    int n_blocks, layers, init, n_features, enc, imprimitive, readout, read_count, read_stride, clamp;
    float pad_value, add_offset, enc_scale, post_scale, clamp_lo, clamp_hi;
    int n_rot;                 // n_blocks * layers * n_qubits
    int merge_post;            // 1: RZ(omega) phases of layer l are merged into layer l+1's RZ(phi) table (CZ entangler)
    // fused patch-unfold (QConv); unfold == 0 -> plain (B, n_in) rows
    int unfold, C, H, W, kh, kw, ph, pw, Hout, Wout;
    int io64;                  // QConv collapse path: image / output / gradient tensors are float64 (else float32)
    long long B;               // circuit instances
    const float *in;           // (B, n_in) features / angles, or NCHW image when unfold
    const int *basis;          // INIT_BASIS start states (may be null -> instance index)
    const float *gates;        // per-layer phase tables + rotation coefficients (prepare_tables_kernel)
    float *out;
    // backward only
    const float *grad_out;
    float *grad_in;            // nullable
    float *partials;           // [grid][n_rot*3] per-CTA sums of the angle gradients (phi, theta, omega)
    // optional (B, 2^n) complex fp32 final states psi_final (before the readout): the forward kernel WRITES them, the adjoint
    // kernel READS them instead of recomputing the forward sweep (a fifth to a quarter of its time)
    float *state;
    // QIDDM_INIT_STATE (density-matrix path): instance cid starts from init_state[cid] (2^n complex fp32); its re-upload
    // angles are row (cid >> in_shift) of `in` (the 2^n columns of one density matrix share their sample's angles)
    const float *init_state;
    int in_shift;
};

struct LaunchInfo {
    int grid, block;
    size_t smem;
};

// qiddm_gate.cu
int gate_rb(int n_qubits, bool backward);
cudaError_t gate_launch_info(int n_qubits, bool backward, const GateParams &p, LaunchInfo *info);
cudaError_t launch_gate_forward(int n_qubits, const GateParams &p, const LaunchInfo &li, cudaStream_t s);
cudaError_t launch_gate_backward(int n_qubits, const GateParams &p, const LaunchInfo &li, cudaStream_t s);
bool gate_state_compatible(int n_qubits, const GateParams &p);   // forward and adjoint kernels agree on psi_final's phase convention
size_t gate_table_bytes(int n_qubits, int n_layers);
size_t gate_partial_floats(int n_qubits, int n_layers);      // per-CTA angle-gradient sums (upper bound over schedules)
cudaError_t launch_prepare_tables(const void *weights, int wdtype, int remap, int n_qubits, bool backward,
                                  const GateParams &p, float *tables, cudaStream_t s);
cudaError_t launch_finalize_grads(const float *partials, int n_partials, const void *weights, int wdtype, int remap,
                                  int n_qubits, const GateParams &p, void *grad_weights, cudaStream_t s);
void count_launch(int n = 1);

// Optional per-kernel timing (bench.py's roofline): CUDA events recorded on the launching stream around
// the launches of one kind; off by default.
enum { TK_GATE_FWD = 0, TK_GATE_BWD = 1, TK_GEMM = 2, TK_OTHER = 3, TK_GEMM_FWD = 4, TK_GEMM_DX = 5, TK_GEMM_DW = 6,
       TK_PREP_X = 7, TK_TRANSPOSE_X = 8, TK_G_BOUND = 9, TK_GRAD_Y = 10, TK_FINISH_DX = 11, TK_ASSEMBLE = 12,
       TK_BUILD_W = 13, TK_CONV_FWD = 14, TK_CONV_BWD = 15, TK_COUNT = 16 };
void timing_begin(int kind, double work, cudaStream_t s);   // work = algorithmic flops (or bytes) of the launch
void timing_end(cudaStream_t s);
void timing_set_gemm_kind(int kind);

// qiddm_gemm.cu — unitary-collapse path (amplitude families)
struct GemmShape {
    int A, F, Fx, Kp, n_out, N, Np, stride;
    float w_scale;
};
GemmShape gemm_shape(const GateParams &gp, int n_qubits);
size_t gemm_collapsed_bytes(const GemmShape &g);
float *gemm_collapsed_ut(const GemmShape &g, void *collapsed);
float *gemm_collapsed_wd(const GemmShape &g, void *collapsed);      // fp32 rows of U for the direct QConv path (qiddm_conv.cu)
int gemm_build_operands(const GemmShape &g, const GateParams &gp, void *collapsed, cudaStream_t s);
size_t gemm_saved_bytes(const GemmShape &g, long long B);
size_t gemm_forward_ws_bytes(const GemmShape &g, long long B);
size_t gemm_backward_ws_bytes(const GemmShape &g, long long B, bool unfold = false);
int gemm_forward(const GemmShape &g, const GateParams &gp, const void *collapsed, const float *x, float *out,
                 void *saved, void *ws, long long B, int n_seg, cudaStream_t s);
int gemm_backward(const GemmShape &g, const GateParams &gp, const void *collapsed, const float *x,
                  const float *grad_out, const void *saved, float *grad_in, float **gut_out, void *ws, long long B,
                  int n_seg, cudaStream_t s);
// fused diffusion step of one amplitude-embedding layer (ladder -> splits, forward GEMM with MSE + dL/dY epilogue, dW GEMM)
size_t gemm_dense_mse_ws_bytes(const GemmShape &g, long long B);
int gemm_dense_mse_step(const GemmShape &g, const GateParams &gp, const void *collapsed, const void *x, const float *eps,
                        const void *w, int io64, long long n_img, int T, float a, float bshift, float c0, float c1,
                        void *loss_out, float **gut_out, void *ws, int n_seg_fwd, int n_seg_bwd, cudaStream_t s);

// qiddm_conv.cu — QConv2d on the collapse path as a direct fp32 convolution (N = 2 out_channels <= 32, 1x1 / 3x3 "same" windows)
int conv_np(int N);
size_t conv_wd_bytes(const GemmShape &g);
int conv_build_wd(const GemmShape &g, const GateParams &gp, const float *UT, float *Wd, cudaStream_t s);
bool conv_direct_supported(const GemmShape &g, const GateParams &gp);
size_t conv_direct_saved_bytes(const GemmShape &g, const GateParams &gp, long long n_images);
size_t conv_direct_ws_bytes(const GemmShape &g, const GateParams &gp, long long n_images);
// `up` (1 x 1 windows only): `img` is the (n, C, h_in, w_in) SOURCE of a bilinear upsample to the unfold geometry's (H, W), the
// interpolation runs inside the staging of the kernels (nn/unet.py:36-41: Upsample -> 1 x 1 Conv2d); grad_img is then the gradient
// w.r.t. the UPSAMPLED image (the caller applies the transpose of the interpolation)
struct ConvUp {
    int h_in, w_in;
    double scale_h, scale_w;
};
int conv_direct_forward(const GemmShape &g, const GateParams &gp, const float *Wd, const void *img, void *out, void *saved,
                        long long n_images, cudaStream_t s, const ConvUp *up = nullptr);
int conv_direct_backward(const GemmShape &g, const GateParams &gp, const float *Wd, const void *img, const void *grad_out,
                         const void *saved, void *grad_img, float **gut_out, void *ws, long long n_images, cudaStream_t s,
                         const ConvUp *up = nullptr);

// qiddm_dm.cu — density-matrix pieces for the mid-circuit noise channels (tau = rho^T, (B, 2^n, 2^n) complex fp32)
size_t dm_state_bytes(int n_qubits, long long B);
int dm_init(float2 *tau, int n, long long B, cudaStream_t s);
int dm_channel(float2 *tau, int n, long long B, float f_off, float m00, float m01, float m10, float m11, cudaStream_t s);
int dm_transpose_conj(const float2 *src, float2 *dst, int n, long long B, cudaStream_t s);
int dm_readout(const float2 *tau, int n, long long B, int readout, int read_count, int read_stride, float post_scale, int clamp,
               float lo, float hi, float *out, cudaStream_t s);

// qiddm_pca.cu — single-CTA Jacobi eigensolver (float64) for the on-device PCA-in-forward
size_t eigh_smem_bytes(int m);
int sym_eigh_f64(const double *A, int m, long long count, double *evals, double *evecs, cudaStream_t s);

// qiddm_glue.cu — UNet glue around QConv2d: bilinear resize (align_corners = False) and BatchNorm2d (NCHW)
size_t batchnorm_ws_bytes(int C);
size_t mse_ws_bytes();
int qconv_reference_map(const void *img, const void *grad_out, void *out, int dtype, bool backward, long long n_images, int C,
                        int H, int W, int kh, int kw, int ph, int pw, int n_ch_out, cudaStream_t s);
int probe_fp32_fma(int iters, float *sink, double *flops, cudaStream_t s);
int prob_channel(const void *p_in, void *p_out, int dtype, long long batch, int n, double m00, double m01, double m10, double m11,
                 cudaStream_t s);
int noise_ladder(const void *x, const float *eps, const void *w, int dtype, long long batch, int P, int tau, void *noisy,
                 void *clean, cudaStream_t s);
int mse_loss_grad(const void *r, const void *t1, const void *t2, int dtype, double a, double b, long long n, void *grad,
                  void *loss, void *ws, cudaStream_t s);
size_t linear_up_mse_ws_bytes(int P, int K);
int linear_up_mse_step(const void *h, const void *W, const void *bias, const void *x, const float *eps, const void *w, int dtype,
                       long long batch, int P, int tau, int K, double a, double b, double c0, double c1, void *loss, void *dW,
                       void *dbias, void *dh, void *ws, cudaStream_t s);
int mse_ladder_loss_grad(const void *r, const void *x, const float *eps, const void *w, int dtype, long long batch, int P, int tau,
                         double a, double b, double c0, double c1, void *grad, void *loss, void *ws, cudaStream_t s);
int upsample_bilinear(const void *in, void *out, int dtype, bool backward, long long planes, int Hin, int Win, int Hout, int Wout,
                      double scale_h, double scale_w, cudaStream_t s);
int batchnorm_forward(const void *x, void *y, int dtype, int N, int C, int HW, const void *gamma, const void *beta,
                      double *save_mean, double *save_rstd, void *running_mean, void *running_var, double momentum, double eps,
                      void *ws, cudaStream_t s, int relu = 0);
int batchnorm_backward(const void *x, const void *dy, void *dx, int dtype, int N, int C, int HW, const void *gamma,
                       const double *save_mean, const double *save_rstd, void *dgamma, void *dbeta, void *ws, cudaStream_t s,
                       const void *beta = nullptr, int relu = 0);
int maxpool2d(const void *x, const void *gy, void *out, int dtype, bool backward, long long planes, int H, int W, int k,
              cudaStream_t s);

// qiddm_linear.cu — Linear layers with one narrow side (linear_down / linear_up of the re-upload families)
size_t skinny_linear_ws_bytes(long long rows, int in_f, int out_f);
int skinny_linear_forward(const void *x, const void *w, const void *bias, void *y, int dtype, long long rows, int in_f, int out_f,
                          cudaStream_t s);
int skinny_linear_backward(const void *x, const void *w, const void *grad_y, void *grad_x, void *grad_w, void *grad_b, int dtype,
                           long long rows, int in_f, int out_f, void *ws, cudaStream_t s);

}  // namespace qiddm
