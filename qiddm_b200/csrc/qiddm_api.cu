// C-ABI entry points of libqiddm_b200.so (see include/qiddm.h).
#include <atomic>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "qiddm_internal.h"

namespace qiddm {
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct TimedSpan {
    cudaEvent_t a, b;
    int kind;
    double work;
};
static std::mutex g_tmutex;
static bool g_timing = false;
static std::vector<TimedSpan> g_spans;
static TimedSpan g_open;
static bool g_has_open = false;
static int g_gemm_kind = TK_GEMM_FWD;

void timing_set_gemm_kind(int kind) { g_gemm_kind = kind; }

void timing_begin(int kind, double work, cudaStream_t s) {
    if (!g_timing) return;
    std::lock_guard<std::mutex> lk(g_tmutex);
    if (cudaEventCreate(&g_open.a) != cudaSuccess || cudaEventCreate(&g_open.b) != cudaSuccess) return;
    g_open.kind = kind;
    g_open.work = work;
    cudaEventRecord(g_open.a, s);
    g_has_open = true;
}
void timing_end(cudaStream_t s) {
    if (!g_timing) return;
    std::lock_guard<std::mutex> lk(g_tmutex);
    if (!g_has_open) return;
    cudaEventRecord(g_open.b, s);
    g_spans.push_back(g_open);
    if (g_open.kind == TK_GEMM) {          // second record under the per-role GEMM kind (shares the events)
        TimedSpan t = g_open;
        t.kind = -g_gemm_kind;             // negative: do not destroy the events twice
        g_spans.push_back(t);
    }
    g_has_open = false;
}
}  // namespace qiddm

struct qiddm_plan {
    qiddm_circuit_desc d;
    int n_rot;
    int dim;
};

namespace {

using namespace qiddm;

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) & ~(kAlign - 1); }

bool desc_valid(const qiddm_circuit_desc *d) {
    if (!d) return false;
    if (d->n_qubits < 1 || d->n_qubits > QIDDM_MAX_QUBITS) return false;
    if (d->n_blocks < 1 || d->layers_per_block < 1) return false;
    const int A = 1 << d->n_qubits;
    if (d->init < QIDDM_INIT_ZERO || d->init > QIDDM_INIT_BASIS) return false;
    if (d->enc < QIDDM_ENC_NONE || d->enc > QIDDM_ENC_RY) return false;
    if (d->init == QIDDM_INIT_AMPLITUDE) {
        if (d->n_features < 1 || d->n_features > A) return false;
        if (d->enc != QIDDM_ENC_NONE) return false;  // one input tensor per stage
    }
    if (d->imprimitive != QIDDM_IMP_CNOT && d->imprimitive != QIDDM_IMP_CZ) return false;
    if (d->remap < QIDDM_REMAP_NONE || d->remap > QIDDM_REMAP_PI_TANH) return false;
    if (d->readout < QIDDM_READ_PROBS || d->readout > QIDDM_READ_STATE) return false;
    if (d->readout == QIDDM_READ_PROBS) {
        if (d->read_count < 1 || d->read_stride < 1) return false;
        if ((long long)(d->read_count - 1) * d->read_stride >= A) return false;
    } else if (d->clamp) {
        return false;
    }
    if (d->path < QIDDM_PATH_AUTO || d->path > QIDDM_PATH_GEMM) return false;
    return true;
}

int n_inputs(const qiddm_circuit_desc *d) {
    if (d->init == QIDDM_INIT_AMPLITUDE) return d->n_features;
    if (d->enc != QIDDM_ENC_NONE) return d->n_qubits;
    return 0;
}
int n_outputs(const qiddm_circuit_desc *d) {
    if (d->readout == QIDDM_READ_PROBS) return d->read_count;
    if (d->readout == QIDDM_READ_EXPVAL_Z) return d->n_qubits;
    return 2 << d->n_qubits;
}

GateParams make_params(const qiddm_plan *pl, const qiddm_unfold_desc *u, long long B) {
    const qiddm_circuit_desc &d = pl->d;
    GateParams p;
    std::memset(&p, 0, sizeof(p));
    p.n_blocks = d.n_blocks;
    p.layers = d.layers_per_block;
    p.init = d.init;
    p.n_features = d.n_features;
    p.enc = d.enc;
    p.imprimitive = d.imprimitive;
    p.readout = d.readout;
    p.read_count = d.read_count;
    p.read_stride = d.read_stride > 0 ? d.read_stride : 1;
    p.clamp = d.clamp;
    p.pad_value = d.pad_value;
    p.add_offset = d.add_offset;
    p.enc_scale = d.enc_scale;
    p.post_scale = d.post_scale;
    p.clamp_lo = d.clamp_lo;
    p.clamp_hi = d.clamp_hi;
    p.n_rot = pl->n_rot;
    p.merge_post = (d.imprimitive == QIDDM_IMP_CZ && d.enc != QIDDM_ENC_RY) ? 1 : 0;
    p.B = B;
    if (u) {
        p.unfold = 1;
        p.C = u->channels; p.H = u->height; p.W = u->width;
        p.kh = u->kernel_h; p.kw = u->kernel_w; p.ph = u->pad_h; p.pw = u->pad_w;
        p.Hout = u->height + 2 * u->pad_h - u->kernel_h + 1;
        p.Wout = u->width + 2 * u->pad_w - u->kernel_w + 1;
    }
    return p;
}

bool unfold_valid(const qiddm_plan *pl, const qiddm_unfold_desc *u) {
    if (!u || pl->d.init != QIDDM_INIT_AMPLITUDE || pl->d.readout != QIDDM_READ_PROBS) return false;
    if (u->channels < 1 || u->height < 1 || u->width < 1 || u->kernel_h < 1 || u->kernel_w < 1) return false;
    if (u->pad_h < 0 || u->pad_w < 0) return false;
    if (u->channels * u->kernel_h * u->kernel_w != pl->d.n_features) return false;
    if (u->height + 2 * u->pad_h - u->kernel_h + 1 < 1 || u->width + 2 * u->pad_w - u->kernel_w + 1 < 1) return false;
    return true;
}

size_t gates_bytes(const qiddm_plan *pl) {
    return align_up(gate_table_bytes(pl->d.n_qubits, pl->d.n_blocks * pl->d.layers_per_block));
}

int forward_impl(const qiddm_plan *pl, const qiddm_unfold_desc *u, const float *in, const int32_t *basis,
                 const void *weights, int wdtype, float *out, void *ws, long long B, cudaStream_t s, float *state = nullptr,
                 const float *init_state = nullptr, int in_shift = 0, int io64 = 0) {
    if (!pl || !weights || !ws || B < 0) return QIDDM_EINVAL;
    if (wdtype != QIDDM_DTYPE_F32 && wdtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    if (B == 0) return QIDDM_OK;
    if (n_inputs(&pl->d) > 0 && !in) return QIDDM_EINVAL;
    if (!out) return QIDDM_EINVAL;
    GateParams p = make_params(pl, u, B);
    float *gates = reinterpret_cast<float *>(ws);
    p.in = in; p.basis = basis; p.gates = gates; p.out = out;
    p.state = (state != nullptr && gate_state_compatible(pl->d.n_qubits, p)) ? state : nullptr;
    p.init_state = init_state; p.in_shift = in_shift;
    p.io64 = (u != nullptr && io64) ? 1 : 0;
    if (pl->d.init == QIDDM_INIT_STATE && !init_state) return QIDDM_EINVAL;
    cudaError_t e = launch_prepare_tables(weights, wdtype, pl->d.remap, pl->d.n_qubits, false, p, gates, s);
    if (e != cudaSuccess) return (int)e;
    LaunchInfo li;
    if ((e = gate_launch_info(pl->d.n_qubits, false, p, &li)) != cudaSuccess) return (int)e;
    if ((e = launch_gate_forward(pl->d.n_qubits, p, li, s)) != cudaSuccess) return (int)e;
    return QIDDM_OK;
}

int backward_impl(const qiddm_plan *pl, const qiddm_unfold_desc *u, const float *in, const int32_t *basis,
                  const void *weights, int wdtype, const float *grad_out, float *grad_in, void *grad_weights,
                  void *ws, long long B, long long grad_in_elems, cudaStream_t s, const float *state = nullptr, int io64 = 0) {
    if (!pl || !weights || !ws || B < 0) return QIDDM_EINVAL;
    if (wdtype != QIDDM_DTYPE_F32 && wdtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    if (B > 0 && (!grad_out || (n_inputs(&pl->d) > 0 && !in))) return QIDDM_EINVAL;
    if (n_inputs(&pl->d) == 0) grad_in = nullptr;
    GateParams p = make_params(pl, u, B);
    float *gates = reinterpret_cast<float *>(ws);
    float *partials = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + gates_bytes(pl));
    p.in = in; p.basis = basis; p.gates = gates; p.out = nullptr;
    p.grad_out = grad_out; p.grad_in = grad_in; p.partials = partials;
    p.state = (state != nullptr && gate_state_compatible(pl->d.n_qubits, p)) ? const_cast<float *>(state) : nullptr;
    p.io64 = (u != nullptr && io64) ? 1 : 0;
    cudaError_t e;
    if (B == 0) {
        if (grad_weights) {
            const size_t esz = wdtype == QIDDM_DTYPE_F64 ? 8 : 4;
            if ((e = cudaMemsetAsync(grad_weights, 0, (size_t)pl->n_rot * 3 * esz, s)) != cudaSuccess) return (int)e;
        }
        return QIDDM_OK;
    }
    if (u && grad_in) {
        if ((e = cudaMemsetAsync(grad_in, 0, (size_t)grad_in_elems * (p.io64 ? sizeof(double) : sizeof(float)), s)) != cudaSuccess)
            return (int)e;
    }
    if ((e = launch_prepare_tables(weights, wdtype, pl->d.remap, pl->d.n_qubits, true, p, gates, s)) != cudaSuccess)
        return (int)e;
    LaunchInfo li;
    if ((e = gate_launch_info(pl->d.n_qubits, true, p, &li)) != cudaSuccess) return (int)e;
    if ((e = launch_gate_backward(pl->d.n_qubits, p, li, s)) != cudaSuccess) return (int)e;
    if (grad_weights) {
        if ((e = launch_finalize_grads(partials, li.grid, weights, wdtype, pl->d.remap, pl->d.n_qubits, p, grad_weights,
                                       s)) != cudaSuccess)
            return (int)e;
    }
    return QIDDM_OK;
}

}  // namespace

extern "C" {

int qiddm_abi_version(void) { return QIDDM_ABI_VERSION; }

const char *qiddm_error_string(int code) {
    switch (code) {
        case QIDDM_OK: return "ok";
        case QIDDM_EINVAL: return "invalid argument or descriptor";
        case QIDDM_EUNSUPPORTED: return "unsupported configuration";
        case QIDDM_ENOMEM: return "out of memory";
        case QIDDM_ENODEVICE: return "no CUDA device";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int qiddm_n_inputs(const qiddm_circuit_desc *d) { return desc_valid(d) ? n_inputs(d) : 0; }
int qiddm_n_outputs(const qiddm_circuit_desc *d) { return desc_valid(d) ? n_outputs(d) : 0; }
int qiddm_n_weights(const qiddm_circuit_desc *d) {
    return desc_valid(d) ? d->n_blocks * d->layers_per_block * d->n_qubits * 3 : 0;
}

int qiddm_plan_create(const qiddm_circuit_desc *desc, qiddm_plan **plan) {
    if (!plan) return QIDDM_EINVAL;
    *plan = nullptr;
    if (!desc_valid(desc)) return QIDDM_EINVAL;
    const long long n_rot = (long long)desc->n_blocks * desc->layers_per_block * desc->n_qubits;
    // the backward kernel keeps three angle-gradient sums per Rot gate in shared memory
    if (n_rot * 12 > 96 * 1024) return QIDDM_EUNSUPPORTED;
    qiddm_plan *p = new (std::nothrow) qiddm_plan;
    if (!p) return QIDDM_ENOMEM;
    p->d = *desc;
    p->n_rot = (int)n_rot;
    p->dim = 1 << desc->n_qubits;
    *plan = p;
    return QIDDM_OK;
}

void qiddm_plan_destroy(qiddm_plan *plan) { delete plan; }

size_t qiddm_workspace_bytes(const qiddm_plan *plan, int64_t batch) {
    if (!plan) return 0;
    // gate matrices + per-CTA partial cotangents of the persistent backward grid
    long long grid = 148 * 16;
    if (batch > 0) {
        GateParams p = make_params(plan, nullptr, batch);
        LaunchInfo li;
        if (gate_launch_info(plan->d.n_qubits, true, p, &li) == cudaSuccess) grid = li.grid;
        else (void)cudaGetLastError();
    }
    return gates_bytes(plan) +
           align_up((size_t)grid * gate_partial_floats(plan->d.n_qubits, plan->d.n_blocks * plan->d.layers_per_block) *
                    sizeof(float));
}

int qiddm_forward(const qiddm_plan *plan, const float *in, const int32_t *basis, const void *weights,
                  int weights_dtype, float *out, void *workspace, int64_t batch, qiddm_stream_t stream) {
    return forward_impl(plan, nullptr, in, basis, weights, weights_dtype, out, workspace, batch, (cudaStream_t)stream);
}

int qiddm_backward(const qiddm_plan *plan, const float *in, const int32_t *basis, const void *weights,
                   int weights_dtype, const float *grad_out, float *grad_in, void *grad_weights, void *workspace,
                   int64_t batch, qiddm_stream_t stream) {
    return backward_impl(plan, nullptr, in, basis, weights, weights_dtype, grad_out, grad_in, grad_weights, workspace,
                         batch, 0, (cudaStream_t)stream);
}

size_t qiddm_state_bytes(const qiddm_plan *plan, int64_t batch) {
    if (!plan || batch < 0) return 0;
    return (size_t)batch * (size_t)plan->dim * 2 * sizeof(float);
}

int qiddm_forward_save(const qiddm_plan *plan, const float *in, const int32_t *basis, const void *weights,
                       int weights_dtype, float *out, float *state, void *workspace, int64_t batch, qiddm_stream_t stream) {
    return forward_impl(plan, nullptr, in, basis, weights, weights_dtype, out, workspace, batch, (cudaStream_t)stream, state);
}

int qiddm_backward_saved(const qiddm_plan *plan, const float *in, const int32_t *basis, const void *weights,
                         int weights_dtype, const float *grad_out, const float *state, float *grad_in, void *grad_weights,
                         void *workspace, int64_t batch, qiddm_stream_t stream) {
    return backward_impl(plan, nullptr, in, basis, weights, weights_dtype, grad_out, grad_in, grad_weights, workspace,
                         batch, 0, (cudaStream_t)stream, state);
}

int qiddm_qconv_forward(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, const float *img,
                        const void *weights, int weights_dtype, float *out, void *workspace, int64_t n_images,
                        qiddm_stream_t stream) {
    if (!plan || !unfold_valid(plan, unfold) || n_images < 0) return QIDDM_EINVAL;
    const long long P = (long long)(unfold->height + 2 * unfold->pad_h - unfold->kernel_h + 1) *
                        (unfold->width + 2 * unfold->pad_w - unfold->kernel_w + 1);
    return forward_impl(plan, unfold, img, nullptr, weights, weights_dtype, out, workspace, n_images * P,
                        (cudaStream_t)stream);
}

int qiddm_qconv_backward(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, const float *img,
                         const void *weights, int weights_dtype, const float *grad_out, float *grad_img,
                         void *grad_weights, void *workspace, int64_t n_images, qiddm_stream_t stream) {
    if (!plan || !unfold_valid(plan, unfold) || n_images < 0) return QIDDM_EINVAL;
    const long long P = (long long)(unfold->height + 2 * unfold->pad_h - unfold->kernel_h + 1) *
                        (unfold->width + 2 * unfold->pad_w - unfold->kernel_w + 1);
    const long long img_elems = n_images * unfold->channels * unfold->height * unfold->width;
    return backward_impl(plan, unfold, img, nullptr, weights, weights_dtype, grad_out, grad_img, grad_weights,
                         workspace, n_images * P, img_elems, (cudaStream_t)stream);
}

int qiddm_qconv_forward_io(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, int io_dtype, const void *img,
                           const void *weights, int weights_dtype, void *out, void *workspace, int64_t n_images,
                           qiddm_stream_t stream) {
    if (!plan || !unfold_valid(plan, unfold) || n_images < 0) return QIDDM_EINVAL;
    if (io_dtype != QIDDM_DTYPE_F32 && io_dtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    const long long P = (long long)(unfold->height + 2 * unfold->pad_h - unfold->kernel_h + 1) *
                        (unfold->width + 2 * unfold->pad_w - unfold->kernel_w + 1);
    return forward_impl(plan, unfold, reinterpret_cast<const float *>(img), nullptr, weights, weights_dtype,
                        reinterpret_cast<float *>(out), workspace, n_images * P, (cudaStream_t)stream, nullptr, nullptr, 0,
                        io_dtype == QIDDM_DTYPE_F64);
}

int qiddm_qconv_backward_io(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, int io_dtype, const void *img,
                            const void *weights, int weights_dtype, const void *grad_out, void *grad_img, void *grad_weights,
                            void *workspace, int64_t n_images, qiddm_stream_t stream) {
    if (!plan || !unfold_valid(plan, unfold) || n_images < 0) return QIDDM_EINVAL;
    if (io_dtype != QIDDM_DTYPE_F32 && io_dtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    const long long P = (long long)(unfold->height + 2 * unfold->pad_h - unfold->kernel_h + 1) *
                        (unfold->width + 2 * unfold->pad_w - unfold->kernel_w + 1);
    const long long img_elems = n_images * unfold->channels * unfold->height * unfold->width;
    return backward_impl(plan, unfold, reinterpret_cast<const float *>(img), nullptr, weights, weights_dtype,
                         reinterpret_cast<const float *>(grad_out), reinterpret_cast<float *>(grad_img), grad_weights, workspace,
                         n_images * P, img_elems, (cudaStream_t)stream, nullptr, io_dtype == QIDDM_DTYPE_F64);
}

int qiddm_build_unitary(const qiddm_plan *plan, const void *weights, int weights_dtype, float *unitary,
                        void *workspace, qiddm_stream_t stream) {
    if (!plan || !weights || !unitary || !workspace) return QIDDM_EINVAL;
    if (plan->d.n_blocks != 1 && plan->d.enc != QIDDM_ENC_NONE) return QIDDM_EUNSUPPORTED;
    // column c of U = circuit(|c>): instance c writes row c of U^T; transpose on the fly is not
    // needed by our consumers, who take U^T (row c = U e_c) — document: unitary[c][k] = U[k][c].
    qiddm_plan tmp = *plan;
    tmp.d.init = QIDDM_INIT_BASIS;
    tmp.d.enc = QIDDM_ENC_NONE;
    tmp.d.readout = QIDDM_READ_STATE;
    tmp.d.clamp = 0;
    return forward_impl(&tmp, nullptr, nullptr, nullptr, weights, weights_dtype, unitary, workspace, tmp.dim,
                        (cudaStream_t)stream);
}

// ----------------------------------------------------------------------------- unitary-collapse (GEMM) path
static bool gemm_eligible(const qiddm_plan *pl) {
    const qiddm_circuit_desc &d = pl->d;
    return d.init == QIDDM_INIT_AMPLITUDE && d.readout == QIDDM_READ_PROBS && d.n_blocks == 1 &&
           d.enc == QIDDM_ENC_NONE && d.n_qubits >= 3;
}
static qiddm_plan basis_plan(const qiddm_plan *pl) {
    qiddm_plan t = *pl;
    t.d.init = QIDDM_INIT_BASIS;
    t.d.enc = QIDDM_ENC_NONE;
    t.d.readout = QIDDM_READ_STATE;
    t.d.clamp = 0;
    t.d.n_features = 0;
    return t;
}
static size_t basis_ws_bytes(const qiddm_plan *pl) {
    qiddm_plan t = basis_plan(pl);
    return qiddm_workspace_bytes(&t, t.dim);
}

int qiddm_gemm_supported(const qiddm_plan *plan) { return plan && gemm_eligible(plan) ? 1 : 0; }

size_t qiddm_gemm_collapsed_bytes(const qiddm_plan *plan) {
    if (!plan || !gemm_eligible(plan)) return 0;
    GateParams gp = make_params(plan, nullptr, 1);
    return gemm_collapsed_bytes(gemm_shape(gp, plan->d.n_qubits));
}

size_t qiddm_gemm_workspace_bytes(const qiddm_plan *plan, int64_t batch) {
    if (!plan || !gemm_eligible(plan) || batch < 0) return 0;
    GateParams gp = make_params(plan, nullptr, 1);
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    return gemm_backward_ws_bytes(g, batch > 0 ? batch : 1) + basis_ws_bytes(plan) + 256;
}

int qiddm_gemm_prepare(const qiddm_plan *plan, const void *weights, int weights_dtype, void *collapsed,
                       void *workspace, qiddm_stream_t stream) {
    if (!plan || !weights || !collapsed || !workspace) return QIDDM_EINVAL;
    if (!gemm_eligible(plan)) return QIDDM_EUNSUPPORTED;
    GateParams gp = make_params(plan, nullptr, 1);
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    qiddm_plan t = basis_plan(plan);
    int rc = forward_impl(&t, nullptr, nullptr, nullptr, weights, weights_dtype, gemm_collapsed_ut(g, collapsed),
                          workspace, t.dim, (cudaStream_t)stream);
    if (rc != QIDDM_OK) return rc;
    return gemm_build_operands(g, gp, collapsed, (cudaStream_t)stream);
}

// Collapse for a layer that runs as a direct convolution: U^T and its fp32 filter rows only (no fp16 GEMM operands)
int qiddm_gemm_prepare_direct(const qiddm_plan *plan, const void *weights, int weights_dtype, void *collapsed,
                              void *workspace, qiddm_stream_t stream) {
    if (!plan || !weights || !collapsed || !workspace) return QIDDM_EINVAL;
    if (!gemm_eligible(plan)) return QIDDM_EUNSUPPORTED;
    GateParams gp = make_params(plan, nullptr, 1);
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    if (conv_wd_bytes(g) == 0) return QIDDM_EUNSUPPORTED;
    qiddm_plan t = basis_plan(plan);
    int rc = forward_impl(&t, nullptr, nullptr, nullptr, weights, weights_dtype, gemm_collapsed_ut(g, collapsed),
                          workspace, t.dim, (cudaStream_t)stream);
    if (rc != QIDDM_OK) return rc;
    return conv_build_wd(g, gp, gemm_collapsed_ut(g, collapsed), gemm_collapsed_wd(g, collapsed), (cudaStream_t)stream);
}

size_t qiddm_gemm_forward_workspace_bytes(const qiddm_plan *plan, int64_t batch) {
    if (!plan || !gemm_eligible(plan) || batch < 0) return 0;
    GateParams gp = make_params(plan, nullptr, 1);
    return gemm_forward_ws_bytes(gemm_shape(gp, plan->d.n_qubits), batch > 0 ? batch : 1) + 256;
}

size_t qiddm_gemm_saved_bytes(const qiddm_plan *plan, int64_t batch) {
    if (!plan || !gemm_eligible(plan) || batch < 0) return 0;
    GateParams gp = make_params(plan, nullptr, 1);
    return gemm_saved_bytes(gemm_shape(gp, plan->d.n_qubits), batch > 0 ? batch : 1);
}

int qiddm_gemm_forward(const qiddm_plan *plan, const void *collapsed, const float *in, float *out, void *saved,
                       void *workspace, int64_t batch, int precision, qiddm_stream_t stream) {
    if (!plan || !collapsed || !workspace || batch < 0) return QIDDM_EINVAL;
    if (!gemm_eligible(plan)) return QIDDM_EUNSUPPORTED;
    if (precision != 1 && precision != 3) return QIDDM_EINVAL;
    if (batch == 0) return QIDDM_OK;
    if (!in || !out) return QIDDM_EINVAL;
    if (batch > 0x7fffffffLL - 256) return QIDDM_EUNSUPPORTED;
    GateParams gp = make_params(plan, nullptr, batch);
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    return gemm_forward(g, gp, collapsed, in, out, saved, workspace, batch, precision, (cudaStream_t)stream);
}

int qiddm_gemm_backward(const qiddm_plan *plan, const void *collapsed, const float *in, const void *weights,
                        int weights_dtype, const float *grad_out, const void *saved, float *grad_in,
                        void *grad_weights, void *workspace, int64_t batch, int precision, qiddm_stream_t stream) {
    if (!plan || !collapsed || !workspace || !weights || batch < 0) return QIDDM_EINVAL;
    if (!gemm_eligible(plan)) return QIDDM_EUNSUPPORTED;
    if (precision != 1 && precision != 3) return QIDDM_EINVAL;
    if (batch > 0x7fffffffLL - 256) return QIDDM_EUNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    if (batch == 0) {
        if (grad_weights) {
            const size_t esz = weights_dtype == QIDDM_DTYPE_F64 ? 8 : 4;
            cudaError_t e = cudaMemsetAsync(grad_weights, 0, (size_t)plan->n_rot * 3 * esz, s);
            if (e != cudaSuccess) return (int)e;
        }
        return QIDDM_OK;
    }
    if (!in || !grad_out) return QIDDM_EINVAL;
    GateParams gp = make_params(plan, nullptr, batch);
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    float *gut = nullptr;
    int rc = gemm_backward(g, gp, collapsed, in, grad_out, saved, grad_in, &gut, workspace, batch, precision, s);
    if (rc != QIDDM_OK) return rc;
    if (!grad_weights) return QIDDM_OK;
    // adjoint sweep on the 2^n basis columns with the READ_STATE cotangent dL/dU^T
    qiddm_plan t = basis_plan(plan);
    char *gate_ws = reinterpret_cast<char *>(workspace) + align_up(gemm_backward_ws_bytes(g, batch));
    // psi_final of basis column c is row c of U^T, which the collapse already holds: the adjoint sweep starts from it
    return backward_impl(&t, nullptr, nullptr, nullptr, weights, weights_dtype, gut, nullptr, grad_weights, gate_ws,
                         t.dim, 0, s, gemm_collapsed_ut(g, const_cast<void *>(collapsed)));
}

// ---- fused diffusion training step of one amplitude-embedding layer (include/qiddm.h)
size_t qiddm_dense_mse_step_workspace_bytes(const qiddm_plan *plan, int64_t n_images, int T) {
    if (!plan || !gemm_eligible(plan) || n_images < 0 || T < 1) return 0;
    GateParams gp = make_params(plan, nullptr, 1);
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    const long long B = (n_images > 0 ? n_images : 1) * (long long)T;
    return align_up(gemm_dense_mse_ws_bytes(g, B)) + basis_ws_bytes(plan) + 256;
}

int qiddm_dense_mse_step(const qiddm_plan *plan, const void *collapsed, const void *x, const float *eps, const void *w,
                         int io_dtype, int64_t n_images, int T, double a, double b, double c0, double c1,
                         const void *weights, int weights_dtype, void *loss, void *grad_weights, void *workspace,
                         int precision, int bwd_precision, qiddm_stream_t stream) {
    if (!plan || !collapsed || !workspace || !weights || !loss || !grad_weights || n_images < 1 || T < 1) return QIDDM_EINVAL;
    if (!x || !eps || !w) return QIDDM_EINVAL;
    if (!gemm_eligible(plan)) return QIDDM_EUNSUPPORTED;
    if ((precision != 1 && precision != 3) || (bwd_precision != 1 && bwd_precision != 3)) return QIDDM_EINVAL;
    if (io_dtype != QIDDM_DTYPE_F32 && io_dtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    const long long B = (long long)n_images * T;
    if (B > 0x7fffffffLL - 256) return QIDDM_EUNSUPPORTED;
    if (plan->d.read_count != plan->d.n_features || plan->d.read_stride != 1) return QIDDM_EUNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    GateParams gp = make_params(plan, nullptr, B);
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    float *gut = nullptr;
    int rc = gemm_dense_mse_step(g, gp, collapsed, x, eps, w, io_dtype == QIDDM_DTYPE_F64 ? 1 : 0, n_images, T, (float)a, (float)b,
                                 (float)c0, (float)c1, loss, &gut, workspace, precision, bwd_precision, s);
    if (rc != QIDDM_OK) return rc;
    qiddm_plan t = basis_plan(plan);
    char *gate_ws = reinterpret_cast<char *>(workspace) + align_up(gemm_dense_mse_ws_bytes(g, B));
    return backward_impl(&t, nullptr, nullptr, nullptr, weights, weights_dtype, gut, nullptr, grad_weights, gate_ws,
                         t.dim, 0, s, gemm_collapsed_ut(g, const_cast<void *>(collapsed)));
}

// ---- QConv on the unitary-collapse path: the same GEMMs with the patch-unfold fused into the operand preparation,
// probabilities written straight to NCHW, grad_out read from NCHW and the col2im of dX in one gather kernel.
static long long unfold_patches(const qiddm_unfold_desc *u) {
    return (long long)(u->height + 2 * u->pad_h - u->kernel_h + 1) * (u->width + 2 * u->pad_w - u->kernel_w + 1);
}

// direct fp32 convolution (qiddm_conv.cu) behind the same entry points when the layer's shape has one
static bool qconv_direct(const qiddm_plan *plan, const qiddm_unfold_desc *unfold) {
    if (!plan || !gemm_eligible(plan) || !unfold_valid(plan, unfold)) return false;
    if (plan->d.path == QIDDM_PATH_GEMM) return false;        // an explicit request for the tcgen05 GEMM is honoured
    GateParams gp = make_params(plan, unfold, 1);
    return conv_direct_supported(gemm_shape(gp, plan->d.n_qubits), gp);
}

int qiddm_qconv_direct_supported(const qiddm_plan *plan, const qiddm_unfold_desc *unfold) { return qconv_direct(plan, unfold) ? 1 : 0; }

size_t qiddm_qconv_gemm_saved_bytes(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, int64_t n_images) {
    if (!plan || !gemm_eligible(plan) || !unfold_valid(plan, unfold) || n_images < 0) return 0;
    if (qconv_direct(plan, unfold)) {
        GateParams gp = make_params(plan, unfold, 1);
        return conv_direct_saved_bytes(gemm_shape(gp, plan->d.n_qubits), gp, n_images > 0 ? n_images : 1);
    }
    return qiddm_gemm_saved_bytes(plan, n_images * unfold_patches(unfold));
}

size_t qiddm_qconv_gemm_workspace_bytes(const qiddm_plan *plan, const qiddm_unfold_desc *unfold, int64_t n_images) {
    if (!plan || !gemm_eligible(plan) || !unfold_valid(plan, unfold) || n_images < 0) return 0;
    const long long B = n_images > 0 ? n_images * unfold_patches(unfold) : 1;
    GateParams gp = make_params(plan, nullptr, 1);
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    if (qconv_direct(plan, unfold)) {
        GateParams gu = make_params(plan, unfold, 1);
        return align_up(conv_direct_ws_bytes(g, gu, n_images > 0 ? n_images : 1)) + basis_ws_bytes(plan) + 256;
    }
    return gemm_backward_ws_bytes(g, B, true) + basis_ws_bytes(plan) + 256;
}

int qiddm_qconv_gemm_forward(const qiddm_plan *plan, const void *collapsed, const qiddm_unfold_desc *unfold,
                             int io_dtype, const void *img, void *out, void *saved, void *workspace, int64_t n_images,
                             int precision, qiddm_stream_t stream) {
    if (!plan || !collapsed || !workspace || n_images < 0) return QIDDM_EINVAL;
    if (!gemm_eligible(plan)) return QIDDM_EUNSUPPORTED;
    if (!unfold_valid(plan, unfold) || (precision != 1 && precision != 3)) return QIDDM_EINVAL;
    if (io_dtype != QIDDM_DTYPE_F32 && io_dtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    if (n_images == 0) return QIDDM_OK;
    if (!img || !out) return QIDDM_EINVAL;
    const long long B = n_images * unfold_patches(unfold);
    if (B > 0x7fffffffLL - 256) return QIDDM_EUNSUPPORTED;
    GateParams gp = make_params(plan, unfold, B);
    gp.io64 = io_dtype == QIDDM_DTYPE_F64 ? 1 : 0;
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    if (qconv_direct(plan, unfold))
        return conv_direct_forward(g, gp, gemm_collapsed_wd(g, const_cast<void *>(collapsed)), img, out, saved, n_images,
                                   (cudaStream_t)stream);
    return gemm_forward(g, gp, collapsed, reinterpret_cast<const float *>(img), reinterpret_cast<float *>(out), saved, workspace,
                        B, precision, (cudaStream_t)stream);
}

int qiddm_qconv_gemm_backward(const qiddm_plan *plan, const void *collapsed, const qiddm_unfold_desc *unfold,
                              int io_dtype, const void *img, const void *weights, int weights_dtype, const void *grad_out,
                              const void *saved, void *grad_img, void *grad_weights, void *workspace,
                              int64_t n_images, int precision, qiddm_stream_t stream) {
    if (!plan || !collapsed || !workspace || !weights || n_images < 0) return QIDDM_EINVAL;
    if (!gemm_eligible(plan)) return QIDDM_EUNSUPPORTED;
    if (!unfold_valid(plan, unfold) || (precision != 1 && precision != 3)) return QIDDM_EINVAL;
    if (io_dtype != QIDDM_DTYPE_F32 && io_dtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_images == 0) {
        if (grad_weights) {
            const size_t esz = weights_dtype == QIDDM_DTYPE_F64 ? 8 : 4;
            cudaError_t e = cudaMemsetAsync(grad_weights, 0, (size_t)plan->n_rot * 3 * esz, s);
            if (e != cudaSuccess) return (int)e;
        }
        return QIDDM_OK;
    }
    if (!img || !grad_out) return QIDDM_EINVAL;
    const long long B = n_images * unfold_patches(unfold);
    if (B > 0x7fffffffLL - 256) return QIDDM_EUNSUPPORTED;
    GateParams gp = make_params(plan, unfold, B);
    gp.io64 = io_dtype == QIDDM_DTYPE_F64 ? 1 : 0;
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    float *gut = nullptr;
    const bool direct = qconv_direct(plan, unfold);
    int rc;
    if (direct) {
        if (!saved) return QIDDM_EINVAL;      // the direct path keeps Y from its forward (no re-materialisation)
        rc = conv_direct_backward(g, gp, gemm_collapsed_wd(g, const_cast<void *>(collapsed)), img, grad_out, saved, grad_img, &gut,
                                  workspace, n_images, s);
    } else {
        rc = gemm_backward(g, gp, collapsed, reinterpret_cast<const float *>(img), reinterpret_cast<const float *>(grad_out), saved,
                           reinterpret_cast<float *>(grad_img), &gut, workspace, B, precision, s);
    }
    if (rc != QIDDM_OK) return rc;
    if (!grad_weights) return QIDDM_OK;
    qiddm_plan t = basis_plan(plan);
    char *gate_ws = reinterpret_cast<char *>(workspace) +
                    (direct ? align_up(conv_direct_ws_bytes(g, gp, n_images)) : align_up(gemm_backward_ws_bytes(g, B, true)));
    return backward_impl(&t, nullptr, nullptr, nullptr, weights, weights_dtype, gut, nullptr, grad_weights, gate_ws,
                         t.dim, 0, s, gemm_collapsed_ut(g, const_cast<void *>(collapsed)));
}

// ---- bilinear Upsample -> 1 x 1 QConv2d (nn/unet.py:36-41) in one pass: `img_src` is the (n, C, h_in, w_in) tensor in front of the
// upsample, `unfold` the geometry of the 1 x 1 convolution on the upsampled (height, width) image.  Direct-convolution layers only.
static bool up_args_ok(const qiddm_plan *plan, const qiddm_unfold_desc *u, int h_in, int w_in, double sh, double sw) {
    return qconv_direct(plan, u) && u->kernel_h == 1 && u->kernel_w == 1 && h_in > 0 && w_in > 0 && sh > 0.0 && sw > 0.0;
}

int qiddm_qconv_up_forward(const qiddm_plan *plan, const void *collapsed, const qiddm_unfold_desc *unfold, int io_dtype,
                           const void *img_src, int h_in, int w_in, double scale_h, double scale_w, void *out, void *saved,
                           int64_t n_images, qiddm_stream_t stream) {
    if (!plan || !collapsed || n_images < 0) return QIDDM_EINVAL;
    if (io_dtype != QIDDM_DTYPE_F32 && io_dtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    if (!gemm_eligible(plan) || !unfold_valid(plan, unfold)) return QIDDM_EINVAL;
    if (!up_args_ok(plan, unfold, h_in, w_in, scale_h, scale_w)) return QIDDM_EUNSUPPORTED;
    if (n_images == 0) return QIDDM_OK;
    if (!img_src || !out) return QIDDM_EINVAL;
    const long long B = n_images * unfold_patches(unfold);
    if (B > 0x7fffffffLL - 256) return QIDDM_EUNSUPPORTED;
    GateParams gp = make_params(plan, unfold, B);
    gp.io64 = io_dtype == QIDDM_DTYPE_F64 ? 1 : 0;
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    const ConvUp up{h_in, w_in, scale_h, scale_w};
    return conv_direct_forward(g, gp, gemm_collapsed_wd(g, const_cast<void *>(collapsed)), img_src, out, saved, n_images,
                               (cudaStream_t)stream, &up);
}

int qiddm_qconv_up_backward(const qiddm_plan *plan, const void *collapsed, const qiddm_unfold_desc *unfold, int io_dtype,
                            const void *img_src, int h_in, int w_in, double scale_h, double scale_w, const void *weights,
                            int weights_dtype, const void *grad_out, const void *saved, void *grad_up, void *grad_weights,
                            void *workspace, int64_t n_images, qiddm_stream_t stream) {
    if (!plan || !collapsed || !workspace || !weights || n_images < 1) return QIDDM_EINVAL;
    if (io_dtype != QIDDM_DTYPE_F32 && io_dtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    if (!gemm_eligible(plan) || !unfold_valid(plan, unfold)) return QIDDM_EINVAL;
    if (!up_args_ok(plan, unfold, h_in, w_in, scale_h, scale_w)) return QIDDM_EUNSUPPORTED;
    if (!img_src || !grad_out || !saved) return QIDDM_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const long long B = n_images * unfold_patches(unfold);
    if (B > 0x7fffffffLL - 256) return QIDDM_EUNSUPPORTED;
    GateParams gp = make_params(plan, unfold, B);
    gp.io64 = io_dtype == QIDDM_DTYPE_F64 ? 1 : 0;
    const GemmShape g = gemm_shape(gp, plan->d.n_qubits);
    const ConvUp up{h_in, w_in, scale_h, scale_w};
    float *gut = nullptr;
    int rc = conv_direct_backward(g, gp, gemm_collapsed_wd(g, const_cast<void *>(collapsed)), img_src, grad_out, saved, grad_up, &gut,
                                  workspace, n_images, s, &up);
    if (rc != QIDDM_OK) return rc;
    if (!grad_weights) return QIDDM_OK;
    qiddm_plan t = basis_plan(plan);
    char *gate_ws = reinterpret_cast<char *>(workspace) + align_up(conv_direct_ws_bytes(g, gp, n_images));
    return backward_impl(&t, nullptr, nullptr, nullptr, weights, weights_dtype, gut, nullptr, grad_weights, gate_ws,
                         t.dim, 0, s, gemm_collapsed_ut(g, const_cast<void *>(collapsed)));
}

// ----------------------------------------------------------------------------- mid-circuit noise (density matrix)
static bool noisy_eligible(const qiddm_plan *pl) {
    const qiddm_circuit_desc &d = pl->d;
    return d.init == QIDDM_INIT_ZERO && d.enc == QIDDM_ENC_RZ &&
           (d.readout == QIDDM_READ_PROBS || d.readout == QIDDM_READ_EXPVAL_Z);
}
static qiddm_plan noisy_block_plan(const qiddm_plan *pl) {
    qiddm_plan t = *pl;
    t.d.init = QIDDM_INIT_STATE;
    t.d.n_blocks = 1;
    t.d.readout = QIDDM_READ_STATE;
    t.d.clamp = 0;
    t.n_rot = pl->d.layers_per_block * pl->d.n_qubits;
    return t;
}

size_t qiddm_noisy_workspace_bytes(const qiddm_plan *plan, int64_t batch) {
    if (!plan || !noisy_eligible(plan) || batch < 0) return 0;
    qiddm_plan t = noisy_block_plan(plan);
    const long long rows = (batch > 0 ? batch : 1) * (long long)plan->dim;
    return 2 * dm_state_bytes(plan->d.n_qubits, batch > 0 ? batch : 1) + qiddm_workspace_bytes(&t, rows) + 256;
}

int qiddm_noisy_forward(const qiddm_plan *plan, const float *in, const void *weights, int weights_dtype, double f_off,
                        double m00, double m01, double m10, double m11, float *out, void *workspace, int64_t batch,
                        qiddm_stream_t stream) {
    if (!plan || !weights || !workspace || batch < 0) return QIDDM_EINVAL;
    if (!noisy_eligible(plan)) return QIDDM_EUNSUPPORTED;
    if (weights_dtype != QIDDM_DTYPE_F32 && weights_dtype != QIDDM_DTYPE_F64) return QIDDM_EINVAL;
    if (batch == 0) return QIDDM_OK;
    if (!in || !out) return QIDDM_EINVAL;
    const int n = plan->d.n_qubits;
    if (batch > 65535 || (batch << n) > 0x7fffffffLL) return QIDDM_EUNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    char *p8 = reinterpret_cast<char *>(workspace);
    float2 *tau = reinterpret_cast<float2 *>(p8); p8 += dm_state_bytes(n, batch);
    float2 *tau2 = reinterpret_cast<float2 *>(p8); p8 += dm_state_bytes(n, batch);
    void *gate_ws = p8;
    qiddm_plan t = noisy_block_plan(plan);
    const size_t esz = weights_dtype == QIDDM_DTYPE_F64 ? 8 : 4;
    const size_t block_stride = (size_t)plan->d.layers_per_block * n * 3 * esz;
    const long long rows = (long long)batch << n;
    int rc = dm_init(tau, n, batch, s);
    for (int i = 0; i < plan->d.n_blocks && rc == QIDDM_OK; ++i) {
        const void *w_i = reinterpret_cast<const char *>(weights) + (size_t)i * block_stride;
        // channels on every wire (RZ(a_j) commutes with them and rides in the unitary part)
        if ((rc = dm_channel(tau, n, batch, (float)f_off, (float)m00, (float)m01, (float)m10, (float)m11, s)) != QIDDM_OK) break;
        // (U rho)^T: U = SEL(W_i) RZ(a) on the rows of tau; then conj(U rho); then conj(U rho) U^T = (U rho U^dagger)^T
        if ((rc = forward_impl(&t, nullptr, in, nullptr, w_i, weights_dtype, reinterpret_cast<float *>(tau2), gate_ws, rows, s,
                               nullptr, reinterpret_cast<const float *>(tau), n)) != QIDDM_OK) break;
        if ((rc = dm_transpose_conj(tau2, tau, n, batch, s)) != QIDDM_OK) break;
        if ((rc = forward_impl(&t, nullptr, in, nullptr, w_i, weights_dtype, reinterpret_cast<float *>(tau2), gate_ws, rows, s,
                               nullptr, reinterpret_cast<const float *>(tau), n)) != QIDDM_OK) break;
        float2 *tmp = tau; tau = tau2; tau2 = tmp;
    }
    if (rc != QIDDM_OK) return rc;
    const qiddm_circuit_desc &d = plan->d;
    return dm_readout(tau, n, batch, d.readout, d.read_count, d.read_stride > 0 ? d.read_stride : 1, d.post_scale, d.clamp,
                      d.clamp_lo, d.clamp_hi, out, s);
}

void qiddm_timing_enable(int enable) {
    std::lock_guard<std::mutex> lk(qiddm::g_tmutex);
    qiddm::g_timing = enable != 0;
}

int qiddm_timing_collect(double *ms_by_kind, double *work_by_kind, int64_t *launches_by_kind) {
    std::lock_guard<std::mutex> lk(qiddm::g_tmutex);
    for (int i = 0; i < qiddm::TK_COUNT; ++i) {
        if (ms_by_kind) ms_by_kind[i] = 0;
        if (work_by_kind) work_by_kind[i] = 0;
        if (launches_by_kind) launches_by_kind[i] = 0;
    }
    int rc = QIDDM_OK;
    for (auto &sp : qiddm::g_spans) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(sp.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, sp.a, sp.b);
        if (e != cudaSuccess) rc = (int)e;
        const int kind = sp.kind < 0 ? -sp.kind : sp.kind;
        if (ms_by_kind) ms_by_kind[kind] += ms;
        if (work_by_kind) work_by_kind[kind] += sp.work;
        if (launches_by_kind) launches_by_kind[kind] += 1;
    }
    for (auto &sp : qiddm::g_spans) {
        if (sp.kind >= 0) {
            cudaEventDestroy(sp.a);
            cudaEventDestroy(sp.b);
        }
    }
    qiddm::g_spans.clear();
    return rc;
}

int qiddm_upsample_bilinear_forward(const void *in, void *out, int dtype, int64_t planes, int h_in, int w_in, int h_out,
                                    int w_out, double scale_h, double scale_w, qiddm_stream_t stream) {
    return qiddm::upsample_bilinear(in, out, dtype, false, planes, h_in, w_in, h_out, w_out, scale_h, scale_w, (cudaStream_t)stream);
}
int qiddm_upsample_bilinear_backward(const void *grad_out, void *grad_in, int dtype, int64_t planes, int h_in, int w_in,
                                     int h_out, int w_out, double scale_h, double scale_w, qiddm_stream_t stream) {
    return qiddm::upsample_bilinear(grad_out, grad_in, dtype, true, planes, h_in, w_in, h_out, w_out, scale_h, scale_w,
                                    (cudaStream_t)stream);
}
size_t qiddm_batchnorm_workspace_bytes(int channels) { return channels > 0 ? qiddm::batchnorm_ws_bytes(channels) : 0; }
int qiddm_batchnorm_forward(const void *x, void *y, int dtype, int n, int c, int hw, const void *gamma, const void *beta,
                            double *save_mean, double *save_rstd, void *running_mean, void *running_var, double momentum,
                            double eps, void *workspace, qiddm_stream_t stream) {
    return qiddm::batchnorm_forward(x, y, dtype, n, c, hw, gamma, beta, save_mean, save_rstd, running_mean, running_var, momentum,
                                    eps, workspace, (cudaStream_t)stream);
}
int qiddm_batchnorm_backward(const void *x, const void *grad_y, void *grad_x, int dtype, int n, int c, int hw, const void *gamma,
                             const double *save_mean, const double *save_rstd, void *grad_gamma, void *grad_beta,
                             void *workspace, qiddm_stream_t stream) {
    return qiddm::batchnorm_backward(x, grad_y, grad_x, dtype, n, c, hw, gamma, save_mean, save_rstd, grad_gamma, grad_beta,
                                     workspace, (cudaStream_t)stream);
}

int qiddm_batchnorm_relu_forward(const void *x, void *y, int dtype, int n, int c, int hw, const void *gamma, const void *beta,
                                 double *save_mean, double *save_rstd, void *running_mean, void *running_var, double momentum,
                                 double eps, int relu_mode, void *workspace, qiddm_stream_t stream) {
    return qiddm::batchnorm_forward(x, y, dtype, n, c, hw, gamma, beta, save_mean, save_rstd, running_mean, running_var, momentum,
                                    eps, workspace, (cudaStream_t)stream, relu_mode);
}
int qiddm_batchnorm_relu_backward(const void *x, const void *grad_y, void *grad_x, int dtype, int n, int c, int hw,
                                  const void *gamma, const void *beta, const double *save_mean, const double *save_rstd,
                                  void *grad_gamma, void *grad_beta, int relu_mode, void *workspace, qiddm_stream_t stream) {
    return qiddm::batchnorm_backward(x, grad_y, grad_x, dtype, n, c, hw, gamma, save_mean, save_rstd, grad_gamma, grad_beta,
                                     workspace, (cudaStream_t)stream, beta, relu_mode);
}
int qiddm_maxpool2d_forward(const void *x, void *y, int dtype, int64_t planes, int h, int w, int kernel, qiddm_stream_t stream) {
    return qiddm::maxpool2d(x, nullptr, y, dtype, false, planes, h, w, kernel, (cudaStream_t)stream);
}
int qiddm_maxpool2d_backward(const void *x, const void *grad_y, void *grad_x, int dtype, int64_t planes, int h, int w, int kernel,
                             qiddm_stream_t stream) {
    return qiddm::maxpool2d(x, grad_y, grad_x, dtype, true, planes, h, w, kernel, (cudaStream_t)stream);
}

size_t qiddm_skinny_linear_workspace_bytes(int64_t rows, int in_features, int out_features) {
    return qiddm::skinny_linear_ws_bytes(rows, in_features, out_features);
}
int qiddm_skinny_linear_forward(const void *x, const void *weight, const void *bias, void *y, int dtype, int64_t rows,
                                int in_features, int out_features, qiddm_stream_t stream) {
    return qiddm::skinny_linear_forward(x, weight, bias, y, dtype, rows, in_features, out_features, (cudaStream_t)stream);
}
int qiddm_skinny_linear_backward(const void *x, const void *weight, const void *grad_y, void *grad_x, void *grad_weight,
                                 void *grad_bias, int dtype, int64_t rows, int in_features, int out_features, void *workspace,
                                 qiddm_stream_t stream) {
    return qiddm::skinny_linear_backward(x, weight, grad_y, grad_x, grad_weight, grad_bias, dtype, rows, in_features, out_features,
                                         workspace, (cudaStream_t)stream);
}

int qiddm_noise_ladder(const void *x, const float *eps, const void *w, int dtype, int64_t batch, int pixels, int tau,
                       void *noisy, void *clean, qiddm_stream_t stream) {
    return qiddm::noise_ladder(x, eps, w, dtype, batch, pixels, tau, noisy, clean, (cudaStream_t)stream);
}
size_t qiddm_mse_workspace_bytes(void) { return qiddm::mse_ws_bytes(); }
int qiddm_mse_loss_grad(const void *pred, const void *target, const void *target_add, int dtype, double scale, double shift,
                        int64_t n, void *grad, void *loss, void *workspace, qiddm_stream_t stream) {
    return qiddm::mse_loss_grad(pred, target, target_add, dtype, scale, shift, n, grad, loss, workspace, (cudaStream_t)stream);
}

int qiddm_mse_ladder_loss_grad(const void *pred, const void *x, const float *eps, const void *w, int dtype, int64_t batch, int pixels,
                               int tau, double scale, double shift, double c0, double c1, void *grad, void *loss, void *workspace,
                               qiddm_stream_t stream) {
    return qiddm::mse_ladder_loss_grad(pred, x, eps, w, dtype, batch, pixels, tau, scale, shift, c0, c1, grad, loss, workspace,
                                       (cudaStream_t)stream);
}

size_t qiddm_linear_up_mse_workspace_bytes(int pixels, int hidden) {
    return (pixels > 0 && hidden > 0 && hidden <= 16) ? qiddm::linear_up_mse_ws_bytes(pixels, hidden) : 0;
}
int qiddm_linear_up_mse_step(const void *h, const void *weight, const void *bias, const void *x, const float *eps, const void *w,
                             int dtype, int64_t batch, int pixels, int tau, int hidden, double scale, double shift, double c0,
                             double c1, void *loss, void *grad_weight, void *grad_bias, void *grad_h, void *workspace,
                             qiddm_stream_t stream) {
    return qiddm::linear_up_mse_step(h, weight, bias, x, eps, w, dtype, batch, pixels, tau, hidden, scale, shift, c0, c1, loss,
                                     grad_weight, grad_bias, grad_h, workspace, (cudaStream_t)stream);
}

int qiddm_readout_channel(const void *probs_in, void *probs_out, int dtype, int64_t batch, int n_qubits, double m00, double m01,
                          double m10, double m11, qiddm_stream_t stream) {
    return qiddm::prob_channel(probs_in, probs_out, dtype, batch, n_qubits, m00, m01, m10, m11, (cudaStream_t)stream);
}

int qiddm_sym_eigh_max_dim(void) {
    int m = 1;
    while (qiddm::eigh_smem_bytes(m + 1) <= 227 * 1024) ++m;
    return m;
}

int qiddm_sym_eigh_f64(const double *a, int m, double *evals, double *evecs, qiddm_stream_t stream) {
    return qiddm::sym_eigh_f64(a, m, 1, evals, evecs, (cudaStream_t)stream);
}

int qiddm_sym_eigh_f64_batched(const double *a, int m, int64_t count, double *evals, double *evecs, qiddm_stream_t stream) {
    return qiddm::sym_eigh_f64(a, m, count, evals, evecs, (cudaStream_t)stream);
}

int64_t qiddm_stream_capture_id(qiddm_stream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo((cudaStream_t)stream, &st, &id) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return st == cudaStreamCaptureStatusActive ? (int64_t)id : 0;
}

int qiddm_qconv_reference_map_forward(const qiddm_unfold_desc *u, int dtype, const void *img, void *out, int out_channels,
                                      int64_t n_images, qiddm_stream_t stream) {
    if (!u) return QIDDM_EINVAL;
    return qiddm::qconv_reference_map(img, nullptr, out, dtype, false, n_images, u->channels, u->height, u->width, u->kernel_h,
                                      u->kernel_w, u->pad_h, u->pad_w, out_channels, (cudaStream_t)stream);
}
int qiddm_qconv_reference_map_backward(const qiddm_unfold_desc *u, int dtype, const void *img, const void *grad_out,
                                       void *grad_img, int out_channels, int64_t n_images, qiddm_stream_t stream) {
    if (!u) return QIDDM_EINVAL;
    return qiddm::qconv_reference_map(img, grad_out, grad_img, dtype, true, n_images, u->channels, u->height, u->width,
                                      u->kernel_h, u->kernel_w, u->pad_h, u->pad_w, out_channels, (cudaStream_t)stream);
}

int qiddm_probe_fp32_fma(int iters, float *sink, double *flops, qiddm_stream_t stream) {
    return qiddm::probe_fp32_fma(iters, sink, flops, (cudaStream_t)stream);
}

int64_t qiddm_launch_count(void) { return (int64_t)qiddm::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
