/* TEST INFRASTRUCTURE ONLY — a second, independent CPU restatement of the reference's circuit semantics, in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and the CPU-baseline legs of bench.py / scripts/ may build, load or call this file;
 * the product (qiddm_b200/) never does.  Where oracle/qiddm_oracle.py mirrors PennyLane's `default.qubit.torch` (batched tensor
 * ops, precomputed ring permutations, autograd), this file mirrors `lightning.qubit`, the device the reference's QIDDM_LL / PL /
 * QNN classes run on (nn/qdense.py:1372-1373, :1568-1569): one complex128 state vector per circuit instance, every gate
 * applied in place, one after the other, in the order the QNode tape lists them.  qc_forward is what the reference runs (it
 * never differentiates through these circuits); qc_backward adds the adjoint method on the same gate list (what
 * lightning.qubit's diff_method="adjoint" does, and what the CUDA gate kernel does on the device), as a check of the torch
 * oracle's autograd gradients that shares nothing with autograd.  The two restatements share no code, so their agreement (tests/
 * test_oracle_c.py, 1e-12) checks the ring composition order, the CNOT direction and the wire order a second time.
 *
 * Conventions (PennyLane 0.29 documentation — PennyLane itself is not installable here):
 *   basis index k = sum_i b_i 2^(n-1-i) (wire 0 = MSB);  RZ(a) = diag(e^{-ia/2}, e^{+ia/2});  RY(a) = [[c,-s],[s,c]];
 *   Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi);  StronglyEntanglingLayers: per layer l, Rot on every wire, then (n > 1)
 *   imprimitive(i, (i + r_l) mod n) for i = 0..n-1 with r_l = (l mod (n-1)) + 1, every SEL call restarting at l = 0
 *   (nn/qdense.py:109, :171, :461, :1612; nn/qconv.py:56);  AmplitudeEmbedding(pad_with, normalize) (nn/qdense.py:41-43, nn/qconv.py:52-54);
 *   probs = |psi_k|^2 in index order (nn/qdense.py:54, :111);  expval(PauliZ(j)) = sum_k (1 - 2 b_j(k)) |psi_k|^2 (nn/qdense.py:1615).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

/* field order = oracle.qiddm_oracle.StageDesc (enums by value: include/qiddm.h) */
typedef struct {
    int n_qubits, n_blocks, layers_per_block;
    int init;            /* 0 |0..0>, 1 amplitude embedding, 2 basis state */
    int n_features;
    double pad_value, add_offset;
    int enc;             /* 0 none, 1 RZ(s a_j), 2 RY(s a_j) before every block */
    double enc_scale;
    int imprimitive;     /* 0 CNOT, 1 CZ */
    int remap;           /* 0 none, 1 tanh, 2 pi tanh */
    int readout;         /* 0 probs, 1 <Z_j>, 2 full state (re, im interleaved) */
    int read_count, read_stride;
    double post_scale;
    int clamp;
    double clamp_lo, clamp_hi;
} qc_desc;

/* number of OpenMP threads of the calls that follow (0: leave the runtime's default) */
void qc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static void apply_1q(cplx *s, int n, int wire, cplx m00, cplx m01, cplx m10, cplx m11) {
    const long A = 1L << n, st = 1L << (n - 1 - wire);
    for (long base = 0; base < A; base += 2 * st)
        for (long j = 0; j < st; ++j) {
            const cplx a = s[base + j], b = s[base + j + st];
            s[base + j] = m00 * a + m01 * b;
            s[base + j + st] = m10 * a + m11 * b;
        }
}

static void apply_rz(cplx *s, int n, int wire, double a) {
    apply_1q(s, n, wire, cexp(-0.5 * I * a), 0, 0, cexp(0.5 * I * a));
}

static void apply_ry(cplx *s, int n, int wire, double a) {
    const double c = cos(a / 2), sn = sin(a / 2);
    apply_1q(s, n, wire, c, -sn, sn, c);
}

static void apply_rot(cplx *s, int n, int wire, double phi, double theta, double omega) {
    apply_rz(s, n, wire, phi);
    apply_ry(s, n, wire, theta);
    apply_rz(s, n, wire, omega);
}

static void apply_cnot(cplx *s, int n, int c, int t) {      /* |c, t> -> |c, t xor c> */
    const long A = 1L << n, cm = 1L << (n - 1 - c), tm = 1L << (n - 1 - t);
    for (long k = 0; k < A; ++k)
        if ((k & cm) && !(k & tm)) {
            const cplx v = s[k];
            s[k] = s[k | tm];
            s[k | tm] = v;
        }
}

static void apply_cz(cplx *s, int n, int a, int b) {
    const long A = 1L << n, am = 1L << (n - 1 - a), bm = 1L << (n - 1 - b);
    for (long k = 0; k < A; ++k)
        if ((k & am) && (k & bm)) s[k] = -s[k];
}

static double remap_w(double w, int remap) {
    if (remap == 1) return tanh(w);
    if (remap == 2) return M_PI * tanh(w);
    return w;
}

/* One QNode evaluation per row.  x: (B, x_stride) doubles (amplitude features, or >= n angles per row; may be NULL when the
 * circuit takes no input), basis_index: (B) for init = 2, weights: (n_blocks, layers_per_block, n, 3) raw, out: (B, n_out) with
 * n_out = read_count (probs), n (expval) or 2 * 2^n (state).  Returns 0, or -1 for an invalid descriptor. */
int qc_forward(const qc_desc *d, const double *x, long x_stride, const long *basis_index, const double *weights, double *out,
               long B) {
    const int n = d->n_qubits;
    if (n < 1 || n > 20 || d->n_blocks < 1 || d->layers_per_block < 1) return -1;
    const long A = 1L << n;
    if (d->init == 1 && (x == NULL || d->n_features < 1 || d->n_features > A)) return -1;
    if (d->init == 2 && basis_index == NULL) return -1;
    if (d->enc != 0 && (x == NULL || x_stride < n)) return -1;
    const long n_out = d->readout == 2 ? 2 * A : (d->readout == 1 ? n : d->read_count);
    if (d->readout == 0 && (long)(d->read_count - 1) * d->read_stride >= A) return -1;
    int rc = 0;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < B; ++b) {
        cplx *s = (cplx *)calloc((size_t)A, sizeof(cplx));
        if (!s) { rc = -1; continue; }
        if (d->init == 1) {
            double nrm = 0;
            for (long k = 0; k < A; ++k) {
                const double v = k < d->n_features ? x[b * x_stride + k] + d->add_offset : d->pad_value;
                s[k] = v;
                nrm += v * v;
            }
            nrm = sqrt(nrm);
            for (long k = 0; k < A; ++k) s[k] /= nrm;
        } else if (d->init == 2) {
            s[basis_index[b]] = 1;
        } else {
            s[0] = 1;
        }
        for (int blk = 0; blk < d->n_blocks; ++blk) {
            if (d->enc == 1)
                for (int j = 0; j < n; ++j) apply_rz(s, n, j, x[b * x_stride + j] * d->enc_scale);
            else if (d->enc == 2)
                for (int j = 0; j < n; ++j) apply_ry(s, n, j, x[b * x_stride + j] * d->enc_scale);
            for (int l = 0; l < d->layers_per_block; ++l) {
                const double *w = weights + ((long)(blk * d->layers_per_block + l) * n) * 3;
                for (int i = 0; i < n; ++i)
                    apply_rot(s, n, i, remap_w(w[3 * i], d->remap), remap_w(w[3 * i + 1], d->remap), remap_w(w[3 * i + 2], d->remap));
                if (n > 1) {
                    const int r = (l % (n - 1)) + 1;
                    for (int i = 0; i < n; ++i) {
                        if (d->imprimitive == 0) apply_cnot(s, n, i, (i + r) % n);
                        else apply_cz(s, n, i, (i + r) % n);
                    }
                }
            }
        }
        double *o = out + b * n_out;
        if (d->readout == 2) {
            for (long k = 0; k < A; ++k) { o[2 * k] = creal(s[k]); o[2 * k + 1] = cimag(s[k]); }
        } else {
            if (d->readout == 1) {
                for (int j = 0; j < n; ++j) {
                    double z = 0;
                    for (long k = 0; k < A; ++k) {
                        const double p = creal(s[k]) * creal(s[k]) + cimag(s[k]) * cimag(s[k]);
                        z += ((k >> (n - 1 - j)) & 1) ? -p : p;
                    }
                    o[j] = z;
                }
            } else {
                for (int m = 0; m < d->read_count; ++m) {
                    const cplx v = s[(long)m * d->read_stride];
                    o[m] = creal(v) * creal(v) + cimag(v) * cimag(v);
                }
            }
            for (long m = 0; m < n_out; ++m) {
                double v = o[m] * d->post_scale;
                if (d->clamp) v = v < d->clamp_lo ? d->clamp_lo : (v > d->clamp_hi ? d->clamp_hi : v);
                o[m] = v;
            }
        }
        free(s);
    }
    return rc;
}

/* ---- adjoint-method backward (gradient of L = sum(out * grad_out) w.r.t. the raw weights and the inputs) ----------------
 * With psi the final state and lambda = dL/d(conj psi), walking the gate list in reverse: for a rotation R_P(a) = exp(-i a P / 2)
 * (P = Z or Y) applied at some point, dL/da = Im <lambda_post | P | psi_post>; then both vectors are taken back through R_P(a)^+.
 * CNOT / CZ are their own inverses.  Amplitude embedding: psi0 = v / |v| is real, dL/dpsi0 = 2 Re(lambda0), projected through
 * the normalisation. */
static double rot_grad_z(const cplx *lam, const cplx *psi, int n, int wire) {
    const long A = 1L << n;
    double acc = 0;
    for (long k = 0; k < A; ++k) {
        const double v = cimag(conj(lam[k]) * psi[k]);
        acc += ((k >> (n - 1 - wire)) & 1) ? -v : v;
    }
    return acc;
}

static double rot_grad_y(const cplx *lam, const cplx *psi, int n, int wire) {       /* Y|0> = i|1>, Y|1> = -i|0> */
    const long A = 1L << n, st = 1L << (n - 1 - wire);
    double acc = 0;
    for (long k = 0; k < A; ++k)
        if (!(k & st)) acc += cimag(conj(lam[k]) * (-I * psi[k | st]) + conj(lam[k | st]) * (I * psi[k]));
    return acc;
}

static double remap_dw(double w, int remap) {
    if (remap == 1) { const double t = tanh(w); return 1 - t * t; }
    if (remap == 2) { const double t = tanh(w); return M_PI * (1 - t * t); }
    return 1;
}

/* grad_x: (B, x_stride) or NULL; grad_w: (n_blocks, layers_per_block, n, 3), ACCUMULATED over the rows (zeroed here). */
int qc_backward(const qc_desc *d, const double *x, long x_stride, const long *basis_index, const double *weights,
                const double *grad_out, double *grad_x, double *grad_w, long B) {
    const int n = d->n_qubits;
    if (d->readout == 2) return -1;                  /* state read-out: not needed by the tests */
    const long A = 1L << n;
    const long n_out = d->readout == 1 ? n : d->read_count;
    const long n_w = (long)d->n_blocks * d->layers_per_block * n * 3;
    double *fwd = (double *)malloc((size_t)(B * n_out) * sizeof(double));
    qc_desc ds = *d;
    ds.readout = 2;                                   /* final state of every row */
    double *st_all = (double *)malloc((size_t)(B * 2 * A) * sizeof(double));
    if (!fwd || !st_all) { free(fwd); free(st_all); return -1; }
    int rc = qc_forward(&ds, x, x_stride, basis_index, weights, st_all, B);
    if (rc == 0) rc = qc_forward(d, x, x_stride, basis_index, weights, fwd, B);
    if (rc != 0) { free(fwd); free(st_all); return rc; }
    memset(grad_w, 0, (size_t)n_w * sizeof(double));
    if (grad_x) memset(grad_x, 0, (size_t)(B * x_stride) * sizeof(double));
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    double *gw_all = (double *)calloc((size_t)nt * n_w, sizeof(double));   /* per-thread partial sums, added in thread order */
    if (!gw_all) { free(fwd); free(st_all); return -1; }
#pragma omp parallel num_threads(nt)
    {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double *grad_w = gw_all + (long)tid * n_w;        /* shadows the output: this thread's partial sum */
#pragma omp for schedule(static)
    for (long b = 0; b < B; ++b) {
        cplx *psi = (cplx *)malloc((size_t)A * sizeof(cplx)), *lam = (cplx *)calloc((size_t)A, sizeof(cplx));
        for (long k = 0; k < A; ++k) psi[k] = st_all[(b * A + k) * 2] + I * st_all[(b * A + k) * 2 + 1];
        for (long m = 0; m < n_out; ++m) {            /* seed: out = clamp(post_scale * q), q = |psi_k|^2 or <Z_j> */
            const double raw = fwd[b * n_out + m];
            double g = grad_out[b * n_out + m] * d->post_scale;
            if (d->clamp && (raw <= d->clamp_lo || raw >= d->clamp_hi)) {
                /* torch.clamp passes the gradient only strictly inside (and at exact equality with a bound it does too:
                 * recompute the unclamped value to decide) */
                double q = 0;
                if (d->readout == 0) { const cplx v = psi[m * d->read_stride]; q = creal(v) * creal(v) + cimag(v) * cimag(v); }
                else for (long k = 0; k < A; ++k) { const double p = creal(psi[k]) * creal(psi[k]) + cimag(psi[k]) * cimag(psi[k]); q += ((k >> (n - 1 - m)) & 1) ? -p : p; }
                const double u = q * d->post_scale;
                if (u < d->clamp_lo || u > d->clamp_hi) g = 0;
            }
            if (d->readout == 0) lam[m * d->read_stride] += g * psi[m * d->read_stride];
            else for (long k = 0; k < A; ++k) lam[k] += (((k >> (n - 1 - m)) & 1) ? -g : g) * psi[k];
        }
        for (int blk = d->n_blocks - 1; blk >= 0; --blk) {
            for (int l = d->layers_per_block - 1; l >= 0; --l) {
                const long wo = ((long)(blk * d->layers_per_block + l) * n) * 3;
                if (n > 1) {
                    const int r = (l % (n - 1)) + 1;
                    for (int i = n - 1; i >= 0; --i) {
                        if (d->imprimitive == 0) { apply_cnot(psi, n, i, (i + r) % n); apply_cnot(lam, n, i, (i + r) % n); }
                        else { apply_cz(psi, n, i, (i + r) % n); apply_cz(lam, n, i, (i + r) % n); }
                    }
                }
                for (int i = n - 1; i >= 0; --i) {
                    const double *w = weights + wo + 3 * i;
                    const double ang[3] = {remap_w(w[0], d->remap), remap_w(w[1], d->remap), remap_w(w[2], d->remap)};
                    /* Rot = RZ(omega) RY(theta) RZ(phi): undo omega, theta, phi in this order */
                    grad_w[wo + 3 * i + 2] += rot_grad_z(lam, psi, n, i) * remap_dw(w[2], d->remap);
                    apply_rz(psi, n, i, -ang[2]); apply_rz(lam, n, i, -ang[2]);
                    grad_w[wo + 3 * i + 1] += rot_grad_y(lam, psi, n, i) * remap_dw(w[1], d->remap);
                    apply_ry(psi, n, i, -ang[1]); apply_ry(lam, n, i, -ang[1]);
                    grad_w[wo + 3 * i + 0] += rot_grad_z(lam, psi, n, i) * remap_dw(w[0], d->remap);
                    apply_rz(psi, n, i, -ang[0]); apply_rz(lam, n, i, -ang[0]);
                }
            }
            if (d->enc != 0)
                for (int j = n - 1; j >= 0; --j) {
                    const double a = x[b * x_stride + j] * d->enc_scale;
                    const double gr = d->enc == 1 ? rot_grad_z(lam, psi, n, j) : rot_grad_y(lam, psi, n, j);
                    if (grad_x) grad_x[b * x_stride + j] += gr * d->enc_scale;
                    if (d->enc == 1) { apply_rz(psi, n, j, -a); apply_rz(lam, n, j, -a); }
                    else { apply_ry(psi, n, j, -a); apply_ry(lam, n, j, -a); }
                }
        }
        if (d->init == 1 && grad_x) {
            double nrm = 0, dot = 0;
            for (long k = 0; k < A; ++k) {
                const double v = k < d->n_features ? x[b * x_stride + k] + d->add_offset : d->pad_value;
                nrm += v * v;
            }
            nrm = sqrt(nrm);
            for (long k = 0; k < A; ++k) dot += creal(psi[k]) * 2 * creal(lam[k]);       /* psi is psi0 (real) again */
            for (long k = 0; k < d->n_features; ++k) grad_x[b * x_stride + k] = (2 * creal(lam[k]) - creal(psi[k]) * dot) / nrm;
        }
        free(psi);
        free(lam);
    }
    }
    for (int t = 0; t < nt; ++t)
        for (long i = 0; i < n_w; ++i) grad_w[i] += gw_all[(long)t * n_w + i];
    free(gw_all);
    free(fwd);
    free(st_all);
    return 0;
}
