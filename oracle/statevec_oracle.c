/* TEST INFRASTRUCTURE ONLY — a second, independent CPU restatement of the reference's circuit semantics, in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and the CPU-baseline legs of bench.py / scripts/ may build, load or call this file;
 * the product (qiddm_b200/) never does.  Where oracle/qiddm_oracle.py mirrors PennyLane's `default.qubit.torch` (batched tensor
 * ops, precomputed ring permutations, autograd), this file mirrors `lightning.qubit`, the device the reference's QIDDM_LL / PL /
 * QNN classes run on (nn/qdense.py:1372-1373, :1568-1569): one complex128 state vector per circuit instance, every gate
 * applied in place, one after the other, in the order the QNode tape lists them — forward only, as the reference never
 * differentiates through these circuits.  The two restatements share no code, so their agreement (tests/
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
