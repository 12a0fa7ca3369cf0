// Gate-by-gate batched state-vector kernels (forward + adjoint backward) for sm_100a.
//
// Replaces PennyLane's per-gate einsum on a (B,2,...,2) complex128 tensor
// (default.qubit.torch; nn/qdense.py:58,:465) and the per-sample lightning.qubit loop
// (nn/qdense.py:1631-1635) of the reference.
//
// Layout.  One circuit instance is simulated by a group of G = 2^n / R threads; the state
// (2^n complex fp32) lives in shared memory and is swept in register tiles of R = 2^RB
// amplitudes per thread ("views"): in view v the RB index bits [lo_v, lo_v+RB) are local to a
// thread.  The state is stored padded, slot(k) = k + (k >> RB), which makes every tile access
// `base_v + r * stride_v` (no per-access index arithmetic) and bank-conflict free in every view.
//
// Arithmetic.  Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi) is applied in factored form: per
// view, ONE complex multiply per amplitude with a table entry that merges the RZ(phi) phases of all
// local wires, a real rotation per local wire (8 FMA-pipe ops per amplitude pair instead of 16 for a
// general complex 2x2), and one multiply with the merged RZ(omega) table: 28 instead of 40 FMA-pipe
// instructions per amplitude and view at RB = 5.  The tables depend on the weights only and are built
// once per call in double precision (prepare_tables_kernel).  With a CZ entangler (diagonal) the
// RZ(omega) phases of layer l are merged into the RZ(phi) table of layer l+1.  The re-upload data gate
// RZ(s a_j) is one more per-instance diagonal table; RY(s a_j) a per-instance real rotation.  The CNOT
// ring of a StronglyEntangling layer is a GF(2)-linear index permutation folded into the last store of
// the layer; the CZ ring is a sign bit looked up from a per-thread word.
// Groups with G <= 32 share a warp (several instances per warp, __syncwarp only).
//
// Backward = adjoint method: recompute psi_final, seed lambda = dL/dpsi* from the readout, then walk
// the layers in reverse applying the inverse factors to both.  Angle gradients come straight from the
// factored form: for a diagonal RZ(alpha) on wire j, dL/dalpha = sum_k z_j(k) Im(conj(lambda_k) psi_k);
// for RY(theta), dL/dtheta = sum_pairs Re(conj(lambda_1) psi_0 - conj(lambda_0) psi_1).  Per view the
// (at most 15) partial sums are reduced over the warp with a reduce-scatter and added to per-CTA
// accumulators in shared memory; per-CTA partials are summed in double precision by
// finalize_grads_kernel (incl. the tanh / pi*tanh re-map chain rule).
#include <atomic>
#include <math_constants.h>
#include <cstdlib>
#include "qiddm_internal.h"

namespace qiddm {

namespace {

template <int NQ, int RB>
struct Cfg {
    static constexpr int A = 1 << NQ;
    static constexpr int R = 1 << RB;
    static constexpr int G = A / R;
    static constexpr int NV = (NQ + RB - 1) / RB;
    static constexpr int T = G > 128 ? G : 128;
    static constexpr int CPB = T / G;
    static constexpr int SPAN = A + (A >> RB);      // padded float2 slots of one state
    // distance between the states of a CTA's instances.  A warp holds 32 / G instances; when G <= 8 their tiles must
    // interleave over the 16 float2 bank pairs: stride = G (mod 16) makes the 32 lanes cover every bank pair exactly twice
    // (n = 6, RB = 3: 72 instead of 73 -- ncu showed 49 % conflicted shared wavefronts with the odd stride).  Otherwise odd.
    static constexpr int STRIDE = (G >= 2 && G <= 8) ? SPAN + ((G - SPAN % 16) + 16) % 16 : (SPAN | 1);
    static constexpr int LO_LAST = NQ - RB;
    static constexpr int NVF = NQ / RB;                 // full views; a partial last view follows when NQ % RB != 0
    static constexpr int Q_TAIL = RB - NQ % RB;         // first owned local bit of the partial view
    static constexpr int TAB_LAYER = NV * 2 * R + ((NQ + 1) & ~1);   // float2 per layer: [view][pre|post][R], (cos,sin)[wire]
};

// Register-tile width (log2 amplitudes per thread).  Valid widths keep every view's low bit at 0 or
// >= RB (so the padded layout is linear per view): nq == rb or nq >= 2 rb, with at most 256 threads per
// instance.  Defaults picked from B200 measurements; QIDDM_RB_FWD / QIDDM_RB_BWD override for tuning.
inline bool rb_valid(int nq, int rb) {
    if (rb < 1 || rb > 5 || rb > nq) return false;
    if (nq != rb && nq < 2 * rb) return false;
    if (rb < 3 && rb != nq) return false;
    return (nq - rb) <= 8;
}
inline int rb_default(int nq, bool bwd) {
    if (nq <= 5) return nq;
    if (nq <= 7) return 3;
    if (bwd) return nq == 9 ? 3 : (nq <= 8 ? 4 : 5);     // n = 9 adjoint: 128-thread instances beat 64 (5.5 vs 8.5 ms / 65 536)
    return nq <= 9 ? 4 : 5;
}
inline int rb_choose(int nq, bool bwd) {
    static int env_f = -2, env_b = -2;
    int &e = bwd ? env_b : env_f;
    if (e == -2) {
        const char *v = getenv(bwd ? "QIDDM_RB_BWD" : "QIDDM_RB_FWD");
        e = v ? atoi(v) : -1;
    }
    if (e > 0 && rb_valid(nq, e)) return e;
    return rb_default(nq, bwd);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {   // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// RY(theta) on an amplitude pair, cs = (cos, sin)(theta/2)
// The real 2 x 2 rotation acts on (re, im) pairs, so it maps one-to-one onto the packed fp32 instructions of sm_100
// (fma.rn.f32x2 / mul.rn.f32x2): same FMA-pipe time, half the issued instructions (the forward kernels are issue-bound).
__device__ __forceinline__ void ry_pair(float2 cs, float2 &x0, float2 &x1) {
    const float2 a = x0, b = x1;
    const float2 c2 = make_float2(cs.x, cs.x), s2 = make_float2(cs.y, cs.y), ns2 = make_float2(-cs.y, -cs.y);
    x0 = __ffma2_rn(c2, a, __fmul2_rn(ns2, b));
    x1 = __ffma2_rn(s2, a, __fmul2_rn(c2, b));
}
__device__ __forceinline__ void ry_pair_t(float2 cs, float2 &x0, float2 &x1) {   // RY(theta)^T
    const float2 a = x0, b = x1;
    const float2 c2 = make_float2(cs.x, cs.x), s2 = make_float2(cs.y, cs.y), ns2 = make_float2(-cs.y, -cs.y);
    x0 = __ffma2_rn(c2, a, __fmul2_rn(s2, b));
    x1 = __ffma2_rn(c2, b, __fmul2_rn(ns2, a));
}

template <int NQ>
__device__ __forceinline__ int ring_f(int k, int ring) {
    // image of basis index k under CNOT(i -> (i+ring) mod NQ), i = 0..NQ-1 (wire 0 = MSB)
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        int t = i + ring;
        if (t >= NQ) t -= NQ;
        k ^= ((k >> (NQ - 1 - i)) & 1) << (NQ - 1 - t);
    }
    return k;
}
template <int NQ>
__device__ __forceinline__ int cz_parity(int k, int ring) {
    const int rot = ((k << ring) | (k >> (NQ - ring))) & ((1 << NQ) - 1);
    return __popc(k & rot) & 1;
}

// Barrier over the G threads that simulate one instance: warp-level for G <= 32, a named barrier per
// instance slot when several multi-warp groups share the CTA, the CTA barrier when one group fills it.
template <int G, int T>
__device__ __forceinline__ void group_sync(int slot) {
    if (G <= 32) {
        __syncwarp();
    } else if (G == T) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(G) : "memory");
    }
}

// Sum over the G threads of a group.  G <= 32: xor shuffles.  G > 32: the group spans whole
// warps; `red` has one slot per warp of the CTA and is reused, so callers sync around it.
template <int G, int T>
__device__ __forceinline__ float group_sum(float v, float *red, int tid) {
    if (G <= 32) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int slot = tid / G;
        group_sync<G, T>(slot);
        if ((tid & 31) == 0) red[tid >> 5] = v;
        group_sync<G, T>(slot);
        float s = 0.f;
        const int w0 = slot * (G / 32);
#pragma unroll
        for (int i = 0; i < G / 32; ++i) s += red[w0 + i];
        return s;
    }
}

// Warp-wide sums of 16 values with a reduce-scatter (16 shuffles): afterwards the even lane 2i holds
// the warp sum of v[i] in v[0].
__device__ __forceinline__ void warp_reduce_scatter16(float (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, o = 16; half >= 1; half >>= 1, o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[half + i];
            const float keep = up ? v[half + i] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// S[q] = sum_r (bit q of r ? -t[r] : t[r]); destroys t
template <int RB>
__device__ __forceinline__ void signed_sums(float (&t)[1 << RB], float (&S)[RB]) {
#pragma unroll
    for (int q = 0; q < RB; ++q) {
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < ((1 << RB) >> (q + 1)); ++j) {
            const float a = t[2 * j], b = t[2 * j + 1];
            d += a - b;
            t[j] = a + b;
        }
        S[q] = d;
    }
}

struct InstanceGeom {  // where instance `cid` reads its features / writes its outputs
    long long in_base, out_base;
    int out_stride;     // stride between consecutive outputs m
    int b, y, x;        // unfold coordinates
};

__device__ __forceinline__ InstanceGeom instance_geom(const GateParams &p, long long cid, int n_in, int n_out) {
    InstanceGeom g;
    if (p.unfold) {
        const int P = p.Hout * p.Wout;
        g.b = (int)(cid / P);
        const int rem = (int)(cid - (long long)g.b * P);
        g.y = rem / p.Wout;
        g.x = rem - g.y * p.Wout;
        g.in_base = (long long)g.b * p.C * p.H * p.W;
        g.out_base = (long long)g.b * n_out * P + rem;
        g.out_stride = P;
    } else {
        g.b = g.y = g.x = 0;
        g.in_base = (cid >> p.in_shift) * n_in;
        g.out_base = cid * n_out;
        g.out_stride = 1;
    }
    return g;
}
// Per-CTA feature table for the fused patch-unfold: feature k = (ch, ky, kx) -> offset inside the image
// relative to the patch origin, and (ky - pad_h, kx - pad_w) packed for the bounds test.
__device__ __forceinline__ void build_unfold_table(const GateParams &p, int *foff, int *fyx, int tid, int nthreads) {
    const int kk = p.kh * p.kw;
    for (int k = tid; k < p.n_features; k += nthreads) {
        const int ch = k / kk;
        const int r = k - ch * kk;
        const int ky = r / p.kw;
        const int kx = r - ky * p.kw;
        foff[k] = (ch * p.H + (ky - p.ph)) * p.W + (kx - p.pw);
        fyx[k] = ((ky - p.ph) << 16) | ((kx - p.pw) & 0xffff);
    }
}
// offset of feature k inside the image, or -1 when it falls in the zero padding
__device__ __forceinline__ long long unfold_offset(const GateParams &p, const InstanceGeom &g, const int *foff,
                                                   const int *fyx, int k) {
    const int yx = fyx[k];
    const int yy = g.y + (yx >> 16), xx = g.x + (int)(short)(yx & 0xffff);
    if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) return -1;
    return g.in_base + (long long)g.y * p.W + g.x + foff[k];
}

template <int RB>
__device__ __forceinline__ int slot_of(int k) {     // padded position of amplitude k
    return k + (k >> RB);
}
// index of a thread's first tile element in the view whose local bits start at `lo`
template <int RB>
__device__ __forceinline__ int tile_k0(int g, int lo) {
    return ((g >> lo) << (lo + RB)) | (g & ((1 << lo) - 1));
}

template <int NQ, int RB>
struct View {
    int lo, q_lo, q_hi, base, stride;
    __device__ __forceinline__ View(int v, int g) {
        constexpr int LO_LAST = NQ - RB;
        lo = (v * RB < LO_LAST) ? v * RB : LO_LAST;
        q_lo = v * RB - lo;                                          // first local bit this view owns
        q_hi = (((v + 1) * RB < NQ) ? (v + 1) * RB : NQ) - lo;       // one past the last
        base = slot_of<RB>(tile_k0<RB>(g, lo));
        stride = lo == 0 ? 1 : ((1 << lo) + (1 << (lo - RB)));
    }
};


// ---------------------------------------------------------------------------------------------
// View cores: everything that happens to a register tile between its load and its store.  Q_LO is the
// first local bit the view owns (0 for a full view; the only partial view is the last one when
// NQ % RB != 0), so the set of active local bits is known at compile time.
// ---------------------------------------------------------------------------------------------
enum { ENCL_NONE = 0, ENCL_RZ = 1, ENCL_RY = 2 };

template <int RB, int Q>
__device__ __forceinline__ void ry_all(float2 cs, float2 (&s)[1 << RB]) {
#pragma unroll
    for (int j = 0; j < (1 << RB) / 2; ++j) {
        const int r0 = ((j >> Q) << (Q + 1)) | (j & ((1 << Q) - 1));
        ry_pair(cs, s[r0], s[r0 | (1 << Q)]);
    }
}
// theta gradient from the post-rotation states, then RY^T on both
template <int RB, int Q>
__device__ __forceinline__ float ry_all_bwd(float2 cs, float2 (&s)[1 << RB], float2 (&l)[1 << RB]) {
    // Re(conj(l1) s0) - Re(conj(l0) s1) summed over the tile: two packed accumulators (re and im products side by side)
    float2 gp = make_float2(0.f, 0.f), gn = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < (1 << RB) / 2; ++j) {
        const int r0 = ((j >> Q) << (Q + 1)) | (j & ((1 << Q) - 1));
        const int r1 = r0 | (1 << Q);
        gp = __ffma2_rn(l[r1], s[r0], gp);
        gn = __ffma2_rn(l[r0], s[r1], gn);
        ry_pair_t(cs, s[r0], s[r1]);
        ry_pair_t(cs, l[r0], l[r1]);
    }
    return (gp.x + gp.y) - (gn.x + gn.y);
}
// s[r] *= table[r] (CONJ: conj(table[r])); two entries per 16-byte uniform load
template <int RB, bool CONJ, bool GLOBAL>
__device__ __forceinline__ void diag_mul(float2 (&s)[1 << RB], const float2 *t) {
    if ((1 << RB) >= 2) {
#pragma unroll
        for (int j = 0; j < (1 << RB) / 2; ++j) {
            const float4 e = GLOBAL ? __ldg(reinterpret_cast<const float4 *>(t) + j) : reinterpret_cast<const float4 *>(t)[j];
            s[2 * j] = CONJ ? cmul_conj(s[2 * j], make_float2(e.x, e.y)) : cmul(s[2 * j], make_float2(e.x, e.y));
            s[2 * j + 1] = CONJ ? cmul_conj(s[2 * j + 1], make_float2(e.z, e.w)) : cmul(s[2 * j + 1], make_float2(e.z, e.w));
        }
    }
}
template <int RB, bool GLOBAL>
__device__ __forceinline__ void diag_mul2_conj(float2 (&s)[1 << RB], float2 (&l)[1 << RB], const float2 *t) {
#pragma unroll
    for (int j = 0; j < (1 << RB) / 2; ++j) {
        const float4 e = GLOBAL ? __ldg(reinterpret_cast<const float4 *>(t) + j) : reinterpret_cast<const float4 *>(t)[j];
        const float2 e0 = make_float2(e.x, e.y), e1 = make_float2(e.z, e.w);
        s[2 * j] = cmul_conj(s[2 * j], e0);
        l[2 * j] = cmul_conj(l[2 * j], e0);
        s[2 * j + 1] = cmul_conj(s[2 * j + 1], e1);
        l[2 * j + 1] = cmul_conj(l[2 * j + 1], e1);
    }
}

// forward: [RY(s a)] -> pre phases -> [RZ(s a) phases] -> RY(theta) per owned wire -> [post phases]
template <int NQ, int RB, int Q_LO>
__device__ __forceinline__ void view_fwd(float2 (&s)[1 << RB], const float2 *tv, const float2 *cst, int lo,
                                         bool has_post, int encl, const float2 *ev, const float2 *ep) {
    constexpr int R = 1 << RB;
    if (encl == ENCL_RY) {
        if (Q_LO <= 0 && RB > 0) ry_all<RB, 0>(ep[NQ - 1 - lo], s);
        if (Q_LO <= 1 && RB > 1) ry_all<RB, (RB > 1 ? 1 : 0)>(ep[NQ - 2 - lo], s);
        if (Q_LO <= 2 && RB > 2) ry_all<RB, (RB > 2 ? 2 : 0)>(ep[NQ - 3 - lo], s);
        if (Q_LO <= 3 && RB > 3) ry_all<RB, (RB > 3 ? 3 : 0)>(ep[NQ - 4 - lo], s);
        if (Q_LO <= 4 && RB > 4) ry_all<RB, (RB > 4 ? 4 : 0)>(ep[NQ - 5 - lo], s);
    }
    diag_mul<RB, false, true>(s, tv);
    if (encl == ENCL_RZ) diag_mul<RB, false, false>(s, ev);
    if (Q_LO <= 0 && RB > 0) ry_all<RB, 0>(__ldg(cst + (NQ - 1 - lo)), s);
    if (Q_LO <= 1 && RB > 1) ry_all<RB, (RB > 1 ? 1 : 0)>(__ldg(cst + (NQ - 2 - lo)), s);
    if (Q_LO <= 2 && RB > 2) ry_all<RB, (RB > 2 ? 2 : 0)>(__ldg(cst + (NQ - 3 - lo)), s);
    if (Q_LO <= 3 && RB > 3) ry_all<RB, (RB > 3 ? 3 : 0)>(__ldg(cst + (NQ - 4 - lo)), s);
    if (Q_LO <= 4 && RB > 4) ry_all<RB, (RB > 4 ? 4 : 0)>(__ldg(cst + (NQ - 5 - lo)), s);
    if (has_post) diag_mul<RB, false, true>(s, tv + R);
}

// backward: inverse of view_fwd on psi (s) and lambda (l); gv[q*3 + {0,1,2}] = this thread's partial
// d/dphi, d/dtheta, d/domega of the wire at local bit q; ga[] += d/d(s a_wire) on re-upload layers.
template <int NQ, int RB, int Q_LO>
__device__ __forceinline__ void view_bwd(float2 (&s)[1 << RB], float2 (&l)[1 << RB], float (&gv)[16],
                                         float (&ga)[NQ], const float2 *tv, const float2 *cst, int lo,
                                         bool has_post, int encl, const float2 *ev, const float2 *ep) {
    constexpr int R = 1 << RB;
#pragma unroll
    for (int i = 0; i < 16; ++i) gv[i] = 0.f;
    if (has_post) {
        float t[R], S[RB];
#pragma unroll
        for (int r = 0; r < R; ++r) t[r] = l[r].x * s[r].y - l[r].y * s[r].x;   // Im(conj(l) s)
        signed_sums<RB>(t, S);
#pragma unroll
        for (int q = Q_LO; q < RB; ++q) gv[q * 3 + 2] = S[q];
        diag_mul2_conj<RB, true>(s, l, tv + R);
    }
    if (Q_LO <= 4 && RB > 4) gv[4 * 3 + 1] = ry_all_bwd<RB, (RB > 4 ? 4 : 0)>(__ldg(cst + (NQ - 5 - lo)), s, l);
    if (Q_LO <= 3 && RB > 3) gv[3 * 3 + 1] = ry_all_bwd<RB, (RB > 3 ? 3 : 0)>(__ldg(cst + (NQ - 4 - lo)), s, l);
    if (Q_LO <= 2 && RB > 2) gv[2 * 3 + 1] = ry_all_bwd<RB, (RB > 2 ? 2 : 0)>(__ldg(cst + (NQ - 3 - lo)), s, l);
    if (Q_LO <= 1 && RB > 1) gv[1 * 3 + 1] = ry_all_bwd<RB, (RB > 1 ? 1 : 0)>(__ldg(cst + (NQ - 2 - lo)), s, l);
    if (Q_LO <= 0 && RB > 0) gv[0 * 3 + 1] = ry_all_bwd<RB, 0>(__ldg(cst + (NQ - 1 - lo)), s, l);
    {
        float t[R], S[RB];
#pragma unroll
        for (int r = 0; r < R; ++r) t[r] = l[r].x * s[r].y - l[r].y * s[r].x;
        signed_sums<RB>(t, S);
#pragma unroll
        for (int q = Q_LO; q < RB; ++q) gv[q * 3] = S[q];
        if (encl == ENCL_RZ) {
#pragma unroll
            for (int q = Q_LO; q < RB; ++q) {
                const int wire = NQ - 1 - (lo + q);
#pragma unroll
                for (int j = 0; j < NQ; ++j) ga[j] += (j == wire) ? S[q] : 0.f;
            }
            diag_mul2_conj<RB, false>(s, l, ev);
        }
        diag_mul2_conj<RB, true>(s, l, tv);
    }
    if (encl == ENCL_RY) {
        float ge[RB];
#pragma unroll
        for (int q = 0; q < RB; ++q) ge[q] = 0.f;
        if (Q_LO <= 4 && RB > 4) ge[(RB > 4 ? 4 : 0)] = ry_all_bwd<RB, (RB > 4 ? 4 : 0)>(ep[NQ - 5 - lo], s, l);
        if (Q_LO <= 3 && RB > 3) ge[(RB > 3 ? 3 : 0)] = ry_all_bwd<RB, (RB > 3 ? 3 : 0)>(ep[NQ - 4 - lo], s, l);
        if (Q_LO <= 2 && RB > 2) ge[(RB > 2 ? 2 : 0)] = ry_all_bwd<RB, (RB > 2 ? 2 : 0)>(ep[NQ - 3 - lo], s, l);
        if (Q_LO <= 1 && RB > 1) ge[(RB > 1 ? 1 : 0)] = ry_all_bwd<RB, (RB > 1 ? 1 : 0)>(ep[NQ - 2 - lo], s, l);
        if (Q_LO <= 0 && RB > 0) ge[0] = ry_all_bwd<RB, 0>(ep[NQ - 1 - lo], s, l);
#pragma unroll
        for (int q = Q_LO; q < RB; ++q) {
            const int wire = NQ - 1 - (lo + q);
#pragma unroll
            for (int j = 0; j < NQ; ++j) ga[j] += (j == wire) ? ge[q] : 0.f;
        }
    }
}

// resident CTAs per SM the register allocation aims for (more warps hide the shared-memory and table-load latency)
template <int NQ, int RB, bool BWD>
constexpr int min_blocks() {
    constexpr int T = Cfg<NQ, RB>::T;
    constexpr int want_regs = BWD ? (RB >= 5 ? 255 : RB == 4 ? 168 : 128) : (RB >= 5 ? 128 : RB == 4 ? 96 : 64);
    constexpr int b = 65536 / (T * want_regs);
    return b < 1 ? 1 : b;
}

// RES ("resident" schedule; NQ == 2 RB, diagonal entangler): a tile stays in registers across a layer
// boundary -- RY(l-1) on its wires, the whole boundary diagonal (own wires from a table, the other half's
// wires as one per-thread scalar, CZ signs, re-upload phases), RY(l) -- so the state crosses shared memory
// once per layer instead of twice and only one phase table is read per layer.
template <int NQ, int RB, bool BWD, bool RES>
__global__ void __launch_bounds__(Cfg<NQ, RB>::T, min_blocks<NQ, RB, BWD>()) gate_kernel(const GateParams p) {
    using C = Cfg<NQ, RB>;
    constexpr int A = C::A, R = C::R, G = C::G, NV = C::NV, T = C::T, CPB = C::CPB, STRIDE = C::STRIDE;
    constexpr int LO_LAST = C::LO_LAST, TAB_LAYER = C::TAB_LAYER, NVF = C::NVF, Q_TAIL = C::Q_TAIL;
    constexpr int NRING = NQ > 1 ? NQ - 1 : 1;
    constexpr int TAB_RES = 2 * R + ((NQ + 1) & ~1);   // RES: float2 per boundary: local[R], scalar[G = R], (cos,sin)[wire]

    extern __shared__ float4 smem_f4[];
    float *sm = reinterpret_cast<float *>(smem_f4);
    // carve-up (float offsets; the per-instance phase tables are read as float4, so they come 16-byte aligned)
    const int n_grad = RES ? (p.n_rot + NQ) * 2 : p.n_rot * 3;   // angle-gradient sums kept per CTA
    const int n_acc = (n_grad + 3) & ~3;
    float *acc_s = sm;                                               // [n_acc] (BWD)
    float2 *et_all = reinterpret_cast<float2 *>(acc_s + (BWD ? n_acc : 0));   // [CPB][NV][R] RZ re-upload phase tables
    float2 *psi_all = et_all + (p.enc == QIDDM_ENC_RZ ? CPB * NV * R : 0);
    float2 *lam_all = psi_all + CPB * STRIDE;
    float2 *ep_all = BWD ? lam_all + CPB * STRIDE : lam_all;         // [CPB][NQ] (cos, sin)(s a_j / 2)
    float *red = reinterpret_cast<float *>(ep_all + CPB * NQ);       // [T/32]
    unsigned int *czw = reinterpret_cast<unsigned int *>(red + T / 32);                          // [NRING][G] (RES: [2][NRING][G])
    int *foff = reinterpret_cast<int *>(czw + (p.imprimitive == QIDDM_IMP_CZ ? (RES ? 2 : 1) * NRING * G : 0));   // [2][n_features] (unfold)
    int *fyx = foff + p.n_features;
    unsigned short *ftab = reinterpret_cast<unsigned short *>(foff + (p.unfold ? 2 * p.n_features : 0));

    const int tid = threadIdx.x, lane = tid & 31;
    const int slot = tid / G, g = tid % G;
    const float2 *tab = reinterpret_cast<const float2 *>(p.gates);
    const int n_layers = p.n_blocks * p.layers;

    if (BWD)
        for (int i = tid; i < n_acc; i += T) acc_s[i] = 0.f;
    if (p.unfold) build_unfold_table(p, foff, fyx, tid, T);
    if (NQ > 1) {
        if (p.imprimitive == QIDDM_IMP_CNOT) {
            for (int i = tid; i < NRING * R; i += T) {
                const int ring = i / R + 1, r = i % R;
                ftab[i] = (unsigned short)ring_f<NQ>(r << LO_LAST, ring);
            }
        } else {
            // bit r of czw[ring-1][g]: parity of the CZ ring on amplitude g | r << LO_LAST (last view's tile);
            // RES: czw[v][ring-1][g] for the tiles of view 0 (g << RB | r) and view 1 (r << RB | g)
            for (int i = tid; i < (RES ? 2 : 1) * NRING * G; i += T) {
                const int v = i / (NRING * G), ring = (i / G) % NRING + 1, gg = i % G;
                unsigned int w = 0;
                for (int r = 0; r < R; ++r) {
                    const int k = RES ? (v == 0 ? ((gg << RB) | r) : ((r << RB) | gg)) : (gg | (r << LO_LAST));
                    w |= (unsigned int)cz_parity<NQ>(k, ring) << r;
                }
                czw[i] = w;
            }
        }
    }
    __syncthreads();

    float2 *psi = psi_all + slot * STRIDE;
    float2 *lam = lam_all + slot * STRIDE;
    float2 *ep = ep_all + slot * NQ;
    float2 *et = et_all + slot * NV * R;
    const int n_in = p.init == QIDDM_INIT_AMPLITUDE ? p.n_features : (p.enc != QIDDM_ENC_NONE ? NQ : 0);
    const int n_out = p.readout == QIDDM_READ_PROBS ? p.read_count : (p.readout == QIDDM_READ_EXPVAL_Z ? NQ : 2 * A);
    const bool post_all = !p.merge_post;    // merge_post (CZ entangler): RZ(omega) phases live in the next layer's table

    for (long long base = (long long)blockIdx.x * CPB; base < p.B; base += (long long)gridDim.x * CPB) {
        const long long cid = base + slot;
        const bool active = cid < p.B;
        const InstanceGeom geo = instance_geom(p, active ? cid : 0, n_in, n_out);

        // ---------------------------------------------------------------- initial state
        float inv_norm = 1.f;
        if (p.init == QIDDM_INIT_AMPLITUDE) {
            // un-normalised features go to shared memory first (each thread re-reads only its own slots)
            float ss = 0.f;
#pragma unroll 1
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                float v = p.pad_value;
                if (k < p.n_features) {
                    float x = 0.f;
                    if (active) {
                        if (p.unfold) {
                            const long long off = unfold_offset(p, geo, foff, fyx, k);
                            // io64: the image is float64 (the reference's UNet), read in place of a cast kernel
                            x = off < 0 ? 0.f : (p.io64 ? (float)__ldg(reinterpret_cast<const double *>(p.in) + off) : __ldg(p.in + off));
                        } else {
                            x = __ldg(p.in + geo.in_base + k);
                        }
                    }
                    v = x + p.add_offset;
                }
                psi[slot_of<RB>(k)] = make_float2(v, 0.f);
                ss += v * v;
            }
            ss = group_sum<G, T>(ss, red, tid);
            inv_norm = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
#pragma unroll
            for (int i = 0; i < R; ++i) psi[slot_of<RB>(g + i * G)].x *= inv_norm;
        } else if (p.init == QIDDM_INIT_STATE) {
            const float2 *src = reinterpret_cast<const float2 *>(p.init_state) + (active ? cid : 0) * A;
#pragma unroll 4
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                psi[slot_of<RB>(k)] = active ? __ldg(src + k) : make_float2(0.f, 0.f);
            }
        } else {
            int start = 0;
            if (p.init == QIDDM_INIT_BASIS) start = p.basis ? (active ? p.basis[cid] : 0) : (int)(cid & (A - 1));
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                psi[slot_of<RB>(k)] = make_float2(k == start ? 1.f : 0.f, 0.f);
            }
        }
        if (p.enc != QIDDM_ENC_NONE) {
            for (int j = g; j < NQ; j += G) {
                const float a = active ? p.enc_scale * __ldg(p.in + geo.in_base + j) : 0.f;
                float s, c;
                sincosf(0.5f * a, &s, &c);
                ep[j] = make_float2(c, s);
            }
            if (p.enc == QIDDM_ENC_RZ) {
                group_sync<G, T>(slot);
                // et[v][r] = prod over the wires view v owns of e^{-+ i s a / 2} (bit 0: -, bit 1: +)
                for (int i = g; i < NV * R; i += G) {
                    const int v = i / R, r = i % R;
                    const View<NQ, RB> vw(v, 0);
                    float2 ph = make_float2(1.f, 0.f);
                    for (int q = vw.q_lo; q < vw.q_hi; ++q) {
                        const float2 cs = ep[NQ - 1 - (vw.lo + q)];
                        ph = cmul(ph, make_float2(cs.x, ((r >> q) & 1) ? cs.y : -cs.y));
                    }
                    et[i] = ph;
                }
            }
        }
        group_sync<G, T>(slot);

        // ---------------------------------------------------------------- forward sweep
        const int enc_mode = p.enc == QIDDM_ENC_RZ ? ENCL_RZ : (p.enc == QIDDM_ENC_RY ? ENCL_RY : ENCL_NONE);
        const bool apply_last = p.readout == QIDDM_READ_STATE;   // a trailing diagonal only matters for a state readout
        if (BWD && p.state != nullptr) {
            // adjoint with psi_final kept by the forward launch (same kernel family, same arithmetic): no recomputation
            const float2 *src = reinterpret_cast<const float2 *>(p.state) + (active ? cid : 0) * A;
#pragma unroll 4
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                psi[slot_of<RB>(k)] = active ? __ldg(src + k) : make_float2(0.f, 0.f);
            }
            group_sync<G, T>(slot);
        } else if constexpr (RES) {
#pragma unroll 1
            for (int j = 0; j <= n_layers; ++j) {
                const int v = j & 1;
                float2 *pp = psi + (v == 0 ? slot_of<RB>(g << RB) : g);
                const int stride = v == 0 ? 1 : R + 1;
                const int w0 = NQ - 1 - v * RB;                       // wire of local bit 0 in this view
                const float2 *tb = tab + (size_t)j * TAB_RES;
                float2 s[R];
#pragma unroll
                for (int r = 0; r < R; ++r) s[r] = pp[r * stride];
                if (j > 0) {
                    const float2 *cst = tb - TAB_RES + 2 * R;
                    if (RB > 0) ry_all<RB, 0>(__ldg(cst + w0), s);
                    if (RB > 1) ry_all<RB, (RB > 1 ? 1 : 0)>(__ldg(cst + w0 - 1), s);
                    if (RB > 2) ry_all<RB, (RB > 2 ? 2 : 0)>(__ldg(cst + w0 - 2), s);
                    if (RB > 3) ry_all<RB, (RB > 3 ? 3 : 0)>(__ldg(cst + w0 - 3), s);
                    if (RB > 4) ry_all<RB, (RB > 4 ? 4 : 0)>(__ldg(cst + w0 - 4), s);
                }
                if (j < n_layers || apply_last) {
                    diag_mul<RB, false, true>(s, tb);
                    float2 x = __ldg(tb + R + g);
                    if (enc_mode == ENCL_RZ && j < n_layers && j % p.layers == 0) {
                        diag_mul<RB, false, false>(s, et + v * R);
                        x = cmul(x, et[(1 - v) * R + g]);
                    }
                    if (j > 0 && NQ > 1) {
                        const unsigned int w = czw[(v * NRING + ((j - 1) % p.layers) % NRING) * G + g];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const unsigned int sg = (w << (31 - r)) & 0x80000000u;
                            const float2 y = cmul(s[r], x);
                            s[r] = make_float2(__uint_as_float(__float_as_uint(y.x) ^ sg), __uint_as_float(__float_as_uint(y.y) ^ sg));
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) s[r] = cmul(s[r], x);
                    }
                }
                if (j < n_layers) {
                    const float2 *cst = tb + 2 * R;
                    if (RB > 0) ry_all<RB, 0>(__ldg(cst + w0), s);
                    if (RB > 1) ry_all<RB, (RB > 1 ? 1 : 0)>(__ldg(cst + w0 - 1), s);
                    if (RB > 2) ry_all<RB, (RB > 2 ? 2 : 0)>(__ldg(cst + w0 - 2), s);
                    if (RB > 3) ry_all<RB, (RB > 3 ? 3 : 0)>(__ldg(cst + w0 - 3), s);
                    if (RB > 4) ry_all<RB, (RB > 4 ? 4 : 0)>(__ldg(cst + w0 - 4), s);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) pp[r * stride] = s[r];
                group_sync<G, T>(slot);
            }
        } else {
#pragma unroll 1
        for (int li = 0; li < n_layers; ++li) {
            const int layer = li % p.layers;
            const int ring = NQ > 1 ? (layer % NRING) + 1 : 0;
            const int encl = layer == 0 ? enc_mode : ENCL_NONE;
            const bool has_post = post_all || li == n_layers - 1;
            const float2 *tl = tab + (size_t)li * TAB_LAYER;
            const float2 *cst = tl + NV * 2 * R;
#pragma unroll 1
            for (int v = 0; v < NV; ++v) {
                const View<NQ, RB> vw(v, g);
                float2 *pp = psi + vw.base;
                float2 s[R];
#pragma unroll
                for (int r = 0; r < R; ++r) s[r] = pp[r * vw.stride];
                if (NVF == NV || v < NVF) view_fwd<NQ, RB, 0>(s, tl + v * 2 * R, cst, vw.lo, has_post, encl, et + v * R, ep);
                else view_fwd<NQ, RB, (NVF == NV ? 0 : Q_TAIL)>(s, tl + v * 2 * R, cst, vw.lo, has_post, encl, et + v * R, ep);
                if (v == NV - 1 && NQ > 1) {
                    if (p.imprimitive == QIDDM_IMP_CNOT) {
                        if (G > 1) group_sync<G, T>(slot);  // every tile is in registers before the scatter
                        const int fk = ring_f<NQ>(g, ring);  // the last view's tile starts at amplitude g
                        const unsigned short *ft = ftab + (ring - 1) * R;
#pragma unroll
                        for (int r = 0; r < R; ++r) psi[slot_of<RB>(fk ^ ft[r])] = s[r];
                    } else {
                        const unsigned int w = czw[(ring - 1) * G + g];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const unsigned int sg = (w << (31 - r)) & 0x80000000u;
                            pp[r * vw.stride] = make_float2(__uint_as_float(__float_as_uint(s[r].x) ^ sg),
                                                            __uint_as_float(__float_as_uint(s[r].y) ^ sg));
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) pp[r * vw.stride] = s[r];
                }
                group_sync<G, T>(slot);
            }
        }
        }

        // ---------------------------------------------------------------- readout
        if (!BWD && p.state != nullptr && active) {
            float2 *dst = reinterpret_cast<float2 *>(p.state) + cid * A;
#pragma unroll 4
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                dst[k] = psi[slot_of<RB>(k)];
            }
        }
        if (!BWD) {
            if (p.readout == QIDDM_READ_PROBS) {
                for (int m = g; m < p.read_count; m += G) {
                    const float2 a = psi[slot_of<RB>(m * p.read_stride)];
                    float v = p.post_scale * (a.x * a.x + a.y * a.y);
                    if (p.clamp) v = fminf(fmaxf(v, p.clamp_lo), p.clamp_hi);
                    if (active) {
                        const long long oi = geo.out_base + (long long)m * geo.out_stride;
                        if (p.unfold && p.io64) reinterpret_cast<double *>(p.out)[oi] = (double)v;
                        else p.out[oi] = v;
                    }
                }
            } else if (p.readout == QIDDM_READ_EXPVAL_Z) {
                float ez[NQ];
#pragma unroll
                for (int j = 0; j < NQ; ++j) ez[j] = 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int k = (g << RB) | r;
                    const float2 a = psi[slot_of<RB>(k)];
                    const float pr = a.x * a.x + a.y * a.y;
#pragma unroll
                    for (int j = 0; j < NQ; ++j) ez[j] += ((k >> (NQ - 1 - j)) & 1) ? -pr : pr;
                }
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const float t = group_sum<G, T>(ez[j], red, tid);
                    if (g == 0 && active) p.out[geo.out_base + j] = p.post_scale * t;
                }
            } else {
                for (int k = g; k < A; k += G)
                    if (active) reinterpret_cast<float2 *>(p.out + geo.out_base)[k] = psi[slot_of<RB>(k)];
            }
            group_sync<G, T>(slot);
            continue;
        }

        // ================================================================ backward
        if (BWD) {
            // seed lambda = dL/dpsi*
            float go[NQ];
            if (p.readout == QIDDM_READ_EXPVAL_Z) {
#pragma unroll
                for (int j = 0; j < NQ; ++j) go[j] = active ? __ldg(p.grad_out + geo.out_base + j) : 0.f;
            }
#pragma unroll 1
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                const float2 a = psi[slot_of<RB>(k)];
                float2 l = make_float2(0.f, 0.f);
                if (p.readout == QIDDM_READ_PROBS) {
                    const int m = k / p.read_stride;
                    if (m * p.read_stride == k && m < p.read_count && active) {
                        const float v = p.post_scale * (a.x * a.x + a.y * a.y);
                        const bool pass = !p.clamp || (v >= p.clamp_lo && v <= p.clamp_hi);
                        const long long gi_ = geo.out_base + (long long)m * geo.out_stride;
                        const float go_ = (p.unfold && p.io64) ? (float)__ldg(reinterpret_cast<const double *>(p.grad_out) + gi_)
                                                               : __ldg(p.grad_out + gi_);
                        const float c = pass ? p.post_scale * go_ : 0.f;
                        l = make_float2(c * a.x, c * a.y);
                    }
                } else if (p.readout == QIDDM_READ_EXPVAL_Z) {
                    float c = 0.f;
#pragma unroll
                    for (int j = 0; j < NQ; ++j) c += ((k >> (NQ - 1 - j)) & 1) ? -go[j] : go[j];
                    c *= p.post_scale;
                    l = make_float2(c * a.x, c * a.y);
                } else if (active) {
                    const float2 gq = reinterpret_cast<const float2 *>(p.grad_out + geo.out_base)[k];
                    l = make_float2(0.5f * gq.x, 0.5f * gq.y);
                }
                lam[slot_of<RB>(k)] = l;
            }
            group_sync<G, T>(slot);

            float ga[NQ];  // per-instance d/d(s a_wire), partial over this thread's amplitudes
#pragma unroll
            for (int j = 0; j < NQ; ++j) ga[j] = 0.f;

            if constexpr (RES) {
                // per-lane destination of a reduce-scattered gradient: lane 2i holds value i = kind * RB + q
                auto push = [&](float (&gv)[16], int v, int jt, int ja) {   // jt: layer of the theta values, ja: boundary of the alphas (-1: none)
                    warp_reduce_scatter16(gv, lane);
                    if ((lane & 1) == 0) {
                        const int i = lane >> 1, kind = i / RB, q = i - kind * RB;
                        const int wire = NQ - 1 - (v * RB + q);
                        if (kind == 0 && jt >= 0) atomicAdd(acc_s + (jt * NQ + wire) * 2 + 1, gv[0]);
                        if (kind == 1 && ja >= 0) atomicAdd(acc_s + (ja * NQ + wire) * 2, gv[0]);
                    }
                };
                auto alpha_sums = [&](float2 (&s)[R], float2 (&l)[R], float (&gv)[16], int v, bool enc_b) {
                    float t[R], S[RB];
#pragma unroll
                    for (int r = 0; r < R; ++r) t[r] = l[r].x * s[r].y - l[r].y * s[r].x;   // Im(conj(l) s)
                    signed_sums<RB>(t, S);
#pragma unroll
                    for (int q = 0; q < RB; ++q) gv[RB + q] = S[q];
                    if (enc_b) {
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            const int wire = NQ - 1 - (v * RB + q);
#pragma unroll
                            for (int jj = 0; jj < NQ; ++jj) ga[jj] += (jj == wire) ? S[q] : 0.f;
                        }
                    }
                };
                if (apply_last) {
                    // the trailing diagonal's phases on the wires that are thread bits in residency n_layers
                    const int v = (n_layers + 1) & 1;
                    const float2 *pp = psi + (v == 0 ? slot_of<RB>(g << RB) : g), *lp = lam + (v == 0 ? slot_of<RB>(g << RB) : g);
                    const int stride = v == 0 ? 1 : R + 1;
                    float2 s[R], l[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        s[r] = pp[r * stride];
                        l[r] = lp[r * stride];
                    }
                    float gv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) gv[i] = 0.f;
                    alpha_sums(s, l, gv, v, false);
                    push(gv, v, -1, n_layers);
                }
#pragma unroll 1
                for (int j = n_layers; j >= 0; --j) {
                    const int v = j & 1;
                    float2 *pp = psi + (v == 0 ? slot_of<RB>(g << RB) : g), *lp = lam + (v == 0 ? slot_of<RB>(g << RB) : g);
                    const int stride = v == 0 ? 1 : R + 1;
                    const int w0 = NQ - 1 - v * RB;
                    const float2 *tb = tab + (size_t)j * TAB_RES;
                    float2 s[R], l[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        s[r] = pp[r * stride];
                        l[r] = lp[r * stride];
                    }
                    float gv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) gv[i] = 0.f;
                    if (j < n_layers) {
                        const float2 *cst = tb + 2 * R;
                        if (RB > 4) gv[(RB > 4 ? 4 : 0)] = ry_all_bwd<RB, (RB > 4 ? 4 : 0)>(__ldg(cst + w0 - 4), s, l);
                        if (RB > 3) gv[(RB > 3 ? 3 : 0)] = ry_all_bwd<RB, (RB > 3 ? 3 : 0)>(__ldg(cst + w0 - 3), s, l);
                        if (RB > 2) gv[(RB > 2 ? 2 : 0)] = ry_all_bwd<RB, (RB > 2 ? 2 : 0)>(__ldg(cst + w0 - 2), s, l);
                        if (RB > 1) gv[(RB > 1 ? 1 : 0)] = ry_all_bwd<RB, (RB > 1 ? 1 : 0)>(__ldg(cst + w0 - 1), s, l);
                        if (RB > 0) gv[0] = ry_all_bwd<RB, 0>(__ldg(cst + w0), s, l);
                    }
                    const bool has_diag = j < n_layers || apply_last;
                    if (has_diag) {
                        const bool enc_b = enc_mode == ENCL_RZ && j < n_layers && j % p.layers == 0;
                        alpha_sums(s, l, gv, v, enc_b);
                        float2 x = __ldg(tb + R + g);
                        if (enc_b) {
                            diag_mul2_conj<RB, false>(s, l, et + v * R);
                            x = cmul(x, et[(1 - v) * R + g]);
                        }
                        unsigned int w = 0;
                        if (j > 0 && NQ > 1) w = czw[(v * NRING + ((j - 1) % p.layers) % NRING) * G + g];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const unsigned int sg = (w << (31 - r)) & 0x80000000u;
                            const float2 ys = cmul_conj(s[r], x), yl = cmul_conj(l[r], x);
                            s[r] = make_float2(__uint_as_float(__float_as_uint(ys.x) ^ sg), __uint_as_float(__float_as_uint(ys.y) ^ sg));
                            l[r] = make_float2(__uint_as_float(__float_as_uint(yl.x) ^ sg), __uint_as_float(__float_as_uint(yl.y) ^ sg));
                        }
                        diag_mul2_conj<RB, true>(s, l, tb);
                    }
                    push(gv, v, j < n_layers ? j : -1, has_diag ? j : -1);
                    if (j > 0) {
                        const float2 *cst = tb - TAB_RES + 2 * R;
#pragma unroll
                        for (int i = 0; i < 16; ++i) gv[i] = 0.f;
                        if (RB > 4) gv[(RB > 4 ? 4 : 0)] = ry_all_bwd<RB, (RB > 4 ? 4 : 0)>(__ldg(cst + w0 - 4), s, l);
                        if (RB > 3) gv[(RB > 3 ? 3 : 0)] = ry_all_bwd<RB, (RB > 3 ? 3 : 0)>(__ldg(cst + w0 - 3), s, l);
                        if (RB > 2) gv[(RB > 2 ? 2 : 0)] = ry_all_bwd<RB, (RB > 2 ? 2 : 0)>(__ldg(cst + w0 - 2), s, l);
                        if (RB > 1) gv[(RB > 1 ? 1 : 0)] = ry_all_bwd<RB, (RB > 1 ? 1 : 0)>(__ldg(cst + w0 - 1), s, l);
                        if (RB > 0) gv[0] = ry_all_bwd<RB, 0>(__ldg(cst + w0), s, l);
                        // the boundary j-1 phases of this view's wires were applied as a per-thread scalar in
                        // residency j-1: Im(lambda^H Z_w psi) is unchanged by gates on other wires, so take it here
                        alpha_sums(s, l, gv, v, enc_mode == ENCL_RZ && (j - 1) % p.layers == 0);
                        push(gv, v, j - 1, j - 1);
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        pp[r * stride] = s[r];
                        lp[r * stride] = l[r];
                    }
                    group_sync<G, T>(slot);
                }
            } else {
#pragma unroll 1
            for (int li = n_layers - 1; li >= 0; --li) {
                const int layer = li % p.layers;
                const int ring = NQ > 1 ? (layer % NRING) + 1 : 0;
                const int encl = layer == 0 ? enc_mode : ENCL_NONE;
                const bool has_post = post_all || li == n_layers - 1;
                const float2 *tl = tab + (size_t)li * TAB_LAYER;
                const float2 *cst = tl + NV * 2 * R;
#pragma unroll 1
                for (int v = NV - 1; v >= 0; --v) {
                    const View<NQ, RB> vw(v, g);
                    float2 *pp = psi + vw.base, *lp = lam + vw.base;
                    float2 s[R], l[R];
                    if (v == NV - 1 && NQ > 1) {
                        if (p.imprimitive == QIDDM_IMP_CNOT) {
                            // pre-ring amplitude k sits at post-ring index f(k)
                            const int fk = ring_f<NQ>(g, ring);
                            const unsigned short *ft = ftab + (ring - 1) * R;
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int a = slot_of<RB>(fk ^ ft[r]);
                                s[r] = psi[a];
                                l[r] = lam[a];
                            }
                            if (G > 1) group_sync<G, T>(slot);
                        } else {
                            const unsigned int w = czw[(ring - 1) * G + g];
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const unsigned int sg = (w << (31 - r)) & 0x80000000u;
                                const float2 a = pp[r * vw.stride], b = lp[r * vw.stride];
                                s[r] = make_float2(__uint_as_float(__float_as_uint(a.x) ^ sg),
                                                   __uint_as_float(__float_as_uint(a.y) ^ sg));
                                l[r] = make_float2(__uint_as_float(__float_as_uint(b.x) ^ sg),
                                                   __uint_as_float(__float_as_uint(b.y) ^ sg));
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            s[r] = pp[r * vw.stride];
                            l[r] = lp[r * vw.stride];
                        }
                    }
                    float gv[16];   // [q*3 + {phi, theta, omega}] partial angle gradients of this view
                    if (NVF == NV || v < NVF) view_bwd<NQ, RB, 0>(s, l, gv, ga, tl + v * 2 * R, cst, vw.lo, has_post, encl, et + v * R, ep);
                    else view_bwd<NQ, RB, (NVF == NV ? 0 : Q_TAIL)>(s, l, gv, ga, tl + v * 2 * R, cst, vw.lo, has_post, encl, et + v * R, ep);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        pp[r * vw.stride] = s[r];
                        lp[r * vw.stride] = l[r];
                    }
                    // reduce the view's angle gradients over the warp (all its instances) into the CTA sums
                    warp_reduce_scatter16(gv, lane);
                    if ((lane & 1) == 0) {
                        const int i = lane >> 1, q = i / 3, kind = i - 3 * q;
                        if (q < RB && q >= vw.q_lo && (kind != 2 || has_post))
                            atomicAdd(acc_s + (li * NQ + (NQ - 1 - (vw.lo + q))) * 3 + kind, gv[0]);
                    }
                    group_sync<G, T>(slot);
                }
            }
            }

            // ------------------------------------------------------------ input gradients
            if (p.grad_in != nullptr) {
                if (p.init == QIDDM_INIT_AMPLITUDE) {
                    // psi is back at psi0 = f/|f| (real); dL/dpsi0_k = 2 Re lambda0_k
                    float dot = 0.f;
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int a = slot_of<RB>(g + i * G);
                        dot += 2.f * lam[a].x * psi[a].x;
                    }
                    dot = group_sum<G, T>(dot, red, tid);
#pragma unroll 1
                    for (int i = 0; i < R; ++i) {
                        const int k = g + i * G;
                        if (k < p.n_features && active) {
                            const int a = slot_of<RB>(k);
                            const float gvv = (2.f * lam[a].x - psi[a].x * dot) * inv_norm;
                            if (p.unfold) {
                                const long long off = unfold_offset(p, geo, foff, fyx, k);
                                if (off >= 0) {
                                    if (p.io64) atomicAdd(reinterpret_cast<double *>(p.grad_in) + off, (double)gvv);
                                    else atomicAdd(p.grad_in + off, gvv);
                                }
                            } else {
                                p.grad_in[geo.in_base + k] = gvv;
                            }
                        }
                    }
                }
                if (p.enc != QIDDM_ENC_NONE) {
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        const float t = group_sum<G, T>(ga[j], red, tid);
                        if (g == 0 && active) p.grad_in[geo.in_base + j] = p.enc_scale * t;
                    }
                }
            }
            group_sync<G, T>(slot);
        }
    }

    if (BWD) {
        __syncthreads();
        float *dst = p.partials + (long long)blockIdx.x * n_grad;
        for (int i = tid; i < n_grad; i += T) dst[i] = acc_s[i];
    }
}

// the resident schedule needs two views of equal width and an all-diagonal layer boundary
template <int NQ, int RB>
constexpr bool can_resident() { return NQ == 2 * RB; }
inline bool wants_resident(int nq, int rb, const GateParams &p) { return nq == 2 * rb && p.merge_post != 0; }

template <int NQ, int RB>
size_t smem_bytes(const GateParams &p, bool bwd, bool res) {
    using C = Cfg<NQ, RB>;
    constexpr int NRING = NQ > 1 ? NQ - 1 : 1;
    size_t floats = 0;
    const size_t n_grad = res ? ((size_t)p.n_rot + NQ) * 2 : (size_t)p.n_rot * 3;
    if (bwd) floats += (n_grad + 3) & ~(size_t)3;
    floats += 2 * (size_t)C::CPB * C::STRIDE;                        // psi
    if (bwd) floats += 2 * (size_t)C::CPB * C::STRIDE;               // lambda
    floats += 2 * (size_t)C::CPB * NQ;                               // (cos, sin) of the re-upload angles
    if (p.enc == QIDDM_ENC_RZ) floats += 2 * (size_t)C::CPB * C::NV * C::R;
    floats += C::T / 32;                                             // red
    if (p.imprimitive == QIDDM_IMP_CZ) floats += (size_t)(res ? 2 : 1) * NRING * C::G;
    if (p.unfold) floats += 2 * (size_t)p.n_features;
    size_t bytes = floats * 4 + (p.imprimitive == QIDDM_IMP_CNOT ? (size_t)NRING * C::R * 2 : 0);
    return (bytes + 15) & ~(size_t)15;
}

template <int NQ, int RB, bool BWD, bool RES>
cudaError_t info_k(const GateParams &p, LaunchInfo *li) {
    using C = Cfg<NQ, RB>;
    auto kern = gate_kernel<NQ, RB, BWD, RES>;
    li->block = C::T;
    li->smem = smem_bytes<NQ, RB>(p, BWD, RES);
    // the attribute / occupancy queries cost ~10 us of host time per launch: remembered per kernel instantiation, thread
    // and device for the last shared-memory size (a module calls the same circuit over and over)
    thread_local size_t c_smem = ~(size_t)0;
    thread_local int c_dev = -1, c_sms = 0, c_per_sm = 0;
    cudaError_t e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (c_smem == li->smem && c_dev == dev) {
        sms = c_sms; per_sm = c_per_sm;
    } else {
        // the opt-in shared-memory limit of a function is process-wide per device: only ever raise it
        static std::atomic<int> attr_max[64];
        const int di = dev & 63;
        if ((int)li->smem > attr_max[di].load(std::memory_order_relaxed)) {
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)li->smem)) != cudaSuccess) return e;
            int cur = attr_max[di].load(std::memory_order_relaxed);
            while ((int)li->smem > cur && !attr_max[di].compare_exchange_weak(cur, (int)li->smem)) {}
        }
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::T, li->smem)) != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        c_smem = li->smem; c_dev = dev; c_sms = sms; c_per_sm = per_sm;
    }
    const long long need = (p.B + C::CPB - 1) / C::CPB;
    const long long cap = (long long)sms * per_sm;
    li->grid = (int)(need < cap ? (need > 0 ? need : 1) : cap);
    return cudaSuccess;
}
template <int NQ, int RB, bool BWD>
cudaError_t info_t(const GateParams &p, LaunchInfo *li) {
    if constexpr (can_resident<NQ, RB>()) {
        if (wants_resident(NQ, RB, p)) return info_k<NQ, RB, BWD, true>(p, li);
    }
    return info_k<NQ, RB, BWD, false>(p, li);
}

template <int NQ, int RB, bool BWD>
cudaError_t launch_t(const GateParams &p, const LaunchInfo &li, cudaStream_t s) {
    // algorithmic flops: 14 * 2^n per Rot; the adjoint costs 3x a forward (un-apply on psi, apply-dagger on
    // lambda, the inner products) plus one more forward when psi_final has to be recomputed
    const double work = (double)p.B * p.n_rot * 14.0 * (double)(1 << NQ) * (BWD ? (p.state != nullptr ? 3.0 : 4.0) : 1.0);
    timing_begin(BWD ? TK_GATE_BWD : TK_GATE_FWD, work, s);
    bool done = false;
    if constexpr (can_resident<NQ, RB>()) {
        if (wants_resident(NQ, RB, p)) {
            gate_kernel<NQ, RB, BWD, true><<<li.grid, li.block, li.smem, s>>>(p);
            done = true;
        }
    }
    if (!done) gate_kernel<NQ, RB, BWD, false><<<li.grid, li.block, li.smem, s>>>(p);
    timing_end(s);
    count_launch();
    return cudaGetLastError();
}

#define QIDDM_DISPATCH_NQ(nq, EXPR)                                                                             \
    switch (nq) {                                                                                               \
        case 1: { constexpr int NQ = 1; constexpr int RB = 1; return EXPR; }                                    \
        case 2: { constexpr int NQ = 2; constexpr int RB = 2; return EXPR; }                                    \
        case 3: { constexpr int NQ = 3; constexpr int RB = 3; return EXPR; }                                    \
        case 4: { constexpr int NQ = 4; constexpr int RB = 4; return EXPR; }                                    \
        case 5: { constexpr int NQ = 5; constexpr int RB = 5; return EXPR; }                                    \
        case 6: { constexpr int NQ = 6; constexpr int RB = 3; return EXPR; }                                    \
        case 7: { constexpr int NQ = 7; constexpr int RB = 3; return EXPR; }                                    \
        case 8: { constexpr int NQ = 8;                                                                         \
                  if (rb == 3) { constexpr int RB = 3; return EXPR; } { constexpr int RB = 4; return EXPR; } }  \
        case 9: { constexpr int NQ = 9;                                                                         \
                  if (rb == 3) { constexpr int RB = 3; return EXPR; } { constexpr int RB = 4; return EXPR; } }  \
        case 10: { constexpr int NQ = 10;                                                                       \
                   if (rb == 3) { constexpr int RB = 3; return EXPR; }                                          \
                   if (rb == 5) { constexpr int RB = 5; return EXPR; } { constexpr int RB = 4; return EXPR; } } \
        case 11: { constexpr int NQ = 11;                                                                       \
                   if (rb == 3) { constexpr int RB = 3; return EXPR; }                                          \
                   if (rb == 5) { constexpr int RB = 5; return EXPR; } { constexpr int RB = 4; return EXPR; } } \
        case 12: { constexpr int NQ = 12;                                                                       \
                   if (rb == 5) { constexpr int RB = 5; return EXPR; } { constexpr int RB = 4; return EXPR; } } \
        default: return cudaErrorInvalidValue;                                                                  \
    }

// ---------------------------------------------------------------------------------------------
// weights -> per-layer phase tables and rotation coefficients; angle gradients -> weight gradients
// (double precision, tiny)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double load_w(const void *w, int dtype, int i) {
    return dtype == QIDDM_DTYPE_F64 ? reinterpret_cast<const double *>(w)[i]
                                    : (double)reinterpret_cast<const float *>(w)[i];
}
__device__ __forceinline__ double remap_fn(double w, int remap) {
    if (remap == QIDDM_REMAP_TANH) return tanh(w);
    if (remap == QIDDM_REMAP_PI_TANH) return CUDART_PI * tanh(w);
    return w;
}
__device__ __forceinline__ double remap_grad(double w, int remap) {
    if (remap == QIDDM_REMAP_NONE) return 1.0;
    const double t = tanh(w);
    const double d = 1.0 - t * t;
    return remap == QIDDM_REMAP_PI_TANH ? CUDART_PI * d : d;
}

// One thread per table entry.  Layer li, view v, local index r:
//   pre[r]  = exp(i/2 sum_q sigma_q(r) phi_w(q))   (+ the previous layer's omega when merge_post)
//   post[r] = exp(i/2 sum_q sigma_q(r) omega_w(q))  (identity when merged into the next layer)
// sigma = -1 for bit 0, +1 for bit 1 (RZ(a) = diag(e^{-ia/2}, e^{+ia/2})); q runs over the bits view v owns.
__global__ void prepare_tables_kernel(const void *weights, int wdtype, int remap, int nq, int rb, int n_layers,
                                      int merge_post, float2 *tab) {
    const int R = 1 << rb, NV = (nq + rb - 1) / rb, lo_last = nq - rb;
    const int tab_layer = NV * 2 * R + ((nq + 1) & ~1);
    const long long total = (long long)n_layers * tab_layer;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int li = (int)(i / tab_layer);
        const int e = (int)(i - (long long)li * tab_layer);
        float2 o;
        if (e >= NV * 2 * R) {
            const int wire = e - NV * 2 * R;
            double s = 0.0, c = 1.0;
            if (wire < nq) sincos(0.5 * remap_fn(load_w(weights, wdtype, (li * nq + wire) * 3 + 1), remap), &s, &c);
            o = make_float2((float)c, (float)s);
        } else {
            const int v = e / (2 * R), post = (e / R) & 1, r = e % R;
            const int lo = (v * rb < lo_last) ? v * rb : lo_last;
            const int q_lo = v * rb - lo;
            const int q_hi = (((v + 1) * rb < nq) ? (v + 1) * rb : nq) - lo;
            double ang = 0.0;
            for (int q = q_lo; q < q_hi; ++q) {
                const int wire = nq - 1 - (lo + q);
                const double sgn = ((r >> q) & 1) ? 0.5 : -0.5;
                if (!post) {
                    ang += sgn * remap_fn(load_w(weights, wdtype, (li * nq + wire) * 3 + 0), remap);
                    if (merge_post && li > 0)
                        ang += sgn * remap_fn(load_w(weights, wdtype, ((li - 1) * nq + wire) * 3 + 2), remap);
                } else if (!merge_post || li == n_layers - 1) {
                    ang += sgn * remap_fn(load_w(weights, wdtype, (li * nq + wire) * 3 + 2), remap);
                }
            }
            double s, c;
            sincos(ang, &s, &c);
            o = make_float2((float)c, (float)s);
        }
        tab[i] = o;
    }
}

// Resident schedule (nq = 2 rb): boundary j = 0..n_layers carries alpha_{j,w} = phi^j_w + omega^{j-1}_w.  Residency j
// works in view v = j & 1 (local bits [v rb, v rb + rb)):  local[r] = phases of the view's own wires, scalar[g] =
// phases of the other half's wires (g = the thread's index), then (cos,sin)(theta^j_w / 2) per wire.
__global__ void prepare_tables_res_kernel(const void *weights, int wdtype, int remap, int nq, int rb, int n_layers,
                                          float2 *tab) {
    const int R = 1 << rb;
    const int tab_b = 2 * R + ((nq + 1) & ~1);
    const long long total = (long long)(n_layers + 1) * tab_b;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i / tab_b);
        const int e = (int)(i - (long long)j * tab_b);
        float2 o;
        if (e >= 2 * R) {
            const int wire = e - 2 * R;
            double s = 0.0, c = 1.0;
            if (wire < nq && j < n_layers)
                sincos(0.5 * remap_fn(load_w(weights, wdtype, (j * nq + wire) * 3 + 1), remap), &s, &c);
            o = make_float2((float)c, (float)s);
        } else {
            const int v = j & 1;
            const int half = (e < R) ? v : 1 - v;     // which half of the wires this entry covers
            const int idx = e % R;
            double ang = 0.0;
            for (int q = 0; q < rb; ++q) {
                const int wire = nq - 1 - (half * rb + q);
                const double sgn = ((idx >> q) & 1) ? 0.5 : -0.5;
                if (j < n_layers) ang += sgn * remap_fn(load_w(weights, wdtype, (j * nq + wire) * 3 + 0), remap);
                if (j > 0) ang += sgn * remap_fn(load_w(weights, wdtype, ((j - 1) * nq + wire) * 3 + 2), remap);
            }
            double s, c;
            sincos(ang, &s, &c);
            o = make_float2((float)c, (float)s);
        }
        tab[i] = o;
    }
}

// partials[b][(li*nq + wire)*3 + {0: d/dphi (+ d/domega of layer li-1 when merged), 1: d/dtheta, 2: d/domega}]
__global__ void finalize_grads_kernel(const float *partials, int n_partials, const void *weights, int wdtype,
                                      int remap, int nq, int n_layers, int merge_post, void *grad_weights) {
    // one warp per output: the lanes sum interleaved subsets of the per-CTA partials in a fixed order (deterministic),
    // then a shuffle tree -- the serial loop over up to ~2 400 partial buffers cost 25 us per launch
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // (li*nq + wire)*3 + kind
    const int lane = threadIdx.x & 31;
    const int n = n_layers * nq * 3;
    if (i >= n) return;
    const int kind = i % 3, gate = i / 3, li = gate / nq;
    int src = i;
    if (kind == 2 && merge_post && li < n_layers - 1) src = (gate + nq) * 3;    // merged into the next layer's phi table
    double acc = 0.0;
    for (int b = lane; b < n_partials; b += 32) acc += (double)partials[(size_t)b * n + src];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane != 0) return;
    const double gval = acc * remap_grad(load_w(weights, wdtype, i), remap);
    if (wdtype == QIDDM_DTYPE_F64) reinterpret_cast<double *>(grad_weights)[i] = gval;
    else reinterpret_cast<float *>(grad_weights)[i] = (float)gval;
}

// resident schedule: partials[b][(j*nq + wire)*2 + {0: d/dalpha_j, 1: d/dtheta_j}], j = 0..n_layers
__global__ void finalize_grads_res_kernel(const float *partials, int n_partials, const void *weights, int wdtype,
                                          int remap, int nq, int n_layers, void *grad_weights) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // (li*nq + wire)*3 + kind; one warp per output
    const int lane = threadIdx.x & 31;
    const int n = n_layers * nq * 3;
    if (i >= n) return;
    const int kind = i % 3, gate = i / 3;
    const int src = kind == 0 ? gate * 2 : (kind == 1 ? gate * 2 + 1 : (gate + nq) * 2);
    const size_t stride = (size_t)(n_layers + 1) * nq * 2;
    double acc = 0.0;
    for (int b = lane; b < n_partials; b += 32) acc += (double)partials[(size_t)b * stride + src];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane != 0) return;
    const double gval = acc * remap_grad(load_w(weights, wdtype, i), remap);
    if (wdtype == QIDDM_DTYPE_F64) reinterpret_cast<double *>(grad_weights)[i] = gval;
    else reinterpret_cast<float *>(grad_weights)[i] = (float)gval;
}

}  // namespace

int gate_rb(int n_qubits, bool backward) { return rb_choose(n_qubits, backward); }

// The resident schedule leaves psi_final without the trailing diagonal (it only matters for a state readout), the plain
// schedule applies it: a saved psi_final is only handed from the forward to the adjoint kernel when both run the same schedule.
bool gate_state_compatible(int n_qubits, const GateParams &p) {
    const int rf = rb_choose(n_qubits, false), rbw = rb_choose(n_qubits, true);
    return wants_resident(n_qubits, rf, p) == wants_resident(n_qubits, rbw, p);
}

size_t gate_table_bytes(int n_qubits, int n_layers) {
    // upper bound over the register-tile widths and schedules: NV * 2 * R + n float2 per layer (NV <= 4, R <= 32),
    // one extra boundary for the resident schedule
    size_t per_layer = 0;
    for (int rb = 1; rb <= 5; ++rb) {
        if (!rb_valid(n_qubits, rb)) continue;
        const size_t nv = (n_qubits + rb - 1) / rb;
        const size_t t = nv * 2 * ((size_t)1 << rb) + ((n_qubits + 1) & ~1);
        if (t > per_layer) per_layer = t;
    }
    return (size_t)(n_layers + 1) * per_layer * sizeof(float2);
}

size_t gate_partial_floats(int n_qubits, int n_layers) {
    const size_t a = (size_t)n_layers * n_qubits * 3, b = (size_t)(n_layers + 1) * n_qubits * 2;
    return a > b ? a : b;
}

cudaError_t gate_launch_info(int n_qubits, bool backward, const GateParams &p, LaunchInfo *info) {
    const int rb = rb_choose(n_qubits, backward);
    if (backward) { QIDDM_DISPATCH_NQ(n_qubits, (info_t<NQ, RB, true>(p, info))) }
    QIDDM_DISPATCH_NQ(n_qubits, (info_t<NQ, RB, false>(p, info)))
}
cudaError_t launch_gate_forward(int n_qubits, const GateParams &p, const LaunchInfo &li, cudaStream_t s) {
    const int rb = rb_choose(n_qubits, false);
    QIDDM_DISPATCH_NQ(n_qubits, (launch_t<NQ, RB, false>(p, li, s)))
}
cudaError_t launch_gate_backward(int n_qubits, const GateParams &p, const LaunchInfo &li, cudaStream_t s) {
    const int rb = rb_choose(n_qubits, true);
    QIDDM_DISPATCH_NQ(n_qubits, (launch_t<NQ, RB, true>(p, li, s)))
}
cudaError_t launch_prepare_tables(const void *weights, int wdtype, int remap, int n_qubits, bool backward,
                                  const GateParams &p, float *tables, cudaStream_t s) {
    const int rb = rb_choose(n_qubits, backward);
    const int n_layers = p.n_blocks * p.layers;
    const int R = 1 << rb, NV = (n_qubits + rb - 1) / rb;
    const bool res = wants_resident(n_qubits, rb, p);
    const long long total = res ? (long long)(n_layers + 1) * (2 * R + ((n_qubits + 1) & ~1))
                                : (long long)n_layers * (NV * 2 * R + ((n_qubits + 1) & ~1));
    const int blocks = (int)((total + 255) / 256);
    const int grid = blocks < 1184 ? blocks : 1184;
    if (res)
        prepare_tables_res_kernel<<<grid, 256, 0, s>>>(weights, wdtype, remap, n_qubits, rb, n_layers,
                                                       reinterpret_cast<float2 *>(tables));
    else
        prepare_tables_kernel<<<grid, 256, 0, s>>>(weights, wdtype, remap, n_qubits, rb, n_layers, p.merge_post,
                                                   reinterpret_cast<float2 *>(tables));
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_finalize_grads(const float *partials, int n_partials, const void *weights, int wdtype, int remap,
                                  int n_qubits, const GateParams &p, void *grad_weights, cudaStream_t s) {
    const int n_layers = p.n_blocks * p.layers;
    const int n = n_layers * n_qubits * 3;
    if (wants_resident(n_qubits, rb_choose(n_qubits, true), p))
        finalize_grads_res_kernel<<<(n + 7) / 8, 256, 0, s>>>(partials, n_partials, weights, wdtype, remap, n_qubits,
                                                                 n_layers, grad_weights);
    else
        finalize_grads_kernel<<<(n + 7) / 8, 256, 0, s>>>(partials, n_partials, weights, wdtype, remap, n_qubits,
                                                             n_layers, p.merge_post, grad_weights);
    count_launch();
    return cudaGetLastError();
}

}  // namespace qiddm
