// Gate-by-gate batched state-vector kernels (forward + adjoint backward) for sm_100a.
//
// Replaces PennyLane's per-gate einsum on a (B,2,...,2) complex128 tensor
// (default.qubit.torch; nn/qdense.py:58,:465) and the per-sample lightning.qubit loop
// (nn/qdense.py:1631-1635) of the reference.
//
// Layout.  One circuit instance is simulated by a group of G = 2^n / R threads; the state
// (2^n complex fp32) lives in shared memory and is swept in register tiles of R = 2^RB
// amplitudes per thread ("views"): in view v the RB index bits [lo_v, lo_v+RB) are local
// to a thread, so every Rot on those wires is pure register FMA work; between views the
// tile goes back through shared memory (XOR-swizzled, bank-conflict free), and the CNOT
// ring of a StronglyEntangling layer is a GF(2)-linear index permutation folded into the
// last store of the layer (CZ ring: a sign computed from popc(k & rotl(k, r))).  The
// re-upload data gate RZ/RY(a_j) that precedes a block's first Rot on the same wire is
// folded into that Rot's 2x2 matrix per instance, so it costs no pass over the state.
// Groups with G <= 32 share a warp (several instances per warp, __syncwarp only).
//
// Backward = adjoint method: recompute psi_final, seed lambda = dL/dpsi* from the readout,
// then walk the gates in reverse applying U^dagger to both while accumulating, per gate,
// the 2x2 cotangent M_ab = sum conj(lambda_post_a) psi_pre_b.  M is reduced over the warp
// with a 9-shuffle reduce-scatter, summed per CTA in shared memory, written as per-CTA
// partials and turned into angle gradients (incl. the tanh / pi*tanh re-map chain rule) in
// double precision by finalize_grads_kernel (deterministic across CTAs).
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
    static constexpr int STRIDE = A + 1;  // float2 slots per state (odd: spreads instances over banks)
    static constexpr int LO_LAST = NQ - RB;
};

// Register-tile width (log2 amplitudes per thread).  Candidates compiled per qubit count:
// min(nq,3), min(nq,4), min(nq,5) (subject to 2^(nq-rb) <= 256 threads per instance); the defaults below were
// picked from B200 measurements and can be overridden with QIDDM_RB_FWD / QIDDM_RB_BWD for tuning.
inline bool rb_valid(int nq, int rb) {
    if (rb < 1 || rb > 5 || rb > nq) return false;
    if (rb < 3 && rb != nq) return false;
    return (nq - rb) <= 8;
}
inline int rb_default(int nq, bool bwd) {
    if (bwd) return nq <= 4 ? nq : nq <= 6 ? 3 : nq <= 8 ? 4 : nq == 9 ? 3 : 4;
    return nq <= 5 ? nq : nq == 6 ? 3 : nq <= 8 ? 4 : nq <= 10 ? 5 : 4;
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

struct Mat {
    float r00, i00, r01, i01, r10, i10, r11, i11;
};

__device__ __forceinline__ Mat load_mat(const float *g) {
    const float4 a = *reinterpret_cast<const float4 *>(g);
    const float4 b = *reinterpret_cast<const float4 *>(g + 4);
    return Mat{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
}
__device__ __forceinline__ Mat adjoint(const Mat &m) {
    return Mat{m.r00, -m.i00, m.r10, -m.i10, m.r01, -m.i01, m.r11, -m.i11};
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// U' = U * E(alpha) with E = RZ (enc==1) or RY (enc==2); cs = (cos(alpha/2), sin(alpha/2)).
__device__ __forceinline__ Mat fold_enc(const Mat &u, float2 cs, int enc) {
    Mat o;
    const float c = cs.x, s = cs.y;
    if (enc == QIDDM_ENC_RZ) {
        // column 0 times e^{-i alpha/2} = (c,-s); column 1 times e^{+i alpha/2} = (c, s)
        o.r00 = u.r00 * c + u.i00 * s;  o.i00 = u.i00 * c - u.r00 * s;
        o.r10 = u.r10 * c + u.i10 * s;  o.i10 = u.i10 * c - u.r10 * s;
        o.r01 = u.r01 * c - u.i01 * s;  o.i01 = u.i01 * c + u.r01 * s;
        o.r11 = u.r11 * c - u.i11 * s;  o.i11 = u.i11 * c + u.r11 * s;
    } else {
        // E = [[c,-s],[s,c]]
        o.r00 = u.r00 * c + u.r01 * s;  o.i00 = u.i00 * c + u.i01 * s;
        o.r01 = u.r01 * c - u.r00 * s;  o.i01 = u.i01 * c - u.i00 * s;
        o.r10 = u.r10 * c + u.r11 * s;  o.i10 = u.i10 * c + u.i11 * s;
        o.r11 = u.r11 * c - u.r10 * s;  o.i11 = u.i11 * c - u.i10 * s;
    }
    return o;
}

__device__ __forceinline__ void apply_pair(const Mat &m, float2 &x0, float2 &x1) {
    const float2 a = x0, b = x1;
    x0.x = m.r00 * a.x - m.i00 * a.y + m.r01 * b.x - m.i01 * b.y;
    x0.y = m.r00 * a.y + m.i00 * a.x + m.r01 * b.y + m.i01 * b.x;
    x1.x = m.r10 * a.x - m.i10 * a.y + m.r11 * b.x - m.i11 * b.y;
    x1.y = m.r10 * a.y + m.i10 * a.x + m.r11 * b.y + m.i11 * b.x;
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
__device__ __forceinline__ float cz_sign(int k, int ring) {
    const int rot = ((k << ring) | (k >> (NQ - ring))) & ((1 << NQ) - 1);
    return (__popc(k & rot) & 1) ? -1.0f : 1.0f;
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

// Warp-wide sum of 8 values with a reduce-scatter (9 shuffles), then 8 lanes add into acc[0..8).
__device__ __forceinline__ void warp_reduce8_add(const float (&v)[8], float *acc, int lane) {
    float a[4], b[2], c;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h16 ? v[i] : v[4 + i];
        const float keep = h16 ? v[4 + i] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h8 ? a[i] : a[2 + i];
        const float keep = h8 ? a[2 + i] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float send = h4 ? b[0] : b[1];
        const float keep = h4 ? b[1] : b[0];
        c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    if ((lane & 3) == 0) atomicAdd(acc + ((h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0)), c);
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
        g.in_base = cid * n_in;
        g.out_base = cid * n_out;
        g.out_stride = 1;
    }
    return g;
}
// offset of feature k inside the image, or -1 when it falls in the zero padding
__device__ __forceinline__ long long unfold_offset(const GateParams &p, const InstanceGeom &g, int k) {
    const int kk = p.kh * p.kw;
    const int ch = k / kk;
    const int r = k - ch * kk;
    const int ky = r / p.kw;
    const int kx = r - ky * p.kw;
    const int yy = g.y + ky - p.ph, xx = g.x + kx - p.pw;
    if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) return -1;
    return g.in_base + ((long long)ch * p.H + yy) * p.W + xx;
}

template <int NQ, int RB>
__device__ __forceinline__ int amp_index(int g, int r, int lo) {
    return ((g >> lo) << (lo + RB)) | (r << lo) | (g & ((1 << lo) - 1));
}
template <int RB>
__device__ __forceinline__ int swz(int k) {
    return k ^ ((k >> RB) & ((1 << RB) - 1));
}

template <int NQ, int RB, bool BWD>
__global__ void __launch_bounds__(Cfg<NQ, RB>::T) gate_kernel(const GateParams p) {
    using C = Cfg<NQ, RB>;
    constexpr int A = C::A, R = C::R, G = C::G, NV = C::NV, T = C::T, CPB = C::CPB, STRIDE = C::STRIDE;
    constexpr int LO_LAST = C::LO_LAST;
    constexpr int NRING = NQ > 1 ? NQ - 1 : 1;

    extern __shared__ float4 smem_f4[];
    float *sm = reinterpret_cast<float *>(smem_f4);
    // carve-up (all offsets multiples of 4 floats)
    const int n_acc = p.n_rot * 8;
    float *gates_s = sm;                                       // [n_acc] when gates_in_smem
    float *acc_s = gates_s + (p.gates_in_smem ? n_acc : 0);    // [n_acc] (BWD)
    float2 *psi_all = reinterpret_cast<float2 *>(acc_s + (BWD ? n_acc : 0));
    float2 *lam_all = psi_all + CPB * STRIDE + 1;              // (+1 keeps float2 alignment irrelevant)
    float2 *ep_all = BWD ? lam_all + CPB * STRIDE + 1 : lam_all;   // [CPB][NQ] (cos, sin)(alpha/2)
    float *misc = reinterpret_cast<float *>(ep_all + CPB * NQ);     // [CPB] inv norms
    float *red = misc + CPB;                                         // [T/32] + [CPB*NQ] scratch
    unsigned short *ftab = reinterpret_cast<unsigned short *>(red + T / 32 + CPB * NQ);  // [NRING][R]

    const int tid = threadIdx.x, lane = tid & 31;
    const int slot = tid / G, g = tid % G;
    const float *gm = p.gates;

    if (p.gates_in_smem) {
        for (int i = tid; i < n_acc; i += T) gates_s[i] = p.gates[i];
        gm = gates_s;
    }
    if (BWD)
        for (int i = tid; i < n_acc; i += T) acc_s[i] = 0.f;
    for (int i = tid; i < NRING * R; i += T) {
        const int ring = i / R + 1, r = i % R;
        ftab[i] = (unsigned short)ring_f<NQ>(r << LO_LAST, ring);
    }
    __syncthreads();

    float2 *psi = psi_all + slot * STRIDE;
    float2 *lam = lam_all + slot * STRIDE;
    float2 *ep = ep_all + slot * NQ;
    const int n_in = p.init == QIDDM_INIT_AMPLITUDE ? p.n_features : (p.enc != QIDDM_ENC_NONE ? NQ : 0);
    const int n_out = p.readout == QIDDM_READ_PROBS ? p.read_count : (p.readout == QIDDM_READ_EXPVAL_Z ? NQ : 2 * A);
    const int k_thread_last = amp_index<NQ, RB>(g, 0, LO_LAST);

    for (long long base = (long long)blockIdx.x * CPB; base < p.B; base += (long long)gridDim.x * CPB) {
        const long long cid = base + slot;
        const bool active = cid < p.B;
        const InstanceGeom geo = instance_geom(p, active ? cid : 0, n_in, n_out);

        // ---------------------------------------------------------------- initial state
        float inv_norm = 1.f;
        if (p.init == QIDDM_INIT_AMPLITUDE) {
            float vals[R];
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                float v = p.pad_value;
                if (k < p.n_features) {
                    float x = 0.f;
                    if (active) {
                        if (p.unfold) {
                            const long long off = unfold_offset(p, geo, k);
                            x = off >= 0 ? __ldg(p.in + off) : 0.f;
                        } else {
                            x = __ldg(p.in + geo.in_base + k);
                        }
                    }
                    v = x + p.add_offset;
                }
                vals[i] = v;
                ss += v * v;
            }
            ss = group_sum<G, T>(ss, red, tid);
            inv_norm = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
#pragma unroll
            for (int i = 0; i < R; ++i) psi[swz<RB>(g + i * G)] = make_float2(vals[i] * inv_norm, 0.f);
        } else {
            int start = 0;
            if (p.init == QIDDM_INIT_BASIS) start = p.basis ? (active ? p.basis[cid] : 0) : (int)(cid & (A - 1));
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                psi[swz<RB>(k)] = make_float2(k == start ? 1.f : 0.f, 0.f);
            }
        }
        if (p.enc != QIDDM_ENC_NONE) {
            for (int j = g; j < NQ; j += G) {
                const float a = active ? p.enc_scale * __ldg(p.in + geo.in_base + j) : 0.f;
                float s, c;
                sincosf(0.5f * a, &s, &c);
                ep[j] = make_float2(c, s);
            }
        }
        group_sync<G, T>(slot);

        // ---------------------------------------------------------------- forward sweep
        for (int blk = 0; blk < p.n_blocks; ++blk) {
            for (int layer = 0; layer < p.layers; ++layer) {
                const int gate_base = (blk * p.layers + layer) * NQ;
                const int ring = NQ > 1 ? (layer % NRING) + 1 : 0;
                const bool encl = (p.enc != QIDDM_ENC_NONE) && layer == 0;
                float2 s[R];
#pragma unroll 1
                for (int v = 0; v < NV; ++v) {
                    const int lo = (v * RB < LO_LAST) ? v * RB : LO_LAST;
                    const int q_lo = v * RB - lo;                                        // first new local bit
                    const int q_hi = (((v + 1) * RB < NQ) ? (v + 1) * RB : NQ) - lo;     // one past the last
                    const int kbase = amp_index<NQ, RB>(g, 0, lo);
#pragma unroll
                    for (int r = 0; r < R; ++r) s[r] = psi[swz<RB>(kbase | (r << lo))];
#pragma unroll
                    for (int q = 0; q < RB; ++q) {
                        if (q < q_lo || q >= q_hi) continue;
                        const int wire = NQ - 1 - (lo + q);
                        Mat m = load_mat(gm + (gate_base + wire) * 8);
                        if (encl) m = fold_enc(m, ep[wire], p.enc);
#pragma unroll
                        for (int j = 0; j < R / 2; ++j) {
                            const int r0 = ((j >> q) << (q + 1)) | (j & ((1 << q) - 1));
                            apply_pair(m, s[r0], s[r0 | (1 << q)]);
                        }
                    }
                    if (v == NV - 1 && NQ > 1) {
                        if (p.imprimitive == QIDDM_IMP_CNOT) {
                            if (G > 1) group_sync<G, T>(slot);  // every tile is in registers before the scatter
                            const int fk = ring_f<NQ>(k_thread_last, ring);
                            const unsigned short *ft = ftab + (ring - 1) * R;
#pragma unroll
                            for (int r = 0; r < R; ++r) psi[swz<RB>(fk ^ ft[r])] = s[r];
                        } else {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int k = kbase | (r << lo);
                                const float sg = cz_sign<NQ>(k, ring);
                                psi[swz<RB>(k)] = make_float2(s[r].x * sg, s[r].y * sg);
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) psi[swz<RB>(kbase | (r << lo))] = s[r];
                    }
                    group_sync<G, T>(slot);
                }
            }
        }

        // ---------------------------------------------------------------- readout
        if (!BWD) {
            if (p.readout == QIDDM_READ_PROBS) {
                for (int m = g; m < p.read_count; m += G) {
                    const float2 a = psi[swz<RB>(m * p.read_stride)];
                    float v = p.post_scale * (a.x * a.x + a.y * a.y);
                    if (p.clamp) v = fminf(fmaxf(v, p.clamp_lo), p.clamp_hi);
                    if (active) p.out[geo.out_base + (long long)m * geo.out_stride] = v;
                }
            } else if (p.readout == QIDDM_READ_EXPVAL_Z) {
                float ez[NQ];
#pragma unroll
                for (int j = 0; j < NQ; ++j) ez[j] = 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int k = amp_index<NQ, RB>(g, r, 0);
                    const float2 a = psi[swz<RB>(k)];
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
                    if (active) reinterpret_cast<float2 *>(p.out + geo.out_base)[k] = psi[swz<RB>(k)];
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
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int k = g + i * G;
                const float2 a = psi[swz<RB>(k)];
                float2 l = make_float2(0.f, 0.f);
                if (p.readout == QIDDM_READ_PROBS) {
                    const int m = k / p.read_stride;
                    if (m * p.read_stride == k && m < p.read_count && active) {
                        const float v = p.post_scale * (a.x * a.x + a.y * a.y);
                        const bool pass = !p.clamp || (v >= p.clamp_lo && v <= p.clamp_hi);
                        const float c = pass ? p.post_scale * __ldg(p.grad_out + geo.out_base + (long long)m * geo.out_stride) : 0.f;
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
                lam[swz<RB>(k)] = l;
            }
            group_sync<G, T>(slot);

            float ga[NQ];  // per-instance d/d(alpha_wire), partial over this thread's amplitudes
#pragma unroll
            for (int j = 0; j < NQ; ++j) ga[j] = 0.f;

            for (int blk = p.n_blocks - 1; blk >= 0; --blk) {
                for (int layer = p.layers - 1; layer >= 0; --layer) {
                    const int gate_base = (blk * p.layers + layer) * NQ;
                    const int ring = NQ > 1 ? (layer % NRING) + 1 : 0;
                    const bool encl = (p.enc != QIDDM_ENC_NONE) && layer == 0;
                    float2 s[R], l[R];
#pragma unroll 1
                    for (int vv = 0; vv < NV; ++vv) {
                        const int v = NV - 1 - vv;
                        const int lo = (v * RB < LO_LAST) ? v * RB : LO_LAST;
                        const int q_lo = v * RB - lo;
                        const int q_hi = (((v + 1) * RB < NQ) ? (v + 1) * RB : NQ) - lo;
                        const int kbase = amp_index<NQ, RB>(g, 0, lo);
                        if (vv == 0 && NQ > 1) {
                            if (p.imprimitive == QIDDM_IMP_CNOT) {
                                // pre-ring amplitude k sits at post-ring index f(k)
                                const int fk = ring_f<NQ>(k_thread_last, ring);
                                const unsigned short *ft = ftab + (ring - 1) * R;
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    const int a = swz<RB>(fk ^ ft[r]);
                                    s[r] = psi[a];
                                    l[r] = lam[a];
                                }
                                if (G > 1) group_sync<G, T>(slot);
                            } else {
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    const int k = kbase | (r << lo);
                                    const float sg = cz_sign<NQ>(k, ring);
                                    const int a = swz<RB>(k);
                                    s[r] = make_float2(psi[a].x * sg, psi[a].y * sg);
                                    l[r] = make_float2(lam[a].x * sg, lam[a].y * sg);
                                }
                            }
                        } else {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int a = swz<RB>(kbase | (r << lo));
                                s[r] = psi[a];
                                l[r] = lam[a];
                            }
                        }
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            if (q < q_lo || q >= q_hi) continue;
                            const int wire = NQ - 1 - (lo + q);
                            const Mat ub = load_mat(gm + (gate_base + wire) * 8);
                            float2 cs = make_float2(1.f, 0.f);
                            Mat u = ub;
                            if (encl) {
                                cs = ep[wire];
                                u = fold_enc(ub, cs, p.enc);
                            }
                            const Mat ud = adjoint(u);
                            float M[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int j = 0; j < R / 2; ++j) {
                                const int r0 = ((j >> q) << (q + 1)) | (j & ((1 << q) - 1));
                                const int r1 = r0 | (1 << q);
                                apply_pair(ud, s[r0], s[r1]);  // psi_pre
                                const float2 l0 = l[r0], l1 = l[r1], x0 = s[r0], x1 = s[r1];
                                M[0] += l0.x * x0.x + l0.y * x0.y;  M[1] += l0.x * x0.y - l0.y * x0.x;  // conj(l0) x0
                                M[2] += l0.x * x1.x + l0.y * x1.y;  M[3] += l0.x * x1.y - l0.y * x1.x;  // conj(l0) x1
                                M[4] += l1.x * x0.x + l1.y * x0.y;  M[5] += l1.x * x0.y - l1.y * x0.x;  // conj(l1) x0
                                M[6] += l1.x * x1.x + l1.y * x1.y;  M[7] += l1.x * x1.y - l1.y * x1.x;  // conj(l1) x1
                                apply_pair(ud, l[r0], l[r1]);  // lambda_pre
                            }
                            if (encl) {
                                const float c = cs.x, sn = cs.y;
                                float Mb[8];
                                float dalpha;
                                if (p.enc == QIDDM_ENC_RZ) {
                                    // d/dalpha = Im(M00 U'00) - Im(M01 U'01) + Im(M10 U'10) - Im(M11 U'11)
                                    dalpha = (M[0] * u.i00 + M[1] * u.r00) - (M[2] * u.i01 + M[3] * u.r01) +
                                             (M[4] * u.i10 + M[5] * u.r10) - (M[6] * u.i11 + M[7] * u.r11);
                                    // M_base = M' E^T, E = diag((c,-s),(c,s))
                                    Mb[0] = M[0] * c + M[1] * sn;  Mb[1] = M[1] * c - M[0] * sn;
                                    Mb[2] = M[2] * c - M[3] * sn;  Mb[3] = M[3] * c + M[2] * sn;
                                    Mb[4] = M[4] * c + M[5] * sn;  Mb[5] = M[5] * c - M[4] * sn;
                                    Mb[6] = M[6] * c - M[7] * sn;  Mb[7] = M[7] * c + M[6] * sn;
                                } else {
                                    // d/dalpha = Re sum_a [ M'a0 (-Ua0 s + Ua1 c) + M'a1 (-Ua0 c - Ua1 s) ]
                                    const float t0r = -ub.r00 * sn + ub.r01 * c, t0i = -ub.i00 * sn + ub.i01 * c;
                                    const float t1r = -ub.r00 * c - ub.r01 * sn, t1i = -ub.i00 * c - ub.i01 * sn;
                                    const float t2r = -ub.r10 * sn + ub.r11 * c, t2i = -ub.i10 * sn + ub.i11 * c;
                                    const float t3r = -ub.r10 * c - ub.r11 * sn, t3i = -ub.i10 * c - ub.i11 * sn;
                                    dalpha = (M[0] * t0r - M[1] * t0i) + (M[2] * t1r - M[3] * t1i) +
                                             (M[4] * t2r - M[5] * t2i) + (M[6] * t3r - M[7] * t3i);
                                    // M_base = M' E^T, E = [[c,-s],[s,c]]
                                    Mb[0] = M[0] * c - M[2] * sn;  Mb[1] = M[1] * c - M[3] * sn;
                                    Mb[2] = M[0] * sn + M[2] * c;  Mb[3] = M[1] * sn + M[3] * c;
                                    Mb[4] = M[4] * c - M[6] * sn;  Mb[5] = M[5] * c - M[7] * sn;
                                    Mb[6] = M[4] * sn + M[6] * c;  Mb[7] = M[5] * sn + M[7] * c;
                                }
                                // wire is a runtime value here: select the accumulator without dynamic indexing
#pragma unroll
                                for (int j = 0; j < NQ; ++j) ga[j] += (j == wire) ? dalpha : 0.f;
                                warp_reduce8_add(Mb, acc_s + (gate_base + wire) * 8, lane);
                            } else {
                                warp_reduce8_add(M, acc_s + (gate_base + wire) * 8, lane);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int a = swz<RB>(kbase | (r << lo));
                            psi[a] = s[r];
                            lam[a] = l[r];
                        }
                        group_sync<G, T>(slot);
                    }
                }
            }

            // ------------------------------------------------------------ input gradients
            if (p.grad_in != nullptr) {
                if (p.init == QIDDM_INIT_AMPLITUDE) {
                    // psi is back at psi0 = f/|f| (real); dL/dpsi0_k = 2 Re lambda0_k
                    float qv[R], dot = 0.f;
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int a = swz<RB>(g + i * G);
                        qv[i] = 2.f * lam[a].x;
                        dot += qv[i] * psi[a].x;
                    }
                    dot = group_sum<G, T>(dot, red, tid);
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int k = g + i * G;
                        if (k < p.n_features && active) {
                            const float gv = (qv[i] - psi[swz<RB>(k)].x * dot) * inv_norm;
                            if (p.unfold) {
                                const long long off = unfold_offset(p, geo, k);
                                if (off >= 0) atomicAdd(p.grad_in + off, gv);
                            } else {
                                p.grad_in[geo.in_base + k] = gv;
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
        float *dst = p.partials + (long long)blockIdx.x * n_acc;
        for (int i = tid; i < n_acc; i += T) dst[i] = acc_s[i];
    }
}

template <int NQ, int RB>
size_t smem_bytes(const GateParams &p, bool bwd) {
    using C = Cfg<NQ, RB>;
    constexpr int NRING = NQ > 1 ? NQ - 1 : 1;
    size_t floats = 0;
    const size_t n_acc = (size_t)p.n_rot * 8;
    if (p.gates_in_smem) floats += n_acc;
    if (bwd) floats += n_acc;
    floats += 2 * ((size_t)C::CPB * C::STRIDE + 1);                 // psi
    if (bwd) floats += 2 * ((size_t)C::CPB * C::STRIDE + 1);        // lambda
    floats += 2 * (size_t)C::CPB * NQ;                              // enc phases
    floats += C::CPB;                                               // misc
    floats += C::T / 32 + C::CPB * NQ;                              // red
    size_t bytes = floats * 4 + (size_t)NRING * C::R * 2;
    return (bytes + 15) & ~(size_t)15;
}

template <int NQ, int RB, bool BWD>
cudaError_t info_t(const GateParams &p, LaunchInfo *li) {
    using C = Cfg<NQ, RB>;
    auto kern = gate_kernel<NQ, RB, BWD>;
    li->block = C::T;
    li->smem = smem_bytes<NQ, RB>(p, BWD);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)li->smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::T, li->smem)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const long long need = (p.B + C::CPB - 1) / C::CPB;
    const long long cap = (long long)sms * per_sm;
    li->grid = (int)(need < cap ? (need > 0 ? need : 1) : cap);
    return cudaSuccess;
}

template <int NQ, int RB, bool BWD>
cudaError_t launch_t(const GateParams &p, const LaunchInfo &li, cudaStream_t s) {
    // algorithmic flops: 14 * 2^n per Rot; the adjoint sweep costs 4x a forward
    const double work = (double)p.B * p.n_rot * 14.0 * (double)(1 << NQ) * (BWD ? 4.0 : 1.0);
    timing_begin(BWD ? TK_GATE_BWD : TK_GATE_FWD, work, s);
    gate_kernel<NQ, RB, BWD><<<li.grid, li.block, li.smem, s>>>(p);
    timing_end(s);
    count_launch();
    return cudaGetLastError();
}

#define QIDDM_RB_CASES(EXPR)                                             \
    switch (rb) {                                                        \
        case 3: { constexpr int RB = NQ < 3 ? NQ : 3; return EXPR; }     \
        case 4: { constexpr int RB = NQ < 4 ? NQ : 4; return EXPR; }     \
        case 5: { constexpr int RB = NQ < 5 ? NQ : 5; return EXPR; }     \
        default: { constexpr int RB = NQ < 3 ? NQ : 3; return EXPR; }    \
    }
#define QIDDM_RB_CASES_BIG(EXPR)                                         \
    switch (rb) {                                                        \
        case 5: { constexpr int RB = 5; return EXPR; }                   \
        default: { constexpr int RB = 4; return EXPR; }                  \
    }
#define QIDDM_DISPATCH_NQ(nq, EXPR)                                      \
    switch (nq) {                                                        \
        case 1: { constexpr int NQ = 1; constexpr int RB = 1; return EXPR; } \
        case 2: { constexpr int NQ = 2; constexpr int RB = 2; return EXPR; } \
        case 3: { constexpr int NQ = 3; constexpr int RB = 3; return EXPR; } \
        case 4: { constexpr int NQ = 4; QIDDM_RB_CASES(EXPR) }           \
        case 5: { constexpr int NQ = 5; QIDDM_RB_CASES(EXPR) }           \
        case 6: { constexpr int NQ = 6; QIDDM_RB_CASES(EXPR) }           \
        case 7: { constexpr int NQ = 7; QIDDM_RB_CASES(EXPR) }           \
        case 8: { constexpr int NQ = 8; QIDDM_RB_CASES(EXPR) }           \
        case 9: { constexpr int NQ = 9; QIDDM_RB_CASES(EXPR) }           \
        case 10: { constexpr int NQ = 10; QIDDM_RB_CASES(EXPR) }         \
        case 11: { constexpr int NQ = 11; QIDDM_RB_CASES(EXPR) }         \
        case 12: { constexpr int NQ = 12; QIDDM_RB_CASES_BIG(EXPR) }     \
        default: return cudaErrorInvalidValue;                           \
    }

// ---------------------------------------------------------------------------------------------
// weights -> 2x2 matrices, and gate cotangents -> weight gradients (double precision, tiny)
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

__global__ void prepare_gates_kernel(const void *weights, int wdtype, int remap, int n_rot, float *gates) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rot) return;
    const double phi = remap_fn(load_w(weights, wdtype, 3 * i + 0), remap);
    const double th = remap_fn(load_w(weights, wdtype, 3 * i + 1), remap);
    const double om = remap_fn(load_w(weights, wdtype, 3 * i + 2), remap);
    double c, s, cp, sp, cm, sm_;
    sincos(0.5 * th, &s, &c);
    sincos(0.5 * (phi + om), &sp, &cp);
    sincos(0.5 * (phi - om), &sm_, &cm);
    float *o = gates + 8 * i;
    o[0] = (float)(cp * c);   o[1] = (float)(-sp * c);     // e^{-i(phi+om)/2} c
    o[2] = (float)(-cm * s);  o[3] = (float)(-sm_ * s);    // -e^{+i(phi-om)/2} s
    o[4] = (float)(cm * s);   o[5] = (float)(-sm_ * s);    // e^{-i(phi-om)/2} s
    o[6] = (float)(cp * c);   o[7] = (float)(sp * c);      // e^{+i(phi+om)/2} c
}

__global__ void finalize_grads_kernel(const float *partials, int n_partials, const void *weights, int wdtype,
                                      int remap, int n_rot, void *grad_weights) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rot) return;
    double M[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t stride = (size_t)n_rot * 8;
    for (int b = 0; b < n_partials; ++b) {
        const float4 x = *reinterpret_cast<const float4 *>(partials + b * stride + 8 * i);
        const float4 y = *reinterpret_cast<const float4 *>(partials + b * stride + 8 * i + 4);
        M[0] += x.x; M[1] += x.y; M[2] += x.z; M[3] += x.w;
        M[4] += y.x; M[5] += y.y; M[6] += y.z; M[7] += y.w;
    }
    const double w0 = load_w(weights, wdtype, 3 * i + 0), w1 = load_w(weights, wdtype, 3 * i + 1),
                 w2 = load_w(weights, wdtype, 3 * i + 2);
    const double phi = remap_fn(w0, remap), th = remap_fn(w1, remap), om = remap_fn(w2, remap);
    double c, s, cp, sp, cm, sm_;
    sincos(0.5 * th, &s, &c);
    sincos(0.5 * (phi + om), &sp, &cp);
    sincos(0.5 * (phi - om), &sm_, &cm);
    // U entries
    const double u00r = cp * c, u00i = -sp * c, u01r = -cm * s, u01i = -sm_ * s;
    const double u10r = cm * s, u10i = -sm_ * s, u11r = cp * c, u11i = sp * c;
    // Im(M_ab U_ab)
    const double im00 = M[0] * u00i + M[1] * u00r, im01 = M[2] * u01i + M[3] * u01r;
    const double im10 = M[4] * u10i + M[5] * u10r, im11 = M[6] * u11i + M[7] * u11r;
    const double dphi = im00 - im01 + im10 - im11;
    const double dom = im00 + im01 - im10 - im11;
    // 2 dU/dtheta = [[-e^{-i(p+o)/2} s, -e^{i(p-o)/2} c], [e^{-i(p-o)/2} c, -e^{i(p+o)/2} s]]
    const double d00r = -cp * s, d00i = sp * s, d01r = -cm * c, d01i = -sm_ * c;
    const double d10r = cm * c, d10i = -sm_ * c, d11r = -cp * s, d11i = -sp * s;
    const double dth = (M[0] * d00r - M[1] * d00i) + (M[2] * d01r - M[3] * d01i) +
                       (M[4] * d10r - M[5] * d10i) + (M[6] * d11r - M[7] * d11i);
    const double g0 = dphi * remap_grad(w0, remap), g1 = dth * remap_grad(w1, remap), g2 = dom * remap_grad(w2, remap);
    if (wdtype == QIDDM_DTYPE_F64) {
        double *o = reinterpret_cast<double *>(grad_weights) + 3 * i;
        o[0] = g0; o[1] = g1; o[2] = g2;
    } else {
        float *o = reinterpret_cast<float *>(grad_weights) + 3 * i;
        o[0] = (float)g0; o[1] = (float)g1; o[2] = (float)g2;
    }
}

}  // namespace

int gate_rb(int n_qubits, bool backward) { return rb_choose(n_qubits, backward); }

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
cudaError_t launch_prepare_gates(const void *weights, int wdtype, int remap, int n_rot, float *gates, cudaStream_t s) {
    prepare_gates_kernel<<<(n_rot + 127) / 128, 128, 0, s>>>(weights, wdtype, remap, n_rot, gates);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_finalize_grads(const float *partials, int n_partials, const void *weights, int wdtype, int remap,
                                  int n_rot, void *grad_weights, cudaStream_t s) {
    finalize_grads_kernel<<<(n_rot + 127) / 128, 128, 0, s>>>(partials, n_partials, weights, wdtype, remap, n_rot,
                                                             grad_weights);
    count_launch();
    return cudaGetLastError();
}

}  // namespace qiddm
