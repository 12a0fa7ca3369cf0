// QConv2d on the unitary-collapse path as a DIRECT fp32 convolution (no materialised patch matrix, no operand splits).
//
// With the SEL block collapsed into U (qiddm_gemm_prepare), `_QConv2d_FAST` (reference nn/qconv.py:51-56, :71-90 with the
// H1 line restored) reads only `out_channels` amplitudes of every patch's state, i.e. N = 2 * out_channels real rows of U:
//     f      = patch + 0.1                                  (F = C kh kw features, zero padding of torch.nn.Unfold)
//     Y[n]   = sum_c f[c] Wd[c][n] + bias[n]                Wd[c][n] = part_n(U^T[c][m stride]), bias = pad * (rows c >= F)
//     out[m] = clamp(post / (|f|^2 + n_pad pad^2) * (Y[2m]^2 + Y[2m+1]^2))
// which is a convolution with N <= 32 output channels, a square, and a division by a box filter of f^2.  For the UNet's
// layers (F = 9 ... 288, N = 2 ... 32) the tcgen05 path spends its time writing and re-reading the fp16 hi / lo patch
// matrix (320 B per patch and pass at F = 72 against 64 B of image), the gate path simulates every patch gate by gate.
// Here a CTA stages a band of image rows (with halo) in shared memory once and every thread keeps the N accumulators of
// two patches in registers; the FP32 FMA pipe is the bound (F N FMAs per patch and pass), the image is read once.
// Every kernel is persistent over (image, band) units; shared-memory footprints and register counts leave two to three
// CTAs per SM, whose staging and compute phases overlap (measured: an explicit cp.async pipeline with raw staging buffers
// lost more to the lower occupancy than it gained, profiles/r2c_conv_summary.md).
//
//   conv_fwd_kernel       out (NCHW, io dtype), optionally one row [Y (NP), 1/|f|^2] per patch for the backward
//   conv_grad_kernel      G = dL/dY per patch from (Y, grad_out, clamp mask) + the normalisation term S / |f|^2: one streaming
//                         pass, rows of NP + 4 floats that both gradient kernels copy into their tiles with 16-byte cp.async
//   conv_bwd_data_kernel  gather form of the transposed convolution over a haloed G tile:
//                         d img = sum_taps (G . Wd) - 2 f sum_taps S / |f|^2
//   conv_bwd_w_kernel     dWd[c][n] = sum_patches f[c] G[n]: a lane owns one (channel, ky) row of the kernel window (kw taps,
//                         sliding along x) x 16 outputs in registers for the whole launch; per-CTA partials
//   conv_reduce_kernel / conv_assemble_kernel  fixed-order fp64 sum of the partials -> READ_STATE cotangent of U^T for the
//                         adjoint gate kernel
#include <atomic>
#include <cstdlib>
#include <cstring>
#include "qiddm_internal.h"

namespace qiddm {

namespace {

constexpr int CONV_MAX_WGRID = 320;      // per-CTA dW partials the workspace is sized for

struct ConvParams {
    const void *img;
    void *out;
    const void *go;
    void *gimg;
    const float *Wd;           // [(F + 1)][NP], row F = bias
    float *Y;                  // [Bp][NP + 4]: Y (columns >= N zero), 1/|f|^2 at column NP
    float *G;                  // [Bp][NP + 4]: dL/dY, S / |f|^2 at column NP
    float *partials;           // [grid][(F + 1) NP]
    long long B;               // patches
    int n_images, C, H, W, F, N, n_out, bands, TH, TC, CS, units, clamp;
    int wr, parts;             // conv_bwd_w_kernel: warps per n-chunk (rows / 32), pixel-row parts
    float add_offset, pad2, post_scale, clamp_lo, clamp_hi;
    // fused bilinear upsample in front of a 1 x 1 window (nn/unet.py:36-41): `img` is the (n, C, Hin, Win) source and (H, W)
    // the upsampled size the convolution runs on; source coordinate = max((dst + 0.5) * scale - 0.5, 0) (align_corners = False)
    int up, Hin, Win;
    double sh, sw;
};

// bilinear taps of one upsampled pixel: four source offsets inside a channel plane and their weights
struct Lerp4 {
    int o00, o01, o10, o11;
    float w00, w01, w10, w11;
};
__device__ __forceinline__ Lerp4 lerp4_of(const ConvParams &p, int iy, int ix) {
    double sy = ((double)iy + 0.5) * p.sh - 0.5, sx = ((double)ix + 0.5) * p.sw - 0.5;
    if (sy < 0.0) sy = 0.0;
    if (sx < 0.0) sx = 0.0;
    int y0 = (int)sy, x0 = (int)sx;
    if (y0 > p.Hin - 1) y0 = p.Hin - 1;
    if (x0 > p.Win - 1) x0 = p.Win - 1;
    const int y1 = y0 + (y0 < p.Hin - 1 ? 1 : 0), x1 = x0 + (x0 < p.Win - 1 ? 1 : 0);
    const double ly1 = sy - (double)y0, lx1 = sx - (double)x0, ly0 = 1.0 - ly1, lx0 = 1.0 - lx1;
    Lerp4 l;
    l.o00 = y0 * p.Win + x0; l.o01 = y0 * p.Win + x1; l.o10 = y1 * p.Win + x0; l.o11 = y1 * p.Win + x1;
    l.w00 = (float)(ly0 * lx0); l.w01 = (float)(ly0 * lx1); l.w10 = (float)(ly1 * lx0); l.w11 = (float)(ly1 * lx1);
    return l;
}
template <typename IO>
__device__ __forceinline__ float lerp_load(const IO *plane, const Lerp4 &l) {
    return l.w00 * (float)__ldg(plane + l.o00) + l.w01 * (float)__ldg(plane + l.o01) + l.w10 * (float)__ldg(plane + l.o10) +
           l.w11 * (float)__ldg(plane + l.o11);
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// band of image rows [y0 - PH, y0 + th + PH) x columns [-PH, W + PH) of all channels, + add_offset (the zero padding of
// torch.nn.Unfold becomes add_offset: the reference adds 0.1 to the unfolded patch).  The rows that exist are one contiguous
// run per channel: a thread keeps three positions of that run (its tile offsets computed once per band) and walks the
// channels with twelve loads in flight; everything else in the tile is add_offset, written first.
template <typename IO, int KS>
__device__ __forceinline__ void stage_image(const ConvParams &p, const IO *ib, int y0, int th, float *tile, int tid, int T) {
    constexpr int PH = KS / 2;
    const int HW = p.H * p.W;
    const int lo = max(y0 - PH, 0), hi = min(y0 + th + PH, p.H);
    const int len = (hi - lo) * p.W, rofs = lo - (y0 - PH);
    {
        const float4 o4 = make_float4(p.add_offset, p.add_offset, p.add_offset, p.add_offset);
        float4 *t4 = reinterpret_cast<float4 *>(tile);
        const int n4 = (p.C * p.CS + 3) >> 2;
        for (int i = tid; i < n4; i += T) t4[i] = o4;
    }
    __syncthreads();
    const IO *src = ib + lo * p.W;
    for (int i0 = tid; i0 < len; i0 += 3 * T) {
        const IO *sp[3];
        float *dp[3];
        bool in[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int idx = i0 + k * T;
            in[k] = idx < len;
            const int r = idx / p.W, x = idx - r * p.W;
            sp[k] = src + (in[k] ? idx : 0);
            dp[k] = tile + (r + rofs) * p.TC + x + PH;
        }
#pragma unroll 4
        for (int ch = 0; ch < p.C; ++ch) {
            float v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                v[k] = (float)__ldg(sp[k]);
                sp[k] += HW;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (in[k]) *dp[k] = v[k] + p.add_offset;
                dp[k] += p.CS;
            }
        }
    }
}

// the same tile for a 1 x 1 window on the bilinear upsample of `ib` (n-th image of the (C, Hin, Win) source): the taps of a
// position are computed once and reused for every channel
template <typename IO>
__device__ __forceinline__ void stage_image_up(const ConvParams &p, const IO *ib, int y0, int th, float *tile, int tid, int T) {
    const int len = th * p.W, HWs = p.Hin * p.Win;
    for (int i0 = tid; i0 < len; i0 += 2 * T) {
        Lerp4 l[2];
        float *dp[2];
        bool in[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int idx = i0 + k * T;
            in[k] = idx < len;
            const int r = idx / p.W, x = idx - r * p.W;
            l[k] = lerp4_of(p, in[k] ? y0 + r : y0, in[k] ? x : 0);
            dp[k] = tile + r * p.TC + x;
        }
#pragma unroll 2
        for (int ch = 0; ch < p.C; ++ch) {
            float v[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) v[k] = lerp_load(ib + (long long)ch * HWs, l[k]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (in[k]) *dp[k] = v[k] + p.add_offset;
                dp[k] += p.CS;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------ forward
template <typename IO, int KS, int NP, bool UP = false>
__global__ void __launch_bounds__(256, NP >= 32 ? 2 : 3) conv_fwd_kernel(const ConvParams p) {
    extern __shared__ float4 conv_smem[];
    constexpr int KK = KS * KS, YS = NP + 4;
    float *wd = reinterpret_cast<float *>(conv_smem);
    float *tile = wd + (((p.F + 1) * NP + 3) & ~3);
    const int T = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < (p.F + 1) * NP; i += T) wd[i] = __ldg(p.Wd + i);
    const int HW = p.H * p.W;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        const int b = unit / p.bands, band = unit - b * p.bands;
        const int y0 = band * p.TH;
        const int th = min(p.TH, p.H - y0);
        __syncthreads();
        if constexpr (UP) {
            stage_image_up<IO>(p, reinterpret_cast<const IO *>(p.img) + (long long)b * p.C * p.Hin * p.Win, y0, th, tile, tid, T);
        } else {
            stage_image<IO, KS>(p, reinterpret_cast<const IO *>(p.img) + (long long)b * p.C * HW, y0, th, tile, tid, T);
        }
        __syncthreads();
        const int npx = th * p.W;
        int q[2], base[2];
        bool ok[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            q[k] = tid + k * T;
            ok[k] = q[k] < npx;
            const int qq = ok[k] ? q[k] : 0;
            const int y = qq / p.W, x = qq - y * p.W;
            base[k] = y * p.TC + x;
        }
        float2 acc[2][NP / 2];
#pragma unroll
        for (int j = 0; j < NP / 2; ++j) acc[0][j] = acc[1][j] = *reinterpret_cast<const float2 *>(wd + p.F * NP + 2 * j);
        float ss[2] = {0.f, 0.f};
#pragma unroll 1
        for (int ch = 0; ch < p.C; ++ch) {
            const float *t0 = tile + ch * p.CS + base[0], *t1 = tile + ch * p.CS + base[1];
            const float4 *w4 = reinterpret_cast<const float4 *>(wd + ch * KK * NP);
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const float f0 = t0[ky * p.TC + kx], f1 = t1[ky * p.TC + kx];
                    ss[0] = fmaf(f0, f0, ss[0]);
                    ss[1] = fmaf(f1, f1, ss[1]);
                    const float2 ff0 = f2(f0, f0), ff1 = f2(f1, f1);
#pragma unroll
                    for (int j = 0; j < NP / 4; ++j) {
                        const float4 w = w4[(ky * KS + kx) * (NP / 4) + j];
                        acc[0][2 * j] = __ffma2_rn(f2(w.x, w.y), ff0, acc[0][2 * j]);
                        acc[0][2 * j + 1] = __ffma2_rn(f2(w.z, w.w), ff0, acc[0][2 * j + 1]);
                        acc[1][2 * j] = __ffma2_rn(f2(w.x, w.y), ff1, acc[1][2 * j]);
                        acc[1][2 * j + 1] = __ffma2_rn(f2(w.z, w.w), ff1, acc[1][2 * j + 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!ok[k]) continue;
            const int pix = y0 * p.W + q[k];
            const long long gp = (long long)b * HW + pix;
            const float s = ss[k] + p.pad2;
            const float inv = s > 0.f ? 1.0f / s : 0.f;
            if (p.Y != nullptr) {
                float4 *yr = reinterpret_cast<float4 *>(p.Y + gp * YS);
#pragma unroll
                for (int j = 0; j < NP / 4; ++j) yr[j] = make_float4(acc[k][2 * j].x, acc[k][2 * j].y, acc[k][2 * j + 1].x, acc[k][2 * j + 1].y);
                yr[NP / 4] = make_float4(inv, 0.f, 0.f, 0.f);
            }
            IO *ob = reinterpret_cast<IO *>(p.out) + (long long)b * p.n_out * HW + pix;
#pragma unroll
            for (int m = 0; m < NP / 2; ++m) {
                if (m < p.n_out) {
                    float v = p.post_scale * inv * (acc[k][m].x * acc[k][m].x + acc[k][m].y * acc[k][m].y);
                    if (p.clamp) v = fminf(fmaxf(v, p.clamp_lo), p.clamp_hi);
                    ob[(long long)m * HW] = (IO)v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ dL/dY per patch
// G[n] = 2 c_m Y[n] / |f|^2 with c_m = post * grad_out_m on the outputs the clamp passes, and S / |f|^2 with
// S = sum_m c_m |Y_m|^2 / |f|^2 (the normalisation term of d out / d f) at column NP.
template <typename IO, int NP>
__global__ void __launch_bounds__(256) conv_grad_kernel(const ConvParams p) {
    constexpr int YS = NP + 4, M = NP / 2;
    const int HW = p.H * p.W;
    for (long long gp = (long long)blockIdx.x * blockDim.x + threadIdx.x; gp < p.B; gp += (long long)gridDim.x * blockDim.x) {
        const long long b = gp / HW;
        const int pix = (int)(gp - b * HW);
        const IO *gob = reinterpret_cast<const IO *>(p.go) + b * p.n_out * HW + pix;
        const float4 *yr = reinterpret_cast<const float4 *>(p.Y + gp * YS);
        float4 y4[NP / 4];
#pragma unroll
        for (int j = 0; j < NP / 4; ++j) y4[j] = __ldg(yr + j);
        const float inv = __ldg(reinterpret_cast<const float *>(yr + NP / 4));
        float go[M];
#pragma unroll
        for (int m = 0; m < M; ++m) go[m] = m < p.n_out ? (float)__ldg(gob + (long long)m * HW) : 0.f;
        float4 *gr = reinterpret_cast<float4 *>(p.G + gp * YS);
        float S = 0.f;
#pragma unroll
        for (int j = 0; j < NP / 4; ++j) {
            float o[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float a = h ? y4[j].z : y4[j].x, bb = h ? y4[j].w : y4[j].y;
                const float pr = a * a + bb * bb;
                const float v = p.post_scale * inv * pr;
                const bool pass = !p.clamp || (v >= p.clamp_lo && v <= p.clamp_hi);
                const float cm = pass ? p.post_scale * go[2 * j + h] : 0.f;
                o[2 * h] = 2.f * cm * inv * a;
                o[2 * h + 1] = 2.f * cm * inv * bb;
                S += cm * inv * pr;
            }
            gr[j] = make_float4(o[0], o[1], o[2], o[3]);
        }
        gr[NP / 4] = make_float4(inv * S, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------- image gradient (gather)
template <typename IO, int KS, int NP, int CT, bool UP = false>
__global__ void __launch_bounds__(256, 2) conv_bwd_data_kernel(const ConvParams p) {
    extern __shared__ float4 conv_smem[];
    constexpr int KK = KS * KS, PH = KS / 2, GS = NP + 4, NC = NP < 16 ? NP : 16, NQ = NC / 4;
    const int c_pad = (p.C + CT - 1) / CT * CT;
    float *wd = reinterpret_cast<float *>(conv_smem);               // [c_pad KK][NP], rows >= F zero
    float *gt = wd + c_pad * KK * NP;                               // [th + KS - 1][W + KS - 1][GS]
    const int T = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < c_pad * KK * NP; i += T) wd[i] = i < p.F * NP ? __ldg(p.Wd + i) : 0.f;
    const int HW = p.H * p.W, tcg = p.W + KS - 1;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        const int b = unit / p.bands, band = unit - b * p.bands;
        const int y0 = band * p.TH;
        const int th = min(p.TH, p.H - y0);
        __syncthreads();
        for (int i = tid; i < (th + KS - 1) * tcg; i += T) {
            const int r = i / tcg, c = i - r * tcg;
            const int py = y0 - PH + r, px = c - PH;
            float4 *g = reinterpret_cast<float4 *>(gt + i * GS);
            if (py >= 0 && py < p.H && px >= 0 && px < p.W) {
                const float4 *src = reinterpret_cast<const float4 *>(p.G + ((long long)b * HW + py * p.W + px) * GS);
#pragma unroll
                for (int j = 0; j < GS / 4; ++j) cp_async16(g + j, src + j);
            } else {
#pragma unroll
                for (int j = 0; j < GS / 4; ++j) g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_wait_all();
        __syncthreads();
        const int npx = th * p.W;
        int q[2], gbase[2];
        bool ok[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            q[k] = tid + k * T;
            ok[k] = q[k] < npx;
            const int qq = ok[k] ? q[k] : 0;
            const int y = qq / p.W, x = qq - y * p.W;
            gbase[k] = ((y + KS - 1) * tcg + (x + KS - 1)) * GS;      // tap (ky, kx) reads the patch at (y - ky + PH, x - kx + PH)
        }
        float nsum[2] = {0.f, 0.f};
#pragma unroll 1
        for (int ct = 0; ct < c_pad; ct += CT) {
            float2 acc[2][CT];
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) acc[0][ch] = acc[1][ch] = f2(0.f, 0.f);
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const float *g0 = gt + gbase[0] - (ky * tcg + kx) * GS, *g1 = gt + gbase[1] - (ky * tcg + kx) * GS;
                    if (ct == 0) {
                        nsum[0] += g0[NP];
                        nsum[1] += g1[NP];
                    }
#pragma unroll
                    for (int nc = 0; nc < NP / NC; ++nc) {
                        float4 G0[NQ], G1[NQ];
#pragma unroll
                        for (int j = 0; j < NQ; ++j) {
                            G0[j] = *reinterpret_cast<const float4 *>(g0 + nc * NC + 4 * j);
                            G1[j] = *reinterpret_cast<const float4 *>(g1 + nc * NC + 4 * j);
                        }
#pragma unroll
                        for (int ch = 0; ch < CT; ++ch) {
                            const float4 *w4 = reinterpret_cast<const float4 *>(wd + ((ct + ch) * KK + ky * KS + kx) * NP + nc * NC);
#pragma unroll
                            for (int j = 0; j < NQ; ++j) {
                                const float4 w = w4[j];
                                acc[0][ch] = __ffma2_rn(f2(G0[j].x, G0[j].y), f2(w.x, w.y), acc[0][ch]);
                                acc[0][ch] = __ffma2_rn(f2(G0[j].z, G0[j].w), f2(w.z, w.w), acc[0][ch]);
                                acc[1][ch] = __ffma2_rn(f2(G1[j].x, G1[j].y), f2(w.x, w.y), acc[1][ch]);
                                acc[1][ch] = __ffma2_rn(f2(G1[j].z, G1[j].w), f2(w.z, w.w), acc[1][ch]);
                            }
                        }
                    }
                }
            }
            // the image values of the normalisation term: every load issued before the first store (the compiler keeps the
            // order of a load behind a store to another global pointer: one DRAM latency per channel otherwise)
            float fv[2][CT];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int pix = y0 * p.W + (ok[k] ? q[k] : 0);
                if constexpr (UP) {       // the convolution's input is the bilinear upsample of p.img: re-interpolated here
                    const int iy = pix / p.W;
                    const Lerp4 l = lerp4_of(p, iy, pix - iy * p.W);
#pragma unroll
                    for (int ch = 0; ch < CT; ++ch) {
                        const int c = ct + ch < p.C ? ct + ch : p.C - 1;
                        fv[k][ch] = lerp_load(reinterpret_cast<const IO *>(p.img) + ((long long)b * p.C + c) * p.Hin * p.Win, l);
                    }
                } else {
#pragma unroll
                    for (int ch = 0; ch < CT; ++ch) {
                        const int c = ct + ch < p.C ? ct + ch : p.C - 1;
                        fv[k][ch] = (float)__ldg(reinterpret_cast<const IO *>(p.img) + ((long long)b * p.C + c) * HW + pix);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (!ok[k]) continue;
                const int pix = y0 * p.W + q[k];
#pragma unroll
                for (int ch = 0; ch < CT; ++ch) {
                    const int c = ct + ch;
                    if (c < p.C) {
                        const long long idx = ((long long)b * p.C + c) * HW + pix;
                        reinterpret_cast<IO *>(p.gimg)[idx] =
                            (IO)((acc[k][ch].x + acc[k][ch].y) - 2.f * (fv[k][ch] + p.add_offset) * nsum[k]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------- weight gradient
template <typename IO, int KS, int NP, bool UP = false>
__global__ void __launch_bounds__(256, 2) conv_bwd_w_kernel(const ConvParams p) {
    extern __shared__ float4 conv_smem[];
    constexpr int KK = KS * KS, GS = NP + 4, NC = NP < 16 ? NP : 16, NQ = NC / 4;
    float *dacc = reinterpret_cast<float *>(conv_smem);             // [(F + 1)][NP] sums of this CTA
    float *tile = dacc + (((p.F + 1) * NP + 3) & ~3);               // [C][CS]
    float *gt = tile + ((p.C * p.CS + 3) & ~3);                     // [TH W][GS]
    const int T = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (p.F + 1) * NP; i += T) dacc[i] = 0.f;
    const int HW = p.H * p.W;
    // warp -> (n-chunk, 32 rows of the kernel window, part of the band's pixel rows)
    const int ig = warp % (p.wr * (NP / NC)), part = warp / (p.wr * (NP / NC));
    const int nchunk = ig / p.wr;
    const int row = (ig % p.wr) * 32 + lane;                        // (channel, ky)
    const bool rowok = row < p.C * KS;
    const int ch = rowok ? row / KS : 0, ky = rowok ? row - ch * KS : 0;
    float2 acc[KS][NC / 2];
#pragma unroll
    for (int kx = 0; kx < KS; ++kx)
#pragma unroll
        for (int j = 0; j < NC / 2; ++j) acc[kx][j] = f2(0.f, 0.f);
    float bsum = 0.f;

    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        const int b = unit / p.bands, band = unit - b * p.bands;
        const int y0 = band * p.TH;
        const int th = min(p.TH, p.H - y0);
        const int npx = th * p.W;
        __syncthreads();
        {   // the band's rows of G are one contiguous run
            const float4 *src = reinterpret_cast<const float4 *>(p.G + ((long long)b * HW + y0 * p.W) * GS);
            float4 *dst = reinterpret_cast<float4 *>(gt);
            for (int i = tid; i < npx * (GS / 4); i += T) cp_async16(dst + i, src + i);
        }
        if constexpr (UP) {
            stage_image_up<IO>(p, reinterpret_cast<const IO *>(p.img) + (long long)b * p.C * p.Hin * p.Win, y0, th, tile, tid, T);
        } else {
            stage_image<IO, KS>(p, reinterpret_cast<const IO *>(p.img) + (long long)b * p.C * HW, y0, th, tile, tid, T);
        }
        cp_async_wait_all();
        __syncthreads();
        {   // bias row: column sums of G, kept per thread until the end of the launch
            const int n = tid % NP, sub = tid / NP, nsub = T / NP;
            if (sub < nsub)
                for (int i = sub; i < npx; i += nsub) bsum += gt[i * GS + n];
        }
        if (part < p.parts) {
            for (int y = part; y < th; y += p.parts) {
                const float *fr = tile + ch * p.CS + (y + ky) * p.TC;
                const float *g = gt + (y * p.W) * GS + nchunk * NC;
                float f0 = fr[0], f1 = KS > 1 ? fr[1] : 0.f;
#pragma unroll 2
                for (int x = 0; x < p.W; ++x) {
                    const float fn = fr[x + KS - 1];
                    float4 G[NQ];
#pragma unroll
                    for (int j = 0; j < NQ; ++j) G[j] = *reinterpret_cast<const float4 *>(g + 4 * j);
                    g += GS;
                    if constexpr (KS == 3) {
                        const float2 a = f2(f0, f0), bb = f2(f1, f1), c = f2(fn, fn);
#pragma unroll
                        for (int j = 0; j < NQ; ++j) {
                            const float2 lo = f2(G[j].x, G[j].y), hi = f2(G[j].z, G[j].w);
                            acc[0][2 * j] = __ffma2_rn(lo, a, acc[0][2 * j]);
                            acc[0][2 * j + 1] = __ffma2_rn(hi, a, acc[0][2 * j + 1]);
                            acc[1][2 * j] = __ffma2_rn(lo, bb, acc[1][2 * j]);
                            acc[1][2 * j + 1] = __ffma2_rn(hi, bb, acc[1][2 * j + 1]);
                            acc[2][2 * j] = __ffma2_rn(lo, c, acc[2][2 * j]);
                            acc[2][2 * j + 1] = __ffma2_rn(hi, c, acc[2][2 * j + 1]);
                        }
                        f0 = f1;
                        f1 = fn;
                    } else {
                        const float2 c = f2(fn, fn);
#pragma unroll
                        for (int j = 0; j < NQ; ++j) {
                            acc[0][2 * j] = __ffma2_rn(f2(G[j].x, G[j].y), c, acc[0][2 * j]);
                            acc[0][2 * j + 1] = __ffma2_rn(f2(G[j].z, G[j].w), c, acc[0][2 * j + 1]);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    // the warps of one pixel-row part own distinct (row, n-chunk) entries: the parts add in turn (no atomics, fixed order)
    for (int pt = 0; pt < p.parts; ++pt) {
        if (part == pt && rowok) {
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                float *d = dacc + ((ch * KK + ky * KS + kx) * NP + nchunk * NC);
#pragma unroll
                for (int j = 0; j < NC / 2; ++j) {
                    d[2 * j] += acc[kx][j].x;
                    d[2 * j + 1] += acc[kx][j].y;
                }
            }
        }
        __syncthreads();
    }
    float *bred = gt + p.TH * p.W * GS;  // [sub][n] partial column sums of G
    bred[tid] = bsum;
    __syncthreads();
    if (tid < NP) {
        float sb = 0.f;
        for (int sub = 0; sub < T / NP; ++sub) sb += bred[sub * NP + tid];
        dacc[p.F * NP + tid] = sb;
    }
    __syncthreads();
    float *dst = p.partials + (long long)blockIdx.x * (p.F + 1) * NP;
    for (int i = tid; i < (p.F + 1) * NP; i += T) dst[i] = dacc[i];
}

// dWd[(F + 1) NP] = sum of the per-CTA partials, fixed order, double precision: 32 outputs x 8 slices of the partials per CTA
__global__ void __launch_bounds__(256) conv_reduce_kernel(const float *partials, int n_part, int total, double *sum) {
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int o = blockIdx.x * 32 + lane;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (o < total) {
        const float *src = partials + o;
        int j = slice;
        for (; j + 56 < n_part; j += 64) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (long long)(j + 8 * u) * total);
            s0 += (double)v[0] + (double)v[4];
            s1 += (double)v[1] + (double)v[5];
            s2 += (double)v[2] + (double)v[6];
            s3 += (double)v[3] + (double)v[7];
        }
        for (; j < n_part; j += 8) s0 += __ldg(src + (long long)j * total);
    }
    red[slice][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (slice == 0 && o < total) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][lane];
        sum[o] = s;
    }
}

// READ_STATE cotangent of U^T: gUT[c][m stride].{re,im} = c < F ? dWd[c][n] : pad * dbias[n], n = 2m + ri
__global__ void __launch_bounds__(256) conv_assemble_kernel(const double *sum, int A, int F, int N, int NP, int stride, float pad,
                                                            float *gUT) {
    const long long total = (long long)A * A * 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ri = (int)(i & 1);
        const long long ck = i >> 1;
        const int c = (int)(ck / A), kk = (int)(ck % A);
        float v = 0.f;
        if (kk % stride == 0) {
            const int n = 2 * (kk / stride) + ri;
            if (n < N) v = (float)(c < F ? sum[c * NP + n] : (double)pad * sum[F * NP + n]);
        }
        gUT[i] = v;
    }
}

// Wd[c][n] = part_n(U^T[c][m stride]) (c < F), row F = pad * sum over the pad rows c >= F; columns n >= N are zero
__global__ void build_wd_kernel(const float2 *UT, int A, int F, int N, int NP, int stride, float pad, float *Wd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (F + 1) * NP) return;
    const int c = i / NP, n = i - c * NP;
    float v = 0.f;
    if (n < N) {
        const int k = (n >> 1) * stride;
        if (c < F) {
            const float2 u = UT[(long long)c * A + k];
            v = (n & 1) ? u.y : u.x;
        } else {
            double s = 0.0;
            for (int r = F; r < A; ++r) {
                const float2 u = UT[(long long)r * A + k];
                s += (n & 1) ? u.y : u.x;
            }
            v = (float)((double)pad * s);
        }
    }
    Wd[i] = v;
}

// ------------------------------------------------------------------------------------------------------ host side
enum { KIND_FWD = 0, KIND_DATA = 1, KIND_W = 2 };

struct ConvTile {
    int TH, T, TC, CS, bands, wr, parts, ct;
    size_t smem;
};

constexpr size_t CONV_SMEM_MAX = 200 * 1024;
constexpr size_t CONV_SMEM_PREF = 72 * 1024;       // three CTAs per SM

// Band height of one kernel: fills the 2-pixels-per-thread CTA (forward / image gradient) or balances the pixel rows over the
// warps (weight gradient) with the least halo, preferring a footprint that leaves three CTAs per SM.
bool conv_tile(const GemmShape &g, const GateParams &gp, int NP, int kind, ConvTile *t) {
    const int KS = gp.kh, GS = NP + 4, NC = NP < 16 ? NP : 16;
    const int wr = (gp.C * KS + 31) / 32, ig = wr * (NP / NC);
    if (ig > 8) return false;
    const int parts = 8 / ig, Tw = 32 * ig * parts;
    const int ct = gp.C > 8 ? 16 : 8;
    const int c_pad = (gp.C + ct - 1) / ct * ct;
    const size_t wdf = (size_t)(((g.F + 1) * NP + 3) & ~3);
    double best[2] = {0.0, 0.0};       // [0]: any footprint, [1]: preferred footprint
    ConvTile cand[2];
    for (int th = 1; th <= gp.H; ++th) {
        const int pu = th * gp.W;
        if (kind != KIND_W && pu > 512) break;
        if (kind == KIND_W && pu > 2048) break;
        int T = ((pu + 63) / 64) * 32;
        if (T < 64) T = 64;
        if (kind == KIND_W) T = Tw;
        const int bands = (gp.H + th - 1) / th;
        const int tc = (gp.W + KS - 1) | 1;
        const int trtc = (th + KS - 1) * tc;
        const int cs = trtc + ((3 - trtc % 32) + 32) % 32;
        const size_t tile = ((size_t)gp.C * cs + 3) & ~(size_t)3;
        size_t smem;
        double eff;
        if (kind == KIND_FWD) {
            smem = 4 * (wdf + tile);
            eff = (double)gp.H * gp.W / ((double)bands * 2 * T) * (1.0 - 0.02 * (KS - 1) / th);
        } else if (kind == KIND_DATA) {
            smem = 4 * ((size_t)c_pad * KS * KS * NP + (size_t)(th + KS - 1) * (gp.W + KS - 1) * GS);
            eff = (double)gp.H * gp.W / ((double)bands * 2 * T) * ((double)th / (th + KS - 1));
        } else {
            smem = 4 * (wdf + tile + (size_t)pu * GS + 256);
            const int rounds = (th + parts - 1) / parts;
            eff = (double)gp.H / ((double)bands * rounds * parts) * ((double)th / (th + KS - 1));
        }
        if (smem > CONV_SMEM_MAX) continue;
        ConvTile c;
        c.TH = th; c.T = T; c.TC = tc; c.CS = cs; c.bands = bands; c.wr = wr; c.parts = parts; c.ct = ct; c.smem = smem;
        // ties go to the taller band (fewer units, less halo)
        if (eff > best[0] - 1e-9) { best[0] = eff > best[0] ? eff : best[0]; cand[0] = c; }
        if (smem <= CONV_SMEM_PREF && eff > best[1] - 1e-9) { best[1] = eff > best[1] ? eff : best[1]; cand[1] = c; }
    }
    if (best[0] <= 0.0) return false;
    *t = (best[1] >= 0.75 * best[0]) ? cand[1] : cand[0];
    return true;
}

template <typename K>
int conv_grid(K kern, int T, size_t smem, int units, int cap_ctas) {
    // attribute / occupancy queries remembered per kernel, device and launch shape (they cost ~10 us of host time)
    struct Entry { const void *k; int dev, T; size_t smem; int per_sm, sms; };
    thread_local Entry cache[96];
    thread_local int n_cache = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const void *kp = reinterpret_cast<const void *>(kern);
    int per_sm = 0, sms = 0;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].k == kp && cache[i].dev == dev && cache[i].T == T && cache[i].smem == smem) { per_sm = cache[i].per_sm; sms = cache[i].sms; }
    if (per_sm == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV_SMEM_MAX) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem) != cudaSuccess || per_sm < 1) return -1;
        if (n_cache < 96) cache[n_cache++] = Entry{kp, dev, T, smem, per_sm, sms};
    }
    long long cap = (long long)per_sm * sms;
    if (cap_ctas > 0 && cap > cap_ctas) cap = cap_ctas;
    return (int)(units < cap ? units : cap);
}

ConvParams conv_params(const GemmShape &g, const GateParams &gp, const ConvTile &t, long long n_images) {
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.n_images = (int)n_images; p.C = gp.C; p.H = gp.H; p.W = gp.W; p.F = g.F; p.N = g.N; p.n_out = g.n_out;
    p.bands = t.bands; p.TH = t.TH; p.TC = t.TC; p.CS = t.CS; p.units = (int)(n_images * t.bands);
    p.clamp = gp.clamp; p.clamp_lo = gp.clamp_lo; p.clamp_hi = gp.clamp_hi; p.post_scale = gp.post_scale;
    p.add_offset = gp.add_offset;
    p.pad2 = (float)(g.A - g.F) * gp.pad_value * gp.pad_value;
    p.wr = t.wr; p.parts = t.parts;
    p.B = n_images * gp.H * gp.W;
    return p;
}

#define CONV_DISPATCH_NP(CALL)                   \
    do {                                         \
        if (NP == 4) { CALL(4); }                \
        else if (NP == 16) { CALL(16); }         \
        else { CALL(32); }                       \
    } while (0)

template <typename IO, int KS>
int conv_forward_t(const ConvParams &p, const ConvTile &t, int NP, cudaStream_t s) {
#define CALL(NP_)                                                                                   \
    {                                                                                               \
        auto k = conv_fwd_kernel<IO, KS, NP_, false>;                                               \
        if constexpr (KS == 1) { if (p.up) k = conv_fwd_kernel<IO, KS, NP_, true>; }                \
        const int grid = conv_grid(k, t.T, t.smem, p.units, 0);                                     \
        if (grid < 1) return QIDDM_EUNSUPPORTED;                                                    \
        k<<<grid, t.T, t.smem, s>>>(p);                                                             \
    }
    CONV_DISPATCH_NP(CALL);
#undef CALL
    return QIDDM_OK;
}

template <typename IO, int KS>
int conv_bwd_data_t(const ConvParams &p, const ConvTile &t, int NP, cudaStream_t s) {
#define CALL(NP_)                                                                                   \
    {                                                                                               \
        if (t.ct == 8) {                                                                            \
            auto k = conv_bwd_data_kernel<IO, KS, NP_, 8, false>;                                   \
            if constexpr (KS == 1) { if (p.up) k = conv_bwd_data_kernel<IO, KS, NP_, 8, true>; }    \
            const int grid = conv_grid(k, t.T, t.smem, p.units, 0);                                 \
            if (grid < 1) return QIDDM_EUNSUPPORTED;                                                \
            k<<<grid, t.T, t.smem, s>>>(p);                                                         \
        } else {                                                                                    \
            auto k = conv_bwd_data_kernel<IO, KS, NP_, 16, false>;                                  \
            if constexpr (KS == 1) { if (p.up) k = conv_bwd_data_kernel<IO, KS, NP_, 16, true>; }   \
            const int grid = conv_grid(k, t.T, t.smem, p.units, 0);                                 \
            if (grid < 1) return QIDDM_EUNSUPPORTED;                                                \
            k<<<grid, t.T, t.smem, s>>>(p);                                                         \
        }                                                                                           \
    }
    CONV_DISPATCH_NP(CALL);
#undef CALL
    return QIDDM_OK;
}

template <typename IO, int KS>
int conv_bwd_w_t(ConvParams &p, const ConvTile &t, int NP, int *grid_out, cudaStream_t s) {
#define CALL(NP_)                                                                                   \
    {                                                                                               \
        auto k = conv_bwd_w_kernel<IO, KS, NP_, false>;                                             \
        if constexpr (KS == 1) { if (p.up) k = conv_bwd_w_kernel<IO, KS, NP_, true>; }              \
        const int grid = conv_grid(k, t.T, t.smem, p.units, CONV_MAX_WGRID);                        \
        if (grid < 1) return QIDDM_EUNSUPPORTED;                                                    \
        *grid_out = grid;                                                                           \
        k<<<grid, t.T, t.smem, s>>>(p);                                                             \
    }
    CONV_DISPATCH_NP(CALL);
#undef CALL
    return QIDDM_OK;
}

}  // namespace

int conv_np(int N) { return N <= 4 ? 4 : (N <= 16 ? 16 : (N <= 32 ? 32 : 0)); }

size_t conv_wd_bytes(const GemmShape &g) {
    const int NP = conv_np(g.N);
    return NP ? (((size_t)(g.F + 1) * NP * 4 + 255) & ~(size_t)255) : 0;
}

int conv_build_wd(const GemmShape &g, const GateParams &gp, const float *UT, float *Wd, cudaStream_t s) {
    const int NP = conv_np(g.N);
    if (!NP) return QIDDM_OK;
    const int total = (g.F + 1) * NP;
    build_wd_kernel<<<(total + 127) / 128, 128, 0, s>>>(reinterpret_cast<const float2 *>(UT), g.A, g.F, g.N, NP, g.stride,
                                                        gp.pad_value, Wd);
    count_launch();
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

bool conv_direct_supported(const GemmShape &g, const GateParams &gp) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("QIDDM_QCONV_DIRECT"); on = (e && e[0] == '0') ? 0 : 1; }
    if (!on || !gp.unfold) return false;
    if (gp.kh != gp.kw || (gp.kh != 1 && gp.kh != 3)) return false;
    if (gp.ph != gp.kh / 2 || gp.pw != gp.kw / 2) return false;       // "same" geometry: one patch per pixel
    if (g.F != gp.C * gp.kh * gp.kw) return false;
    const int NP = conv_np(g.N);
    if (!NP) return false;
    ConvTile t;
    for (int kind = KIND_FWD; kind <= KIND_W; ++kind)
        if (!conv_tile(g, gp, NP, kind, &t)) return false;
    return true;
}

size_t conv_direct_saved_bytes(const GemmShape &g, const GateParams &gp, long long n_images) {
    return (size_t)n_images * gp.H * gp.W * (conv_np(g.N) + 4) * 4 + 256;        // one row [Y (NP), 1/|f|^2, pad] per patch
}

static size_t conv_partials_bytes(const GemmShape &g, int NP) { return ((size_t)CONV_MAX_WGRID * (g.F + 1) * NP * 4 + 255) & ~(size_t)255; }
static size_t conv_gut_bytes(const GemmShape &g) { return ((size_t)g.A * g.A * 8 + 255) & ~(size_t)255; }
static size_t conv_sum_bytes(const GemmShape &g, int NP) { return ((size_t)(g.F + 1) * NP * 8 + 255) & ~(size_t)255; }

size_t conv_direct_ws_bytes(const GemmShape &g, const GateParams &gp, long long n_images) {
    const int NP = conv_np(g.N);
    return conv_partials_bytes(g, NP) + conv_gut_bytes(g) + conv_sum_bytes(g, NP) +
           (((size_t)n_images * gp.H * gp.W * (NP + 4) * 4 + 255) & ~(size_t)255);                 // G rows
}

// out (NCHW, io dtype); `saved` non-null (training): Y and 1/|f|^2 are kept for conv_direct_backward
static int conv_set_up(ConvParams &p, const GateParams &gp, const ConvUp *up) {
    if (up == nullptr || up->h_in <= 0) return QIDDM_OK;
    if (gp.kh != 1 || up->w_in <= 0 || !(up->scale_h > 0.0) || !(up->scale_w > 0.0)) return QIDDM_EUNSUPPORTED;
    p.up = 1; p.Hin = up->h_in; p.Win = up->w_in; p.sh = up->scale_h; p.sw = up->scale_w;
    return QIDDM_OK;
}

int conv_direct_forward(const GemmShape &g, const GateParams &gp, const float *Wd, const void *img, void *out, void *saved,
                        long long n_images, cudaStream_t s, const ConvUp *up) {
    const int NP = conv_np(g.N);
    ConvTile t;
    if (!NP || !conv_tile(g, gp, NP, KIND_FWD, &t)) return QIDDM_EUNSUPPORTED;
    ConvParams p = conv_params(g, gp, t, n_images);
    if (conv_set_up(p, gp, up) != QIDDM_OK) return QIDDM_EUNSUPPORTED;
    p.img = img; p.out = out; p.Wd = Wd;
    p.Y = reinterpret_cast<float *>(saved);
    timing_begin(TK_CONV_FWD, 2.0 * (double)n_images * gp.H * gp.W * g.F * g.N, s);
    int rc;
    if (gp.io64) rc = gp.kh == 3 ? conv_forward_t<double, 3>(p, t, NP, s) : conv_forward_t<double, 1>(p, t, NP, s);
    else rc = gp.kh == 3 ? conv_forward_t<float, 3>(p, t, NP, s) : conv_forward_t<float, 1>(p, t, NP, s);
    timing_end(s);
    count_launch();
    if (rc != QIDDM_OK) return rc;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

// grad_img (nullable, overwritten) and the READ_STATE cotangent gUT (inside `ws`) for the adjoint gate kernel
int conv_direct_backward(const GemmShape &g, const GateParams &gp, const float *Wd, const void *img, const void *grad_out,
                         const void *saved, void *grad_img, float **gut_out, void *ws, long long n_images, cudaStream_t s,
                         const ConvUp *up) {
    const int NP = conv_np(g.N);
    ConvTile td, tw;
    if (!NP || saved == nullptr || !conv_tile(g, gp, NP, KIND_DATA, &td) || !conv_tile(g, gp, NP, KIND_W, &tw))
        return QIDDM_EUNSUPPORTED;
    char *w8 = reinterpret_cast<char *>(ws);
    float *partials = reinterpret_cast<float *>(w8); w8 += conv_partials_bytes(g, NP);
    float *gUT = reinterpret_cast<float *>(w8); w8 += conv_gut_bytes(g);
    double *sum = reinterpret_cast<double *>(w8); w8 += conv_sum_bytes(g, NP);
    float *G = reinterpret_cast<float *>(w8);
    *gut_out = gUT;
    int rc = QIDDM_OK;
    timing_begin(TK_CONV_BWD, (grad_img ? 4.0 : 2.0) * (double)n_images * gp.H * gp.W * g.F * g.N, s);
    {   // dL/dY rows, shared by the two gradient kernels
        ConvParams p = conv_params(g, gp, td, n_images);
        p.go = grad_out; p.Y = reinterpret_cast<float *>(const_cast<void *>(saved)); p.G = G;
        const long long blocks = (p.B + 255) / 256;
        const unsigned grid = (unsigned)(blocks < 148 * 8 ? blocks : 148 * 8);
#define CALL(NP_)                                                               \
        {                                                                       \
            if (gp.io64) conv_grad_kernel<double, NP_><<<grid, 256, 0, s>>>(p); \
            else conv_grad_kernel<float, NP_><<<grid, 256, 0, s>>>(p);          \
        }
        CONV_DISPATCH_NP(CALL);
#undef CALL
        count_launch();
    }
    if (grad_img != nullptr) {
        ConvParams p = conv_params(g, gp, td, n_images);
        if (conv_set_up(p, gp, up) != QIDDM_OK) return QIDDM_EUNSUPPORTED;
        p.img = img; p.gimg = grad_img; p.Wd = Wd; p.G = G;
        if (gp.io64) rc = gp.kh == 3 ? conv_bwd_data_t<double, 3>(p, td, NP, s) : conv_bwd_data_t<double, 1>(p, td, NP, s);
        else rc = gp.kh == 3 ? conv_bwd_data_t<float, 3>(p, td, NP, s) : conv_bwd_data_t<float, 1>(p, td, NP, s);
        count_launch();
    }
    int wgrid = 0;
    if (rc == QIDDM_OK) {
        ConvParams p = conv_params(g, gp, tw, n_images);
        if (conv_set_up(p, gp, up) != QIDDM_OK) return QIDDM_EUNSUPPORTED;
        p.img = img; p.G = G; p.partials = partials;
        if (gp.io64) rc = gp.kh == 3 ? conv_bwd_w_t<double, 3>(p, tw, NP, &wgrid, s) : conv_bwd_w_t<double, 1>(p, tw, NP, &wgrid, s);
        else rc = gp.kh == 3 ? conv_bwd_w_t<float, 3>(p, tw, NP, &wgrid, s) : conv_bwd_w_t<float, 1>(p, tw, NP, &wgrid, s);
        count_launch();
    }
    timing_end(s);
    if (rc != QIDDM_OK) return rc;
    const int total = (g.F + 1) * NP;
    conv_reduce_kernel<<<(total + 31) / 32, 256, 0, s>>>(partials, wgrid, total, sum);
    const long long elems = (long long)g.A * g.A * 2;
    conv_assemble_kernel<<<(unsigned)((elems + 255) / 256 < 592 ? (elems + 255) / 256 : 592), 256, 0, s>>>(
        sum, g.A, g.F, g.N, NP, g.stride, gp.pad_value, gUT);
    count_launch(2);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

}  // namespace qiddm
