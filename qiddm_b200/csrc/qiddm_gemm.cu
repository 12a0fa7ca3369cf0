// Unitary-collapse path for the amplitude-embedding families (QDenseUndirected_old[_noise],
// QConv2d): the weight-only circuit is collapsed once per optimizer step into its unitary U (gate
// kernel on the 2^n basis states), after which every circuit instance is one row of a dense real GEMM
//     Y[b, 2m+{0,1}] = sum_c f[b,c] * {Re,Im} U[k_m, c]            (k_m = m * read_stride)
// on the 5th-gen tensor cores: tcgen05.mma (kind::f16, fp32 accumulate in TMEM), operands staged by
// TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 4-stage mbarrier ring, warp-specialised
// (TMA warp / MMA warp / TMEM-alloc warp / 4 epilogue warps), persistent over output tiles with a
// double-buffered TMEM accumulator so the epilogue of tile i overlaps the MMAs of tile i+1.  The
// readout  p = |Y|^2 / |f|^2 * scale -> clamp  is fused into the epilogue (tcgen05.ld -> registers).
//
// Precision: "x3" mode splits every fp32 operand into fp16 hi + lo (22 mantissa bits) and runs the
// three cross terms (hi*hi, lo*hi, hi*lo) as three K-segments of the same accumulator — fp32-grade
// results at 3x the executed flops; "x1" runs hi*hi only (about 3e-4 relative).  Per-step ("activation")
// operands X and G carry two arrays (hi, lo' = lo * 2^11, kept in fp16's normal range); the static
// ("weight-side") operands W and X^T carry three (hi, hi * 2^-11 to pair with lo', and the plain lo, whose
// subnormal rounding is an absolute 2^-25 on O(1) values).
//
// Backward (same GEMM kernel): the forward keeps Y; G = dL/dY is formed in one streaming pass; then
//   dX = G W^T (normalisation term fused in the epilogue) and dW^T = G^T X, where G and X are both consumed
// row-major as MN-major UMMA operands (no transposed copies; split-K over the batch, 16-byte fp32 reductions); dW is pushed
// back through the circuit with the adjoint gate kernel run on the 2^n basis columns (READ_STATE cotangent).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "qiddm_internal.h"

namespace qiddm {

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // 64 fp16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int PAIR_THREADS = 256;   // CTA-pair kernel: 4 role warps + 4 epilogue warps
constexpr int EPI_STAGE_BYTES = 3072;   // per epilogue warp and buffer: 32 x 64 B (Y / dX chunk) + 32 x 32 B (probabilities)
constexpr int ACC_COLS = 256;    // TMEM columns per accumulator stage (2 stages = 512 columns)

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane; the registers are valid after tc_ld_wait()
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// waits for every outstanding tcgen05.ld of this thread; the operands tie the consumers of `r` to the wait
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row groups of 1024 B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}

// K-major, SWIZZLE_64B (32 fp16 per row; 8-row groups of 512 B)
__device__ __forceinline__ uint64_t make_smem_desc_k32(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
    return d;
}

// MN-major, SWIZZLE_128B: each K row is 128 B of 64 contiguous M elements; 8 K rows form a 1024-B atom
// (SBO); the next 64 M elements start LBO = 8192 B later (second TMA box of the 128-row tile).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr, uint32_t lbo = 8192) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ---- CTA-pair (cta_group::2) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrive on the leader's tempty barrier.  RELAXED: what the MMA issuer must observe is that this warp's tcgen05.ld of the
// accumulator have completed, which tcgen05.wait::ld + tcgen05.fence::before_thread_sync order before the arrive; a
// release arrive compiles to MEMBAR.ALL.CTA and makes the warp wait for its global stores of the tile to drain first
// (ncu: 8 % of the epilogue warps' time, and the accumulator stage stays blocked meanwhile).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory that signals an mbarrier of either CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
// the same load with an L2 eviction-priority policy (createpolicy)
__device__ __forceinline__ void tma_load_2d_pair_hint(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1,
                                                      uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

}  // namespace

// ------------------------------------------------------------------------------------------
// GEMM: D[M,N] (+)= sum_seg A_seg[M,K] * B_seg[N,K]^T     (fp16 in, fp32 accumulate)
// ------------------------------------------------------------------------------------------
enum { EPI_STORE = 0, EPI_PROBS = 1, EPI_DX = 2 };

struct GemmParams {
    alignas(64) CUtensorMap a_map[2];   // activation side: hi, lo
    alignas(64) CUtensorMap b_map[2];   // weight side: hi, lo
    int n_seg;            // 1: (a0,b0);  3: (a0,b0), (a1,b0), (a0,b1)
    int a_mn;             // A is MN-major: its tensor map is over the row-major (K rows, M cols) array
    int b_mn;             // pair kernel, with a_mn: B is MN-major too, its map is over the row-major (K rows, N cols) array
    int M, N, K;          // K per segment, elements
    int bn;               // BLOCK_N: multiple of 16, <= 256
    int bn_last;          // pair kernel: width of the LAST N tile (multiple of 16, <= bn): N = (tiles_n - 1) * bn + (<= bn_last)
    int stages;
    int dual_n;           // pair kernel, MN-major operands: a work item = TWO neighbouring N tiles fed from one staged A tile
                          // (both TMEM accumulators live at once, no epilogue overlap: for long split-K items)
    int k_splits;         // > 1: fp32 atomics into `out` (must be zeroed)
    int epi;
    float *out;
    long long ldo;
    float out_scale;      // EPI_STORE: out = acc * out_scale
    // EPI_PROBS
    const float *bias;       // [N] (may be null)
    const float *row_scale;  // [M]
    float post_scale, clamp_lo, clamp_hi;
    int clamp;
    int n_out;               // number of (re,im) pairs that are real outputs
    float *y_out;            // optional (EPI_PROBS): Y + bias stored as fp32 (M, N) for the backward pass
    // EPI_DX: out = acc / gsc - 2 (x + add_offset) inv_n2[row] S[row]
    const float *dx_x, *dx_S;
    const unsigned int *gmax_bits;
    float add_offset;
    int dbg_pfd, dbg_nostore;   // tuning knobs (QIDDM_GEMM_PFD, QIDDM_GEMM_NOSTORE)
    int dbg_skip;               // QIDDM_GEMM_SKIP_LOADS (timing experiments)
    int l2_hints;               // bit 0: epilogue TMA stores evict_first; bit 1: weight-side (B) operand loads evict_last;
                                // bit 2: activation-side (A) operand loads evict_first; bit 3: LSU epilogue stores st.global.cs
    int out_f64;             // EPI_PROBS, QConv: `out` is a float64 tensor
    int out_P;               // EPI_PROBS, QConv: > 0 -> row = (image b, patch r), out[(b * n_out + m) * out_P + r] (NCHW)
    // EPI_PROBS fused with the diffusion step's MSE loss and dL/dY (tma_epi == 3, epilogue_tile_mse): row = (image b, level t),
    // target recomputed from the image and its noise draw (src/noise.py:105-126), no out / Y / grad_out arrays
    struct {
        const void *x;            // (n_images, P) images, float32 / float64 (f64)
        const float *eps;         // (n_images, P) noise draw
        const void *w;            // (T + 1) level weights, dtype of x
        int T, f64, P, want_lo;   // want_lo: the gradient GEMMs run the 3-term split and need the lo part of G
        float a, b, c0, c1;       // d = a out + b - (c0 level_t + c1 level_{t+1})
        double kk;                // grad_out = kk d  (= 2 a / numel)
        __half *gh, *gl;          // G = dL/dY' splits, (M, ldg) row-major
        long long ldg;
        double *loss_partial;     // [gridDim.x * 4]: per epilogue warp sum of d^2
    } mse;
    // pair kernel: epilogue through shared memory + TMA stores (fp32 boxes of 32 rows)
    int tma_epi;
    alignas(64) CUtensorMap y_map;      // EPI_PROBS: Y (M, N), box 16 x 32, SWIZZLE_64B
    alignas(64) CUtensorMap o_map;      // EPI_PROBS: out (M, n_out), box 8 x 32; EPI_DX: dX (M, N), box 16 x 32, SWIZZLE_64B
};

// fp32 -> fp16 hi + fp16 lo (22 mantissa bits; the tensor cores take fp16 subnormals, so lo needs no scaling --
// measured: 7e-6 rel-to-max on the n = 10, K = 784 layer with inputs down to 2^-17)
__device__ __forceinline__ void split_act(float v, __half &hi, __half &lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// power-of-two scale that maps the bound on max |G| into [2^13, 2^14)
__device__ __forceinline__ float g_scale_from_max(unsigned int bits) {
    const float mx = __uint_as_float(bits);
    if (!(mx > 0.f)) return 1.f;
    int e;
    frexpf(mx, &e);                 // mx = f * 2^e, f in [0.5, 1)
    return ldexpf(1.f, 14 - e);
}
// The scale every consumer of G uses.  bits[0]: the PROVISIONAL bound (sampled rows x margin, g_bound_kernel) the first
// grad_y pass worked with; bits[1]: the exact bound over all rows, a by-product of that pass.  Only if the exact bound
// exceeds the provisional one (an outlier row outside the sample) did grad_y run again with it.
__device__ __forceinline__ float g_scale_final(const unsigned int *bits) {
    const unsigned int b0 = bits[0], b1 = bits[1];
    return g_scale_from_max(b1 > b0 ? b1 : b0);
}

// Epilogue of one accumulator tile for one warp: TMEM lanes [q*32, q*32+32), columns [cbeg, cend) -> registers ->
// global.  The TMEM load of chunk i+1 is in flight while chunk i is processed.
__device__ __forceinline__ void epilogue_chunk(const GemmParams &p, const uint32_t (&r)[16], int row, int col, float rs,
                                               float dx_a, float dx_b) {
    if (row >= p.M || col >= p.N || p.dbg_nostore) return;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    if (p.epi == EPI_PROBS) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float re = v[2 * j], im = v[2 * j + 1];
            if (p.bias != nullptr && col + 2 * j + 1 < p.N) {
                re += __ldg(p.bias + col + 2 * j);
                im += __ldg(p.bias + col + 2 * j + 1);
            }
            v[2 * j] = re;
            v[2 * j + 1] = im;
            float pr = (re * re + im * im) * rs;
            if (p.clamp) pr = fminf(fmaxf(pr, p.clamp_lo), p.clamp_hi);
            o[j] = pr;
        }
        if (p.y_out != nullptr) {
            float *yd = p.y_out + (long long)row * p.N + col;
            if (col + 16 <= p.N && ((p.N & 3) == 0)) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    reinterpret_cast<float4 *>(yd)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (col + j < p.N) yd[j] = v[j];
            }
        }
        if (p.out == nullptr) return;
        const int m0 = col >> 1;
        if (p.out_P > 0) {      // QConv: consecutive rows (patches) of one image are consecutive addresses per channel
            const int b = row / p.out_P, r = row - b * p.out_P;
            const long long o0 = ((long long)b * p.n_out + m0) * p.out_P + r;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (m0 + j >= p.n_out) break;
                if (p.out_f64) reinterpret_cast<double *>(p.out)[o0 + (long long)j * p.out_P] = (double)o[j];
                else p.out[o0 + (long long)j * p.out_P] = o[j];
            }
            return;
        }
        float *dst = p.out + (long long)row * p.ldo + m0;
        if (m0 + 8 <= p.n_out && ((p.ldo & 3) == 0)) {
            reinterpret_cast<float4 *>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<float4 *>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (m0 + j < p.n_out) dst[j] = o[j];
        }
    } else if (p.epi == EPI_DX) {
        if (p.dx_x == nullptr) {       // QConv: dX stored TRANSPOSED (feature-major, ld = ldo): consecutive rows (patches) are
                                       // consecutive addresses, for this store and for fold_rows_kernel's gather
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (col + j < p.N) p.out[(long long)(col + j) * p.ldo + row] = v[j] * dx_a;
            return;
        }
        float *dst = p.out + (long long)row * p.ldo + col;
        const float *xs = p.dx_x + (long long)row * p.ldo + col;
        if (col + 16 <= p.N && ((p.ldo & 3) == 0)) {
            float4 xv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) xv[j] = __ldg(reinterpret_cast<const float4 *>(xs) + j);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                reinterpret_cast<float4 *>(dst)[j] =
                    make_float4(v[4 * j] * dx_a + dx_b * (xv[j].x + p.add_offset), v[4 * j + 1] * dx_a + dx_b * (xv[j].y + p.add_offset),
                                v[4 * j + 2] * dx_a + dx_b * (xv[j].z + p.add_offset), v[4 * j + 3] * dx_a + dx_b * (xv[j].w + p.add_offset));
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (col + j < p.N) dst[j] = v[j] * dx_a + dx_b * (__ldg(xs + j) + p.add_offset);
        }
    } else {
        float *dst = p.out + (long long)row * p.ldo + col;
        if (p.k_splits > 1) {
            if (col + 16 <= p.N && ((p.ldo & 3) == 0)) {     // 16-byte vector reductions: a quarter of the L2 atomic transactions
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(v[4 * j] * p.out_scale),
                                 "f"(v[4 * j + 1] * p.out_scale), "f"(v[4 * j + 2] * p.out_scale), "f"(v[4 * j + 3] * p.out_scale)
                                 : "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (col + j < p.N) atomicAdd(dst + j, v[j] * p.out_scale);
            }
        } else if (col + 16 <= p.N && ((p.ldo & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                reinterpret_cast<float4 *>(dst)[j] = make_float4(v[4 * j] * p.out_scale, v[4 * j + 1] * p.out_scale,
                                                                 v[4 * j + 2] * p.out_scale, v[4 * j + 3] * p.out_scale);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (col + j < p.N) dst[j] = v[j] * p.out_scale;
        }
    }
}

__device__ __forceinline__ void epilogue_tile(const GemmParams &p, uint32_t tmem_acc, int row0, int n0, int q, int lane,
                                              int cbeg, int cend) {
    const int row = row0 + q * 32 + lane;
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
    const float rs = (p.epi == EPI_PROBS && row < p.M) ? p.row_scale[row] * p.post_scale : 0.f;
    float dx_a = 0.f, dx_b = 0.f;
    if (p.epi == EPI_DX && row < p.M) {
        dx_a = 1.f / g_scale_final(p.gmax_bits);
        dx_b = p.dx_x != nullptr ? -2.f * p.row_scale[row] * p.dx_S[row] : 0.f;
    }
    uint32_t ra[16], rb[16];
    if (cbeg < cend) tc_ld16_issue(taddr + cbeg, ra);
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
        tc_ld_wait(ra);
        if (c0 + 16 < cend) tc_ld16_issue(taddr + c0 + 16, rb);
        epilogue_chunk(p, ra, row, n0 + c0, rs, dx_a, dx_b);
        if (c0 + 16 < cend) {
            tc_ld_wait(rb);
            if (c0 + 32 < cend) tc_ld16_issue(taddr + c0 + 32, ra);
            epilogue_chunk(p, rb, row, n0 + c0 + 16, rs, dx_a, dx_b);
        }
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_bytes = BM * BK * 2;
    const uint32_t b_bytes = (uint32_t)p.bn * BK * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
    // barriers: full[stages], empty[stages], tfull[2], tempty[2]; then the TMEM base slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int tiles_m = (p.M + BM - 1) / BM;
    const int tiles_n = (p.N + p.bn - 1) / p.bn;
    const int KB = (p.K + BK - 1) / BK;
    const int KT = p.n_seg * KB;
    const long long tiles_mn = (long long)tiles_m * tiles_n;
    const long long total = tiles_mn * p.k_splits;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long w = blockIdx.x; w < total; w += gridDim.x) {
                const int split = (int)(w / tiles_mn);
                const long long tile = w % tiles_mn;
                const int tm = (int)(tile / tiles_n), tn = (int)(tile % tiles_n);
                const int it0 = (int)((long long)split * KT / p.k_splits);
                const int it1 = (int)((long long)(split + 1) * KT / p.k_splits);
                for (int it = it0; it < it1; ++it) {
                    const int seg = it / KB, kb = it - seg * KB;
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), stage_bytes);
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const CUtensorMap *am = &p.a_map[seg == 1 ? 1 : 0];
                    if (p.a_mn) {   // two (64 M) x (64 K) boxes from the row-major (K, M) array
                        tma_load_2d(sa, am, full_bar(stage), tm * BM, kb * BK);
                        tma_load_2d(sa + a_bytes / 2, am, full_bar(stage), tm * BM + 64, kb * BK);
                    } else {
                        tma_load_2d(sa, am, full_bar(stage), kb * BK, tm * BM);
                    }
                    tma_load_2d(sa + a_bytes, &p.b_map[seg == 2 ? 1 : 0], full_bar(stage), kb * BK, tn * p.bn);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=f16, K-major both, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24) |
                                   (p.a_mn ? (1u << 15) : 0u);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long w = blockIdx.x; w < total; w += gridDim.x) {
                const int split = (int)(w / tiles_mn);
                const int it0 = (int)((long long)split * KT / p.k_splits);
                const int it1 = (int)((long long)(split + 1) * KT / p.k_splits);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
                for (int it = it0; it < it1; ++it) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint64_t adesc = p.a_mn ? make_smem_desc_mn(sa) : make_smem_desc(sa);
                    const uint64_t bdesc = make_smem_desc(sa + a_bytes);
                    // K advance per MMA (16 elements): K-major +32 B inside the swizzle row (+2 in 16-B units);
                    // MN-major +16 K rows = two 1024-B atoms (+128)
                    const uint64_t a_step = p.a_mn ? 128 : 2;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        tc_mma_f16(d_tmem, adesc + a_step * k, bdesc + 2 * k, idesc, (it > it0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (TMEM -> registers -> global) =====================
        const int q = warp & 3;   // TMEM lane quarter this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long w = blockIdx.x; w < total; w += gridDim.x) {
            const long long tile = w % tiles_mn;
            const int tm = (int)(tile / tiles_n), tn = (int)(tile % tiles_n);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            epilogue_tile(p, tmem_base + (uint32_t)acc * ACC_COLS, tm * BM, tn * p.bn, q, lane, 0, p.bn);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05.mma.cta_group::2, UMMA M = 256): the two CTAs of a cluster own the two
// 128-row halves of a 256 x bn tile.  Each loads its own A rows and HALF of the B tile; the tensor cores
// of the pair read both halves, so B crosses L2 -> SM once per pair.  In x3 mode one pipeline stage holds
// A hi, A lo and the three B splits of a k-block (A hi is used by two of the three products), i.e.
// 2 x 16 KB + 3 x bn/2 x 128 B per CTA for three 128 x bn x 64 products: 25 KB per product instead of
// 44 KB, which is what the L2 -> SM throughput (~42 B/cycle/SM) sustains at the tensor peak.
// Barriers: full[s] lives in the leader (rank 0) and collects both CTAs' TMA bytes; empty[s] and tfull[a]
// are signalled in both CTAs by multicast commits; tempty[a] lives in the leader and counts the epilogue
// warps of both CTAs.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap *map, uint32_t src, int c0, int c1, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_shared_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Epilogue of one accumulator tile for one warp through shared memory: each lane writes its row of a 16-column chunk
// (64-byte-swizzled) into the warp's staging buffer, one lane hands the 32-row box to the TMA (coalesced, clipped at the
// matrix edges, asynchronous).  Two buffers alternate, so a chunk is staged while the previous one drains.
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams &p, uint32_t tmem_acc, int row0, int n0, int q, int lane,
                                                  uint32_t stage0, int &buf, int bn_t) {
    const int rbase = row0 + q * 32;
    const int row = rbase + lane;
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
    const bool row_ok = row < p.M;
    const float rs = (p.epi == EPI_PROBS && row_ok) ? p.row_scale[row] * p.post_scale : 0.f;
    float dx_a = 0.f, dx_b = 0.f;
    if (p.epi == EPI_DX && row_ok) {
        dx_a = 1.f / g_scale_final(p.gmax_bits);
        dx_b = p.dx_x != nullptr ? -2.f * p.row_scale[row] * p.dx_S[row] : 0.f;
    }
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);       // SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
    auto chunk = [&](const uint32_t (&r)[16], int c0) {
        const int col = n0 + c0;
        if (col >= p.N || rbase >= p.M) return;             // warp-uniform
        const uint32_t sb = stage0 + (uint32_t)buf * EPI_STAGE_BYTES;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the buffer used two chunks ago is free
        __syncwarp();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        const uint32_t yrow = sb + (uint32_t)lane * 64u;
        if (p.epi == EPI_PROBS) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float pr = (v[2 * j] * v[2 * j] + v[2 * j + 1] * v[2 * j + 1]) * rs;
                if (p.clamp) pr = fminf(fmaxf(pr, p.clamp_lo), p.clamp_hi);
                o[j] = pr;
            }
            if (p.y_out != nullptr) {
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    st_shared_f4(yrow + ((j ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (p.out != nullptr) {
                const uint32_t orow = sb + 2048u + (uint32_t)lane * 32u;
                st_shared_f4(orow, o[0], o[1], o[2], o[3]);
                st_shared_f4(orow + 16u, o[4], o[5], o[6], o[7]);
            }
        } else if (p.dx_x == nullptr) {   // EPI_DX, QConv: feature-major dX -> stage [feature][patch] (16 x 32 floats)
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb + (j * 32u + (uint32_t)lane) * 4u), "f"(v[j] * dx_a) : "memory");
        } else {   // EPI_DX
            const float *xs = p.dx_x + (long long)row * p.ldo + col;
            float x[16];
            if (row_ok && col + 16 <= p.N) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 t = __ldg(reinterpret_cast<const float4 *>(xs) + j);
                    x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = (row_ok && col + j < p.N) ? __ldg(xs + j) : 0.f;
            }
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                st_shared_f4(yrow + ((j ^ sw) << 4), v[4 * j] * dx_a + dx_b * (x[4 * j] + p.add_offset),
                             v[4 * j + 1] * dx_a + dx_b * (x[4 * j + 1] + p.add_offset),
                             v[4 * j + 2] * dx_a + dx_b * (x[4 * j + 2] + p.add_offset),
                             v[4 * j + 3] * dx_a + dx_b * (x[4 * j + 3] + p.add_offset));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            if (p.l2_hints & 1) {       // results are not re-read by this launch: first in line for eviction
                const uint64_t pol = l2_policy_evict_first();
                if (p.epi == EPI_PROBS) {
                    if (p.y_out != nullptr) tma_store_2d_hint(&p.y_map, sb, col, rbase, pol);
                    if (p.out != nullptr) tma_store_2d_hint(&p.o_map, sb + 2048u, col >> 1, rbase, pol);
                } else if (p.dx_x == nullptr) {
                    tma_store_2d_hint(&p.o_map, sb, rbase, col, pol);
                } else {
                    tma_store_2d_hint(&p.o_map, sb, col, rbase, pol);
                }
            } else if (p.epi == EPI_PROBS) {
                if (p.y_out != nullptr) tma_store_2d(&p.y_map, sb, col, rbase);
                if (p.out != nullptr) tma_store_2d(&p.o_map, sb + 2048u, col >> 1, rbase);
            } else if (p.dx_x == nullptr) {
                tma_store_2d(&p.o_map, sb, rbase, col);     // (patch, feature) coordinates of the feature-major dX
            } else {
                tma_store_2d(&p.o_map, sb, col, rbase);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf ^= 1;
    };
    uint32_t ra[16], rb[16];
    tc_ld16_issue(taddr, ra);
    for (int c0 = 0; c0 < bn_t; c0 += 32) {
        tc_ld_wait(ra);
        if (c0 + 16 < bn_t) tc_ld16_issue(taddr + c0 + 16, rb);
        chunk(ra, c0);
        if (c0 + 16 < bn_t) {
            tc_ld_wait(rb);
            if (c0 + 32 < bn_t) tc_ld16_issue(taddr + c0 + 32, ra);
            chunk(rb, c0 + 16);
        }
    }
}

__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// EPI_PROBS epilogue through shared memory and plain global stores (tma_epi == 2).  The TMA-store epilogue shares the SM's
// TMA unit with the operand loads: ncu shows its UTMASTG issue retried and every 16-column chunk waiting ~1.4 k cycles for
// the store two chunks back to leave its staging buffer -- the forward GEMM ran at 67 % tensor activity and 25 % faster with
// the stores removed, whether or not Y was written.  Here a lane stages its row of a 32-column chunk (128 B of Y, 64 B of
// probabilities, 16-byte pieces XOR-swizzled: conflict-free both ways) and the warp writes the block back with 16-byte
// stores, 8 lanes per 128-byte row segment: full-line, no asynchronous proxy, no fences, one __syncwarp each way.
__device__ __forceinline__ void epilogue_tile_lsu(const GemmParams &p, uint32_t tmem_acc, int row0, int n0, int q, int lane,
                                                  uint32_t stage0, int bn_t) {
    const int rbase = row0 + q * 32;
    const int row = rbase + lane;
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
    const float rs = row < p.M ? p.row_scale[row] * p.post_scale : 0.f;
    const uint32_t ybase = stage0, obase = stage0 + 4096u;
    uint32_t ra[32], rb[32];
    auto chunk = [&](const uint32_t (&r)[32], int c0) {
        const int col = n0 + c0;
        if (col >= p.N || rbase >= p.M) return;             // warp-uniform
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float re = __uint_as_float(r[2 * j]), im = __uint_as_float(r[2 * j + 1]);
            float pr = (re * re + im * im) * rs;
            if (p.clamp) pr = fminf(fmaxf(pr, p.clamp_lo), p.clamp_hi);
            o[j] = pr;
        }
        __syncwarp();                                       // the previous chunk's block has been read back by every lane
        if (p.y_out != nullptr) {
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j)
                st_shared_f4(ybase + (uint32_t)lane * 128u + ((j ^ ((uint32_t)lane & 7u)) << 4), __uint_as_float(r[4 * j]),
                             __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
        if (p.out != nullptr) {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                st_shared_f4(obase + (uint32_t)lane * 64u + ((j ^ (((uint32_t)lane >> 1) & 3u)) << 4), o[4 * j], o[4 * j + 1],
                             o[4 * j + 2], o[4 * j + 3]);
        }
        __syncwarp();
        if (p.y_out != nullptr) {
            const uint32_t c = (uint32_t)lane & 7u;
            const int gcol = col + 4 * (int)c;
#pragma unroll
            for (uint32_t it = 0; it < 8; ++it) {
                const uint32_t r_ = it * 4u + ((uint32_t)lane >> 3);
                const int grow = rbase + (int)r_;
                const float4 v = ld_shared_f4(ybase + r_ * 128u + ((c ^ (r_ & 7u)) << 4));
                if (grow < p.M && gcol < p.N && c0 + 4 * (int)c < bn_t) {
                    float4 *dst = reinterpret_cast<float4 *>(p.y_out + (long long)grow * p.N + gcol);
                    if (p.l2_hints & 8) __stcs(dst, v); else *dst = v;      // streaming: Y is not read again before the backward
                }
            }
        }
        if (p.out != nullptr) {
            const uint32_t c = (uint32_t)lane & 3u;
            const int gcol = (col >> 1) + 4 * (int)c;
#pragma unroll
            for (uint32_t it = 0; it < 4; ++it) {
                const uint32_t r_ = it * 8u + ((uint32_t)lane >> 2);
                const int grow = rbase + (int)r_;
                const float4 v = ld_shared_f4(obase + r_ * 64u + ((c ^ ((r_ >> 1) & 3u)) << 4));
                if (grow < p.M && gcol < p.n_out && c0 + 8 * (int)c < bn_t) {
                    float4 *dst = reinterpret_cast<float4 *>(p.out + (long long)grow * p.ldo + gcol);
                    if (p.l2_hints & 8) __stcs(dst, v); else *dst = v;
                }
            }
        }
    };
    tc_ld32_issue(taddr, ra);
    for (int c0 = 0; c0 < bn_t; c0 += 64) {
        tc_ld_wait32(ra);
        if (c0 + 32 < bn_t) tc_ld32_issue(taddr + c0 + 32, rb);
        chunk(ra, c0);
        if (c0 + 32 < bn_t) {
            tc_ld_wait32(rb);
            if (c0 + 64 < bn_t) tc_ld32_issue(taddr + c0 + 64, ra);
            chunk(rb, c0 + 32);
        }
    }
}

// EPI_PROBS + MSE + dL/dY in the epilogue (the diffusion training step of a single amplitude-embedding layer:
// src/models.py:65-67 / :95-99 around nn/qdense.py:95-111).  Per output (row = (b, t), m):
//   out = clamp(|Y'|^2 rs),  target = c0 level_t + c1 level_{t+1},  level_k = clamp(x (1 - w_k) + eps w_k, 0, 1),
//   d = a out + b - target,  loss += d^2,  g = kk d (0 where the clamp is active),  G[2m + ri] = 2 g rs gsc Y'[2m + ri]
// and G leaves as the scaled fp16 (hi, lo) operand of the dW GEMM: what the unfused path does with the `out` store, the
// MSE pass (read out + clean, write grad) and grad_y_kernel (read Y + grad, write G) -- 13 GB of HBM traffic per step at the
// bench size -- happens on the accumulator tile in registers.  The scale gsc comes from an analytic bound (the clamp bounds
// |d|), so no pass over the gradient is needed.
template <bool F64>
__device__ __forceinline__ void epilogue_tile_mse(const GemmParams &p, uint32_t tmem_acc, int row0, int n0, int q, int lane,
                                                  uint32_t stage0, int bn_t, double &loss_acc) {
    typedef typename std::conditional<F64, double, float>::type T;
    const int rbase = row0 + q * 32;
    const int row = rbase + lane;
    const bool row_ok = row < p.M;
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
    const float rs = row_ok ? p.row_scale[row] * p.post_scale : 0.f;
    const float kq = 2.f * rs * g_scale_final(p.gmax_bits) * (float)p.mse.kk;      // G = kq d Y' on unclamped outputs
    const long long b = row_ok ? row / p.mse.T : 0;
    const int t = row_ok ? row - (int)b * p.mse.T : 0;
    const T *wl = reinterpret_cast<const T *>(p.mse.w);
    const T w0 = wl[t], w1 = wl[t + 1], omw0 = (T)1 - w0, omw1 = (T)1 - w1;
    const T ka = (T)p.mse.a, kb = (T)p.mse.b, kc0 = (T)p.mse.c0, kc1 = (T)p.mse.c1;
    const bool do_clamp = p.clamp != 0, two_levels = p.mse.c1 != 0.f;
    const T *xr = reinterpret_cast<const T *>(p.mse.x) + b * p.mse.P;
    const float *er = p.mse.eps + b * p.mse.P;
    const uint32_t hbase = stage0, lbase = stage0 + 2048u;
    const bool vec = (p.mse.P & 3) == 0 && ((((uintptr_t)p.mse.x) | ((uintptr_t)p.mse.eps)) & 15) == 0;
    uint32_t ra[32], rb[32];
    T tile_part = (T)0;       // float32 models: one float64 add per TILE (ncu: the per-chunk conversion + DADD stalled the
                              // epilogue warps on the fp64 pipe for 9 % of their time)
    // image pixels and noise draw of this lane's row for the 16 outputs of a chunk; issued one chunk ahead (L2 latency)
    auto load_xe = [&](int c0, T (&xv)[16], float (&ev)[16]) {
        const int col = n0 + c0;
        if (col >= (int)p.mse.ldg || rbase >= p.M) return;             // warp-uniform
        const int m0 = col >> 1;
        if (row_ok && vec && m0 + 16 <= p.n_out) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 e4 = __ldg(reinterpret_cast<const float4 *>(er + m0) + j);
                ev[4 * j] = e4.x; ev[4 * j + 1] = e4.y; ev[4 * j + 2] = e4.z; ev[4 * j + 3] = e4.w;
            }
            if (F64) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double2 x2 = __ldg(reinterpret_cast<const double2 *>(xr + m0) + j);
                    xv[2 * j] = (T)x2.x; xv[2 * j + 1] = (T)x2.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 x4 = __ldg(reinterpret_cast<const float4 *>(xr + m0) + j);
                    xv[4 * j] = (T)x4.x; xv[4 * j + 1] = (T)x4.y; xv[4 * j + 2] = (T)x4.z; xv[4 * j + 3] = (T)x4.w;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const bool ok = row_ok && m0 + j < p.n_out;
                xv[j] = ok ? __ldg(xr + m0 + j) : (T)0;
                ev[j] = ok ? __ldg(er + m0 + j) : 0.f;
            }
        }
    };
    auto chunk = [&](const uint32_t (&r)[32], int c0, const T (&xv)[16], const float (&ev)[16]) {
        const int col = n0 + c0;
        if (col >= (int)p.mse.ldg || rbase >= p.M) return;             // warp-uniform
        const int m0 = col >> 1;
        __half2 hh[16], ll[16];
        // float32 models: d in float32 (the difference of two float32 numbers of similar size is exact or rounds at 6e-8 of the
        // larger), the 16 squares of a chunk summed in float32 and added to the float64 accumulator once per chunk -- float64
        // arithmetic and conversions per OUTPUT made this epilogue the bottleneck of the forward GEMM.  The instruction count
        // per output is what bounds the fused forward (4 epilogue warps per CTA, one per scheduler): bounds hoisted to one
        // count per chunk, constants folded per row, packed fp16 conversions.
        T part = (T)0;
        int nlive = p.n_out - m0;                                   // outputs of this chunk that exist and lie inside the tile
        if ((bn_t - c0) >> 1 < nlive) nlive = (bn_t - c0) >> 1;
        if (!row_ok) nlive = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float re = __uint_as_float(r[2 * j]), im = __uint_as_float(r[2 * j + 1]);
            const float pr = (re * re + im * im) * rs;
            const float outv = do_clamp ? fminf(fmaxf(pr, p.clamp_lo), p.clamp_hi) : pr;
            T l0 = xv[j] * omw0 + (T)ev[j] * w0;
            l0 = l0 < (T)0 ? (T)0 : (l0 > (T)1 ? (T)1 : l0);
            T tgt = kc0 * l0;
            if (two_levels) {
                T l1 = xv[j] * omw1 + (T)ev[j] * w1;
                l1 = l1 < (T)0 ? (T)0 : (l1 > (T)1 ? (T)1 : l1);
                tgt += kc1 * l1;
            }
            const T d = j < nlive ? ka * (T)outv + kb - tgt : (T)0;
            part += d * d;
            // the clamp passes the gradient where it left the value alone (inclusive bounds, like torch.clamp)
            const float coef = outv == pr ? kq * (float)d : 0.f;
            const __half2 h2 = __floats2half2_rn(coef * re, coef * im);
            const float2 hf = __half22float2(h2);
            hh[j] = h2;
            ll[j] = __floats2half2_rn(coef * re - hf.x, coef * im - hf.y);
        }
        tile_part += part;
        __syncwarp();                                       // the previous chunk's block has been read back by every lane
        const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t off = (uint32_t)lane * 64u + ((j ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hbase + off), "r"(*reinterpret_cast<uint32_t *>(&hh[4 * j])),
                         "r"(*reinterpret_cast<uint32_t *>(&hh[4 * j + 1])), "r"(*reinterpret_cast<uint32_t *>(&hh[4 * j + 2])),
                         "r"(*reinterpret_cast<uint32_t *>(&hh[4 * j + 3])) : "memory");
            if (p.mse.want_lo)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lbase + off), "r"(*reinterpret_cast<uint32_t *>(&ll[4 * j])),
                             "r"(*reinterpret_cast<uint32_t *>(&ll[4 * j + 1])), "r"(*reinterpret_cast<uint32_t *>(&ll[4 * j + 2])),
                             "r"(*reinterpret_cast<uint32_t *>(&ll[4 * j + 3])) : "memory");
        }
        __syncwarp();
        const uint32_t c = (uint32_t)lane & 3u;
        const int gcol = col + 8 * (int)c;                  // 8 fp16 columns = 16 bytes
#pragma unroll
        for (uint32_t it = 0; it < 4; ++it) {
            const uint32_t r_ = it * 8u + ((uint32_t)lane >> 2);
            const int grow = rbase + (int)r_;
            if (grow < p.M && gcol < (int)p.mse.ldg && c0 + 8 * (int)c < bn_t) {
                const uint32_t off = r_ * 64u + ((c ^ ((r_ >> 1) & 3u)) << 4);
                const float4 vh = ld_shared_f4(hbase + off);
                *reinterpret_cast<float4 *>(p.mse.gh + (long long)grow * p.mse.ldg + gcol) = vh;
                if (p.mse.want_lo) {
                    const float4 vl = ld_shared_f4(lbase + off);
                    *reinterpret_cast<float4 *>(p.mse.gl + (long long)grow * p.mse.ldg + gcol) = vl;
                }
            }
        }
    };
    T xa[16], xb[16];
    float ea[16], eb[16];
    tc_ld32_issue(taddr, ra);
    load_xe(0, xa, ea);
    for (int c0 = 0; c0 < bn_t; c0 += 64) {
        tc_ld_wait32(ra);
        if (c0 + 32 < bn_t) { tc_ld32_issue(taddr + c0 + 32, rb); load_xe(c0 + 32, xb, eb); }
        chunk(ra, c0, xa, ea);
        if (c0 + 32 < bn_t) {
            tc_ld_wait32(rb);
            if (c0 + 64 < bn_t) { tc_ld32_issue(taddr + c0 + 64, ra); load_xe(c0 + 64, xa, ea); }
            chunk(rb, c0 + 32, xb, eb);
        }
    }
    loss_acc += (double)tile_part;
}

// The image pixels / noise draw the NEXT tile of this CTA will read in its epilogue, pulled into L2 a tile ahead: they are
// first touched here (330 MB at the bench size, 5 % of the launch's traffic) and their DRAM latency stalled the epilogue
// warps for 21 % of their time (ncu, long scoreboard at the first use).
__device__ __forceinline__ void prefetch_tile_xe(const GemmParams &p, int row0, int n0, int q, int lane, int bn_t) {
    const int row = row0 + q * 32 + lane;
    if (row >= p.M) return;
    const long long b = row / p.mse.T;
    if (lane > 0 && (row - 1) / p.mse.T == b) return;                  // one lane per image
    const int esz = p.mse.f64 ? 8 : 4;
    const char *xs = reinterpret_cast<const char *>(p.mse.x) + (b * p.mse.P + (n0 >> 1)) * esz;
    const char *es = reinterpret_cast<const char *>(p.mse.eps + b * p.mse.P + (n0 >> 1));
    int m1 = bn_t >> 1;
    if ((n0 >> 1) + m1 > p.n_out) m1 = p.n_out - (n0 >> 1);
    for (int o = 0; o < m1 * esz; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + o));
    for (int o = 0; o < m1 * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(es + o));
}

// DUAL (MN-major operands only): a work item is TWO neighbouring N tiles fed from one staged A tile -- both TMEM accumulator
// stages live at once, plain epilogue after the item (long split-K items: the dW GEMM of big batches).
// MSE (K-major operands only): 1 / 2 = the fused readout + MSE + dL/dY epilogue for float32 / float64 images (own
// instantiations: its register needs must not weigh on the plain forward kernel).
template <int NSEG, bool AMN, int BKT, bool DUAL = false, int MSE = 0>
__global__ void __launch_bounds__(PAIR_THREADS, 1) gemm_pair_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    static_assert(BKT == 64 || BKT == 32, "K-major tiles: 64 or 32 fp16 per row; MN-major: 64 or 32 K rows per box");
    constexpr uint32_t N_A = NSEG > 1 ? 2 : 1, N_B = NSEG > 1 ? 2 : 1;   // hi (and lo) tiles of each operand per stage
    constexpr uint32_t a_bytes = BM * BKT * 2;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // this CTA's half of the B tile: K-major rows of BKT elements, or (MN-major) whole 64-column boxes of 64 K rows
    const bool bmn = AMN && p.b_mn != 0;
    const uint32_t b_boxes = (uint32_t)(p.bn / 2 + 63) / 64;
    constexpr uint32_t mn_box = 64u * BKT * 2u;                       // one (64 M or N) x (BKT K rows) box
    const uint32_t b_bytes = bmn ? b_boxes * mn_box : (uint32_t)(p.bn / 2) * BKT * 2;
    static_assert(!DUAL || AMN, "dual-N items: MN-major operands");
    constexpr bool dual = DUAL;
    const uint32_t stage_bytes = N_A * a_bytes + N_B * b_bytes * (dual ? 2u : 1u);
    const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
    const uint32_t epi_base = (tmem_slot + 4u + 511u) & ~511u;      // staging of the TMA-store epilogue (tma_epi)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);      // the leader's arrive(expect_tx) for the bytes of both CTAs
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 8);    // 4 epilogue warps in each CTA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);       // 256-row tiles
    const int tiles_n_all = (p.N + p.bn - 1) / p.bn;
    const int tiles_n = dual ? (tiles_n_all + 1) / 2 : tiles_n_all;     // dual: items along N, each two tiles wide
    const int KB = (p.K + BKT - 1) / BKT;
    const long long tiles_mn = (long long)tiles_m * tiles_n;
    const long long total = tiles_mn * p.k_splits;
    const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; one elected lane issues) =====================
        const bool issuer = elect_one();
        const uint32_t fb0 = mapa_rank(full_bar(0), 0);      // the leader's full barriers (shared::cluster address)
        int stage = 0;
        uint32_t phase = 0;
        const int PFD = p.dbg_pfd;                           // L2 prefetch distance of the A operand, in k-blocks (0: off)
        const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
        for (long long w = pair; w < total; w += n_pairs) {
            const int split = (int)(w / tiles_mn);
            const long long tile = w % tiles_mn;
            const int tm = (int)(tile / tiles_n), tn = (int)((tile % tiles_n + tm) % tiles_n);   // skewed by tm (see the MMA warp)
            const int m0 = tm * 2 * BM + (int)rank * BM;
            const int bn_t = tn == tiles_n - 1 ? p.bn_last : p.bn;      // the last N tile may be narrower (no padded MMA columns)
            const int nb0 = (dual ? 2 * tn : tn) * p.bn + (int)rank * (bn_t / 2);
            const uint32_t nt = (dual && 2 * tn + 1 < tiles_n_all) ? 2u : 1u;   // N tiles of this item
            const int kb0 = (int)((long long)split * KB / p.k_splits);
            const int kb1 = (int)((long long)(split + 1) * KB / p.k_splits);
            // where this pair's next work item starts (for prefetching across the tile boundary)
            const long long wn = w + n_pairs;
            int nm0 = -1, nkb0 = 0;
            if (wn < total) {
                const int nsplit = (int)(wn / tiles_mn);
                nm0 = (int)((wn % tiles_mn) / tiles_n) * 2 * BM + (int)rank * BM;
                nkb0 = (int)((long long)nsplit * KB / p.k_splits);
            }
            if (w == pair && issuer) {
                for (int i = 0; i < PFD && kb0 + i < kb1; ++i)
                    for (uint32_t a = 0; a < N_A; ++a) {
                        if (AMN) { tma_prefetch_l2_2d(&p.a_map[a], m0, (kb0 + i) * BKT); tma_prefetch_l2_2d(&p.a_map[a], m0 + 64, (kb0 + i) * BKT); }
                        else tma_prefetch_l2_2d(&p.a_map[a], (kb0 + i) * BKT, m0);
                    }
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                if (issuer) {
                    const uint32_t fb = fb0 + 8u * stage;
                    // timing experiments only (QIDDM_GEMM_SKIP_LOADS: bit 0 = A operand, bit 1 = B operand; results are garbage)
                    const bool skip_a = (p.dbg_skip & 1) != 0, skip_b = (p.dbg_skip & 2) != 0;
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * ((skip_a ? 0u : N_A * a_bytes) + (skip_b ? 0u : nt * N_B * b_bytes)));
                    const uint32_t sa = smem_base + stage * stage_bytes;
#pragma unroll
                    for (uint32_t i = 0; i < (skip_a ? 0u : N_A); ++i) {
                        const uint32_t dst = sa + i * a_bytes;
                        if (AMN) {   // two (64 M) x (64 K) boxes from the row-major (K, M) array
                            tma_load_2d_pair(dst, &p.a_map[i], fb, m0, kb * BKT);
                            tma_load_2d_pair(dst + a_bytes / 2, &p.a_map[i], fb, m0 + 64, kb * BKT);
                        } else if (p.l2_hints & 4) {
                            tma_load_2d_pair_hint(dst, &p.a_map[i], fb, kb * BKT, m0, pol_first);
                        } else {
                            tma_load_2d_pair(dst, &p.a_map[i], fb, kb * BKT, m0);
                        }
                    }
#pragma unroll
                    for (uint32_t i = 0; i < (skip_b ? 0u : N_B); ++i) {
                        const uint32_t dst = sa + N_A * a_bytes + i * b_bytes;
                        if (bmn) {   // (64 N) x (BKT K) boxes from the row-major (K, N) array; columns past N/2 are not read
                            for (uint32_t t = 0; t < nt; ++t)
                                for (uint32_t j = 0; j < b_boxes; ++j)
                                    tma_load_2d_pair(dst + t * N_B * b_bytes + j * mn_box, &p.b_map[i], fb,
                                                     nb0 + (int)t * p.bn + 64 * (int)j, kb * BKT);
                        } else if (p.l2_hints & 2) {
                            tma_load_2d_pair_hint(dst, &p.b_map[i], fb, kb * BKT, nb0, pol_last);
                        } else {
                            tma_load_2d_pair(dst, &p.b_map[i], fb, kb * BKT, nb0);
                        }
                    }
                    // pull the A tile PFD k-blocks ahead (possibly in this pair's next tile) into L2
                    int pk = kb + PFD, pm = m0;
                    if (pk >= kb1) { pk = nkb0 + (pk - kb1); pm = nm0; }
                    if (PFD > 0 && pm >= 0 && pk < KB) {
#pragma unroll
                        for (uint32_t a = 0; a < N_A; ++a) {
                            if (AMN) { tma_prefetch_l2_2d(&p.a_map[a], pm, pk * BKT); tma_prefetch_l2_2d(&p.a_map[a], pm + 64, pk * BKT); }
                            else tma_prefetch_l2_2d(&p.a_map[a], pk * BKT, pm);
                        }
                    }
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only; the whole warp runs the loop so descriptors stay
        // warp-uniform, one elected lane issues) =====================
        if (leader) {
            const bool issuer = elect_one();
            // instruction descriptor: D=f32, A=B=f16, N>>3 at [17,23), M>>4 at [24,29) with M = 256 for the pair
            const uint32_t idesc0 = (1u << 4) | ((uint32_t)((2 * BM) >> 4) << 24) | (AMN ? (1u << 15) : 0u) | (bmn ? (1u << 16) : 0u);
            constexpr uint64_t a_step = AMN ? 128 : 2;
            const uint64_t b_step = bmn ? 128 : 2;
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long long w = pair; w < total; w += n_pairs) {
                const int split = (int)(w / tiles_mn);
                const int kb0 = (int)((long long)split * KB / p.k_splits);
                const int kb1 = (int)((long long)(split + 1) * KB / p.k_splits);
                // N tile of item w, skewed by its M tile: with a narrower last N tile the items are not equally long, and the
                // persistent round-robin (pair p takes w = p, p + n_pairs, ...) must not hand some pairs only the wide ones
                const long long tile_w = w % tiles_mn;
                const int tn_w = (int)((tile_w % tiles_n + tile_w / tiles_n) % tiles_n);
                const int bn_t = tn_w == tiles_n - 1 ? p.bn_last : p.bn;
                const bool two = dual && 2 * tn_w + 1 < tiles_n_all;          // second N tile of a dual item
                const uint32_t idesc = idesc0 | ((uint32_t)(bn_t >> 3) << 17);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint64_t ad0 = AMN ? make_smem_desc_mn(sa, mn_box) : (BKT == 64 ? make_smem_desc(sa) : make_smem_desc_k32(sa));
                    const uint64_t ad1 = AMN ? make_smem_desc_mn(sa + a_bytes, mn_box)
                                             : (BKT == 64 ? make_smem_desc(sa + a_bytes) : make_smem_desc_k32(sa + a_bytes));
                    const uint64_t bd0 = bmn ? make_smem_desc_mn(sa + N_A * a_bytes, mn_box)
                                             : (BKT == 64 ? make_smem_desc(sa + N_A * a_bytes) : make_smem_desc_k32(sa + N_A * a_bytes));
                    const uint64_t bstep = (uint64_t)(b_bytes >> 4);
                    const uint64_t btile = (uint64_t)((N_B * b_bytes) >> 4);        // dual: the second N tile's operand
                    // the last k-block of the operands only issues the 16-wide steps that hold data (the rest is zero fill)
                    const int ksteps = kb == KB - 1 ? (p.K - kb * BKT + UMMA_K - 1) / UMMA_K : BKT / UMMA_K;
                    if (issuer) {
#pragma unroll
                        for (int seg = 0; seg < NSEG; ++seg) {
                            const uint64_t adesc = seg == 1 ? ad1 : ad0;
                            const uint64_t bdesc = seg == 2 ? bd0 + bstep : bd0;
#pragma unroll
                            for (int k = 0; k < BKT / UMMA_K; ++k)
                                if (k < ksteps) {
                                    const uint32_t accum = (kb > kb0 || seg > 0 || k > 0) ? 1u : 0u;
                                    tc_mma_f16_pair(d_tmem, adesc + a_step * k, bdesc + b_step * k, idesc, accum);
                                    if (two) tc_mma_f16_pair(d_tmem + ACC_COLS, adesc + a_step * k, bdesc + btile + b_step * k, idesc, accum);
                                }
                        }
                        tc_commit_pair(empty_bar(stage));
                        if (kb == kb1 - 1) tc_commit_pair(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (dual) acc_phase ^= 1;                    // one accumulator stage holding both tiles
                else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own 128 rows) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const uint32_t te0 = mapa_rank(tempty_bar(0), 0);
        const uint32_t stage0 = epi_base + (uint32_t)q * 2u * EPI_STAGE_BYTES;
        int acc = 0, buf = 0;
        uint32_t acc_phase = 0;
        double loss_acc = 0.0;                        // tma_epi == 3: this lane's sum of d^2 over its rows of every tile
        for (long long w = pair; w < total; w += n_pairs) {
            const long long tile = w % tiles_mn;
            const int tm = (int)(tile / tiles_n), tn = (int)((tile % tiles_n + tm) % tiles_n);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int bn_t = tn == tiles_n - 1 ? p.bn_last : p.bn;
            if constexpr (DUAL) {
                const int row0 = tm * 2 * BM + (int)rank * BM;
#pragma unroll 1
                for (int t = 0; t < 2; ++t)        // not unrolled: one copy of the epilogue body
                    if (2 * tn + t < tiles_n_all)
                        epilogue_tile(p, tmem_base + (uint32_t)t * ACC_COLS, row0, (2 * tn + t) * p.bn, q, lane, 0, p.bn);
            } else if constexpr (MSE != 0) {
                if (w + n_pairs < total) {
                    const long long tile2 = (w + n_pairs) % tiles_mn;
                    const int tm2 = (int)(tile2 / tiles_n), tn2 = (int)((tile2 % tiles_n + tm2) % tiles_n);
                    prefetch_tile_xe(p, tm2 * 2 * BM + (int)rank * BM, tn2 * p.bn, q, lane, tn2 == tiles_n - 1 ? p.bn_last : p.bn);
                }
                epilogue_tile_mse<MSE == 2>(p, tmem_base + (uint32_t)acc * ACC_COLS, tm * 2 * BM + (int)rank * BM, tn * p.bn, q, lane, stage0, bn_t, loss_acc);
            } else if (p.tma_epi == 2)
                epilogue_tile_lsu(p, tmem_base + (uint32_t)acc * ACC_COLS, tm * 2 * BM + (int)rank * BM, tn * p.bn, q, lane, stage0, bn_t);
            else if (p.tma_epi)
                epilogue_tile_tma(p, tmem_base + (uint32_t)acc * ACC_COLS, tm * 2 * BM + (int)rank * BM, tn * p.bn, q, lane, stage0, buf, bn_t);
            else
                epilogue_tile(p, tmem_base + (uint32_t)acc * ACC_COLS, tm * 2 * BM + (int)rank * BM, tn * p.bn, q, lane, 0, bn_t);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(te0 + 8u * acc);
            if (dual) acc_phase ^= 1;
            else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.tma_epi == 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if constexpr (MSE != 0) {  // one slot per epilogue warp of the persistent grid: summed in a fixed order afterwards
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
            if (lane == 0) p.mse.loss_partial[(long long)blockIdx.x * 4 + q] = loss_acc;
        }
    }

    tc_fence_before();
    cluster_sync_all();       // the peer's shared memory and barriers stay valid until both CTAs are done
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------

// x (B,F) fp32 -> Xh/Xl (B,Kp) fp16 (hi, lo*2^11) and inv_n2[b] = 1 / (sum f^2 + n_pad * pad^2).  One warp per row.
// Column F (when the state has constant pad rows) is a column of ones: it meets zero weights in the forward
// and dX GEMMs and yields sum_b G[b,n] (the pad rows of dW) in the dW GEMM.
__device__ __forceinline__ float ld_io(const void *p, long long i, int io64) {
    return io64 ? (float)__ldg(reinterpret_cast<const double *>(p) + i) : __ldg(reinterpret_cast<const float *>(p) + i);
}

struct UnfoldGeom {      // fused patch-unfold (QConv, nn/qconv.py:76-77): row = (image, y, x), feature = (ch, ky, kx)
    int on, C, H, W, kh, kw, ph, pw, Hout, Wout;
};
__global__ void prep_x_kernel(const float *x, long long B, int F, int Kp, int n_pad, float add_offset, float pad,
                              __half *Xh, __half *Xl, float *inv_n2, int want_lo, const UnfoldGeom u, int gw, int io64) {
    // QConv: per-CTA feature table (offset inside the image relative to the patch origin, packed (dy, dx) for the
    // bounds test), so the per-element work is one table read instead of three integer divisions
    extern __shared__ int ftab[];          // [2 * F] when u.on
    if (u.on) {
        const int kk = u.kh * u.kw;
        for (int c = threadIdx.x; c < F; c += blockDim.x) {
            const int ch = c / kk, q = c - ch * kk;
            const int ky = q / u.kw, kx = q - ky * u.kw;
            ftab[c] = (ch * u.H + (ky - u.ph)) * u.W + (kx - u.pw);
            ftab[F + c] = ((ky - u.ph) << 16) | ((kx - u.pw) & 0xffff);
        }
        __syncthreads();
    }
    // a group of `gw` lanes (power of two <= 32) per row, 32 / gw rows per warp: short QConv rows keep the lanes busy
    const int lane = threadIdx.x & 31, sub = lane & (gw - 1), rpw = 32 / gw;
    const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw + lane / gw;
    const bool active = row < B;
    long long base = row * F;
    int py = 0, px = 0;
    if (u.on && active) {
        const int P = u.Hout * u.Wout;
        const long long b = row / P;
        const int r = (int)(row - b * P);
        py = r / u.Wout;
        px = r - py * u.Wout;
        base = b * u.C * u.H * u.W + (long long)py * u.W + px;
    }
    float ss = 0.f;
    // two consecutive features per lane: one 4-byte store per array
    for (int c = 2 * sub; c < Kp && active; c += 2 * gw) {
        float f[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int cc = c + j;
            if (cc < F) {
                if (u.on) {
                    const int yx = ftab[F + cc];
                    const int iy = py + (yx >> 16), ix = px + (int)(short)(yx & 0xffff);
                    const bool in = iy >= 0 && iy < u.H && ix >= 0 && ix < u.W;       // zero padding of torch.nn.Unfold
                    f[j] = (in ? ld_io(x, base + ftab[cc], io64) : 0.f) + add_offset;
                } else {
                    f[j] = __ldg(x + base + cc) + add_offset;
                }
            }
        }
        __half h0, l0, h1, l1;
        split_act((c == F && n_pad > 0) ? 1.f : f[0], h0, l0);
        split_act((c + 1 == F && n_pad > 0) ? 1.f : f[1], h1, l1);
        *reinterpret_cast<__half2 *>(Xh + row * Kp + c) = __halves2half2(h0, h1);       // Kp is a multiple of 8
        if (want_lo) *reinterpret_cast<__half2 *>(Xl + row * Kp + c) = __halves2half2(l0, l1);
        ss += f[0] * f[0] + f[1] * f[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        if (o < gw) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (sub == 0 && active) {
        ss += (float)n_pad * pad * pad;
        inv_n2[row] = ss > 0.f ? 1.0f / ss : 0.f;
    }
}

// Dense rows with 16-byte aligned fp32 features (F % 4 == 0): the same outputs as prep_x_kernel, one warp per row, a lane
// takes 8 consecutive features per step (two 16-byte loads, one 16-byte store per split) - a pure streaming pass.
__global__ void __launch_bounds__(256) prep_x_dense_kernel(const float *x, long long B, int F, int Kp, int n_pad, float add_offset,
                                                           float pad, __half *Xh, __half *Xl, float *inv_n2, int want_lo) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = warp0; row < B; row += nwarps) {
        const float *src = x + row * F;
        float ss = 0.f;
        for (int c = 8 * lane; c < Kp; c += 256) {
            float f[8];
            if (c + 8 <= F) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(src + c)), b = __ldg(reinterpret_cast<const float4 *>(src + c) + 1);
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += add_offset;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = c + j < F ? __ldg(src + c + j) + add_offset : 0.f;
            }
            __half hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ss += f[j] * f[j];
                split_act((c + j == F && n_pad > 0) ? 1.f : f[j], hi[j], lo[j]);
            }
            *reinterpret_cast<uint4 *>(Xh + row * Kp + c) = *reinterpret_cast<const uint4 *>(hi);       // Kp is a multiple of 8
            if (want_lo) *reinterpret_cast<uint4 *>(Xl + row * Kp + c) = *reinterpret_cast<const uint4 *>(lo);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) {
            ss += (float)n_pad * pad * pad;
            inv_n2[row] = ss > 0.f ? 1.0f / ss : 0.f;
        }
    }
}

// Training forward, dense rows: operand splits and their transposes in one pass (x is read once): a CTA owns 64 rows and
// walks the column tiles; it writes the row-major splits X (B,Kp), stages each 64 x 64 tile in shared memory for the
// transposed splits XT (Kp,Bp), and keeps the row sums of squares in registers (deterministic: 8 lanes per row).
__global__ void __launch_bounds__(256) prep_xt_kernel(const float *x, long long B, int F, int Kp, long long Bp, int n_pad,
                                                      float add_offset, float pad, __half *Xh, __half *Xl, __half *XTh,
                                                      __half *XTl, float *inv_n2) {
    __shared__ __half th[64][66], tl[64][66];      // 33-word rows: the transposed 2-byte reads spread over the banks
    const long long r0 = (long long)blockIdx.x * 64;
    const int t = threadIdx.x;
    const bool vec = (F & 3) == 0 && ((uintptr_t)x & 15) == 0;
    float ss[2] = {0.f, 0.f};
    for (int c0 = 0; c0 < Kp; c0 += 64) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = t + i * 256;            // row = idx / 8, 8-column piece = idx % 8
            const int rr = idx >> 3, pc = idx & 7;
            const long long r = r0 + rr;
            const int c = c0 + pc * 8;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
            if (r < B && c < F) {
                const float *src = x + r * F + c;
                if (vec && c + 8 <= F) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(src)), b = __ldg(reinterpret_cast<const float4 *>(src) + 1);
                    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += add_offset;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c + j < F) f[j] = __ldg(src + j) + add_offset;
                }
            }
            __half hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ss[i] += f[j] * f[j];
                const bool one = (c + j == F) && n_pad > 0 && r < B;      // the ones column (pad-row column sums)
                split_act(one ? 1.f : f[j], hi[j], lo[j]);
            }
            if (r < B && c < Kp) {
                *reinterpret_cast<uint4 *>(Xh + r * Kp + c) = *reinterpret_cast<const uint4 *>(hi);
                *reinterpret_cast<uint4 *>(Xl + r * Kp + c) = *reinterpret_cast<const uint4 *>(lo);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<__half2 *>(&th[rr][pc * 8 + 2 * j]) = __halves2half2(hi[2 * j], hi[2 * j + 1]);
                *reinterpret_cast<__half2 *>(&tl[rr][pc * 8 + 2 * j]) = __halves2half2(lo[2 * j], lo[2 * j + 1]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = t + i * 256;            // output row (feature) = idx / 8, piece of 8 batch rows = idx % 8
            const int cc = idx >> 3, pr = idx & 7;
            const int c = c0 + cc;
            const long long r = r0 + pr * 8;
            if (c < Kp && r < Bp) {
                __half oh[8], ol[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    oh[j] = th[pr * 8 + j][cc];
                    ol[j] = tl[pr * 8 + j][cc];
                }
                *reinterpret_cast<uint4 *>(XTh + (long long)c * Bp + r) = *reinterpret_cast<const uint4 *>(oh);
                *reinterpret_cast<uint4 *>(XTl + (long long)c * Bp + r) = *reinterpret_cast<const uint4 *>(ol);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float v = ss[i];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        const long long r = r0 + ((t + i * 256) >> 3);
        if ((t & 7) == 0 && r < B) {
            v += (float)n_pad * pad * pad;
            inv_n2[r] = v > 0.f ? 1.0f / v : 0.f;
        }
    }
}

// Training forward, QConv: the same outputs straight from the NCHW image (patch-unfold fused).  A CTA owns 64 consecutive
// patches; a thread keeps ONE patch and walks the features, so for a given feature the 32 lanes of a warp read 32
// neighbouring pixels (coalesced).  The 64 x 64 (feature, patch) tile goes through shared memory once: XT rows are
// written as 128-byte runs of patches, X rows as 16-byte pieces of features.
__global__ void __launch_bounds__(256) prep_xt_unfold_kernel(const void *img, int io64, long long B, int F, int Kp, long long Bp,
                                                             int n_pad, float add_offset, float pad, __half *Xh, __half *Xl,
                                                             __half *XTh, __half *XTl, float *inv_n2, const UnfoldGeom u) {
    __shared__ __half th[64][66], tl[64][66];     // [feature of the tile][patch]; 33-word rows: conflict-free both ways
    __shared__ float ssum[4][64];
    extern __shared__ int ftab[];                 // [2 * F] feature table (see prep_x_kernel)
    const long long r0 = (long long)blockIdx.x * 64;
    const int t = threadIdx.x, rr = t & 63, cg = t >> 6;
    const long long r = r0 + rr;
    const bool rowok = r < B;
    long long base = 0;
    int py = 0, px = 0;
    if (rowok) {
        const int P = u.Hout * u.Wout;
        const long long b = r / P;
        const int q = (int)(r - b * P);
        py = q / u.Wout;
        px = q - py * u.Wout;
        base = b * u.C * u.H * u.W + (long long)py * u.W + px;
    }
    {
        const int kk = u.kh * u.kw;
        for (int c = t; c < F; c += 256) {
            const int ch = c / kk, q = c - ch * kk;
            const int ky = q / u.kw, kx = q - ky * u.kw;
            ftab[c] = (ch * u.H + (ky - u.ph)) * u.W + (kx - u.pw);
            ftab[F + c] = ((ky - u.ph) << 16) | ((kx - u.pw) & 0xffff);
        }
    }
    __syncthreads();
    float ss = 0.f;
    for (int c0 = 0; c0 < Kp; c0 += 64) {
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const int cc = cg + 4 * k, c = c0 + cc;
            float f = 0.f;
            if (rowok && c < F) {
                const int yx = ftab[F + c];
                const int iy = py + (yx >> 16), ix = px + (int)(short)(yx & 0xffff);
                const bool in = iy >= 0 && iy < u.H && ix >= 0 && ix < u.W;       // zero padding of torch.nn.Unfold
                f = (in ? ld_io(img, base + ftab[c], io64) : 0.f) + add_offset;
            }
            ss += f * f;
            __half hi, lo;
            split_act((c == F && n_pad > 0 && rowok) ? 1.f : f, hi, lo);
            th[cc][rr] = hi;
            tl[cc][rr] = lo;
        }
        __syncthreads();
        if (XTh != nullptr) {   // XT: feature rows of 64 patches = 32 half2 per row
            const int lane = t & 31, w = t >> 5;
            const long long r2 = r0 + 2 * lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int cc = w + 8 * k, c = c0 + cc;
                if (c < Kp && r2 < Bp) {
                    *reinterpret_cast<__half2 *>(XTh + (long long)c * Bp + r2) = *reinterpret_cast<const __half2 *>(&th[cc][2 * lane]);
                    *reinterpret_cast<__half2 *>(XTl + (long long)c * Bp + r2) = *reinterpret_cast<const __half2 *>(&tl[cc][2 * lane]);
                }
            }
        }
        {   // X: patch rows; 4 threads per row, two 8-feature pieces each
            const int row = t >> 2, q = t & 3;
            const long long rx = r0 + row;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cc0 = q * 16 + h * 8, c = c0 + cc0;
                if (rx < B && c < Kp) {
                    __half oh[8], ol[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        oh[j] = th[cc0 + j][row];
                        ol[j] = tl[cc0 + j][row];
                    }
                    *reinterpret_cast<uint4 *>(Xh + rx * Kp + c) = *reinterpret_cast<const uint4 *>(oh);
                    *reinterpret_cast<uint4 *>(Xl + rx * Kp + c) = *reinterpret_cast<const uint4 *>(ol);
                }
            }
        }
        __syncthreads();
    }
    ssum[cg][rr] = ss;
    __syncthreads();
    if (t < 64 && r0 + t < B) {
        const float v = ssum[0][t] + ssum[1][t] + ssum[2][t] + ssum[3][t] + (float)n_pad * pad * pad;
        inv_n2[r0 + t] = v > 0.f ? 1.0f / v : 0.f;
    }
}

// QConv col2im of dX (stored feature-major: dxT[f][patch], ld = ldt) with the normalisation term of the amplitude
// embedding, gather form (one thread per image element, no atomics; neighbouring pixels read neighbouring patches):
//   grad_img[b,ch,iy,ix] = sum_{(ky,kx): patch p=(b, iy-ky+ph, ix-kx+pw) exists} dxT[(ch,ky,kx)][p] - 2 (img + add_offset) inv_n2[p] S[p]
template <int KH, int KW>      // > 0: compile-time window (fully unrolled gather, loads in flight together); 0: runtime
__global__ void __launch_bounds__(256) fold_rows_kernel(const float *dxT, long long ldt, const void *img, int io64,
                                                        const float *inv_n2, const float *S, long long n_elems, float add_offset,
                                                        const UnfoldGeom u, void *grad_img) {
    const int kh = KH > 0 ? KH : u.kh, kw = KW > 0 ? KW : u.kw;
    const int P = u.Hout * u.Wout, kk = kh * kw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % u.W);
        long long t = i / u.W;
        const int iy = (int)(t % u.H);
        t /= u.H;
        const int ch = (int)(t % u.C);
        const long long b = t / u.C;
        const float f2 = -2.f * (ld_io(img, i, io64) + add_offset);
        const float *dch = dxT + (long long)ch * kk * ldt + b * P;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < kh; ++ky) {
            const int y = iy - ky + u.ph;
#pragma unroll
            for (int kx = 0; kx < kw; ++kx) {
                const int xx = ix - kx + u.pw;
                const bool ok = y >= 0 && y < u.Hout && xx >= 0 && xx < u.Wout;
                const int pr = ok ? y * u.Wout + xx : 0;
                const float d = __ldg(dch + (long long)(ky * kw + kx) * ldt + pr);
                const float ns = __ldg(inv_n2 + b * P + pr) * __ldg(S + b * P + pr);
                if (ok) acc += d + f2 * ns;
            }
        }
        if (io64) reinterpret_cast<double *>(grad_img)[i] = (double)acc;
        else reinterpret_cast<float *>(grad_img)[i] = acc;
    }
}

// From UT (row c = U|c>, complex fp32) build the weight-side GEMM operands (scaled by w_scale):
//   Wn[n][c] (N x Kp), Wt[c][n] (F x Np), n = 2m + {re,im}, value = part(UT[c][m*stride]);
//   bias[n] = pad * sum_{c >= F} value.
__global__ void __launch_bounds__(256) build_w_kernel(const float2 *UT, int A, int F, int Kp, int N, int Np, int stride,
                                                      float w_scale, float pad, __half *Wn_h, __half *Wn_l, __half *Wt_h,
                                                      __half *Wt_l, float *bias) {
    // one CTA per output column n: the pad-row sum is reduced in a fixed order (warp shuffles, then the 8 warp sums in
    // sequence), so two collapses of the same weights give bit-identical operands (no float atomics)
    __shared__ float wsum[8];
    const int n = blockIdx.x;
    const int m = n >> 1, ri = n & 1;
    float bsum = 0.f;
    for (int c = threadIdx.x; c < max(A, Kp); c += blockDim.x) {
        float v = 0.f;
        if (c < A) {
            const float2 u = UT[(long long)c * A + (long long)m * stride];
            v = (ri ? u.y : u.x) * w_scale;
        }
        if (c < Kp) {
            __half hi, lo;
            split_act(c < F ? v : 0.f, hi, lo);
            Wn_h[(long long)n * Kp + c] = hi;
            Wn_l[(long long)n * Kp + c] = lo;
            if (c < F) {
                Wt_h[(long long)c * Np + n] = hi;
                Wt_l[(long long)c * Np + n] = lo;
            }
        }
        if (c >= F && c < A) bsum += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = bsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wsum[w];
        bias[n] = t * pad;
    }
}

// The constant pad rows of the state contribute bias[n] = pad * sum_{c >= F} W'[n][c] to every row of Y.  X carries a
// column of ones at index F, so the bias becomes that column's weight: no bias add in the GEMM epilogue.
__global__ void fold_bias_kernel(const float *bias, int N, int Kp, int F, __half *Wn_h, __half *Wn_l) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    __half hi, lo;
    split_act(bias[n], hi, lo);
    Wn_h[(long long)n * Kp + F] = hi;
    Wn_l[(long long)n * Kp + F] = lo;
}

// Upper bound of max |G| (for the fp16 range) without touching Y: |Y'| <= w_scale * |f| (U is unitary), so
// |G[b,n]| = 2 |g| scale inv_n2 |Y'| <= 2 max_m|g[b,m]| * scale * sqrt(inv_n2[b]) * w_scale.
// Only every `row_stride`-th row is read (B = number of SAMPLED rows) and the result is multiplied by `margin`: a provisional
// bound -- fp16 hi/lo operands keep 22 bits of anything within 2^16 of the largest element, so a scale that is a few binades
// conservative costs nothing, and grad_y_kernel checks it against the exact bound as it streams every row anyway.
__global__ void __launch_bounds__(256) g_bound_kernel(const float *go, const float *inv_n2, long long B, int n_out,
                                                      float scale, float w_scale, unsigned int *gmax_bits, int go_P, int gw, int go_f64,
                                                      long long row_stride, float margin) {
    // a group of `gw` lanes (power of two <= 32) per row, 32 / gw rows per warp (small QConv rows keep the lanes busy)
    const int lane = threadIdx.x & 31, sub = lane & (gw - 1), rpw = 32 / gw;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const bool vec = gw == 32 && go_P == 0 && (n_out & 3) == 0 && ((uintptr_t)go & 15) == 0;
    float best = 0.f;
    for (long long rbase = warp0 * rpw; rbase < B; rbase += nwarps * rpw) {
        const long long srow = rbase + lane / gw;
        const bool active = srow < B;
        const long long row = srow * row_stride;
        float mx = 0.f;
        if (!active) {
        } else if (go_P > 0) {        // QConv: grad_out is NCHW, element (row, m) at ((b * n_out + m) * P + r)
            const long long b = row / go_P, r = row - b * go_P;
            for (int m = sub; m < n_out; m += gw) mx = fmaxf(mx, fabsf(ld_io(go, (b * n_out + m) * go_P + r, go_f64)));
        } else if (vec) {
            const float *g = go + row * n_out;
            for (int i = lane; i < (n_out >> 2); i += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i);
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
            }
        } else {
            const float *g = go + row * n_out;
            for (int m = sub; m < n_out; m += gw) mx = fmaxf(mx, fabsf(__ldg(g + m)));
        }
        if (active) best = fmaxf(best, 2.f * mx * scale * sqrtf(inv_n2[row]) * w_scale);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    best *= margin;
    if (lane == 0 && best > 0.f && best < 3.0e38f) atomicMax(gmax_bits, __float_as_uint(best));
}

// One streaming pass over Y (= X W' + bias', saved by the forward GEMM) and grad_out, one warp per row:
//   out = scale inv_n2 |Y|^2, mask = !clamp || lo <= out <= hi,  G[2m+ri] = 2 g mask scale inv_n2 Y[2m+ri]
// written as scaled fp16 (hi, lo) row-major (B,Np); S[b] = sum_m g mask out (normalisation term of dX).
// Scale of G: pass 0 (retry == 0) works with the provisional bound gmax_bits[0] and, reading every row of grad_out anyway,
// leaves the exact bound (same formula as g_bound_kernel) in gmax_bits[1]; pass 1 (retry == 1, launched only when the
// provisional bound came from a sample of the rows) returns at once unless the exact bound turned out larger, and then
// redoes the pass with it.  Consumers use g_scale_final().
__global__ void __launch_bounds__(256) grad_y_kernel(const float *Y, const float *go, const float *inv_n2, long long B,
                                                     int N, int Np, int n_out, float scale, int clamp, float lo,
                                                     float hi, unsigned int *gmax_bits, __half *Gh, __half *Gl,
                                                     float *S, int want_lo, int go_P, int gw, int go_f64, float w_scale,
                                                     int retry) {
    // a group of `gw` lanes (power of two <= 32) per row, 32 / gw rows per warp
    const int lane = threadIdx.x & 31, sub = lane & (gw - 1), rpw = 32 / gw;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    if (retry && gmax_bits[1] <= gmax_bits[0]) return;
    const float gsc = g_scale_from_max(gmax_bits[retry ? 1 : 0]);
    float bound = 0.f;      // pass 0: this lane's share of max_b 2 max_m |g[b,m]| scale sqrt(inv_n2[b]) w_scale
    // 4 outputs (8 columns of Y / G) per lane and iteration when everything is 16-byte aligned
    const bool vec = gw == 32 && go_P == 0 && (n_out & 3) == 0 && Np == N &&
                     (((uintptr_t)Y | (uintptr_t)go | (uintptr_t)Gh | (uintptr_t)Gl) & 15) == 0;
    for (long long rbase = warp0 * rpw; rbase < B; rbase += nwarps * rpw) {
        const long long r = rbase + lane / gw;
        const bool active = r < B;
        const float in2 = active ? inv_n2[r] : 0.f;
        const float k0 = scale * in2, k1 = 2.f * scale * in2 * gsc;
        float s_part = 0.f, mx = 0.f;
        if (vec) {
            const float4 *y4 = reinterpret_cast<const float4 *>(Y + r * N);
            const float4 *g4 = reinterpret_cast<const float4 *>(go + r * n_out);
            uint4 *gh4 = reinterpret_cast<uint4 *>(Gh + r * Np), *gl4 = reinterpret_cast<uint4 *>(Gl + r * Np);
            for (int i = lane; i < (n_out >> 2); i += 32) {
                const float4 ya = __ldg(y4 + 2 * i), yb = __ldg(y4 + 2 * i + 1), gg = __ldg(g4 + i);
                const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                const float gv[4] = {gg.x, gg.y, gg.z, gg.w};
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(gg.x), fabsf(gg.y))), fmaxf(fabsf(gg.z), fabsf(gg.w)));
                __half2 hh[4], ll[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float re = yv[2 * j], im = yv[2 * j + 1];
                    const float outv = k0 * (re * re + im * im);
                    const bool pass = !clamp || (outv >= lo && outv <= hi);
                    const float g = pass ? gv[j] : 0.f;
                    const float coef = k1 * g;
                    s_part += g * outv;
                    __half h0, l0, h1, l1;
                    split_act(coef * re, h0, l0);
                    split_act(coef * im, h1, l1);
                    hh[j] = __halves2half2(h0, h1);
                    ll[j] = __halves2half2(l0, l1);
                }
                gh4[i] = make_uint4(*reinterpret_cast<unsigned int *>(&hh[0]), *reinterpret_cast<unsigned int *>(&hh[1]),
                                    *reinterpret_cast<unsigned int *>(&hh[2]), *reinterpret_cast<unsigned int *>(&hh[3]));
                if (want_lo)
                    gl4[i] = make_uint4(*reinterpret_cast<unsigned int *>(&ll[0]), *reinterpret_cast<unsigned int *>(&ll[1]),
                                        *reinterpret_cast<unsigned int *>(&ll[2]), *reinterpret_cast<unsigned int *>(&ll[3]));
            }
        } else if (active) {
            for (int m = sub; 2 * m < Np; m += gw) {
                float gre = 0.f, gim = 0.f;
                if (m < n_out) {
                    const float2 y = *reinterpret_cast<const float2 *>(Y + r * N + 2 * m);
                    const float outv = k0 * (y.x * y.x + y.y * y.y);
                    const bool pass = !clamp || (outv >= lo && outv <= hi);
                    const long long gi = go_P > 0 ? ((r / go_P) * n_out + m) * go_P + (r % go_P) : r * n_out + m;
                    const float graw = go_P > 0 ? ld_io(go, gi, go_f64) : __ldg(go + gi);
                    mx = fmaxf(mx, fabsf(graw));
                    const float g = pass ? graw : 0.f;
                    gre = k1 * g * y.x;
                    gim = k1 * g * y.y;
                    s_part += g * outv;
                }
                __half h0, l0, h1, l1;
                split_act(gre, h0, l0);
                split_act(gim, h1, l1);
                *reinterpret_cast<__half2 *>(Gh + r * Np + 2 * m) = __halves2half2(h0, h1);
                if (want_lo) *reinterpret_cast<__half2 *>(Gl + r * Np + 2 * m) = __halves2half2(l0, l1);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            if (o < gw) s_part += __shfl_xor_sync(0xffffffffu, s_part, o);
        if (sub == 0 && active) S[r] = s_part;
        if (active) bound = fmaxf(bound, 2.f * mx * scale * sqrtf(in2) * w_scale);   // lanes of a row group hold partial maxima
    }
    if (!retry) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, o));
        if (lane == 0 && bound > 0.f && bound < 3.0e38f) atomicMax(gmax_bits + 1, __float_as_uint(bound));
    }
}

// Fused diffusion step, operand side: the noisy rows (level t + 1 of image b's ladder, src/noise.py:105-126 as sliced by
// src/models.py:46-63) are never written as fp32 -- a warp forms row (b, t) from the image and its noise draw and stores
// the fp16 (hi, lo) splits and 1 / |f|^2 the forward GEMM consumes (what qiddm_noise_ladder + prep_x_dense_kernel do in two
// passes through HBM).  in2max_bits: max over the rows of 1 / |f|^2 (for the analytic bound on |G|).
// One warp per IMAGE: its pixels and noise draw stay in registers (ITS x 8 values per lane) while the T rows are formed
// and stored -- the images are read once (an image's rows re-reading them through L2 made the pass L2-bound).
template <typename T, int ITS>
__global__ void __launch_bounds__(256) ladder_prep_kernel(const T *x, const float *eps, const T *w, long long n_img, int steps,
                                                          int F, int Kp, int n_pad, float add_offset, float pad, __half *Xh,
                                                          __half *Xl, float *inv_n2, int want_lo, unsigned int *in2max_bits) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const bool vec = (F & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)eps)) & 15) == 0;
    float best = 0.f;
    for (long long b = warp0; b < n_img; b += nwarps) {
        const T *xs = x + b * F;
        const float *es = eps + b * F;
        T xv[ITS][8];
        float ev[ITS][8];
#pragma unroll
        for (int it = 0; it < ITS; ++it) {
            const int c = 8 * lane + 256 * it;
            if (vec && c + 8 <= F) {
                const float4 e0 = __ldg(reinterpret_cast<const float4 *>(es + c)), e1 = __ldg(reinterpret_cast<const float4 *>(es + c) + 1);
                ev[it][0] = e0.x; ev[it][1] = e0.y; ev[it][2] = e0.z; ev[it][3] = e0.w;
                ev[it][4] = e1.x; ev[it][5] = e1.y; ev[it][6] = e1.z; ev[it][7] = e1.w;
                if (sizeof(T) == 8) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 t2 = __ldg(reinterpret_cast<const double2 *>(xs + c) + j);
                        xv[it][2 * j] = (T)t2.x; xv[it][2 * j + 1] = (T)t2.y;
                    }
                } else {
                    const float4 a0 = __ldg(reinterpret_cast<const float4 *>(xs + c)), a1 = __ldg(reinterpret_cast<const float4 *>(xs + c) + 1);
                    xv[it][0] = (T)a0.x; xv[it][1] = (T)a0.y; xv[it][2] = (T)a0.z; xv[it][3] = (T)a0.w;
                    xv[it][4] = (T)a1.x; xv[it][5] = (T)a1.y; xv[it][6] = (T)a1.z; xv[it][7] = (T)a1.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xv[it][j] = c + j < F ? __ldg(xs + c + j) : (T)0;
                    ev[it][j] = c + j < F ? __ldg(es + c + j) : 0.f;
                }
            }
        }
        for (int t = 0; t < steps; ++t) {
            const T wt = w[t + 1], omw = (T)1 - wt;
            const long long row = b * steps + t;
            float ss = 0.f;
#pragma unroll
            for (int it = 0; it < ITS; ++it) {
                const int c = 8 * lane + 256 * it;
                if (c < Kp) {
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        f[j] = 0.f;
                        if (c + j < F) {
                            T v = xv[it][j] * omw + (T)ev[it][j] * wt;
                            v = v < (T)0 ? (T)0 : (v > (T)1 ? (T)1 : v);
                            f[j] = (float)v + add_offset;
                        }
                        ss += f[j] * f[j];
                    }
                    if (n_pad > 0 && c <= F && F < c + 8) f[F - c] = 1.f;       // the ones column (one lane of one chunk per row)
                    __half2 hi[4], lo[4];                                        // packed conversions: two features per instruction
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        hi[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                        const float2 hf = __half22float2(hi[j]);
                        lo[j] = __floats2half2_rn(f[2 * j] - hf.x, f[2 * j + 1] - hf.y);
                    }
                    *reinterpret_cast<uint4 *>(Xh + row * Kp + c) = *reinterpret_cast<const uint4 *>(hi);       // Kp is a multiple of 8
                    if (want_lo) *reinterpret_cast<uint4 *>(Xl + row * Kp + c) = *reinterpret_cast<const uint4 *>(lo);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            ss += (float)n_pad * pad * pad;
            const float in2 = ss > 0.f ? 1.0f / ss : 0.f;
            if (lane == 0) inv_n2[row] = in2;
            best = fmaxf(best, in2);
        }
    }
    if (lane == 0 && best > 0.f && best < 3.0e38f) atomicMax(in2max_bits, __float_as_uint(best));
}

// Wide rows (Kp > 1280: n = 11, 12): one warp per ROW, the image re-read per level through L1 / L2 (no register staging).
template <typename T>
__global__ void __launch_bounds__(256) ladder_prep_rows_kernel(const T *x, const float *eps, const T *w, long long n_img, int steps,
                                                               int F, int Kp, int n_pad, float add_offset, float pad, __half *Xh,
                                                               __half *Xl, float *inv_n2, int want_lo, unsigned int *in2max_bits) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long B = n_img * steps;
    float best = 0.f;
    for (long long row = warp0; row < B; row += nwarps) {
        const long long b = row / steps;
        const T wt = w[(int)(row - b * steps) + 1], omw = (T)1 - wt;
        const T *xs = x + b * F;
        const float *es = eps + b * F;
        float ss = 0.f;
        for (int c = 8 * lane; c < Kp; c += 256) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                f[j] = 0.f;
                if (c + j < F) {
                    T v = __ldg(xs + c + j) * omw + (T)__ldg(es + c + j) * wt;
                    v = v < (T)0 ? (T)0 : (v > (T)1 ? (T)1 : v);
                    f[j] = (float)v + add_offset;
                }
                ss += f[j] * f[j];
            }
            if (n_pad > 0 && c <= F && F < c + 8) f[F - c] = 1.f;
            __half2 hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hi[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                const float2 hf = __half22float2(hi[j]);
                lo[j] = __floats2half2_rn(f[2 * j] - hf.x, f[2 * j + 1] - hf.y);
            }
            *reinterpret_cast<uint4 *>(Xh + row * Kp + c) = *reinterpret_cast<const uint4 *>(hi);
            if (want_lo) *reinterpret_cast<uint4 *>(Xl + row * Kp + c) = *reinterpret_cast<const uint4 *>(lo);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        ss += (float)n_pad * pad * pad;
        const float in2 = ss > 0.f ? 1.0f / ss : 0.f;
        if (lane == 0) inv_n2[row] = in2;
        best = fmaxf(best, in2);
    }
    if (lane == 0 && best > 0.f && best < 3.0e38f) atomicMax(in2max_bits, __float_as_uint(best));
}

// bits[0] = bound on max |G| = c sqrt(max_rows 1/|f|^2)  (|Y'| <= w_scale |f|; c holds the analytic bound on |grad_out|), bits[1] = 0
__global__ void set_g_bound_kernel(unsigned int *bits, const unsigned int *in2max_bits, float c) {
    const float v = c * sqrtf(__uint_as_float(*in2max_bits));
    bits[0] = (v > 0.f && v < 3.0e38f) ? __float_as_uint(v) : 0u;
    bits[1] = 0u;
}

// loss = sum(partial) / n in a fixed order (one warp); `loss` is float32 or float64
__global__ void loss_finalize_kernel(const double *partial, int n_partial, double n, void *loss, int f64) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 32) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
        if (f64) *reinterpret_cast<double *>(loss) = s / n;
        else *reinterpret_cast<float *>(loss) = (float)(s / n);
    }
}

// Assemble the READ_STATE cotangent of UT for the adjoint gate kernel:
//   gUT[c][k_m].{re,im} = w_scale / gsc * ( c < F ? dWT[n][c] : pad * dWT[n][F] )   (n = 2m+ri), 0 elsewhere;
// dWT[n][F] is the ones-column entry = sum_b G[b,n].
__global__ void assemble_gut_kernel(const float *dWT, const unsigned int *gmax_bits, int A, int F, int ldw, int N,
                                    int stride, float w_scale, float pad, float *gUT) {
    const float k = w_scale / g_scale_final(gmax_bits);
    const long long total = (long long)A * A * 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ri = (int)(i & 1);
        const long long ck = i >> 1;
        const int c = (int)(ck / A), kk = (int)(ck % A);
        float v = 0.f;
        if (kk % stride == 0) {
            const int m = kk / stride;
            const int n = 2 * m + ri;
            if (n < N) v = k * (c < F ? dWT[(long long)n * ldw + c] : pad * dWT[(long long)n * ldw + F]);
        }
        gUT[i] = v;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

// 2-D row-major (rows x cols, pitch elements) tensor map with a (box_cols x box_rows) box.
int make_map_ex(CUtensorMap *map, CUtensorMapDataType dt, int elt_bytes, const void *ptr, long long rows, long long cols,
                long long pitch, int box_cols, int box_rows, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return QIDDM_EUNSUPPORTED;
    if (((uintptr_t)ptr & 15) != 0 || ((pitch * elt_bytes) & 15) != 0) return QIDDM_EINVAL;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * elt_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? QIDDM_OK : QIDDM_EINVAL;
}
// fp16 operand map: rows of `bk` K elements (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B)
int make_map(CUtensorMap *map, const void *ptr, long long rows, long long cols, long long pitch, int box_rows, int bk = BK) {
    return make_map_ex(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, rows, cols, pitch, bk, box_rows,
                       bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

int pick_bn(int N) {
    // multiple of 16 <= 256 that tiles N with the least padding (small tiles pay more per-tile overhead)
    static int forced = -1;
    if (forced < 0) { const char *e = getenv("QIDDM_GEMM_BN"); forced = e ? atoi(e) : 0; }
    if (forced >= 16 && forced <= 256 && forced % 16 == 0) return forced;
    int best = 16;
    double best_cost = 1e30;
    for (int bn = 256; bn >= 16; bn -= 16) {
        const int tiles = (N + bn - 1) / bn;
        const double waste = (double)tiles * bn / N;
        const double cost = waste * (1.0 + 24.0 / bn);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

// B operand MN-major (64-column boxes): the tile that minimises padded MMA work and the operand bytes staged per column
int pick_bn_mn(int N) {
    static int forced = -1;
    if (forced < 0) { const char *e = getenv("QIDDM_GEMM_DW_BN"); forced = e ? atoi(e) : 0; }
    if (forced >= 16 && forced <= 256 && forced % 16 == 0) return forced;
    int best = 16;
    double best_cost = 1e30;
    for (int bn = 256; bn >= 16; bn -= 16) {
        const int tiles = (N + bn - 1) / bn;
        const double waste = (double)tiles * bn / N;
        const double kb_per_col = (32.0 + 16.0 * ((bn / 2 + 63) / 64)) / bn;      // A (256 rows) + B boxes, hi and lo, per stage
        const double cost = waste * (1.0 + 24.0 / bn) + 0.25 * (kb_per_col / 0.325 - 1.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

struct ActOperand { const __half *h, *l; };          // hi, lo
struct WgtOperand { const __half *h, *l; };          // hi, lo

bool use_pair_kernel() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QIDDM_GEMM_PAIR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// The transposed operand splits X^T (Kp, Bp) are only materialised for the cta_group::1 kernel (or QIDDM_GEMM_XT=1); the
// pair kernel reads X row-major as an MN-major B operand in the dW GEMM.
bool use_xt() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QIDDM_GEMM_XT");
        v = (!use_pair_kernel() || (e && e[0] == '1')) ? 1 : 0;
    }
    return v == 1;
}

// D[M,N] = A B^T over the precision segments.  a_mn: A is given as the row-major (K, M) array (MN-major operand).
// b_mn (pair kernel, with a_mn): B is given as the row-major (K, N) array of b_rows x b_cols elements.
int run_gemm(const ActOperand &A, long long a_rows, long long a_cols, long long a_pitch, bool a_mn,
             const WgtOperand &Bm, long long b_rows, long long b_pitch, int M, int N, int K, int n_seg, int k_splits,
             GemmParams &p, cudaStream_t s, bool b_mn = false, long long b_cols = 0) {
    p.M = M; p.N = N; p.K = K;
    p.n_seg = n_seg;
    p.a_mn = a_mn ? 1 : 0;
    p.b_mn = b_mn ? 1 : 0;
    p.bn = b_mn ? pick_bn_mn(N) : pick_bn(N);
    {
        // the last N tile only issues MMA columns that hold data (rounded up to the UMMA granularity of 16)
        // measured (profiles/r2_gemm_summary.md): no gain on dX (its tiles are bound by the A-operand feed, which a narrower tile
        // does not shrink) and the MN-major dW GEMM 12 % SLOWER -- off by default
        static int vartail = -1;
        if (vartail < 0) { const char *e = getenv("QIDDM_GEMM_VARTAIL"); vartail = e ? atoi(e) : 0; }
        const int tiles_n = (N + p.bn - 1) / p.bn;
        const int rest = N - (tiles_n - 1) * p.bn;
        // rounded to 32: each CTA of the pair then holds a multiple of 16 columns of the B tile (176 = 2 x 88 columns ran the
        // MN-major dW GEMM 15 % slower than the padded 208)
        p.bn_last = vartail ? ((rest + 31) / 32) * 32 : p.bn;
        if (p.bn_last > p.bn) p.bn_last = p.bn;
    }
    if (b_mn && !(a_mn && use_pair_kernel())) return QIDDM_EINVAL;
    p.k_splits = k_splits;
    const bool pair = use_pair_kernel();
    {
        static int pfd = -1, nost = -1;
        if (pfd < 0) { const char *e = getenv("QIDDM_GEMM_PFD"); pfd = e ? atoi(e) : 0; }
        if (nost < 0) { const char *e = getenv("QIDDM_GEMM_NOSTORE"); nost = e ? atoi(e) : 0; }
        p.dbg_pfd = pfd; p.dbg_nostore = nost;
        static int hints = -1;
        if (hints < 0) { const char *e = getenv("QIDDM_GEMM_L2_HINTS"); hints = e ? atoi(e) : 0; }
        p.l2_hints = hints;
        static int skip = -1;
        if (skip < 0) { const char *e = getenv("QIDDM_GEMM_SKIP_LOADS"); skip = e ? atoi(e) : 0; }
        p.dbg_skip = skip;
    }
    static int bk32 = -1;
    if (bk32 < 0) { const char *e = getenv("QIDDM_GEMM_BK32"); bk32 = e ? atoi(e) : 0; }
    // MN-major dual-N items (p.dual_n, set by the caller): 32-row k-blocks keep four 48 KB stages in flight where 64-row
    // ones would leave two of 96 KB; K-major: optional 32-wide k-blocks (finer stages; measured slower)
    static int dw_bk = -1;
    if (dw_bk < 0) { const char *e = getenv("QIDDM_GEMM_DW_BK"); dw_bk = e ? atoi(e) : 64; }
    if (!(b_mn && pair)) p.dual_n = 0;
    if (p.dual_n) p.bn_last = p.bn;
    const int bkt = a_mn ? ((pair && p.dual_n && dw_bk == 32) ? 32 : BK) : ((pair && bk32) ? 32 : BK);
    const __half *as[2] = {A.h, A.l};
    const __half *bs[2] = {Bm.h, Bm.l};
    int rc;
    for (int i = 0; i < (n_seg > 1 ? 2 : 1); ++i)
        if ((rc = a_mn ? make_map(&p.a_map[i], as[i], a_rows, a_cols, a_pitch, bkt, 64)      // (64 M) x (bkt K rows) boxes
                       : make_map(&p.a_map[i], as[i], a_rows, a_cols, a_pitch, BM, bkt)) != QIDDM_OK) return rc;
    for (int i = 0; i < (n_seg > 1 ? 2 : 1); ++i)
        if ((rc = b_mn ? make_map(&p.b_map[i], bs[i], b_rows, b_cols, b_pitch, bkt, 64)
                       : make_map(&p.b_map[i], bs[i], b_rows, K, b_pitch, pair ? p.bn / 2 : p.bn, bkt)) != QIDDM_OK) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    timing_begin(TK_GEMM, 2.0 * (double)M * (double)N * (double)K, s);   // single-pass (algorithmic) flops
    cudaError_t e;
    if (pair) {
        // epilogue through shared memory + TMA stores when the outputs are TMA-addressable
        p.tma_epi = 0;
        static int tma_epi_on = -1;
        if (tma_epi_on < 0) { const char *ev = getenv("QIDDM_GEMM_TMA_EPI"); tma_epi_on = ev ? atoi(ev) : 1; }
        if (tma_epi_on && k_splits == 1 && p.epi == EPI_DX && p.dx_x == nullptr) {
            // QConv dX, feature-major (N rows of ldo patches): boxes of 32 patches x 16 features
            p.tma_epi = make_map_ex(&p.o_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.out, N, p.ldo, p.ldo, 32, 16,
                                    CU_TENSOR_MAP_SWIZZLE_NONE) == QIDDM_OK ? 1 : 0;
        } else if (tma_epi_on && k_splits == 1 && p.out_P == 0 && (p.epi == EPI_PROBS || p.epi == EPI_DX)) {
            bool ok = true;
            if (p.epi == EPI_PROBS) {
                if (p.y_out) ok = ok && make_map_ex(&p.y_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.y_out, M, N, N, 16, 32,
                                                    CU_TENSOR_MAP_SWIZZLE_64B) == QIDDM_OK;
                if (p.out) ok = ok && make_map_ex(&p.o_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.out, M, p.n_out, p.ldo, 8, 32,
                                                  CU_TENSOR_MAP_SWIZZLE_NONE) == QIDDM_OK;
            } else {
                ok = make_map_ex(&p.o_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.out, M, N, p.ldo, 16, 32,
                                 CU_TENSOR_MAP_SWIZZLE_64B) == QIDDM_OK;
            }
            p.tma_epi = ok ? 1 : 0;
            // readout epilogue: staged in shared memory, written back with plain 16-byte stores (see epilogue_tile_lsu)
            static int lsu_epi = -1;
            if (lsu_epi < 0) { const char *ev = getenv("QIDDM_GEMM_LSU_EPI"); lsu_epi = ev ? atoi(ev) : 1; }
            if (lsu_epi && p.epi == EPI_PROBS && (N & 3) == 0 && (p.n_out & 3) == 0 && (p.ldo & 3) == 0 &&
                (((uintptr_t)p.y_out | (uintptr_t)p.out) & 15) == 0)
                p.tma_epi = 2;
        }
        if (p.mse.gh != nullptr) {
            if (p.epi != EPI_PROBS || a_mn || bkt != 64 || k_splits != 1) return QIDDM_EINVAL;
            p.tma_epi = 3;      // readout + MSE + dL/dY in the epilogue (epilogue_tile_mse)
        }
        const int epi_bytes = p.tma_epi ? 4 * 2 * EPI_STAGE_BYTES + 512 : 0;
        const int b_tile_bytes = b_mn ? ((p.bn / 2 + 63) / 64) * (64 * bkt * 2) : (p.bn / 2) * bkt * 2;
        const int stage_bytes = (n_seg > 1 ? 2 : 1) * (BM * bkt * 2 + b_tile_bytes * (p.dual_n ? 2 : 1));
        int stages = (226 * 1024 - 1024 - 256 - epi_bytes) / stage_bytes;
        static int max_stages = -1;
        if (max_stages < 0) { const char *ev = getenv("QIDDM_GEMM_STAGES"); max_stages = ev ? atoi(ev) : 8; }
        if (stages > max_stages && max_stages >= 2) stages = max_stages;
        if (stages > 8) stages = 8;
        if (stages < 2) stages = 2;
        p.stages = stages;
        const size_t smem = (size_t)stages * stage_bytes + 1024 + 256 + epi_bytes;
        void (*kern)(const GemmParams);
        if (p.tma_epi == 3)
            kern = p.mse.f64 ? (n_seg > 1 ? gemm_pair_kernel<3, false, 64, false, 2> : gemm_pair_kernel<1, false, 64, false, 2>)
                             : (n_seg > 1 ? gemm_pair_kernel<3, false, 64, false, 1> : gemm_pair_kernel<1, false, 64, false, 1>);
        else if (a_mn && p.dual_n && bkt == 32) kern = n_seg > 1 ? gemm_pair_kernel<3, true, 32, true> : gemm_pair_kernel<1, true, 32, true>;
        else if (a_mn && p.dual_n) kern = n_seg > 1 ? gemm_pair_kernel<3, true, 64, true> : gemm_pair_kernel<1, true, 64, true>;
        else if (a_mn) kern = n_seg > 1 ? gemm_pair_kernel<3, true, 64> : gemm_pair_kernel<1, true, 64>;
        else if (bkt == 32) kern = n_seg > 1 ? gemm_pair_kernel<3, false, 32> : gemm_pair_kernel<1, false, 32>;
        else kern = n_seg > 1 ? gemm_pair_kernel<3, false, 64> : gemm_pair_kernel<1, false, 64>;
        // the opt-in shared-memory limit of a function is per device: remembered per (device, instantiation)
        static std::atomic<bool> attr_set2[64][14];
        const int ki = p.tma_epi == 3 ? 10 + (p.mse.f64 ? 2 : 0) + (n_seg > 1 ? 1 : 0) : p.dual_n ? (bkt == 32 ? 6 : 8) + (n_seg > 1 ? 1 : 0) : (n_seg > 1 ? 3 : 0) + (a_mn ? 2 : (bkt == 32 ? 1 : 0));
        if (!attr_set2[dev & 63][ki].load(std::memory_order_acquire)) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return (int)e;
            attr_set2[dev & 63][ki].store(true, std::memory_order_release);
        }
        const int tiles_n = (N + p.bn - 1) / p.bn;
        const long long tiles = (long long)((M + 2 * BM - 1) / (2 * BM)) * (p.dual_n ? (tiles_n + 1) / 2 : tiles_n) * k_splits;
        const int pairs = (int)(tiles < sms / 2 ? tiles : sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(PAIR_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, p);
    } else {
        const int stage_bytes = BM * BK * 2 + p.bn * BK * 2;
        int stages = (200 * 1024) / stage_bytes;
        if (stages > 8) stages = 8;
        if (stages < 2) stages = 2;
        p.stages = stages;
        const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
        static std::atomic<bool> attr_set[64];
        if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
            e = cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return (int)e;
            attr_set[dev & 63].store(true, std::memory_order_release);
        }
        const long long tiles = (long long)((M + BM - 1) / BM) * ((N + p.bn - 1) / p.bn) * k_splits;
        const int grid = (int)(tiles < sms ? tiles : sms);
        gemm_kernel<<<grid, GEMM_THREADS, smem, s>>>(p);
        e = cudaGetLastError();
    }
    timing_end(s);
    count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// ------------------------------------------------------------------------------------------
// layout of the collapsed operator, the saved forward state and the per-call workspace
// ------------------------------------------------------------------------------------------
int gemm_dw_and_assemble(const GemmShape &g, const GateParams &gp, __half *const X[2], __half *const XT[2], __half *const Gs[2],
                         const unsigned int *gmax, float *dWT, float *gUT, long long B, int n_seg, cudaStream_t s);

GemmShape gemm_shape(const GateParams &gp, int n_qubits) {
    GemmShape g;
    g.A = 1 << n_qubits;
    g.F = gp.n_features;
    g.Fx = g.F < g.A ? g.F + 1 : g.F;        // features + the ones column (pad-row column sums)
    g.Kp = (g.Fx + 7) & ~7;
    g.n_out = gp.read_count;
    g.N = 2 * g.n_out;
    g.Np = (g.N + 7) & ~7;
    g.stride = gp.read_stride;
    g.w_scale = ldexpf(1.f, (n_qubits + 1) / 2);     // keeps |W'| ~ O(1): |U| entries are ~ 2^(-n/2)
    return g;
}

size_t gemm_collapsed_bytes(const GemmShape &g) {
    size_t b = 0;
    b += al((size_t)g.A * g.A * 8);                  // UT
    b += 2 * al((size_t)g.N * g.Kp * 2);             // Wn hi/lo
    b += 2 * al((size_t)g.F * g.Np * 2);             // Wt hi/lo
    b += al((size_t)g.N * 4);                        // bias
    b += conv_wd_bytes(g);                           // fp32 rows of U for the direct QConv path
    return b;
}

struct CollapsedView {
    float2 *UT;
    __half *Wn[2], *Wt[2];
    float *bias;
    float *Wd;
};
static CollapsedView collapsed_view(const GemmShape &g, void *buf) {
    CollapsedView v;
    char *p = reinterpret_cast<char *>(buf);
    v.UT = reinterpret_cast<float2 *>(p); p += al((size_t)g.A * g.A * 8);
    for (int i = 0; i < 2; ++i) { v.Wn[i] = reinterpret_cast<__half *>(p); p += al((size_t)g.N * g.Kp * 2); }
    for (int i = 0; i < 2; ++i) { v.Wt[i] = reinterpret_cast<__half *>(p); p += al((size_t)g.F * g.Np * 2); }
    v.bias = reinterpret_cast<float *>(p); p += al((size_t)g.N * 4);
    v.Wd = reinterpret_cast<float *>(p);
    return v;
}

int gemm_build_operands(const GemmShape &g, const GateParams &gp, void *collapsed, cudaStream_t s) {
    CollapsedView v = collapsed_view(g, collapsed);
    cudaError_t e;
    timing_begin(TK_BUILD_W, 0.0, s);
    build_w_kernel<<<g.N, 256, 0, s>>>(v.UT, g.A, g.F, g.Kp, g.N, g.Np, g.stride, g.w_scale, gp.pad_value, v.Wn[0],
                                        v.Wn[1], v.Wt[0], v.Wt[1], v.bias);
    if (g.Fx > g.F) {
        fold_bias_kernel<<<(g.N + 127) / 128, 128, 0, s>>>(v.bias, g.N, g.Kp, g.F, v.Wn[0], v.Wn[1]);
        count_launch();
    }
    timing_end(s);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    return conv_build_wd(g, gp, reinterpret_cast<const float *>(v.UT), v.Wd, s);
}

float *gemm_collapsed_wd(const GemmShape &g, void *collapsed) { return collapsed_view(g, collapsed).Wd; }
float *gemm_collapsed_ut(const GemmShape &g, void *collapsed) { return reinterpret_cast<float *>(collapsed_view(g, collapsed).UT); }

size_t gemm_saved_bytes(const GemmShape &g, long long B) {
    const long long Bp = (B + 7) & ~7LL;
    return 2 * al((size_t)B * g.Kp * 2) + (use_xt() ? 2 * al((size_t)g.Kp * Bp * 2) : 0) + al((size_t)B * 4) +
           al((size_t)B * g.N * 4);
}

size_t gemm_forward_ws_bytes(const GemmShape &g, long long B) {
    return 2 * al((size_t)B * g.Kp * 2) + al((size_t)B * 4);
}

static UnfoldGeom unfold_geom(const GateParams &gp) {
    UnfoldGeom u;
    u.on = gp.unfold; u.C = gp.C; u.H = gp.H; u.W = gp.W; u.kh = gp.kh; u.kw = gp.kw; u.ph = gp.ph; u.pw = gp.pw;
    u.Hout = gp.Hout; u.Wout = gp.Wout;
    return u;
}

size_t gemm_backward_ws_bytes(const GemmShape &g, long long B, bool unfold) {
    size_t b = gemm_saved_bytes(g, B);                // used when the forward did not save
    if (unfold) b += al((size_t)((B + 7) & ~7LL) * g.F * 4);   // dX (feature-major) before the col2im
    b += al((size_t)B * 4) + al(256);                 // S, gmax
    b += 2 * al((size_t)B * g.Np * 2);                // G splits (row-major)
    b += al((size_t)g.N * ((g.Fx + 3) & ~3) * 4);     // dWT (+ ones column; rows padded to 16 bytes)
    b += al((size_t)g.A * g.A * 8);                   // gUT
    return b;
}

namespace {
struct SavedView {
    __half *X[2], *XT[2];
    float *inv_n2, *Y;
    char *end;
};
// full == false: only X splits + inv_n2 (inference forward)
SavedView saved_view(const GemmShape &g, long long B, void *buf, bool full) {
    SavedView w;
    const long long Bp = (B + 7) & ~7LL;
    char *p = reinterpret_cast<char *>(buf);
    for (int i = 0; i < 2; ++i) { w.X[i] = reinterpret_cast<__half *>(p); p += al((size_t)B * g.Kp * 2); }
    for (int i = 0; i < 2; ++i) {
        const bool xt = full && use_xt();
        w.XT[i] = xt ? reinterpret_cast<__half *>(p) : nullptr;
        if (xt) p += al((size_t)g.Kp * Bp * 2);
    }
    w.inv_n2 = reinterpret_cast<float *>(p); p += al((size_t)B * 4);
    w.Y = nullptr;
    if (full) { w.Y = reinterpret_cast<float *>(p); p += al((size_t)B * g.N * 4); }
    w.end = p;
    return w;
}
}  // namespace

// Forward.  `out` may be null (backward re-materialisation).  `saved` (gemm_saved_bytes) non-null: the X (with use_xt(): and
// X^T) splits, 1/|f|^2 and Y are kept there for the backward pass; null: they live in `ws` and only `out` is produced.
int gemm_forward(const GemmShape &g, const GateParams &gp, const void *collapsed, const float *x, float *out,
                 void *saved, void *ws, long long B, int n_seg, cudaStream_t s) {
    CollapsedView v = collapsed_view(g, const_cast<void *>(collapsed));
    const bool keep = saved != nullptr;
    SavedView w = saved_view(g, B, keep ? saved : ws, keep);
    const long long Bp = (B + 7) & ~7LL;
    const int warps = 8;
    if (keep && gp.unfold) {
        // training forward, QConv: X (X^T only when use_xt()) and the norms straight from the NCHW image
        timing_begin(TK_PREP_X, 0.0, s);
        prep_xt_unfold_kernel<<<(unsigned)((Bp + 63) / 64), 256, 2 * g.F * sizeof(int), s>>>(
            x, gp.io64, B, g.F, g.Kp, Bp, g.A - g.F, gp.add_offset, gp.pad_value, w.X[0], w.X[1], w.XT[0], w.XT[1], w.inv_n2,
            unfold_geom(gp));
        timing_end(s);
        count_launch();
    } else if (keep && use_xt()) {
        // training forward, dense rows: X, X^T and the norms in one pass over x
        timing_begin(TK_PREP_X, 0.0, s);
        prep_xt_kernel<<<(unsigned)((Bp + 63) / 64), 256, 0, s>>>(x, B, g.F, g.Kp, Bp, g.A - g.F, gp.add_offset, gp.pad_value,
                                                                 w.X[0], w.X[1], w.XT[0], w.XT[1], w.inv_n2);
        timing_end(s);
        count_launch();
    } else {
        int gw = 32;                                   // lanes per row: smallest power of two >= Kp / 2
        while (gw > 1 && gw / 2 >= g.Kp / 2) gw >>= 1;
        const long long row_warps = (B * gw + 31) / 32;
        timing_begin(TK_PREP_X, 0.0, s);
        if (!gp.unfold && (g.F & 3) == 0 && g.Kp >= 128 && ((uintptr_t)x & 15) == 0) {
            const long long blocks = (B + warps - 1) / warps;
            prep_x_dense_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), warps * 32, 0, s>>>(
                x, B, g.F, g.Kp, g.A - g.F, gp.add_offset, gp.pad_value, w.X[0], w.X[1], w.inv_n2, keep || n_seg > 1);
        } else {
            prep_x_kernel<<<(unsigned)((row_warps + warps - 1) / warps), warps * 32, gp.unfold ? 2 * g.F * sizeof(int) : 0, s>>>(
                x, B, g.F, g.Kp, g.A - g.F, gp.add_offset, gp.pad_value, w.X[0], w.X[1], w.inv_n2, keep || n_seg > 1,
                unfold_geom(gp), gw, gp.unfold ? gp.io64 : 0);
        }
        timing_end(s);
        count_launch();
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.epi = EPI_PROBS;
    p.out = out; p.ldo = g.n_out; p.out_scale = 1.f;
    p.y_out = w.Y;
    p.bias = nullptr; p.row_scale = w.inv_n2;      // the bias rides in the ones column of X (fold_bias_kernel)
    p.post_scale = gp.post_scale / (g.w_scale * g.w_scale);
    p.clamp = gp.clamp; p.clamp_lo = gp.clamp_lo; p.clamp_hi = gp.clamp_hi;
    p.n_out = g.n_out;
    p.out_P = gp.unfold ? gp.Hout * gp.Wout : 0;    // QConv: probabilities go straight to the NCHW output
    p.out_f64 = gp.unfold ? gp.io64 : 0;
    ActOperand A{w.X[0], w.X[1]};
    WgtOperand Bm{v.Wn[0], v.Wn[1]};
    timing_set_gemm_kind(TK_GEMM_FWD);
    return run_gemm(A, B, g.Kp, g.Kp, false, Bm, g.N, g.Kp, (int)B, g.N, g.Kp, n_seg, 1, p, s);
}

// Produces grad_in (B,F) (nullable) and the READ_STATE cotangent gUT (A x 2A fp32) for the adjoint gate kernel.
// `saved` null: X splits / Y are re-materialised first (one extra GEMM).
int gemm_backward(const GemmShape &g, const GateParams &gp, const void *collapsed, const float *x,
                  const float *grad_out, const void *saved, float *grad_in, float **gut_out, void *ws, long long B,
                  int n_seg, cudaStream_t s) {
    CollapsedView v = collapsed_view(g, const_cast<void *>(collapsed));
    const long long Bp = (B + 7) & ~7LL;
    char *p8 = reinterpret_cast<char *>(ws);
    void *saved_buf = const_cast<void *>(saved);
    if (saved_buf == nullptr) {
        saved_buf = p8;
        int rc0 = gemm_forward(g, gp, collapsed, x, nullptr, saved_buf, nullptr, B, n_seg, s);
        if (rc0 != QIDDM_OK) return rc0;
    }
    // n_seg here may be 1 behind an n_seg = 3 forward ("x3 forward, x1 gradients": the saved Y and X splits are fp32-grade, only
    // the dX / dW GEMMs run single-pass; stated gradient bound in DESIGN.md 4.2)
    p8 += gemm_saved_bytes(g, B);
    SavedView w = saved_view(g, B, saved_buf, true);
    float *S = reinterpret_cast<float *>(p8); p8 += al((size_t)B * 4);
    unsigned int *gmax = reinterpret_cast<unsigned int *>(p8); p8 += al(256);
    __half *Gs[2];
    for (int i = 0; i < 2; ++i) { Gs[i] = reinterpret_cast<__half *>(p8); p8 += al((size_t)B * g.Np * 2); }
    const int ldw = (g.Fx + 3) & ~3;                  // 16-byte rows: vector reductions in the split-K epilogue
    float *dWT = reinterpret_cast<float *>(p8); p8 += al((size_t)g.N * ldw * 4);
    float *gUT = reinterpret_cast<float *>(p8); p8 += al((size_t)g.A * g.A * 8);
    *gut_out = gUT;
    float *dx_rows = reinterpret_cast<float *>(p8);   // QConv only (gemm_backward_ws_bytes(.., unfold = true))
    const int go_P = gp.unfold ? gp.Hout * gp.Wout : 0;

    cudaError_t e;
    const float eff_scale = gp.post_scale / (g.w_scale * g.w_scale);
    // (1) scale bound, then one streaming pass: G splits (row-major) and S
    if ((e = cudaMemsetAsync(gmax, 0, 8, s)) != cudaSuccess) return (int)e;
    const int warps = 8;
    timing_begin(TK_G_BOUND, 0.0, s);
    int gw = 32;                                       // lanes per row: smallest power of two >= Np / 2
    while (gw > 1 && gw / 2 >= g.Np / 2) gw >>= 1;
    const long long row_warps = (B * gw + 31) / 32;
    const unsigned ew_grid = (unsigned)((row_warps + warps - 1) / warps < 148 * 8 ? (row_warps + warps - 1) / warps : 148 * 8);
    // large batches: the provisional bound reads every 64th row only (x 2^6 margin: the hi/lo operands have 2^16 of slack)
    // and grad_y verifies it against the exact bound it accumulates on the way (second, normally empty, launch below)
    static int gsample = -1;
    if (gsample < 0) { const char *ev = getenv("QIDDM_GEMM_GSAMPLE"); gsample = ev ? atoi(ev) : 64; if (gsample < 1) gsample = 1; }
    const long long g_stride = (gsample > 1 && B >= 16384) ? gsample : 1;
    const long long Bs = (B + g_stride - 1) / g_stride;
    const long long s_warps = (Bs * gw + 31) / 32;
    const unsigned s_grid = (unsigned)((s_warps + warps - 1) / warps < 148 * 8 ? (s_warps + warps - 1) / warps : 148 * 8);
    g_bound_kernel<<<s_grid, warps * 32, 0, s>>>(grad_out, w.inv_n2, Bs, g.n_out, eff_scale, g.w_scale, gmax, go_P, gw,
                                                 gp.unfold ? gp.io64 : 0, g_stride, g_stride > 1 ? 64.f : 1.f);
    timing_end(s);
    count_launch();
    timing_begin(TK_GRAD_Y, 0.0, s);
    for (int retry = 0; retry < (g_stride > 1 ? 2 : 1); ++retry) {
        grad_y_kernel<<<ew_grid, warps * 32, 0, s>>>(
            w.Y, grad_out, w.inv_n2, B, g.N, g.Np, g.n_out, eff_scale, gp.clamp, gp.clamp_lo, gp.clamp_hi, gmax, Gs[0],
            Gs[1], S, n_seg > 1, go_P, gw, gp.unfold ? gp.io64 : 0, g.w_scale, retry);
        count_launch();
    }
    timing_end(s);
    GemmParams p;
    int rc;
    ActOperand Go{Gs[0], Gs[1]};
    // (2) dX = G W^T / gsc - 2 f inv_n2 S   (normalisation term fused in the epilogue)
    if (grad_in != nullptr) {
        memset(&p, 0, sizeof(p));
        p.epi = EPI_DX; p.out = gp.unfold ? dx_rows : grad_in; p.ldo = gp.unfold ? Bp : g.F; p.out_scale = 1.f;
        p.row_scale = w.inv_n2; p.dx_S = S; p.dx_x = gp.unfold ? nullptr : x; p.gmax_bits = gmax;
        p.add_offset = gp.add_offset;
        WgtOperand Wt{v.Wt[0], v.Wt[1]};
        timing_set_gemm_kind(TK_GEMM_DX);
        rc = run_gemm(Go, B, g.Np, g.Np, false, Wt, g.F, g.Np, (int)B, g.F, g.N, n_seg, 1, p, s);
        if (rc != QIDDM_OK) return rc;
        if (gp.unfold) {     // col2im + normalisation term: image gradient in NCHW
            const long long n_elems = (B / go_P) * gp.C * gp.H * gp.W;
            const unsigned fg = (unsigned)((n_elems + 255) / 256 < 148 * 16 ? (n_elems + 255) / 256 : 148 * 16);
            timing_begin(TK_FINISH_DX, 0.0, s);
            if (gp.kh == 3 && gp.kw == 3)
                fold_rows_kernel<3, 3><<<fg, 256, 0, s>>>(dx_rows, Bp, x, gp.io64, w.inv_n2, S, n_elems, gp.add_offset,
                                                          unfold_geom(gp), grad_in);
            else if (gp.kh == 1 && gp.kw == 1)
                fold_rows_kernel<1, 1><<<fg, 256, 0, s>>>(dx_rows, Bp, x, gp.io64, w.inv_n2, S, n_elems, gp.add_offset,
                                                          unfold_geom(gp), grad_in);
            else
                fold_rows_kernel<0, 0><<<fg, 256, 0, s>>>(dx_rows, Bp, x, gp.io64, w.inv_n2, S, n_elems, gp.add_offset,
                                                          unfold_geom(gp), grad_in);
            timing_end(s);
            count_launch();
        }
    }
    return gemm_dw_and_assemble(g, gp, w.X, w.XT, Gs, gmax, dWT, gUT, B, n_seg, s);
}

// dW^T[n][c] = sum_b G[b,n] f[b,c] (G consumed row-major as an MN-major operand, split-K over the batch), then the READ_STATE
// cotangent gUT for the adjoint gate kernel.  Shared by gemm_backward and the fused diffusion step.
int gemm_dw_and_assemble(const GemmShape &g, const GateParams &gp, __half *const X[2], __half *const XT[2], __half *const Gs[2],
                         const unsigned int *gmax, float *dWT, float *gUT, long long B, int n_seg, cudaStream_t s) {
    const long long Bp = (B + 7) & ~7LL;
    const int ldw = (g.Fx + 3) & ~3;
    cudaError_t e;
    int rc;
    GemmParams p;
    ActOperand Go{Gs[0], Gs[1]};
    if ((e = cudaMemsetAsync(dWT, 0, (size_t)g.N * ldw * 4, s)) != cudaSuccess) return (int)e;
    {
        memset(&p, 0, sizeof(p));
        p.epi = EPI_STORE; p.out = dWT; p.ldo = ldw; p.out_scale = 1.f;
        const bool xt = use_xt();
        WgtOperand XTo{xt ? XT[0] : X[0], xt ? XT[1] : X[1]};
        const int bn = xt ? pick_bn(g.Fx) : pick_bn_mn(g.Fx);
        const bool pair = use_pair_kernel();
        const int bm = pair ? 2 * BM : BM;
        // long split-K items (big batches) of the MN-major GEMM: two neighbouring N tiles per item share one staged G tile
        // (25 % fewer operand bytes per MMA through L2 -> shared memory, and G is read once per tile PAIR)
        static int dw_dual = -1;
        if (dw_dual < 0) { const char *ev = getenv("QIDDM_GEMM_DW_DUAL"); dw_dual = ev ? atoi(ev) : 1; }
        const int tiles_n = (g.Fx + bn - 1) / bn;
        const bool dual = dw_dual && !xt && pair && tiles_n >= 2 && Bp >= 65536;
        p.dual_n = dual ? 1 : 0;
        const int tiles = ((g.N + bm - 1) / bm) * (dual ? (tiles_n + 1) / 2 : tiles_n);
        const long long kt = pair ? (Bp + BK - 1) / BK : (long long)n_seg * ((Bp + BK - 1) / BK);
        // split-K over the batch: fill whole waves of the persistent grid (items = tiles * splits close to a multiple of the
        // number of CTAs / pairs), with a small price per split for its fp32 atomics
        const int workers = pair ? 74 : 148;
        int splits = 1;
        double best = 1e30;
        // ... and, for the pair kernel, a price on the LENGTH of an item's accumulation chain: every tcgen05.mma accumulate
        // truncates, so a same-sign sum over r batch rows comes out low by about 3.5e-8 r relative (measured through the
        // parameter gradient, profiles/r2b_summary.md); 74 instead of 37 splits at 2^19 rows cost 0.5 % of the launch
        const int max_splits = pair ? 128 : 64;
        for (int sp = 1; sp <= max_splits && sp <= kt; ++sp) {
            const long long items = (long long)tiles * sp;
            const long long waves = (items + workers - 1) / workers;
            const double per_split = pair ? 0.0005 * sp + 4e-6 * (double)Bp / sp : 0.004 * sp;
            const double cost = (double)(waves * workers) / (double)items + per_split + (items < workers ? 10.0 : 0.0);
            if (cost < best) { best = cost; splits = sp; }
        }
        {
            static int forced = -1;
            if (forced < 0) { const char *ev = getenv("QIDDM_GEMM_DW_SPLITS"); forced = ev ? atoi(ev) : 0; }
            if (forced > 0 && forced <= kt) splits = forced;
        }
        timing_set_gemm_kind(TK_GEMM_DW);
        if (xt) rc = run_gemm(Go, B, g.Np, g.Np, true, XTo, g.Fx, Bp, g.N, g.Fx, (int)Bp, n_seg, splits, p, s);
        else rc = run_gemm(Go, B, g.Np, g.Np, true, XTo, B, g.Kp, g.N, g.Fx, (int)Bp, n_seg, splits, p, s, true, g.Kp);
        if (rc != QIDDM_OK) return rc;
    }
    timing_begin(TK_ASSEMBLE, 0.0, s);
    assemble_gut_kernel<<<1184, 256, 0, s>>>(dWT, gmax, g.A, g.F, ldw, g.N, g.stride, g.w_scale, gp.pad_value, gUT);
    timing_end(s);
    count_launch();
    e = cudaGetLastError();
    return e == cudaSuccess ? QIDDM_OK : (int)e;
}


// ------------------------------------------------------------------------------------------
// Fused diffusion training step of ONE amplitude-embedding layer (Diffusion(QDenseUndirected_old[_noise]) of
// src/models.py:44-104 around nn/qdense.py:95-111): ladder -> splits, forward GEMM with the MSE loss and dL/dY in its
// epilogue, dW GEMM, cotangent of U^T.  Rows = (image b, level t), t < T; no grad_in (the layer is the first one).
// ------------------------------------------------------------------------------------------
size_t gemm_dense_mse_ws_bytes(const GemmShape &g, long long B) {
    size_t b = 2 * al((size_t)B * g.Kp * 2) + al((size_t)B * 4) + al(256);      // X splits, 1/|f|^2, gmax words
    b += 2 * al((size_t)B * g.Np * 2);                                          // G splits
    b += al((size_t)g.N * ((g.Fx + 3) & ~3) * 4);                               // dWT
    b += al((size_t)g.A * g.A * 8);                                             // gUT
    b += al((size_t)8 * 4 * 2 * 148);                                           // loss partials (one per epilogue warp)
    return b;
}

int gemm_dense_mse_step(const GemmShape &g, const GateParams &gp, const void *collapsed, const void *x, const float *eps,
                        const void *w, int io64, long long n_img, int T, float a, float bshift, float c0, float c1,
                        void *loss_out, float **gut_out, void *ws, int n_seg_fwd, int n_seg_bwd, cudaStream_t s) {
    if (gp.unfold || g.n_out != g.F || !use_pair_kernel() || use_xt() || T < 1) return QIDDM_EUNSUPPORTED;
    CollapsedView v = collapsed_view(g, const_cast<void *>(collapsed));
    const long long B = n_img * T;
    char *p8 = reinterpret_cast<char *>(ws);
    __half *X[2], *XT[2] = {nullptr, nullptr}, *Gs[2];
    for (int i = 0; i < 2; ++i) { X[i] = reinterpret_cast<__half *>(p8); p8 += al((size_t)B * g.Kp * 2); }
    float *inv_n2 = reinterpret_cast<float *>(p8); p8 += al((size_t)B * 4);
    unsigned int *gmax = reinterpret_cast<unsigned int *>(p8); p8 += al(256);
    for (int i = 0; i < 2; ++i) { Gs[i] = reinterpret_cast<__half *>(p8); p8 += al((size_t)B * g.Np * 2); }
    const int ldw = (g.Fx + 3) & ~3;
    float *dWT = reinterpret_cast<float *>(p8); p8 += al((size_t)g.N * ldw * 4);
    float *gUT = reinterpret_cast<float *>(p8); p8 += al((size_t)g.A * g.A * 8);
    double *partial = reinterpret_cast<double *>(p8);
    *gut_out = gUT;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > 2 * 148) return QIDDM_EUNSUPPORTED;

    cudaError_t e;
    if ((e = cudaMemsetAsync(gmax, 0, 256, s)) != cudaSuccess) return (int)e;
    // (1) noisy rows straight to the GEMM operand splits
    const bool want_lo = n_seg_fwd > 1 || n_seg_bwd > 1;
    const long long blocks = (n_img + 7) / 8;
    const unsigned grid = (unsigned)(blocks < 148 * 8 ? blocks : 148 * 8);
    const int its = (g.Kp + 255) / 256;              // 8 values per lane and iteration
    timing_begin(TK_PREP_X, 0.0, s);
    if (its > 5) {                                   // n = 11, 12: one warp per row
        const long long rblocks = (n_img * T + 7) / 8;
        const unsigned rgrid = (unsigned)(rblocks < 148 * 16 ? rblocks : 148 * 16);
        if (io64)
            ladder_prep_rows_kernel<double><<<rgrid, 256, 0, s>>>(reinterpret_cast<const double *>(x), eps, reinterpret_cast<const double *>(w),
                                                                  n_img, T, g.F, g.Kp, g.A - g.F, gp.add_offset, gp.pad_value, X[0], X[1],
                                                                  inv_n2, want_lo, gmax + 2);
        else
            ladder_prep_rows_kernel<float><<<rgrid, 256, 0, s>>>(reinterpret_cast<const float *>(x), eps, reinterpret_cast<const float *>(w),
                                                                 n_img, T, g.F, g.Kp, g.A - g.F, gp.add_offset, gp.pad_value, X[0], X[1],
                                                                 inv_n2, want_lo, gmax + 2);
    } else
#define QIDDM_LADDER_PREP(TT, ITS_)                                                                                              \
    ladder_prep_kernel<TT, ITS_><<<grid, 256, 0, s>>>(reinterpret_cast<const TT *>(x), eps, reinterpret_cast<const TT *>(w), n_img, \
                                                      T, g.F, g.Kp, g.A - g.F, gp.add_offset, gp.pad_value, X[0], X[1], inv_n2,   \
                                                      want_lo, gmax + 2)
    if (io64) {
        if (its <= 1) QIDDM_LADDER_PREP(double, 1); else if (its == 2) QIDDM_LADDER_PREP(double, 2);
        else if (its <= 4) QIDDM_LADDER_PREP(double, 4); else QIDDM_LADDER_PREP(double, 5);
    } else {
        if (its <= 1) QIDDM_LADDER_PREP(float, 1); else if (its == 2) QIDDM_LADDER_PREP(float, 2);
        else if (its <= 4) QIDDM_LADDER_PREP(float, 4); else QIDDM_LADDER_PREP(float, 5);
    }
#undef QIDDM_LADDER_PREP
    timing_end(s);
    count_launch();
    // (2) analytic bound on |G|: |d| <= |a| max|out| + |b| + |c0| + |c1| (levels are clamped to [0, 1]), |grad_out| = |kk d|,
    //     |G| = 2 |grad_out| post inv_n2 |Y'| <= 2 |grad_out| post w_scale sqrt(inv_n2)
    const float eff_scale = gp.post_scale / (g.w_scale * g.w_scale);
    const double out_max = gp.clamp ? fmax(fabs((double)gp.clamp_lo), fabs((double)gp.clamp_hi)) : fabs((double)gp.post_scale);
    const double d_max = fabs((double)a) * out_max + fabs((double)bshift) + fabs((double)c0) + fabs((double)c1);
    const double kk = 2.0 * (double)a / ((double)B * (double)g.n_out);
    const double cbound = 2.0 * fabs(kk) * d_max * (double)eff_scale * (double)g.w_scale;
    set_g_bound_kernel<<<1, 1, 0, s>>>(gmax, gmax + 2, (float)cbound);
    count_launch();
    // (3) forward GEMM, epilogue = readout + MSE + dL/dY -> G splits, loss partials
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.epi = EPI_PROBS;
    p.out = nullptr; p.ldo = g.n_out; p.out_scale = 1.f;
    p.y_out = nullptr;
    p.bias = nullptr; p.row_scale = inv_n2;
    p.post_scale = eff_scale;
    p.clamp = gp.clamp; p.clamp_lo = gp.clamp_lo; p.clamp_hi = gp.clamp_hi;
    p.n_out = g.n_out;
    p.gmax_bits = gmax;
    p.mse.x = x; p.mse.eps = eps; p.mse.w = w; p.mse.T = T; p.mse.f64 = io64; p.mse.P = g.F;
    p.mse.a = a; p.mse.b = bshift; p.mse.c0 = c0; p.mse.c1 = c1; p.mse.kk = kk;
    p.mse.gh = Gs[0]; p.mse.gl = Gs[1]; p.mse.ldg = g.Np; p.mse.want_lo = n_seg_bwd > 1 ? 1 : 0;
    p.mse.loss_partial = partial;
    if ((e = cudaMemsetAsync(partial, 0, (size_t)8 * 4 * sms, s)) != cudaSuccess) return (int)e;
    ActOperand A{X[0], X[1]};
    WgtOperand Bm{v.Wn[0], v.Wn[1]};
    timing_set_gemm_kind(TK_GEMM_FWD);
    // the G splits carry what the gradient GEMMs need: with single-pass gradients (n_seg_bwd == 1) only the hi part is written
    int rc = run_gemm(A, B, g.Kp, g.Kp, false, Bm, g.N, g.Kp, (int)B, g.N, g.Kp, n_seg_fwd, 1, p, s);
    if (rc != QIDDM_OK) return rc;
    loss_finalize_kernel<<<1, 32, 0, s>>>(partial, 4 * sms, (double)B * (double)g.n_out, loss_out, io64);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    // (4) dW GEMM + cotangent of U^T
    return gemm_dw_and_assemble(g, gp, X, XT, Gs, gmax, dWT, gUT, B, n_seg_bwd, s);
}

}  // namespace qiddm
