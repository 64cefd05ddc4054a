// A3, tensor-core tier (prec==1): padding-masked self-attention over the packed token stream with the three
// contractions (Q K^T, P V and their five backward counterparts) on warp-level tensor-core MMAs (m16n8k8, TF32
// operands, fp32 accumulate) and the softmax in registers -- the T x T matrix lives only in accumulator fragments.
//
// Why warp-level mma and not tcgen05 here: the head dimension is 8 or 16, so one 16x8 score tile costs a single
// MMA while its softmax costs ~6 instructions per score; the kernel is bound by the exp/scale pipeline, and a
// TMEM round trip per score tile (tcgen05.ld / st) would add traffic without removing any of that work.
//
// One CTA per (sequence, head), 4 warps.  Forward: a warp owns 16 query rows, streams the keys in guard-free blocks of
// 64 / 32 (zero-padded) with an online softmax.  Backward: phase 1 (dQ, warp owns 16 queries, streams keys) and phase 2 (dK/dV, warp owns 16
// keys, streams queries) recompute probabilities from the saved log-sum-exp; no atomics, deterministic.
// The probability / dS accumulator fragments are fed back as the A operand of the next MMA without any shuffle by
// relabelling the contraction index (k = t <-> column 2t, k = t+4 <-> column 2t+1) consistently on the B side.
#include <stdlib.h>

#include "common.cuh"

namespace mvn {
namespace {

constexpr int ATC_THREADS = 128;
constexpr int CH = 256;                       // keys (or queries) staged in shared memory at a time
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float TRUNC1 = 1.0f + 3.52e-4f;     // mean shrink of one operand the tensor core truncates to TF32 (see gemm_tc.cu::trunc_comp)

// Round-to-nearest TF32 without touching the XU pipe (cvt.rna.tf32 issues there, next to the exp2 the softmax needs):
// the MMA ignores the low 13 mantissa bits of its operands, so adding half an ulp of TF32 to the bit pattern and letting
// the hardware truncate rounds to nearest (ties away from zero).  One integer add on the ALU pipe.
__device__ __forceinline__ float tf32r(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
// 2^x on the SFU, flush-to-zero: one MUFU.EX2 (exp2f() wraps it in a denormal-range rescale the softmax does not need;
// ex2(-inf) = +0, which is what masked / padded positions rely on)
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// D(16x8) += A(16x8, row) * B(8x8, col);  fragment layouts (g = lane/4, t = lane%4):
//   a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4) ; b0:(k=t,n=g) b1:(k=t+4,n=g) ; c0:(g,2t) c1:(g,2t+1) c2:(g+8,2t) c3:(g+8,2t+1)
__device__ __forceinline__ void mma_tf32(float* c, const float* a, float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// rows [row0, row0+cnt) of a column slice (HD floats at column `col`) of a row-major [*, ld] matrix -> smem
// [cnt_pad8][HD+4] rounded to TF32 and multiplied by `mul`; rows up to the next multiple of 8 are zero-filled.
template <int HD>
__device__ __forceinline__ void load_slice(float* dst, const float* __restrict__ src, size_t ld, int col, int row0, int cnt, float mul) {
    constexpr int LD = HD + 4;
    const int pad = (cnt + 7) & ~7;
    for (int idx = threadIdx.x; idx < pad * (HD / 4); idx += ATC_THREADS) {
        const int j = idx / (HD / 4), d = (idx % (HD / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < cnt) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(row0 + j) * ld + col + d));
        *reinterpret_cast<float4*>(dst + j * LD + d) = make_float4(tf32r(v.x * mul), tf32r(v.y * mul), tf32r(v.z * mul), tf32r(v.w * mul));
    }
}
// cp.async: 16 bytes global -> shared without passing through registers; src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// ---- skewed rows ----------------------------------------------------------------------------------------------------------
// The backward reads each staged operand in BOTH fragment forms (K-form: row g, HD/4 contiguous floats at t*HD/4, a 64- / 128-bit
// load; V-form: rows 2t and 2t+1, HD/8 floats at g*HD/8).  No constant row stride is conflict-free for both (HD+4 made one
// shared-memory wavefront in three of the head-dim-8 backward a bank conflict), a row-dependent skew is:
//   HD = 8 : row r at 8*(r + r/4)   -> K-form rows r..r+3 on banks {0,8,16,24}+, V-form rows 0,2,4,6 of a tile on {0,16,8,24}+
//   HD = 16: row r at 16*r + 8*(r/2) -> K-form row pairs on disjoint 16-bank halves, V-form rows 0,2,4,6 on {0,8,16,24}+
template <int HD>
__device__ __forceinline__ int rowoff(int r) { return HD == 8 ? 8 * (r + (r >> 2)) : 16 * r + 8 * (r >> 1); }
template <int HD>
constexpr int skew_floats(int rows) { return HD == 8 ? 8 * (rows + rows / 4) : 16 * rows + 8 * (rows / 2); }
// like load_slice (rows up to the next multiple of 8 zero-filled), skewed rows.  A thread walks one 16-byte column slot down the
// rows in steps of ATC_THREADS / (HD/4) rows (a multiple of 8, over which rowoff is linear): source and destination advance by
// constants, no per-element index arithmetic.
template <int HD>
__device__ __forceinline__ void load_slice_skew(float* dst, const float* __restrict__ src, size_t ld, int col, int row0, int cnt, float mul) {
    constexpr int Q = HD / 4, STEP = ATC_THREADS / Q, U = 4;
    const int pad = (cnt + 7) & ~7;
    int j = threadIdx.x / Q;
    const int d = (threadIdx.x % Q) * 4;
    const float* p = src + (size_t)(row0 + j) * ld + col + d;
    float* q = dst + rowoff<HD>(j) + d;
    const size_t pstep = (size_t)STEP * ld;
    constexpr int QSTEP = skew_floats<HD>(STEP);
    for (; j < pad; j += U * STEP, p += U * pstep, q += U * QSTEP) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {                               // U loads in flight before the first store
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j + u * STEP < cnt) v[u] = __ldg(reinterpret_cast<const float4*>(p + u * pstep));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u * STEP < pad)
                *reinterpret_cast<float4*>(q + u * QSTEP) = make_float4(tf32r(v[u].x * mul), tf32r(v[u].y * mul), tf32r(v[u].z * mul), tf32r(v[u].w * mul));
    }
}
// asynchronous version (cp.async, raw fp32 bits: the tensor core truncates them, callers fold TRUNC1 into the other operand / the
// output scale); rows up to the next multiple of 8 zero-filled
template <int HD>
__device__ __forceinline__ void load_slice_skew_async(float* dst, const float* __restrict__ src, size_t ld, int col, int row0, int cnt) {
    const int pad = (cnt + 7) & ~7;
    for (int idx = threadIdx.x; idx < pad * (HD / 4); idx += ATC_THREADS) {
        const int j = idx / (HD / 4), d = (idx % (HD / 4)) * 4;
        const int jr = j < cnt ? j : cnt - 1;
        cp_async16(dst + rowoff<HD>(j) + d, src + (size_t)(row0 + jr) * ld + col + d, j < cnt ? 16 : 0);
    }
}
// ---- operand relabelling -------------------------------------------------------------------------------------------
// An MMA's contraction slots and output columns can be assigned to head dimensions in any order as long as both operands
// (resp. the consumer of the accumulator) agree.  Two assignments make every shared-memory fragment load a vector load:
//   pi    (contraction over the head dim, "K-form" B operand = row j of K / V / Q / dO):  slot (ks, t) <-> dim t*(HD/4) + 2ks,
//         slot (ks, t+4) <-> dim t*(HD/4) + 2ks + 1   => lane t reads the HD/4 contiguous floats [t*HD/4, (t+1)*HD/4) of row j;
//   sigma (output over the head dim, "V-form" B operand = rows 2t, 2t+1 of V / K / dO / Q):  column (nt, g) <-> dim g*(HD/8) + nt
//         => lane g reads the HD/8 contiguous floats [g*HD/8, (g+1)*HD/8) of each of its two rows, and the accumulator
//         fragment of lane t covers the contiguous dims [2t*HD/8, (2t+2)*HD/8) of rows g and g+8.
template <int HD>
__device__ __forceinline__ void lds_vec(float* dst, const float* src);       // HD/4 floats (K-form)
template <>
__device__ __forceinline__ void lds_vec<8>(float* dst, const float* src) { const float2 v = *reinterpret_cast<const float2*>(src); dst[0] = v.x; dst[1] = v.y; }
template <>
__device__ __forceinline__ void lds_vec<16>(float* dst, const float* src) { const float4 v = *reinterpret_cast<const float4*>(src); dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
template <int HD>
__device__ __forceinline__ void lds_half(float* dst, const float* src);      // HD/8 floats (V-form)
template <>
__device__ __forceinline__ void lds_half<8>(float* dst, const float* src) { dst[0] = *src; }
template <>
__device__ __forceinline__ void lds_half<16>(float* dst, const float* src) { const float2 v = *reinterpret_cast<const float2*>(src); dst[0] = v.x; dst[1] = v.y; }

// Lane (g, t)'s share of rows row0+g / row0+g+8 under pi: the HD/4 contiguous floats [t*HD/4, (t+1)*HD/4) of each row, straight from
// global memory; rows >= limit read as 0.
template <int HD>
__device__ __forceinline__ void load_rows_pi(float* lo, float* hi, const float* __restrict__ src, size_t ld, int col, int row0, int limit, int g, int t) {
    constexpr int W = HD / 4;
#pragma unroll
    for (int i = 0; i < W; ++i) lo[i] = hi[i] = 0.f;
    const int r_lo = row0 + g, r_hi = row0 + g + 8;
    const float* p = src + (size_t)r_lo * ld + col + t * W;
    if (r_lo < limit) {
        if constexpr (W == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); lo[0] = v.x; lo[1] = v.y; lo[2] = v.z; lo[3] = v.w; }
        else { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); lo[0] = v.x; lo[1] = v.y; }
    }
    if (r_hi < limit) {
        p += 8 * ld;
        if constexpr (W == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); hi[0] = v.x; hi[1] = v.y; hi[2] = v.z; hi[3] = v.w; }
        else { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); hi[0] = v.x; hi[1] = v.y; }
    }
}
// A-operand fragments under pi from a lane's row shares (load_rows_pi)
template <int HD>
__device__ __forceinline__ void afrag_from_rows(float (*a)[4], const float* lo, const float* hi, float mul) {
#pragma unroll
    for (int ks = 0; ks < HD / 8; ++ks) {
        a[ks][0] = tf32r(lo[2 * ks] * mul); a[ks][1] = tf32r(hi[2 * ks] * mul);
        a[ks][2] = tf32r(lo[2 * ks + 1] * mul); a[ks][3] = tf32r(hi[2 * ks + 1] * mul);
    }
}
// A-operand fragments under pi (16 rows starting at row0) straight from global memory; rows >= limit read as 0.
template <int HD>
__device__ __forceinline__ void load_afrag_pi(float (*a)[4], const float* __restrict__ src, size_t ld, int col, int row0, int limit, float mul, int g, int t) {
    constexpr int W = HD / 4;
    float lo[W], hi[W];
#pragma unroll
    for (int i = 0; i < W; ++i) lo[i] = hi[i] = 0.f;
    const int r_lo = row0 + g, r_hi = row0 + g + 8;
    if (r_lo < limit) {
        const float* p = src + (size_t)r_lo * ld + col + t * W;
        if constexpr (W == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); lo[0] = v.x; lo[1] = v.y; lo[2] = v.z; lo[3] = v.w; }
        else { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); lo[0] = v.x; lo[1] = v.y; }
    }
    if (r_hi < limit) {
        const float* p = src + (size_t)r_hi * ld + col + t * W;
        if constexpr (W == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); hi[0] = v.x; hi[1] = v.y; hi[2] = v.z; hi[3] = v.w; }
        else { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); hi[0] = v.x; hi[1] = v.y; }
    }
    afrag_from_rows<HD>(a, lo, hi, mul);
}
// accumulator tiles whose output columns are under sigma: lane t holds dims [2t*KS, 2t*KS + 2*KS) of rows g (c0,c1) and g+8 (c2,c3)
template <int HD>
__device__ __forceinline__ void store_sigma(float* __restrict__ p_lo, float* __restrict__ p_hi, const float (*acc)[4], float mul, bool w_lo, bool w_hi, int t) {
    constexpr int KS = HD / 8;
    if constexpr (KS == 2) {
        if (w_lo) *reinterpret_cast<float4*>(p_lo + 4 * t) = make_float4(acc[0][0] * mul, acc[1][0] * mul, acc[0][1] * mul, acc[1][1] * mul);
        if (w_hi) *reinterpret_cast<float4*>(p_hi + 4 * t) = make_float4(acc[0][2] * mul, acc[1][2] * mul, acc[0][3] * mul, acc[1][3] * mul);
    } else {
        if (w_lo) *reinterpret_cast<float2*>(p_lo + 2 * t) = make_float2(acc[0][0] * mul, acc[0][1] * mul);
        if (w_hi) *reinterpret_cast<float2*>(p_hi + 2 * t) = make_float2(acc[0][2] * mul, acc[0][3] * mul);
    }
}
// like load_slice, but rows up to the next multiple of 32 are zero-filled (branch-free 32-key blocks).  (The constant-stride walk of
// load_slice_skew was measured here too: it costs the head-dim-8 forward 8 registers and 12 % -- kept as the index loop.)
template <int HD, int LD = HD + 4>
__device__ __forceinline__ void load_slice32(float* dst, const float* __restrict__ src, size_t ld, int col, int row0, int cnt, float mul) {
    const int pad = (cnt + 31) & ~31;
    for (int idx = threadIdx.x; idx < pad * (HD / 4); idx += ATC_THREADS) {
        const int j = idx / (HD / 4), d = (idx % (HD / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < cnt) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(row0 + j) * ld + col + d));
        *reinterpret_cast<float4*>(dst + j * LD + d) = make_float4(tf32r(v.x * mul), tf32r(v.y * mul), tf32r(v.z * mul), tf32r(v.w * mul));
    }
}

// ---- asynchronous staging --------------------------------------------------------------------------------------------------
// Half of the forward's warp-time was spent before its first MMA (ncu source sampling: 45 % of the samples in the prologue, stalled on
// three dependent global round trips: cu -> Q rows -> K / V rows -> st.shared -> barrier).  K and V now go global -> shared with
// cp.async (no registers, no conversion) and are in flight together with the Q loads.  The raw fp32 bits are truncated to TF32 by the
// tensor core instead of being rounded here; the mean shrink of that truncation (TRUNC1) is folded into the Q scale (K) and the
// output scale (V), which leaves the same zero-mean error as rounding to nearest.
// rows [row0, row0+cnt) of a column slice -> smem [pad32(cnt)][LD], rows past cnt zero-filled (src-size 0)
template <int HD, int LD>
__device__ __forceinline__ void load_slice32_async(float* dst, const float* __restrict__ src, size_t ld, int col, int row0, int cnt) {
    const int pad = (cnt + 31) & ~31;
    for (int idx = threadIdx.x; idx < pad * (HD / 4); idx += ATC_THREADS) {
        const int j = idx / (HD / 4), d = (idx % (HD / 4)) * 4;
        const int jr = j < cnt ? j : cnt - 1;                          // keep the (unread) address in bounds
        cp_async16(dst + j * LD + d, src + (size_t)(row0 + jr) * ld + col + d, j < cnt ? 16 : 0);
    }
}

// One block of NT key tiles (8 keys each) of the streaming softmax, no per-tile guards: K/V rows past the sequence end are
// zero in shared memory and only the block that straddles the end (`tail`) masks its scores.
template <int HD, int NT>
__device__ __forceinline__ void fwd_block(const float* __restrict__ Ks, const float* __restrict__ Vs, int kb, int kn, const float (*qa)[4],
                                          float (*o)[4], float& m_lo, float& m_hi, float& l_lo, float& l_hi, int g, int t) {
    // K is only read in K-form (row g, HD/4 contiguous floats at t*HD/4): an unpadded row stride HD puts the lanes of one
    // 64- / 128-bit request on disjoint banks (HD+4 made rows 0 and 3 collide at HD = 8: one wavefront in three was a conflict);
    // V is only read in V-form (rows 2t, 2t+1, HD/8 floats at g*HD/8), conflict-free at HD+4.
    constexpr int LDK = HD, LD = HD + 4, KS = HD / 8, W = HD / 4;
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        float kf[W];
        lds_vec<HD>(kf, Ks + (kb + j * 8 + g) * LDK + t * W);
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma_tf32(s[j], qa[ks], kf[2 * ks], kf[2 * ks + 1]);
    }
    if (kb + NT * 8 > kn) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int kc = kb + j * 8 + 2 * t;
            if (kc >= kn) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (kc + 1 >= kn) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
        }
    }
    float bm_lo = fmaxf(s[0][0], s[0][1]), bm_hi = fmaxf(s[0][2], s[0][3]);
#pragma unroll
    for (int j = 1; j < NT; ++j) {
        bm_lo = fmaxf(bm_lo, fmaxf(s[j][0], s[j][1]));
        bm_hi = fmaxf(bm_hi, fmaxf(s[j][2], s[j][3]));
    }
    bm_lo = quad_max(bm_lo); bm_hi = quad_max(bm_hi);          // finite: key kb < kn is always live
    const float mn_lo = fmaxf(m_lo, bm_lo), mn_hi = fmaxf(m_hi, bm_hi);
    const float c_lo = ex2(m_lo - mn_lo), c_hi = ex2(m_hi - mn_hi);
    m_lo = mn_lo; m_hi = mn_hi;
    l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
    for (int i = 0; i < KS; ++i) { o[i][0] *= c_lo; o[i][1] *= c_lo; o[i][2] *= c_hi; o[i][3] *= c_hi; }
    // The softmax denominator rides on the tensor core: column HD of the V tile holds 1 for live keys (0 on padding), so one more
    // MMA per key tile accumulates l = sum_j p_j in an accumulator whose columns all carry the row sum (4 FADDs and the final quad
    // reduction less).  P goes in as raw fp32 bits: numerator and denominator see the SAME truncated probabilities, so the
    // truncation cancels in o / l and no rounding add is needed.
    float la[4] = {l_lo, l_lo, l_hi, l_hi};
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const float pa[4] = {ex2(s[j][0] - mn_lo), ex2(s[j][2] - mn_hi), ex2(s[j][1] - mn_lo), ex2(s[j][3] - mn_hi)};      // k=t <-> key 2t, k=t+4 <-> key 2t+1
        const float* vr = Vs + (kb + j * 8 + 2 * t) * LD;
        float v0[KS], v1[KS];
        lds_half<HD>(v0, vr + g * KS);
        lds_half<HD>(v1, vr + LD + g * KS);
#pragma unroll
        for (int nt = 0; nt < KS; ++nt) mma_tf32(o[nt], pa, v0[nt], v1[nt]);
        mma_tf32(la, pa, vr[HD], vr[LD + HD]);
    }
    l_lo = la[0]; l_hi = la[2];
}

template <int HD>
__global__ void __launch_bounds__(ATC_THREADS) attn_fwd_mma_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu,
                                                                   float* __restrict__ out, float* __restrict__ lse, int E, int H, float scale,
                                                                   const int32_t* __restrict__ order) {
    pdl_trigger();
    pdl_wait();                                  // launched as a programmatic dependent: nothing above touches global memory
    constexpr int LD = HD + 4, KS = HD / 8;
    __shared__ __align__(16) float Ks[CH * LD];
    __shared__ __align__(16) float Vs[CH * LD];
    const int b = order ? order[blockIdx.x / H] : blockIdx.x / H, h = blockIdx.x % H;      // order: longest sequences first (see seq_order_kernel)
    const int r0 = cu[b], n = cu[b + 1] - r0;
    if (n <= 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const size_t ld = 3 * (size_t)E;
    const float* base = qkv + (size_t)r0 * ld;
    const int nchunks = (n + CH - 1) / CH;
    const int nrounds = (n + 63) / 64;

    for (int rd = 0; rd < nrounds; ++rd) {
        const int q0 = rd * 64 + warp * 16;
        const bool active = q0 < n;
        float q_lo_[HD / 4], q_hi_[HD / 4];
        load_rows_pi<HD>(q_lo_, q_hi_, base, ld, h * HD, q0, n, g, t);      // issued first: their latency runs under the cp.async issue below
        if (nchunks == 1 && rd == 0) {                  // the usual case (n <= CH): K / V in flight while the Q rows are fetched
            load_slice32_async<HD, HD>(Ks, base, ld, E + h * HD, 0, n);
            load_slice32_async<HD, LD>(Vs, base, ld, 2 * E + h * HD, 0, n);
            cp_async_commit();
            for (int j = threadIdx.x; j < ((n + 31) & ~31); j += ATC_THREADS) Vs[j * LD + HD] = j < n ? 1.0f : 0.0f;      // the ones column (fwd_block)
        }
        float qa[KS][4];
        afrag_from_rows<HD>(qa, q_lo_, q_hi_, scale * LOG2E * TRUNC1);      // scores come out in log2 units; TRUNC1: K is truncated by the MMA
        float o[KS][4];
#pragma unroll
        for (int i = 0; i < KS; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

        for (int c = 0; c < nchunks; ++c) {
            const int k0c = c * CH, kn = min(CH, n - k0c);
            if (nchunks > 1) {
                __syncthreads();
                load_slice32_async<HD, HD>(Ks, base, ld, E + h * HD, k0c, kn);
                load_slice32_async<HD, LD>(Vs, base, ld, 2 * E + h * HD, k0c, kn);
                cp_async_commit();
                for (int j = threadIdx.x; j < ((kn + 31) & ~31); j += ATC_THREADS) Vs[j * LD + HD] = j < kn ? 1.0f : 0.0f;
                cp_async_wait_all();
                __syncthreads();
            } else if (rd == 0) {
                cp_async_wait_all();
                __syncthreads();
            }
            if (!active) continue;
            const int kpad = (kn + 31) & ~31;
            int kb = 0;
            for (; kb + 64 <= kpad; kb += 64) fwd_block<HD, 8>(Ks, Vs, kb, kn, qa, o, m_lo, m_hi, l_lo, l_hi, g, t);
            if (kb < kpad) fwd_block<HD, 4>(Ks, Vs, kb, kn, qa, o, m_lo, m_hi, l_lo, l_hi, g, t);
        }
        if (!active) continue;
        const float i_lo = TRUNC1 / l_lo, i_hi = TRUNC1 / l_hi;      // l comes out of the ones-column MMA already summed over the row; TRUNC1: V is truncated by the MMA
        const int q_lo = q0 + g, q_hi = q0 + g + 8;
        // accumulator columns under sigma: lane t holds dims [2t*KS, 2t*KS + 2*KS) of rows g (c0,c1) and g+8 (c2,c3)
        float* o_lo = out + (size_t)(r0 + q_lo) * E + h * HD + 2 * t * KS;
        float* o_hi = out + (size_t)(r0 + q_hi) * E + h * HD + 2 * t * KS;
        if constexpr (KS == 2) {
            if (q_lo < n) *reinterpret_cast<float4*>(o_lo) = make_float4(o[0][0] * i_lo, o[1][0] * i_lo, o[0][1] * i_lo, o[1][1] * i_lo);
            if (q_hi < n) *reinterpret_cast<float4*>(o_hi) = make_float4(o[0][2] * i_hi, o[1][2] * i_hi, o[0][3] * i_hi, o[1][3] * i_hi);
        } else {
            if (q_lo < n) *reinterpret_cast<float2*>(o_lo) = make_float2(o[0][0] * i_lo, o[0][1] * i_lo);
            if (q_hi < n) *reinterpret_cast<float2*>(o_hi) = make_float2(o[0][2] * i_hi, o[0][3] * i_hi);
        }
        if (t == 0) {
            if (q_lo < n) lse[(size_t)(r0 + q_lo) * H + h] = (m_lo + log2f(l_lo * TRUNC1)) * LN2;      // l was summed from truncated p: undo the mean shrink
            if (q_hi < n) lse[(size_t)(r0 + q_hi) * H + h] = (m_hi + log2f(l_hi * TRUNC1)) * LN2;
        }
    }
}

// ---- forward, two query tiles per warp ------------------------------------------------------------------------------
// Profiling the kernel above at the C4 light-curve shape (n ~ 120 tokens, head dim 8) showed that only 47 % of the executed
// instructions sit in the key-block body: the rest is per-round (Q fragments, normalisation, stores) and per-CTA work, and every
// K / V fragment load feeds a single 16-row tile.  Here a warp owns 32 query rows (two MMA row tiles): half the rounds, and each
// K / V fragment is loaded once for two tiles.
template <int HD, int NT>
__device__ __forceinline__ void fwd_block2(const float* __restrict__ Ks, const float* __restrict__ Vs, int kb, int kn, const float (*qa)[HD / 8][4],
                                           float (*o)[HD / 8][4], float* m_lo, float* m_hi, float* l_lo, float* l_hi, int g, int t) {
    constexpr int MQ = 2, LDK = HD, LD = HD + 4, KS = HD / 8, W = HD / 4;
    float s[MQ][NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        float kf[W];
        lds_vec<HD>(kf, Ks + (kb + j * 8 + g) * LDK + t * W);
#pragma unroll
        for (int q = 0; q < MQ; ++q) {
            s[q][j][0] = s[q][j][1] = s[q][j][2] = s[q][j][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) mma_tf32(s[q][j], qa[q][ks], kf[2 * ks], kf[2 * ks + 1]);
        }
    }
    if (kb + NT * 8 > kn) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int kc = kb + j * 8 + 2 * t;
#pragma unroll
            for (int q = 0; q < MQ; ++q) {
                if (kc >= kn) { s[q][j][0] = -INFINITY; s[q][j][2] = -INFINITY; }
                if (kc + 1 >= kn) { s[q][j][1] = -INFINITY; s[q][j][3] = -INFINITY; }
            }
        }
    }
    float mn_lo[MQ], mn_hi[MQ];
#pragma unroll
    for (int q = 0; q < MQ; ++q) {
        float bm_lo = fmaxf(s[q][0][0], s[q][0][1]), bm_hi = fmaxf(s[q][0][2], s[q][0][3]);
#pragma unroll
        for (int j = 1; j < NT; ++j) {
            bm_lo = fmaxf(bm_lo, fmaxf(s[q][j][0], s[q][j][1]));
            bm_hi = fmaxf(bm_hi, fmaxf(s[q][j][2], s[q][j][3]));
        }
        bm_lo = quad_max(bm_lo); bm_hi = quad_max(bm_hi);
        mn_lo[q] = fmaxf(m_lo[q], bm_lo); mn_hi[q] = fmaxf(m_hi[q], bm_hi);
        const float c_lo = ex2(m_lo[q] - mn_lo[q]), c_hi = ex2(m_hi[q] - mn_hi[q]);
        m_lo[q] = mn_lo[q]; m_hi[q] = mn_hi[q];
        l_lo[q] *= c_lo; l_hi[q] *= c_hi;
#pragma unroll
        for (int i = 0; i < KS; ++i) { o[q][i][0] *= c_lo; o[q][i][1] *= c_lo; o[q][i][2] *= c_hi; o[q][i][3] *= c_hi; }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        float v0[KS], v1[KS];
        lds_half<HD>(v0, Vs + (kb + j * 8 + 2 * t) * LD + g * KS);
        lds_half<HD>(v1, Vs + (kb + j * 8 + 2 * t + 1) * LD + g * KS);
#pragma unroll
        for (int q = 0; q < MQ; ++q) {
            const float p0 = ex2(s[q][j][0] - mn_lo[q]), p1 = ex2(s[q][j][1] - mn_lo[q]);
            const float p2 = ex2(s[q][j][2] - mn_hi[q]), p3 = ex2(s[q][j][3] - mn_hi[q]);
            l_lo[q] += p0 + p1; l_hi[q] += p2 + p3;
            const float pa[4] = {tf32r(p0), tf32r(p2), tf32r(p1), tf32r(p3)};
#pragma unroll
            for (int nt = 0; nt < KS; ++nt) mma_tf32(o[q][nt], pa, v0[nt], v1[nt]);
        }
    }
}

template <int HD>
__global__ void __launch_bounds__(ATC_THREADS) attn_fwd_mma2_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu,
                                                                    float* __restrict__ out, float* __restrict__ lse, int E, int H, float scale) {
    pdl_trigger();
    pdl_wait();                                  // launched as a programmatic dependent: nothing above touches global memory
    constexpr int LD = HD + 4, KS = HD / 8, MQ = 2, RQ = 64 * MQ;          // queries per round
    __shared__ __align__(16) float Ks[CH * HD];
    __shared__ __align__(16) float Vs[CH * LD];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int r0 = cu[b], n = cu[b + 1] - r0;
    if (n <= 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const size_t ld = 3 * (size_t)E;
    const float* base = qkv + (size_t)r0 * ld;
    const int nchunks = (n + CH - 1) / CH;
    const int nrounds = (n + RQ - 1) / RQ;
    for (int rd = 0; rd < nrounds; ++rd) {
        const int q0 = rd * RQ + warp * 16 * MQ;
        const bool active = q0 < n;
        float qa[MQ][KS][4], o[MQ][KS][4];
        float m_lo[MQ], m_hi[MQ], l_lo[MQ], l_hi[MQ];
#pragma unroll
        for (int q = 0; q < MQ; ++q) {
            load_afrag_pi<HD>(qa[q], base, ld, h * HD, q0 + 16 * q, n, scale * LOG2E, g, t);
#pragma unroll
            for (int i = 0; i < KS; ++i) o[q][i][0] = o[q][i][1] = o[q][i][2] = o[q][i][3] = 0.f;
            m_lo[q] = m_hi[q] = -INFINITY; l_lo[q] = l_hi[q] = 0.f;
        }
        for (int c = 0; c < nchunks; ++c) {
            const int k0c = c * CH, kn = min(CH, n - k0c);
            if (nchunks > 1 || rd == 0) {
                __syncthreads();
                load_slice32<HD, HD>(Ks, base, ld, E + h * HD, k0c, kn, 1.0f);
                load_slice32<HD>(Vs, base, ld, 2 * E + h * HD, k0c, kn, 1.0f);
                __syncthreads();
            }
            if (!active) continue;
            const int kpad = (kn + 31) & ~31;
            for (int kb = 0; kb < kpad; kb += 32) fwd_block2<HD, 4>(Ks, Vs, kb, kn, qa, o, m_lo, m_hi, l_lo, l_hi, g, t);
        }
        if (!active) continue;
#pragma unroll
        for (int q = 0; q < MQ; ++q) {
            const float ll = quad_sum(l_lo[q]), lh = quad_sum(l_hi[q]);
            const float i_lo = 1.0f / ll, i_hi = 1.0f / lh;
            const int q_lo = q0 + 16 * q + g, q_hi = q_lo + 8;
            float* o_lo = out + (size_t)(r0 + q_lo) * E + h * HD + 2 * t * KS;
            float* o_hi = out + (size_t)(r0 + q_hi) * E + h * HD + 2 * t * KS;
            if constexpr (KS == 2) {
                if (q_lo < n) *reinterpret_cast<float4*>(o_lo) = make_float4(o[q][0][0] * i_lo, o[q][1][0] * i_lo, o[q][0][1] * i_lo, o[q][1][1] * i_lo);
                if (q_hi < n) *reinterpret_cast<float4*>(o_hi) = make_float4(o[q][0][2] * i_hi, o[q][1][2] * i_hi, o[q][0][3] * i_hi, o[q][1][3] * i_hi);
            } else {
                if (q_lo < n) *reinterpret_cast<float2*>(o_lo) = make_float2(o[q][0][0] * i_lo, o[q][0][1] * i_lo);
                if (q_hi < n) *reinterpret_cast<float2*>(o_hi) = make_float2(o[q][0][2] * i_hi, o[q][0][3] * i_hi);
            }
            if (t == 0) {
                if (q_lo < n) lse[(size_t)(r0 + q_lo) * H + h] = (m_lo[q] + log2f(ll)) * LN2;
                if (q_hi < n) lse[(size_t)(r0 + q_hi) * H + h] = (m_hi[q] + log2f(lh)) * LN2;
            }
        }
    }
}

// head dim 16: six CTAs per SM (80 registers, 8 bytes of spill) measured 137.3 vs 139.3 us; head dim 8 keeps the natural allocation
template <int HD>
__global__ void __launch_bounds__(ATC_THREADS, HD == 16 ? 6 : 0) attn_bwd_mma_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu,
                                                                   const float* __restrict__ out, const float* __restrict__ lse,
                                                                   const float* __restrict__ dout, float* __restrict__ dqkv,
                                                                   int E, int H, float scale, const int32_t* __restrict__ order) {
    pdl_trigger();
    pdl_wait();                                  // launched as a programmatic dependent: nothing above touches global memory
    __shared__ __align__(16) float As[skew_floats<HD>(CH)];     // phase 1: K        phase 2: Q      (skewed rows: rowoff<HD>)
    __shared__ __align__(16) float Bs[skew_floats<HD>(CH)];     // phase 1: V        phase 2: dO
    __shared__ __align__(8) float lse_s[CH];        // phase 2: -lse_i * log2e (-inf on padding rows)
    __shared__ __align__(8) float D_s[CH];                       // phase 2: -D_i = -dO_i . O_i
    const int b = order ? order[blockIdx.x / H] : blockIdx.x / H, h = blockIdx.x % H;
    const int r0 = cu[b], n = cu[b + 1] - r0;
    if (n <= 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const size_t ld = 3 * (size_t)E;
    const float* base = qkv + (size_t)r0 * ld;
    const float* obase = out + (size_t)r0 * E + h * HD;
    const float* gbase = dout + (size_t)r0 * E + h * HD;
    const int nchunks = (n + CH - 1) / CH;
    const int nrounds = (n + 63) / 64;
    // lane-constant fragment addresses (tile 0): K-form = row g, HD/4 floats at t*HD/4; V-form = rows 2t / 2t+1, HD/8 floats at g*HD/8
    constexpr int TS = skew_floats<HD>(8);                          // one 8-row tile further
    const float* const kform_a = As + rowoff<HD>(g) + t * (HD / 4);
    const float* const kform_b = Bs + rowoff<HD>(g) + t * (HD / 4);
    const float* const vform_a0 = As + rowoff<HD>(2 * t) + g * (HD / 8);
    const float* const vform_a1 = As + rowoff<HD>(2 * t + 1) + g * (HD / 8);
    const float* const vform_b0 = Bs + rowoff<HD>(2 * t) + g * (HD / 8);
    const float* const vform_b1 = Bs + rowoff<HD>(2 * t + 1) + g * (HD / 8);

    // ---- phase 1: dQ_i = scale * sum_j P_ij (dO_i.V_j - D_i) K_j ; warp owns 16 queries -------------------------
    for (int rd = 0; rd < nrounds; ++rd) {
        const int q0 = rd * 64 + warp * 16;
        const bool active = q0 < n;
        const int q_lo = q0 + g, q_hi = q0 + g + 8;
        if (rd == 0 && nchunks == 1) {                            // the usual case (n <= CH): K / V in flight while this warp's rows are fetched
            load_slice_skew_async<HD>(As, base, ld, E + h * HD, 0, n);
            load_slice_skew_async<HD>(Bs, base, ld, 2 * E + h * HD, 0, n);
            cp_async_commit();
        }
        float qa[HD / 8][4], ga[HD / 8][4];                       // contraction slots relabelled by pi (see the forward)
        load_afrag_pi<HD>(qa, base, ld, h * HD, q0, n, scale * LOG2E * TRUNC1, g, t);      // TRUNC1: K / V arrive as raw bits and are truncated by the MMA
        // dO rows once: the A fragments of dP = dO V^T and, against the matching share of O, D = dO . O (each lane of the quad holds
        // HD/4 of the dims)
        float D_lo = 0.f, D_hi = 0.f, L_lo = INFINITY, L_hi = INFINITY;
        {
            float g_lo[HD / 4], g_hi[HD / 4], o_lo[HD / 4], o_hi[HD / 4];
            load_rows_pi<HD>(g_lo, g_hi, gbase, (size_t)E, 0, q0, n, g, t);
            load_rows_pi<HD>(o_lo, o_hi, obase, (size_t)E, 0, q0, n, g, t);
            afrag_from_rows<HD>(ga, g_lo, g_hi, TRUNC1);
#pragma unroll
            for (int d = 0; d < HD / 4; ++d) { D_lo = fmaf(g_lo[d], o_lo[d], D_lo); D_hi = fmaf(g_hi[d], o_hi[d], D_hi); }
        }
        D_lo = quad_sum(D_lo); D_hi = quad_sum(D_hi);
        if (q_lo < n) L_lo = __ldg(lse + (size_t)(r0 + q_lo) * H + h) * LOG2E;
        if (q_hi < n) L_hi = __ldg(lse + (size_t)(r0 + q_hi) * H + h) * LOG2E;
        float dq[HD / 8][4];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

        for (int c = 0; c < nchunks; ++c) {
            const int k0c = c * CH, kn = min(CH, n - k0c);
            if (nchunks > 1) {
                __syncthreads();
                load_slice_skew_async<HD>(As, base, ld, E + h * HD, k0c, kn);
                load_slice_skew_async<HD>(Bs, base, ld, 2 * E + h * HD, k0c, kn);
                cp_async_commit();
                cp_async_wait_all();
                __syncthreads();
            } else if (rd == 0) {
                cp_async_wait_all();
                __syncthreads();
            }
            if (!active) continue;
            const int ntile = (kn + 7) >> 3;
            // rowoff is linear over whole 8-row tiles, so every fragment address is (lane constant) + tile * TS: the unrolled loop
            // addresses shared memory with immediates (the index arithmetic was half of this loop's instructions)
            const float* pk = kform_a;
            const float* pv = kform_b;
            const float* pk0 = vform_a0;
            const float* pk1 = vform_a1;
#pragma unroll 4
            for (int j = 0; j < ntile; ++j, pk += TS, pv += TS, pk0 += TS, pk1 += TS) {
                // The log-sum-exp and D of the tile's two rows are known up front, so they ride in the accumulators' initial values:
                // the MMAs deliver s - L and dP - D directly (8 FADDs per tile less).  dS feeds the next MMA as raw fp32 bits (the
                // tensor core truncates them to TF32); the mean shrink of that truncation is folded into the store scale below.
                float s[4] = {-L_lo, -L_lo, -L_hi, -L_hi}, dp[4] = {-D_lo, -D_lo, -D_hi, -D_hi};
                float kf[HD / 4], vf[HD / 4];
                lds_vec<HD>(kf, pk);
                lds_vec<HD>(vf, pv);
#pragma unroll
                for (int ks = 0; ks < HD / 8; ++ks) {
                    mma_tf32(s, qa[ks], kf[2 * ks], kf[2 * ks + 1]);
                    mma_tf32(dp, ga[ks], vf[2 * ks], vf[2 * ks + 1]);
                }
                // no key mask needed: rows of K past the sequence end are zero-filled in shared memory, so whatever (finite) dS
                // they get multiplies a zero row in the dQ MMA below
                const float da[4] = {ex2(s[0]) * dp[0], ex2(s[2]) * dp[2], ex2(s[1]) * dp[1], ex2(s[3]) * dp[3]};
                float k0[HD / 8], k1[HD / 8];                      // output columns relabelled by sigma
                lds_half<HD>(k0, pk0);
                lds_half<HD>(k1, pk1);
#pragma unroll
                for (int nt = 0; nt < HD / 8; ++nt) mma_tf32(dq[nt], da, k0[nt], k1[nt]);
            }
        }
        if (!active) continue;
        store_sigma<HD>(dqkv + (size_t)(r0 + q_lo) * ld + h * HD, dqkv + (size_t)(r0 + q_hi) * ld + h * HD, dq, scale * TRUNC1 * TRUNC1, q_lo < n, q_hi < n, t);      // dS and K both truncated
    }

    // ---- phase 2: dV_j = sum_i P_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i ; warp owns 16 keys -----------------
    for (int rd = 0; rd < nrounds; ++rd) {
        const int k0 = rd * 64 + warp * 16;
        const bool active = k0 < n;
        float ka[HD / 8][4], va[HD / 8][4];
        load_afrag_pi<HD>(ka, base, ld, E + h * HD, k0, n, scale * LOG2E * TRUNC1, g, t);      // TRUNC1: Q / dO arrive as raw bits
        load_afrag_pi<HD>(va, base, ld, 2 * E + h * HD, k0, n, TRUNC1, g, t);
        float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }

        for (int c = 0; c < nchunks; ++c) {
            const int q0c = c * CH, qn = min(CH, n - q0c);
            if (nchunks > 1 || rd == 0) {
                __syncthreads();                    // also orders phase-1 readers of As/Bs before the overwrite
                load_slice_skew_async<HD>(As, base, ld, h * HD, q0c, qn);
                load_slice_skew_async<HD>(Bs, gbase, (size_t)E, 0, q0c, qn);
                cp_async_commit();
                const int pad = (qn + 7) & ~7;
                for (int i = threadIdx.x; i < pad; i += ATC_THREADS) {
                    float Di = 0.f, Li = INFINITY;                      // stored negated: accumulator initial values of the tile loop
                    if (i < qn) {
#pragma unroll
                        for (int d = 0; d < HD; d += 4) {
                            const float4 gv = __ldg(reinterpret_cast<const float4*>(gbase + (size_t)(q0c + i) * E + d));
                            const float4 ov = __ldg(reinterpret_cast<const float4*>(obase + (size_t)(q0c + i) * E + d));
                            Di = fmaf(gv.x, ov.x, fmaf(gv.y, ov.y, fmaf(gv.z, ov.z, fmaf(gv.w, ov.w, Di))));
                        }
                        Li = __ldg(lse + (size_t)(r0 + q0c + i) * H + h) * LOG2E;
                    }
                    D_s[i] = -Di; lse_s[i] = -Li;
                }
                cp_async_wait_all();
                __syncthreads();
            }
            if (!active) continue;
            const int ntile = (qn + 7) >> 3;
            const float* pq = kform_a;
            const float* pg = kform_b;
            const float* pq0 = vform_a0;
            const float* pq1 = vform_a1;
            const float* pg0 = vform_b0;
            const float* pg1 = vform_b1;
            const float* pl = lse_s + 2 * t;
            const float* pd = D_s + 2 * t;
#pragma unroll 4
            for (int j = 0; j < ntile; ++j, pq += TS, pg += TS, pq0 += TS, pq1 += TS, pg0 += TS, pg1 += TS, pl += 8, pd += 8) {
                // transposed tiles: rows = keys, cols = queries; -lse and -D of the tile's two query columns (stored negated) are the
                // accumulators' initial values
                const float2 L01 = *reinterpret_cast<const float2*>(pl), D01 = *reinterpret_cast<const float2*>(pd);
                float s[4] = {L01.x, L01.y, L01.x, L01.y}, dp[4] = {D01.x, D01.y, D01.x, D01.y};
                float qf[HD / 4], gf[HD / 4];
                lds_vec<HD>(qf, pq);
                lds_vec<HD>(gf, pg);
#pragma unroll
                for (int ks = 0; ks < HD / 8; ++ks) {
                    mma_tf32(s, ka[ks], qf[2 * ks], qf[2 * ks + 1]);
                    mma_tf32(dp, va[ks], gf[2 * ks], gf[2 * ks + 1]);
                }
                const float p0 = ex2(s[0]), p1 = ex2(s[1]), p2 = ex2(s[2]), p3 = ex2(s[3]);          // 0 on padding (-lse = -inf)
                const float pa[4] = {p0, p2, p1, p3};                                                 // raw fp32 bits: truncated by the MMA,
                const float da[4] = {p0 * dp[0], p2 * dp[2], p1 * dp[1], p3 * dp[3]};                 // compensated in the store scale
                float g0[HD / 8], g1[HD / 8], q0v[HD / 8], q1v[HD / 8];
                lds_half<HD>(g0, pg0);
                lds_half<HD>(g1, pg1);
                lds_half<HD>(q0v, pq0);
                lds_half<HD>(q1v, pq1);
#pragma unroll
                for (int nt = 0; nt < HD / 8; ++nt) {
                    mma_tf32(dv[nt], pa, g0[nt], g1[nt]);
                    mma_tf32(dk[nt], da, q0v[nt], q1v[nt]);
                }
            }
        }
        if (!active) continue;
        const int k_lo = k0 + g, k_hi = k0 + g + 8;
        float* row_lo = dqkv + (size_t)(r0 + k_lo) * ld + h * HD;
        float* row_hi = dqkv + (size_t)(r0 + k_hi) * ld + h * HD;
        store_sigma<HD>(row_lo + E, row_hi + E, dk, scale * TRUNC1 * TRUNC1, k_lo < n, k_hi < n, t);       // dS / P and Q / dO all truncated
        store_sigma<HD>(row_lo + 2 * E, row_hi + 2 * E, dv, TRUNC1 * TRUNC1, k_lo < n, k_hi < n, t);
    }
}

// order[r] = index of the sequence with the r-th largest token count (ties by index): thread b counts the sequences ahead of it.
// CTAs are dispatched in blockIdx order, so (sequence, head) work items sorted longest-first fill the last wave with the SHORTEST
// items -- with ~2.5 waves of very unequal CTAs (spectra: 110..220 tokens, cost ~ n^2) a random order left a tail of up to one long
// CTA on every launch.
__global__ void __launch_bounds__(256) seq_order_kernel(const int32_t* __restrict__ cu, int B, int32_t* __restrict__ order) {
    extern __shared__ int32_t len_s[];
    for (int i = threadIdx.x; i < B; i += 256) len_s[i] = cu[i + 1] - cu[i];
    __syncthreads();
    const int b = blockIdx.x * 256 + threadIdx.x;
    if (b >= B) return;
    const int nb = len_s[b];
    int rank = 0;
    for (int i = 0; i < B; ++i) rank += (len_s[i] > nb) || (len_s[i] == nb && i < b);
    order[rank] = b;
}
const int32_t* g_order = nullptr;       // set by the encoder around its attention launches (host-side, one training thread per device)

}  // namespace

void set_attention_order(const int32_t* order) { g_order = order; }
int launch_seq_order(const int32_t* cu, int B, int32_t* order, cudaStream_t st) {
    if ((size_t)B * 4 > 160 * 1024) return MVN_E_UNSUPPORTED;
    static bool configured = false;
    if (!configured) { MVN_CUDA(cudaFuncSetAttribute(seq_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); configured = true; }
    seq_order_kernel<<<cdiv(B, 256), 256, (size_t)B * 4, st>>>(cu, B, order);
    MVN_LAUNCH_CHECK();
    return 0;
}

// Returns MVN_E_UNSUPPORTED for head dims the MMA kernels are not built for (caller falls back to the fp32 kernel).
int launch_attention_fwd_tc(const float* qkv, const int32_t* cu, float* out, float* lse, int B, int E, int H, float scale, cudaStream_t st) {
    const int hd = E / H;
    // MVN_ATTN_FWD=2: two query tiles per warp.  Measured on B200 at the C4 shapes: 108.6 vs 106.5 us (light curve), 73.7 vs 73.7 us
    // (spectra) -- no gain: the kernel is instruction-issue bound (61 % of issue slots, HMMA is 3.8 % of the instructions) and the
    // per-query-tile work (Q fragments, normalisation, stores) is the same in both, so the default stays one tile per warp.
    static const int variant = getenv("MVN_ATTN_FWD") ? atoi(getenv("MVN_ATTN_FWD")) : 1;
    if (variant == 2 && hd == 8) attn_fwd_mma2_kernel<8><<<B * H, ATC_THREADS, 0, st>>>(qkv, cu, out, lse, E, H, scale);
    else if (variant == 2 && hd == 16) attn_fwd_mma2_kernel<16><<<B * H, ATC_THREADS, 0, st>>>(qkv, cu, out, lse, E, H, scale);
    else if (hd == 8) MVN_CUDA(launch_dependent(attn_fwd_mma_kernel<8>, dim3(B * H), dim3(ATC_THREADS), 0, st, qkv, cu, out, lse, E, H, scale, g_order));
    else if (hd == 16) MVN_CUDA(launch_dependent(attn_fwd_mma_kernel<16>, dim3(B * H), dim3(ATC_THREADS), 0, st, qkv, cu, out, lse, E, H, scale, g_order));
    else return MVN_E_UNSUPPORTED;
    MVN_LAUNCH_CHECK();
    return 0;
}
int launch_attention_bwd_tc(const float* qkv, const int32_t* cu, const float* out, const float* lse, const float* dout, float* dqkv,
                            int B, int E, int H, float scale, cudaStream_t st) {
    const int hd = E / H;
    if (hd == 8) MVN_CUDA(launch_dependent(attn_bwd_mma_kernel<8>, dim3(B * H), dim3(ATC_THREADS), 0, st, qkv, cu, out, lse, dout, dqkv, E, H, scale, g_order));
    else if (hd == 16) MVN_CUDA(launch_dependent(attn_bwd_mma_kernel<16>, dim3(B * H), dim3(ATC_THREADS), 0, st, qkv, cu, out, lse, dout, dqkv, E, H, scale, g_order));
    else return MVN_E_UNSUPPORTED;
    MVN_LAUNCH_CHECK();
    return 0;
}

}  // namespace mvn
