// Fused feed-forward half of a TransformerBlock (src/transformer_utils.py:101-105, 113-115), prec == 2 tier:
//
//   forward :  x2 = dropout( LayerNorm( x1 + W2 relu(W1 x1 + b1) + b2 ) * gamma + beta )
//   backward:  LayerNorm backward, both input-gradient GEMMs, both weight-gradient GEMMs, all bias / affine gradients
//
// in ONE kernel each.  The hidden activation h [tokens, 4E] never reaches HBM: the forward keeps each 8-column chunk
// of h in accumulator registers and feeds it straight back as the A operand of the second GEMM; the backward
// recomputes h from x1 and keeps h, dh, dz and x1 of a token tile in shared memory, where the weight-gradient
// contraction (over the tile's tokens) reads them.  Per token-layer the pair moves 3E + 4E floats instead of the
// 12E + 28E of the layer-by-layer kernels (gemm_tc.cu + rowops.cu).
//
// Contractions: warp-level mma.sync m16n8k8, TF32 operands (round-to-nearest for weights, x1 and h; activations that
// only feed gradient GEMMs are truncated by the MMA), fp32 accumulate.  Why not tcgen05 here: the backward needs h, dh,
// dz and x1 each as a K-major AND an MN-major operand plus three weight layouts; at E = 64 that is > 227 KB of shared
// memory and > 512 TMEM columns per SM, while register fragments read any layout from one copy.
//
// Fragment layouts (g = lane / 4, t = lane % 4):
//   A(16x8): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B(8x8): b0 (k=t, n=g) b1 (k=t+4, n=g);
//   C(16x8): c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
// An accumulator tile is reused as an A operand without shuffles by relabelling the contraction index
// (slot t <-> column 2t, slot t+4 <-> column 2t+1): a = {c0, c2, c1, c3}, and the B side reads rows 2t, 2t+1.
#include <stdlib.h>

#include "common.cuh"

namespace mvn {
namespace {

constexpr int FF_WARPS = 16;
constexpr int FF_THREADS = FF_WARPS * 32;

__device__ __forceinline__ float tf32r(float x) {            // round to nearest TF32 (ties away), one integer add
    return __uint_as_float(__float_as_uint(x) + 0x1000u);      // the MMA ignores the low 13 mantissa bits
}
__device__ __forceinline__ void mma_tf32(float* c, const float* a, float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Shared-memory geometry.  Row strides are chosen so that every fragment load is bank-conflict free:
//   W1s [F][E+4]  : read as B with rows = n = g, cols = k = t  (4g + t)   and with rows = k = 2t, cols = n = g  (8t + g)
//   W2s [E][F+8]  : read as B with rows = n = g, 64-bit at col 2t (8g + 2t per half warp) and with rows = k = t, cols = g (8t + g)
template <int E, int THREADS = 512>
struct FfnGeom {
    static constexpr int F = 4 * E;
    static constexpr int P1 = E + 4;
    static constexpr int P2 = F + 8;
    static constexpr int W_FLOATS = F * P1 + E * P2;
    static constexpr int VEC_FLOATS = F + 3 * E;                    // b1 | b2 | gamma | beta
    // backward token tile
    static constexpr int TT = THREADS * 4 / E;                      // tokens per tile: one float4 per thread in the LayerNorm-backward phase
    static constexpr int PX = E + 8;                                // dz / x1 tiles: B-fragment reads (rows = t): 8t + g
    static constexpr int PH = F + 8;                                // h tile: A-fragment reads in the weight phase (rows = t)
    static constexpr int PD = F + 8;                                // dh tile: 64-bit A-fragment reads of the dx GEMM (rows = g, col 2t) and rows = t reads
    static constexpr int TILE_FLOATS = 2 * TT * PX + TT * PH + TT * PD;
    static constexpr size_t SMEM_FWD = (size_t)(W_FLOATS + VEC_FLOATS) * 4;
    static constexpr size_t SMEM_BWD = (size_t)(W_FLOATS + VEC_FLOATS + TILE_FLOATS) * 4;
};

template <int E, int NTHR>
__device__ __forceinline__ void stage_weights(float* W1s, float* W2s, float* vec, const float* __restrict__ W1, const float* __restrict__ b1,
                                              const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ gamma,
                                              const float* __restrict__ beta) {
    using G = FfnGeom<E>;
    constexpr int F = G::F;
    for (int i = threadIdx.x; i < F * E / 4; i += NTHR) {
        const int f = i / (E / 4), e = (i % (E / 4)) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(W1) + i);
        *reinterpret_cast<float4*>(W1s + f * G::P1 + e) = make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w));
    }
    for (int i = threadIdx.x; i < E * F / 4; i += NTHR) {
        const int e = i / (F / 4), f = (i % (F / 4)) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(W2) + i);
        *reinterpret_cast<float4*>(W2s + e * G::P2 + f) = make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w));
    }
    for (int i = threadIdx.x; i < F; i += NTHR) vec[i] = b1 ? b1[i] : 0.f;
    for (int i = threadIdx.x; i < E; i += NTHR) {
        vec[F + i] = b2 ? b2[i] : 0.f;
        vec[F + E + i] = gamma ? gamma[i] : 1.f;
        vec[F + 2 * E + i] = beta ? beta[i] : 0.f;
    }
}

struct FfnFwdArgs {
    const float *X, *W1, *b1, *W2, *b2, *gamma, *beta;
    float *Y, *xhat, *rstd;
    const int32_t* n_rows_dev;
    int M_cap;
    float eps;
    DropCfg drop;
};

// Every warp owns blocks of 16 * MT rows of the packed token stream (no cross-warp dependency after the weights are staged);
// MT = 2 halves the shared-memory traffic per MMA (each weight fragment feeds two row tiles).
template <int E, int MT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) ffn_fwd_kernel(const FfnFwdArgs a) {
    using G = FfnGeom<E>;
    constexpr int F = G::F, KS = E / 8, NT = E / 8, NC = F / 8, P1 = G::P1, P2 = G::P2, RPB = 16 * MT;
    extern __shared__ __align__(16) float smem_f[];
    float* W1s = smem_f;
    float* W2s = W1s + F * P1;
    float* vec = W2s + E * P2;
    pdl_trigger();
    stage_weights<E, WARPS * 32>(W1s, W2s, vec, a.W1, a.b1, a.W2, a.b2, a.gamma, a.beta);
    __syncthreads();
    pdl_wait();                                              // weights are parameters; the token stream is the predecessor's output
    const int rows = a.n_rows_dev ? min(__ldg(a.n_rows_dev), a.M_cap) : a.M_cap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nblk = (rows + RPB - 1) / RPB;
    const float* b1s = vec; const float* b2s = vec + F; const float* gs = vec + F + E; const float* bs = vec + F + 2 * E;
    const float inv_e = 1.0f / (float)E;

    // block -> (CTA-minor) warp: the blocks of the last, partial round are spread over all SMs instead of filling a few CTAs
    for (int blk = warp * gridDim.x + blockIdx.x; blk < nblk; blk += gridDim.x * WARPS) {
        float xa[MT][KS][4];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int r_lo = blk * RPB + 16 * m + g, r_hi = r_lo + 8;
            const bool v_lo = r_lo < rows, v_hi = r_hi < rows;
            const float* x_lo = a.X + (size_t)r_lo * E;
            const float* x_hi = a.X + (size_t)r_hi * E;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                xa[m][ks][0] = v_lo ? tf32r(__ldg(x_lo + 8 * ks + t)) : 0.f;
                xa[m][ks][1] = v_hi ? tf32r(__ldg(x_hi + 8 * ks + t)) : 0.f;
                xa[m][ks][2] = v_lo ? tf32r(__ldg(x_lo + 8 * ks + t + 4)) : 0.f;
                xa[m][ks][3] = v_hi ? tf32r(__ldg(x_hi + 8 * ks + t + 4)) : 0.f;
            }
        }
        float o[MT][NT][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) o[m][nt][0] = o[m][nt][1] = o[m][nt][2] = o[m][nt][3] = 0.f;
#pragma unroll 2
        for (int c = 0; c < NC; ++c) {
            float hc[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                hc[m][0] = hc[m][2] = b1s[8 * c + 2 * t];
                hc[m][1] = hc[m][3] = b1s[8 * c + 2 * t + 1];
            }
            const float* w1p = W1s + (8 * c + g) * P1 + t;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const float b0 = w1p[8 * ks], b1v = w1p[8 * ks + 4];
#pragma unroll
                for (int m = 0; m < MT; ++m) mma_tf32(hc[m], xa[m][ks], b0, b1v);
            }
            float ha[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                ha[m][0] = tf32r(fmaxf(hc[m][0], 0.f)); ha[m][1] = tf32r(fmaxf(hc[m][2], 0.f));
                ha[m][2] = tf32r(fmaxf(hc[m][1], 0.f)); ha[m][3] = tf32r(fmaxf(hc[m][3], 0.f));
            }
            const float* w2p = W2s + g * P2 + 8 * c + 2 * t;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 w = *reinterpret_cast<const float2*>(w2p + 8 * nt * P2);
#pragma unroll
                for (int m = 0; m < MT; ++m) mma_tf32(o[m][nt], ha[m], w.x, w.y);
            }
        }
        // z = acc + b2 + x1 (exact fp32 residual, re-read: the block's rows are L1-resident), then LayerNorm over E
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int r_lo = blk * RPB + 16 * m + g, r_hi = r_lo + 8;
            const bool v_lo = r_lo < rows, v_hi = r_hi < rows;
            const float* x_lo = a.X + (size_t)r_lo * E;
            const float* x_hi = a.X + (size_t)r_hi * E;
            float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = 8 * nt + 2 * t;
                const float2 rl = v_lo ? __ldg(reinterpret_cast<const float2*>(x_lo + col)) : make_float2(0.f, 0.f);
                const float2 rh = v_hi ? __ldg(reinterpret_cast<const float2*>(x_hi + col)) : make_float2(0.f, 0.f);
                o[m][nt][0] += b2s[col] + rl.x; o[m][nt][1] += b2s[col + 1] + rl.y;
                o[m][nt][2] += b2s[col] + rh.x; o[m][nt][3] += b2s[col + 1] + rh.y;
                s_lo += o[m][nt][0] + o[m][nt][1]; s_hi += o[m][nt][2] + o[m][nt][3];
            }
            const float m_lo = quad_sum(s_lo) * inv_e, m_hi = quad_sum(s_hi) * inv_e;
            float q_lo = 0.f, q_hi = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                o[m][nt][0] -= m_lo; o[m][nt][1] -= m_lo; o[m][nt][2] -= m_hi; o[m][nt][3] -= m_hi;
                q_lo = fmaf(o[m][nt][0], o[m][nt][0], fmaf(o[m][nt][1], o[m][nt][1], q_lo));
                q_hi = fmaf(o[m][nt][2], o[m][nt][2], fmaf(o[m][nt][3], o[m][nt][3], q_hi));
            }
            const float rs_lo = rsqrtf(quad_sum(q_lo) * inv_e + a.eps), rs_hi = rsqrtf(quad_sum(q_hi) * inv_e + a.eps);
            if (a.rstd && t == 0) {
                if (v_lo) a.rstd[r_lo] = rs_lo;
                if (v_hi) a.rstd[r_hi] = rs_hi;
            }
            uint32_t rk_lo = 0, rk_hi = 0;
            if (a.drop.thresh) { rk_lo = drop_rowkey(a.drop, (uint32_t)r_lo); rk_hi = drop_rowkey(a.drop, (uint32_t)r_hi); }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = 8 * nt + 2 * t;
                const float h0 = o[m][nt][0] * rs_lo, h1 = o[m][nt][1] * rs_lo, h2 = o[m][nt][2] * rs_hi, h3 = o[m][nt][3] * rs_hi;
                float y0 = fmaf(h0, gs[col], bs[col]), y1 = fmaf(h1, gs[col + 1], bs[col + 1]);
                float y2 = fmaf(h2, gs[col], bs[col]), y3 = fmaf(h3, gs[col + 1], bs[col + 1]);
                if (a.drop.thresh) {
                    y0 *= drop_scale(a.drop, rk_lo, (uint32_t)col); y1 *= drop_scale(a.drop, rk_lo, (uint32_t)col + 1);
                    y2 *= drop_scale(a.drop, rk_hi, (uint32_t)col); y3 *= drop_scale(a.drop, rk_hi, (uint32_t)col + 1);
                }
                if (v_lo) {
                    *reinterpret_cast<float2*>(a.Y + (size_t)r_lo * E + col) = make_float2(y0, y1);
                    if (a.xhat) *reinterpret_cast<float2*>(a.xhat + (size_t)r_lo * E + col) = make_float2(h0, h1);
                }
                if (v_hi) {
                    *reinterpret_cast<float2*>(a.Y + (size_t)r_hi * E + col) = make_float2(y2, y3);
                    if (a.xhat) *reinterpret_cast<float2*>(a.xhat + (size_t)r_hi * E + col) = make_float2(h2, h3);
                }
            }
        }
    }
}

struct FfnBwdArgs {
    const float *dY, *xhat, *rstd, *X, *W1, *b1, *W2, *gamma;
    float* dX;
    const int32_t* n_rows_dev;
    int M_cap;
    DropCfg drop;
    float* partial; size_t pstride;                 // slab s = blockIdx.x at partial + s * pstride
    float* partial2; size_t pstride2;               // CTAs >= kSlabs (two-CTA-per-SM launch): FFN-only slabs, offsets relative to o_w1
    size_t o_w1, o_b1, o_w2, o_b2, o_g, o_b;        // float offsets inside a slab
};

// One CTA = 16 warps, persistent over token tiles of TT rows.  Per tile:
//   P0  all threads : LayerNorm backward of the tile's rows -> dz (exact fp32) and x1 into shared memory; per-thread partial
//                     column sums for dgamma, dbeta, db2 (a thread owns the same 4 columns in every tile).
//   PA  warp (rb, fr): recompute h = relu(x1 W1^T + b1) and dh = (dz W2) * (h > 0) for 16 rows x (F / FR) hidden columns -> smem.
//   PB  warp w      : (i) one 16 x 8 tile of dx1 = dz + dh W1 (contraction over F) -> global;
//                     (ii) its 16 hidden rows of dW2^T += h^T dz and dW1 += dh^T x1 (contraction over the tile's tokens),
//                     accumulators live in registers for the CTA's whole token range; db1 comes from the A fragments.
// The CTA finally writes one slab of partial sums (deterministic fixed-order reduction by launch_reduce_partials).
// WARPS = 16: one CTA per SM (E = 64 needs 225 KB of shared memory).  WARPS = 8 (E = 32, 81 KB): two CTAs per SM, so one CTA's
// load / barrier bubbles are filled by the other's MMAs; CTAs >= kSlabs write their partial sums to the second slab set.
template <int E, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 8 ? 2 : 1) ffn_bwd_kernel(const FfnBwdArgs a) {
    constexpr int FF_WARPS = WARPS, FF_THREADS = WARPS * 32;
    using G = FfnGeom<E, FF_THREADS>;
    constexpr int F = G::F, KS = E / 8, NT = E / 8, P1 = G::P1, P2 = G::P2, TT = G::TT, PX = G::PX, PH = G::PH, PD = G::PD;
    constexpr int RB = TT / 16, FR = FF_WARPS / RB, CPW = (F / 8) / FR;          // row blocks, hidden ranges, 8-col chunks per warp
    constexpr int LPR = E / 4;                                                    // lanes per row in P0 (float4 each)
    constexpr int MTILES = F / 16;                                                // 16-row tiles of the weight gradients
    constexpr int NACC = 2 * MTILES / FF_WARPS;                                   // weight-gradient kinds per warp (2: both, 1: one of them)
    static_assert(FF_THREADS / LPR == TT, "P0 covers the tile in one pass");
    static_assert(RB * NT == FF_WARPS, "one dx tile per warp");
    static_assert(NACC == 1 || NACC == 2, "weight-gradient tiles must divide over the warps");
    extern __shared__ __align__(16) float smem_f[];
    float* W1s = smem_f;
    float* W2s = W1s + F * P1;
    float* vec = W2s + E * P2;
    float* dzs = vec + G::VEC_FLOATS;
    float* x1s = dzs + TT * PX;
    float* hs = x1s + TT * PX;
    float* dhs = hs + TT * PH;
    pdl_trigger();
    stage_weights<E, FF_THREADS>(W1s, W2s, vec, a.W1, a.b1, a.W2, nullptr, a.gamma, nullptr);
    __syncthreads();
    pdl_wait();
    const int rows = a.n_rows_dev ? min(__ldg(a.n_rows_dev), a.M_cap) : a.M_cap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ntiles = (rows + TT - 1) / TT;
    const float* b1s = vec; const float* gs = vec + F + E;
    const float inv_e = 1.0f / (float)E;

    // P0 ownership: row p0r of the tile, columns p0c .. p0c+3
    const int p0r = threadIdx.x / LPR, p0c = (threadIdx.x % LPR) * 4;
    float dg[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f}, dbz[4] = {0.f, 0.f, 0.f, 0.f};
    const float4 gam = *reinterpret_cast<const float4*>(gs + p0c);
    // PA ownership
    const int rb = warp / FR, fr = warp % FR;
    // PB ownership: dx tile (rows xrb*16.., cols xnt*8..) ; weight rows f0 .. f0+15
    const int xrb = warp / NT, xnt = warp % NT;
    const int wkind = (NACC == 2) ? 0 : warp / MTILES;            // NACC == 1: the first MTILES warps own dW2^T, the others dW1
    const int f0 = ((NACC == 2) ? warp : (warp % MTILES)) * 16;
    float acc[NACC][NT][4];
#pragma unroll
    for (int k = 0; k < NACC; ++k)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[k][nt][0] = acc[k][nt][1] = acc[k][nt][2] = acc[k][nt][3] = 0.f;
    float db1_lo = 0.f, db1_hi = 0.f;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * TT;
        // ---- P0: LayerNorm backward ------------------------------------------------------------------------------
        {
            const int row = row0 + p0r;
            float4 dy = make_float4(0.f, 0.f, 0.f, 0.f), xh = dy, xv = dy;
            float rs = 0.f;
            if (row < rows) {
                dy = __ldg(reinterpret_cast<const float4*>(a.dY + (size_t)row * E + p0c));
                xh = __ldg(reinterpret_cast<const float4*>(a.xhat + (size_t)row * E + p0c));
                xv = __ldg(reinterpret_cast<const float4*>(a.X + (size_t)row * E + p0c));
                rs = __ldg(a.rstd + row);
                if (a.drop.thresh) {
                    const uint32_t rk = drop_rowkey(a.drop, (uint32_t)row);
                    dy.x *= drop_scale(a.drop, rk, (uint32_t)p0c); dy.y *= drop_scale(a.drop, rk, (uint32_t)p0c + 1);
                    dy.z *= drop_scale(a.drop, rk, (uint32_t)p0c + 2); dy.w *= drop_scale(a.drop, rk, (uint32_t)p0c + 3);
                }
            }
            const float g0 = dy.x * gam.x, g1 = dy.y * gam.y, g2 = dy.z * gam.z, g3 = dy.w * gam.w;
            const float s1 = group_sum<LPR>((g0 + g1) + (g2 + g3)) * inv_e;
            const float s2 = group_sum<LPR>(fmaf(g0, xh.x, fmaf(g1, xh.y, fmaf(g2, xh.z, g3 * xh.w)))) * inv_e;
            const float4 dz = make_float4(rs * (g0 - s1 - xh.x * s2), rs * (g1 - s1 - xh.y * s2), rs * (g2 - s1 - xh.z * s2),
                                          rs * (g3 - s1 - xh.w * s2));
            dg[0] = fmaf(dy.x, xh.x, dg[0]); dg[1] = fmaf(dy.y, xh.y, dg[1]); dg[2] = fmaf(dy.z, xh.z, dg[2]); dg[3] = fmaf(dy.w, xh.w, dg[3]);
            db[0] += dy.x; db[1] += dy.y; db[2] += dy.z; db[3] += dy.w;
            dbz[0] += dz.x; dbz[1] += dz.y; dbz[2] += dz.z; dbz[3] += dz.w;
            *reinterpret_cast<float4*>(dzs + p0r * PX + p0c) = dz;
            *reinterpret_cast<float4*>(x1s + p0r * PX + p0c) = xv;
        }
        __syncthreads();
        // ---- PA: h and dh of (row block rb, hidden range fr) --------------------------------------------------------
        {
            float hc[CPW][4], dc[CPW][4];
#pragma unroll
            for (int ci = 0; ci < CPW; ++ci) {
                const int col = 8 * (fr * CPW + ci) + 2 * t;
                hc[ci][0] = hc[ci][2] = b1s[col];
                hc[ci][1] = hc[ci][3] = b1s[col + 1];
                dc[ci][0] = dc[ci][1] = dc[ci][2] = dc[ci][3] = 0.f;
            }
            const float* xr = x1s + (rb * 16 + g) * PX + t;
            const float* zr = dzs + (rb * 16 + g) * PX + t;
            const float* w1p = W1s + (8 * fr * CPW + g) * P1 + t;               // B[k = e, n = f] = W1[f][e]
            const float* w2p = W2s + t * P2 + 8 * fr * CPW + g;                 // B[k = e, n = f] = W2[e][f]
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const float xa[4] = {tf32r(xr[8 * ks]), tf32r(xr[8 * PX + 8 * ks]), tf32r(xr[8 * ks + 4]), tf32r(xr[8 * PX + 8 * ks + 4])};
                const float za[4] = {zr[8 * ks], zr[8 * PX + 8 * ks], zr[8 * ks + 4], zr[8 * PX + 8 * ks + 4]};
#pragma unroll
                for (int ci = 0; ci < CPW; ++ci) {
                    mma_tf32(hc[ci], xa, w1p[8 * ci * P1 + 8 * ks], w1p[8 * ci * P1 + 8 * ks + 4]);
                    mma_tf32(dc[ci], za, w2p[8 * ks * P2 + 8 * ci], w2p[(8 * ks + 4) * P2 + 8 * ci]);
                }
            }
#pragma unroll
            for (int ci = 0; ci < CPW; ++ci) {
                const int col = 8 * (fr * CPW + ci) + 2 * t;
                const float h0 = fmaxf(hc[ci][0], 0.f), h1 = fmaxf(hc[ci][1], 0.f), h2 = fmaxf(hc[ci][2], 0.f), h3 = fmaxf(hc[ci][3], 0.f);
                *reinterpret_cast<float2*>(hs + (rb * 16 + g) * PH + col) = make_float2(tf32r(h0), tf32r(h1));
                *reinterpret_cast<float2*>(hs + (rb * 16 + g + 8) * PH + col) = make_float2(tf32r(h2), tf32r(h3));
                *reinterpret_cast<float2*>(dhs + (rb * 16 + g) * PD + col) = make_float2(h0 > 0.f ? dc[ci][0] : 0.f, h1 > 0.f ? dc[ci][1] : 0.f);
                *reinterpret_cast<float2*>(dhs + (rb * 16 + g + 8) * PD + col) = make_float2(h2 > 0.f ? dc[ci][2] : 0.f, h3 > 0.f ? dc[ci][3] : 0.f);
            }
        }
        __syncthreads();
        // ---- PB (i): dx1 tile = dz + dh W1 ------------------------------------------------------------------------
        {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            // contraction slots relabelled (slot t <-> f = 8c + 2t, slot t+4 <-> 8c + 2t + 1): 64-bit A loads, conflict-free B rows
            const float* ar = dhs + (xrb * 16 + g) * PD + 2 * t;                // A[m = token, k = f]
            const float* br = W1s + 2 * t * P1 + 8 * xnt + g;                   // B[k = f, n = e] = W1[f][e]
#pragma unroll 8
            for (int c = 0; c < F / 8; ++c) {
                const float2 lo = *reinterpret_cast<const float2*>(ar + 8 * c);
                const float2 hi = *reinterpret_cast<const float2*>(ar + 8 * PD + 8 * c);
                const float af[4] = {lo.x, hi.x, lo.y, hi.y};
                mma_tf32(d, af, br[8 * c * P1], br[(8 * c + 1) * P1]);
            }
            const int col = 8 * xnt + 2 * t;
            const int r_lo = row0 + xrb * 16 + g, r_hi = r_lo + 8;
            const float2 z_lo = *reinterpret_cast<const float2*>(dzs + (xrb * 16 + g) * PX + col);
            const float2 z_hi = *reinterpret_cast<const float2*>(dzs + (xrb * 16 + g + 8) * PX + col);
            if (r_lo < rows) *reinterpret_cast<float2*>(a.dX + (size_t)r_lo * E + col) = make_float2(d[0] + z_lo.x, d[1] + z_lo.y);
            if (r_hi < rows) *reinterpret_cast<float2*>(a.dX + (size_t)r_hi * E + col) = make_float2(d[2] + z_hi.x, d[3] + z_hi.y);
        }
        // ---- PB (ii): weight gradients over the tile's tokens --------------------------------------------------------
#pragma unroll 2
        for (int ks = 0; ks < TT / 8; ++ks) {
            const int k_lo = 8 * ks + t, k_hi = k_lo + 4;
            if (NACC == 2 || wkind == 0) {                                      // dW2^T[f, e] += h[tok, f] dz[tok, e]
                const float af[4] = {hs[k_lo * PH + f0 + g], hs[k_lo * PH + f0 + g + 8], hs[k_hi * PH + f0 + g], hs[k_hi * PH + f0 + g + 8]};
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[0][nt], af, dzs[k_lo * PX + 8 * nt + g], dzs[k_hi * PX + 8 * nt + g]);
            }
            if (NACC == 2 || wkind == 1) {                                      // dW1[f, e] += dh[tok, f] x1[tok, e]
                const float af[4] = {dhs[k_lo * PD + f0 + g], dhs[k_lo * PD + f0 + g + 8], dhs[k_hi * PD + f0 + g], dhs[k_hi * PD + f0 + g + 8]};
                db1_lo += af[0] + af[2]; db1_hi += af[1] + af[3];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[NACC - 1][nt], af, x1s[k_lo * PX + 8 * nt + g], x1s[k_hi * PX + 8 * nt + g]);
            }
        }
        __syncthreads();                                                        // the next tile's P0 overwrites dzs / x1s
    }

    // ---- slab write -------------------------------------------------------------------------------------------------
    // slab of this CTA: the first kSlabs CTAs use the caller's per-layer slabs, the others the second (FFN-only) slab set, whose
    // offsets are relative to o_w1
    float* slab = (int)blockIdx.x < kSlabs ? a.partial + (size_t)blockIdx.x * a.pstride
                                           : a.partial2 + (size_t)(blockIdx.x - kSlabs) * a.pstride2 - a.o_w1;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int e = 8 * nt + 2 * t;
        if (NACC == 2 || wkind == 0) {                                          // dW2 is [E][F]: transpose on the way out
            float* p = slab + a.o_w2;
            p[(size_t)e * F + f0 + g] = acc[0][nt][0]; p[(size_t)(e + 1) * F + f0 + g] = acc[0][nt][1];
            p[(size_t)e * F + f0 + g + 8] = acc[0][nt][2]; p[(size_t)(e + 1) * F + f0 + g + 8] = acc[0][nt][3];
        }
        if (NACC == 2 || wkind == 1) {                                          // dW1 is [F][E]
            float* p = slab + a.o_w1;
            *reinterpret_cast<float2*>(p + (size_t)(f0 + g) * E + e) = make_float2(acc[NACC - 1][nt][0], acc[NACC - 1][nt][1]);
            *reinterpret_cast<float2*>(p + (size_t)(f0 + g + 8) * E + e) = make_float2(acc[NACC - 1][nt][2], acc[NACC - 1][nt][3]);
        }
    }
    if (NACC == 2 || wkind == 1) {
        db1_lo = quad_sum(db1_lo); db1_hi = quad_sum(db1_hi);
        if (t == 0) { slab[a.o_b1 + f0 + g] = db1_lo; slab[a.o_b1 + f0 + g + 8] = db1_hi; }
    }
    // column sums kept per thread in P0 layout: reduce the TT rows of partials through shared memory (fixed order)
    float* red = dzs;                                                            // >= 3 * TT * E floats are free here (tile buffers)
    static_assert(3 * TT * E <= G::TILE_FLOATS, "reduction scratch fits in the tile buffers");
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[(0 * TT + p0r) * E + p0c + j] = dg[j];
        red[(1 * TT + p0r) * E + p0c + j] = db[j];
        red[(2 * TT + p0r) * E + p0c + j] = dbz[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * E; i += FF_THREADS) {
        const int which = i / E, e = i % E;
        float s = 0.f;
        for (int r = 0; r < TT; ++r) s += red[(which * TT + r) * E + e];
        slab[(which == 0 ? a.o_g : which == 1 ? a.o_b : a.o_b2) + e] = s;
    }
}

template <int E, int MT, int WARPS>
int launch_fwd_t(const FfnFwdArgs& a, cudaStream_t st) {
    using G = FfnGeom<E>;
    static bool configured = false;
    if (!configured) {
        MVN_CUDA(cudaFuncSetAttribute(ffn_fwd_kernel<E, MT, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_FWD));
        configured = true;
    }
    const int blocks = cdiv(a.M_cap, 16 * MT * WARPS);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks < num_sms() ? blocks : num_sms()); cfg.blockDim = dim3(WARPS * 32); cfg.dynamicSmemBytes = G::SMEM_FWD; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    MVN_CUDA(cudaLaunchKernelEx(&cfg, ffn_fwd_kernel<E, MT, WARPS>, a));
    MVN_LAUNCH_CHECK();
    return 0;
}
template <int E, int WARPS>
int launch_bwd_t(const FfnBwdArgs& a, cudaStream_t st) {
    using G = FfnGeom<E, WARPS * 32>;
    static bool configured = false;
    if (!configured) {
        MVN_CUDA(cudaFuncSetAttribute(ffn_bwd_kernel<E, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BWD));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(WARPS == 8 ? 2 * kSlabs : kSlabs);        // every slab is written
    cfg.blockDim = dim3(WARPS * 32); cfg.dynamicSmemBytes = G::SMEM_BWD; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    MVN_CUDA(cudaLaunchKernelEx(&cfg, ffn_bwd_kernel<E, WARPS>, a));
    MVN_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// ffn_fused_tc.cu / gemm_tc.cu
int launch_ffn_fused_fwd_tc(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, const float* gamma, const float* beta,
                            float* Y, float* xhat, float* rstd, const int32_t* n_rows_dev, int M_cap, int E, float eps, float wscale, const DropCfg& drop,
                            cudaStream_t st);
float tc_trunc_comp();

bool ffn_fused_supported(int E, int ff_mult) { return ff_mult == 4 && (E == 32 || E == 64); }
// floats of one FFN-only slab (w1 | b1 | w2 | b2 | gamma | beta, the order of the flat parameter layout)
size_t ffn_fused_slab_floats(int E) { return (size_t)8 * E * E + 4 * E + 3 * E; }
// number of slab sets the backward writes for this width: 2 when it runs two CTAs per SM
int ffn_fused_bwd_slab_sets(int E) {
    // MVN_FFN_BWD=2: two 8-warp CTAs per SM at E = 32.  Measured no better than one 16-warp CTA inside the step (fused_bwd 2.32 vs 2.36 ms)
    // and it doubles the partial-sum slabs to reduce, so the default stays one CTA per SM for every width.
    static const int v = getenv("MVN_FFN_BWD") ? atoi(getenv("MVN_FFN_BWD")) : 1;
    return (E == 32 && v == 2) ? 2 : 1;
}

int launch_ffn_fused_fwd(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, const float* gamma,
                         const float* beta, float* Y, float* xhat, float* rstd, const int32_t* n_rows_dev, int M_cap, int E, float eps,
                         const DropCfg& drop, cudaStream_t st) {
    MVN_CHECK_ARG(X && W1 && W2 && gamma && beta && Y && M_cap > 0, "ffn_fused_fwd: null pointer or empty input");
    MVN_CHECK_ARG(aligned16(X) && aligned16(W1) && aligned16(W2) && aligned16(Y) && (!xhat || aligned16(xhat)), "ffn_fused_fwd: buffers must be 16-byte aligned");
    MVN_UNSUPPORTED(E == 32 || E == 64, "ffn_fused: emb=%d not in {32, 64}", E);
    ProfScope prof(PROF_FUSED_FWD, st);
    count_tier(TIER_FUSED);
    FfnFwdArgs a;
    a.X = X; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.gamma = gamma; a.beta = beta; a.Y = Y; a.xhat = xhat; a.rstd = rstd;
    a.n_rows_dev = n_rows_dev; a.M_cap = M_cap; a.eps = eps; a.drop = drop;
    // MVN_FFN_FWD: -1 (default) per-width choice measured on B200 (scripts/bench_fused.py, C4 token counts): E = 64 -> the tcgen05 kernel
    // of ffn_fused_tc.cu (73.7 us, the warp-MMA kernel 75.8 us); E = 32 -> the warp-MMA kernel (36.9 us vs 51.2 us: with only four 32-column
    // chunks per tile the tcgen05 pipeline is all hand-off latency).  3: tcgen05 for both; 1 / 0 / 2: warp-MMA variants (A/B runs).
    static const int variant = getenv("MVN_FFN_FWD") ? atoi(getenv("MVN_FFN_FWD")) : -1;
    if ((variant == 3 || (variant == -1 && E == 64)) && aligned16(X) && aligned16(Y) && (!xhat || aligned16(xhat))) {
        const int r = launch_ffn_fused_fwd_tc(X, W1, b1, W2, b2, gamma, beta, Y, xhat, rstd, n_rows_dev, M_cap, E, eps, tc_trunc_comp(), drop, st);
        if (r != MVN_E_UNSUPPORTED) return r;               // fewer than 128 rows: the warp-MMA kernel below
    }
    if (variant == 0) return E == 64 ? launch_fwd_t<64, 1, 16>(a, st) : launch_fwd_t<32, 1, 16>(a, st);
    // measured (scripts/bench_fused.py, C4 token counts): E = 64 is faster with one row tile per warp and 16 warps (73.7 vs 80.9 us),
    // E = 32 with two row tiles per warp (36.9 vs 38.9 us)
    if (variant == 2) return E == 64 ? launch_fwd_t<64, 2, 8>(a, st) : launch_fwd_t<32, 2, 16>(a, st);
    return E == 64 ? launch_fwd_t<64, 1, 16>(a, st) : launch_fwd_t<32, 2, 16>(a, st);
}

int launch_ffn_fused_bwd(const float* dY, const float* xhat, const float* rstd, const float* X, const float* W1, const float* b1,
                         const float* W2, const float* gamma, float* dX, const int32_t* n_rows_dev, int M_cap, int E, const DropCfg& drop,
                         float* partial, size_t pstride, size_t o_w1, size_t o_b1, size_t o_w2, size_t o_b2, size_t o_g, size_t o_b,
                         float* partial2, cudaStream_t st, size_t pstride2) {
    MVN_CHECK_ARG(dY && xhat && rstd && X && W1 && W2 && gamma && dX && partial && M_cap > 0, "ffn_fused_bwd: null pointer or empty input");
    MVN_CHECK_ARG(ffn_fused_bwd_slab_sets(E) == 1 || partial2, "ffn_fused_bwd: the second slab set is missing");
    MVN_CHECK_ARG(o_b1 == o_w1 + (size_t)4 * E * E && o_w2 == o_b1 + (size_t)4 * E && o_b2 == o_w2 + (size_t)4 * E * E && o_g == o_b2 + E && o_b == o_g + E,
                  "ffn_fused_bwd: parameter offsets must follow the flat layout w1 | b1 | w2 | b2 | gamma | beta");
    MVN_CHECK_ARG(aligned16(dY) && aligned16(xhat) && aligned16(X) && aligned16(W1) && aligned16(W2) && aligned16(dX) && aligned16(partial) &&
                      pstride % 4 == 0 && o_w1 % 2 == 0,
                  "ffn_fused_bwd: buffers must be 16-byte aligned");
    MVN_UNSUPPORTED(E == 32 || E == 64, "ffn_fused: emb=%d not in {32, 64}", E);
    ProfScope prof(PROF_FUSED_BWD, st);
    count_tier(TIER_FUSED);
    FfnBwdArgs a;
    a.dY = dY; a.xhat = xhat; a.rstd = rstd; a.X = X; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.gamma = gamma; a.dX = dX;
    a.n_rows_dev = n_rows_dev; a.M_cap = M_cap; a.drop = drop; a.partial = partial; a.pstride = pstride;
    a.o_w1 = o_w1; a.o_b1 = o_b1; a.o_w2 = o_w2; a.o_b2 = o_b2; a.o_g = o_g; a.o_b = o_b;
    a.partial2 = partial2; a.pstride2 = pstride2 ? pstride2 : ffn_fused_slab_floats(E);
    if (E == 64) return launch_bwd_t<64, 16>(a, st);
    return ffn_fused_bwd_slab_sets(E) == 2 ? launch_bwd_t<32, 8>(a, st) : launch_bwd_t<32, 16>(a, st);
}

}  // namespace mvn

using namespace mvn;

extern "C" int mvn_ffn_fused_fwd(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, const float* gamma,
                                 const float* beta, float* Y, float* xhat, float* rstd, const int32_t* n_rows_dev, int M_cap, int E,
                                 int ff_mult, float eps, float dropout_p, uint64_t seed, int site, void* stream) {
    MVN_UNSUPPORTED(ffn_fused_supported(E, ff_mult), "ffn_fused_fwd: emb=%d ff_mult=%d (needs emb in {32,64}, ff_mult 4)", E, ff_mult);
    return launch_ffn_fused_fwd(X, W1, b1, W2, b2, gamma, beta, Y, xhat, rstd, n_rows_dev, M_cap, E, eps, make_drop(dropout_p, seed, (uint32_t)site),
                                (cudaStream_t)stream);
}

extern "C" size_t mvn_ffn_fused_bwd_workspace_bytes(int E, int ff_mult) {
    if (!ffn_fused_supported(E, ff_mult)) return 0;
    return (size_t)2 * kSlabs * ffn_fused_slab_floats(E) * sizeof(float) + 256;
}

extern "C" int mvn_ffn_fused_bwd(const float* dY, const float* xhat, const float* rstd, const float* X, const float* W1, const float* b1,
                                 const float* W2, const float* gamma, float* dX, float* dW1, float* db1, float* dW2, float* db2, float* dgamma,
                                 float* dbeta, const int32_t* n_rows_dev, int M_cap, int E, int ff_mult, float dropout_p, uint64_t seed, int site,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    MVN_UNSUPPORTED(ffn_fused_supported(E, ff_mult), "ffn_fused_bwd: emb=%d ff_mult=%d (needs emb in {32,64}, ff_mult 4)", E, ff_mult);
    MVN_CHECK_ARG(dW1 && db1 && dW2 && db2 && dgamma && dbeta && workspace, "ffn_fused_bwd: null gradient pointer");
    const size_t F = 4 * (size_t)E;
    const size_t o_w1 = 0, o_b1 = F * E, o_w2 = o_b1 + F, o_b2 = o_w2 + E * F, o_g = o_b2 + E, o_b = o_g + E, ps = o_b + E;
    float* part = (float*)align_up((size_t)workspace, 256);
    if ((char*)part + (size_t)2 * kSlabs * ps * sizeof(float) > (char*)workspace + workspace_bytes) { set_error("ffn_fused_bwd: workspace too small"); return MVN_E_WORKSPACE; }
    float* part2 = part + (size_t)kSlabs * ps;
    cudaStream_t st = (cudaStream_t)stream;
    MVN_TRY(launch_ffn_fused_bwd(dY, xhat, rstd, X, W1, b1, W2, gamma, dX, n_rows_dev, M_cap, E, make_drop(dropout_p, seed, (uint32_t)site), part, ps,
                                 o_w1, o_b1, o_w2, o_b2, o_g, o_b, part2, st));
    const int sets = ffn_fused_bwd_slab_sets(E);
    float* outs[6] = {dW1, db1, dW2, db2, dgamma, dbeta};
    const size_t offs[6] = {o_w1, o_b1, o_w2, o_b2, o_g, o_b}, lens[6] = {F * E, F, E * F, (size_t)E, (size_t)E, (size_t)E};
    for (int i = 0; i < 6; ++i) {
        MVN_TRY(launch_reduce_partials(part + offs[i], ps, lens[i], outs[i], 0, st));
        if (sets == 2) MVN_TRY(launch_reduce_partials(part2 + offs[i], ps, lens[i], outs[i], 1, st));
    }
    return 0;
}
