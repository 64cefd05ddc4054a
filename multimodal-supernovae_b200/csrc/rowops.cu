// HBM-bound row kernels: ragged pack plan, fused time/band/magnitude embedding, LayerNorm backward,
// masked pooling, L2 normalisation, pack/unpack.  One warp per token row, coalesced along E.
#include "common.cuh"

namespace mvn {
namespace {

// ---------------------------------------------------------------------------------------------------
// pack plan: (1) per-sequence valid counts, (2) single-CTA exclusive scan, (3) per-sequence index write
__global__ void count_valid_kernel(const uint8_t* __restrict__ mask, int B, int T, int valid_only, int32_t* __restrict__ counts) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    int c = 0;
    if (!valid_only || mask == nullptr) {
        c = T;
    } else {
        const uint8_t* m = mask + (size_t)warp * T;
        for (int t = lane; t < T; t += 32) c += m[t] != 0;
        c = (int)warp_sum((float)c);      // T <= 2^24 so exact in fp32
    }
    if (lane == 0) counts[warp] = c;
}

__global__ void __launch_bounds__(1024) scan_kernel(const int32_t* __restrict__ counts, int B, int32_t* __restrict__ cu) {
    // exclusive scan of counts[B] into cu[B+1]; one CTA, chunked.
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int i = base + tid;
        const int v = i < B ? counts[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = x + (wid > 0 ? warp_tot[wid - 1] : 0) + carry;
        if (i < B) cu[i] = incl - v;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
    if (tid == 0) cu[B] = carry_s;
}

__global__ void write_index_kernel(const uint8_t* __restrict__ mask, int B, int T, int valid_only, const int32_t* __restrict__ cu,
                                   int32_t* __restrict__ tok_src, uint8_t* __restrict__ keyvalid) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int b = warp;
    const int base = cu[b];
    if (!valid_only) {
        for (int t = lane; t < T; t += 32) {
            tok_src[base + t] = b * T + t;
            keyvalid[base + t] = mask ? (mask[(size_t)b * T + t] != 0) : 1;
        }
        return;
    }
    int run = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const bool v = t < T && (mask == nullptr || mask[(size_t)b * T + t] != 0);
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        if (v) {
            const int pos = base + run + __popc(bal & ((1u << lane) - 1u));
            tok_src[pos] = b * T + t;
            keyvalid[pos] = 1;
        }
        run += __popc(bal);
    }
}

__global__ void fill_tail_kernel(const int32_t* __restrict__ cu, int B, int BT, int32_t* __restrict__ tok_src, uint8_t* __restrict__ keyvalid) {
    const int n = cu[B];
    for (int i = n + blockIdx.x * blockDim.x + threadIdx.x; i < BT; i += gridDim.x * blockDim.x) { tok_src[i] = -1; keyvalid[i] = 0; }
}

// ---------------------------------------------------------------------------------------------------
// embed: out[m, e] = x*w[e] + b[e] + pe(t)[e] + band_emb[band][e]
// pe[2i] = sin(t*div[i]), pe[2i+1] = cos(t*div[i]); t*div is ONE fp32 multiply, full-range sinf/cosf.
__global__ void __launch_bounds__(256) embed_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                        const int32_t* __restrict__ cu, const int32_t* __restrict__ tok_src,
                                                        const float* __restrict__ div_term, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ band_emb,
                                                        int B, int T, int E, int nband, float* __restrict__ out, const DropCfg drop) {
    // One lane per (sin, cos) pair: sincosf shares the full-range argument reduction between the two, the pair is stored as one
    // 64-bit word, and a warp covers 32 / (E/2) tokens per pass (2 tokens at E = 32), so every store instruction writes full lines.
    const int rows = cu[B];
    const int lane = threadIdx.x & 31;
    const int per_band = T / (nband > 0 ? nband : 1);
    const int half = E >> 1;
    const int lpt = half < 32 ? half : 32;                 // lanes per token (E is a power of two >= 16 here)
    const int tpw = 32 / lpt, sub = lane / lpt, l = lane % lpt;
    const int nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int m = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * tpw + sub; m < rows && sub < tpw; m += nwarp * tpw) {
        const int src = tok_src[m];
        const float xv = x[src], tv = t[src];
        const int band = (nband > 1) ? min((src % T) / per_band, nband - 1) : 0;
        const uint32_t rk = drop.thresh ? drop_rowkey(drop, (uint32_t)m) : 0u;
        for (int i = l; i < half; i += lpt) {
            const float arg = __fmul_rn(tv, div_term[i]);
            float sn, cs;
            sincosf(arg, &sn, &cs);
            const float2 wv = *reinterpret_cast<const float2*>(w + 2 * i), bv = *reinterpret_cast<const float2*>(bias + 2 * i);
            float v0 = fmaf(xv, wv.x, bv.x) + sn, v1 = fmaf(xv, wv.y, bv.y) + cs;
            if (nband > 1) {
                const float2 be = *reinterpret_cast<const float2*>(band_emb + band * E + 2 * i);
                v0 += be.x; v1 += be.y;
            }
            if (drop.thresh) { v0 *= drop_scale(drop, rk, (uint32_t)(2 * i)); v1 *= drop_scale(drop, rk, (uint32_t)(2 * i + 1)); }
            *reinterpret_cast<float2*>(out + (size_t)m * E + 2 * i) = make_float2(v0, v1);
        }
    }
}

// partial[s][off + e] = sum dout*x ; [off+E+e] = sum dout ; [off+2E + band*E + e] = sum dout over band
// 16 warps per CTA, 4 rows in flight per warp (index -> x -> dout is a dependent chain: latency, not bytes, bounds it).
constexpr int EMB_WARPS = 16;
template <int PER>
__global__ void __launch_bounds__(EMB_WARPS * 32) embed_bwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ tok_src,
                                                                   const float* __restrict__ dout, const int32_t* n_rows_dev, int M_cap,
                                                                   int T, int E, int nband, float* __restrict__ partial, size_t pstride, size_t off,
                                                                   const DropCfg drop) {
    constexpr int MAXB = 4, U = 4;
    __shared__ float red[EMB_WARPS][(2 + MAXB) * 32 * PER];
    const int rows = n_rows_dev ? min(*n_rows_dev, M_cap) : M_cap;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int per_band = T / (nband > 0 ? nband : 1);
    float dw[PER], db[PER], dbd[MAXB][PER];
#pragma unroll
    for (int p = 0; p < PER; ++p) { dw[p] = 0.f; db[p] = 0.f;
#pragma unroll
        for (int k = 0; k < MAXB; ++k) dbd[k][p] = 0.f; }
    const int stride = gridDim.x * EMB_WARPS;
    for (int m0 = blockIdx.x * EMB_WARPS + wid; m0 < rows; m0 += U * stride) {
        int src[U]; float xv[U]; float g[U][PER];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int m = m0 + u * stride; src[u] = m < rows ? tok_src[m] : 0; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = m0 + u * stride;
            xv[u] = m < rows ? x[src[u]] : 0.f;
#pragma unroll
            for (int p = 0; p < PER; ++p) { const int e = lane + 32 * p; g[u][p] = (m < rows && e < E) ? dout[(size_t)m * E + e] : 0.f; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = m0 + u * stride;
            const int band = (nband > 1) ? min((src[u] % T) / per_band, nband - 1) : 0;
            const uint32_t rk = drop.thresh ? drop_rowkey(drop, (uint32_t)m) : 0u;
#pragma unroll
            for (int p = 0; p < PER; ++p) {
                float gv = g[u][p];
                if (drop.thresh) gv *= drop_scale(drop, rk, (uint32_t)(lane + 32 * p));
                dw[p] = fmaf(gv, xv[u], dw[p]);
                db[p] += gv;
#pragma unroll
                for (int k = 0; k < MAXB; ++k) dbd[k][p] += (k == band) ? gv : 0.f;
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PER; ++p) {
        red[wid][0 * 32 * PER + lane + 32 * p] = dw[p];
        red[wid][1 * 32 * PER + lane + 32 * p] = db[p];
#pragma unroll
        for (int k = 0; k < MAXB; ++k) red[wid][(2 + k) * 32 * PER + lane + 32 * p] = dbd[k][p];
    }
    __syncthreads();
    float* pp = partial + (size_t)blockIdx.x * pstride + off;
    const int nsec = 2 + (nband > 1 ? nband : 0);
    for (int i = threadIdx.x; i < nsec * 32 * PER; i += blockDim.x) {
        const int sec = i / (32 * PER), e = i % (32 * PER);
        if (e >= E) continue;
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < EMB_WARPS; ++w8) s += red[w8][i];
        pp[(size_t)sec * E + e] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward: dz = rstd * (g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma; partial dgamma/dbeta per CTA
constexpr int LNB_WARPS = 32;      // 1024-thread CTAs: one CTA per slab keeps the partial layout, 32 warps keep HBM busy
template <int PER>
__global__ void __launch_bounds__(LNB_WARPS * 32) ln_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ xhat,
                                                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                float* __restrict__ dZ, const int32_t* n_rows_dev, int M_cap, int E,
                                                                float* __restrict__ partial, size_t pstride, size_t goff, size_t boff,
                                                                const DropCfg drop) {
    __shared__ float red[LNB_WARPS][2 * 32 * PER];
    pdl_trigger();
    const int rows = n_rows_dev ? min(*n_rows_dev, M_cap) : M_cap;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float gam[PER], dg[PER], dbt[PER];
#pragma unroll
    for (int p = 0; p < PER; ++p) { const int e = lane + 32 * p; gam[p] = e < E ? gamma[e] : 0.f; dg[p] = 0.f; dbt[p] = 0.f; }
    const float invE = 1.0f / (float)E;
    const int stride = gridDim.x * LNB_WARPS;
    for (int m0 = blockIdx.x * LNB_WARPS + wid; m0 < rows; m0 += 2 * stride) {
        // two rows in flight per warp (independent loads issued before either reduction)
        float dy[2][PER], xh[2][PER], rs[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int m = m0 + u * stride;
            const bool live = m < rows;
            rs[u] = live ? rstd[m] : 0.f;
            const uint32_t rk = drop.thresh ? drop_rowkey(drop, (uint32_t)m) : 0u;
#pragma unroll
            for (int p = 0; p < PER; ++p) {
                const int e = lane + 32 * p;
                dy[u][p] = (live && e < E) ? dY[(size_t)m * E + e] : 0.f;
                if (drop.thresh) dy[u][p] *= drop_scale(drop, rk, (uint32_t)e);
                xh[u][p] = (live && e < E) ? xhat[(size_t)m * E + e] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int m = m0 + u * stride;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int p = 0; p < PER; ++p) {
                const float g = dy[u][p] * gam[p];
                s1 += g; s2 = fmaf(g, xh[u][p], s2);
                dg[p] = fmaf(dy[u][p], xh[u][p], dg[p]);
                dbt[p] += dy[u][p];
            }
            s1 = warp_sum(s1) * invE; s2 = warp_sum(s2) * invE;
            if (m < rows) {
#pragma unroll
                for (int p = 0; p < PER; ++p) {
                    const int e = lane + 32 * p;
                    if (e < E) dZ[(size_t)m * E + e] = rs[u] * (dy[u][p] * gam[p] - s1 - xh[u][p] * s2);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PER; ++p) { red[wid][lane + 32 * p] = dg[p]; red[wid][32 * PER + lane + 32 * p] = dbt[p]; }
    __syncthreads();
    float* pp = partial + (size_t)blockIdx.x * pstride;
    for (int i = threadIdx.x; i < 2 * 32 * PER; i += blockDim.x) {
        const int sec = i / (32 * PER), e = i % (32 * PER);
        if (e >= E) continue;
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < LNB_WARPS; ++w8) s += red[w8][i];
        pp[(sec == 0 ? goff : boff) + e] = s;
    }
}

// Vectorised variant for E in {16, 32, 64, 128}: a lane owns 4 consecutive columns (one 16-byte access), E/4 lanes share
// a row, so one warp-wide load instruction covers 32/(E/4) rows and U such row groups are in flight per warp before the
// first reduction -- 1024 threads x U x 32 B of loads outstanding per SM, which is what it takes to keep HBM3e busy from
// 128 CTAs.  Same slab layout and (fixed) summation order for the dgamma / dbeta partials as the scalar kernel.
template <int LPR, int U>
__global__ void __launch_bounds__(LNB_WARPS * 32) ln_bwd_vec_kernel(const float* __restrict__ dY, const float* __restrict__ xhat,
                                                                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                    float* __restrict__ dZ, const int32_t* n_rows_dev, int M_cap,
                                                                    float* __restrict__ partial, size_t pstride, size_t goff, size_t boff,
                                                                    const DropCfg drop) {
    constexpr int RPW = 32 / LPR, E = LPR * 4;
    __shared__ float red[LNB_WARPS][32][8];
    pdl_trigger();
    pdl_wait();
    const int rows = n_rows_dev ? min(*n_rows_dev, M_cap) : M_cap;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sub = lane / LPR, cl = lane % LPR;
    const float4 gam = *reinterpret_cast<const float4*>(gamma + cl * 4);
    float dg[4] = {0.f, 0.f, 0.f, 0.f}, dbt[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr float invE = 1.0f / (float)E;
    const int stride = gridDim.x * LNB_WARPS * RPW;
    for (int base = (blockIdx.x * LNB_WARPS + wid) * RPW; base < rows; base += U * stride) {
        float4 dy[U], xh[U];
        float rs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = base + u * stride + sub;
            const bool live = m < rows;
            rs[u] = live ? rstd[m] : 0.f;
            dy[u] = live ? *reinterpret_cast<const float4*>(dY + (size_t)m * E + cl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            xh[u] = live ? *reinterpret_cast<const float4*>(xhat + (size_t)m * E + cl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = base + u * stride + sub;
            if (drop.thresh) {
                const uint32_t rk = drop_rowkey(drop, (uint32_t)m);
                dy[u].x *= drop_scale(drop, rk, (uint32_t)(cl * 4)); dy[u].y *= drop_scale(drop, rk, (uint32_t)(cl * 4 + 1));
                dy[u].z *= drop_scale(drop, rk, (uint32_t)(cl * 4 + 2)); dy[u].w *= drop_scale(drop, rk, (uint32_t)(cl * 4 + 3));
            }
            const float g0 = dy[u].x * gam.x, g1 = dy[u].y * gam.y, g2 = dy[u].z * gam.z, g3 = dy[u].w * gam.w;
            float s1 = (g0 + g1) + (g2 + g3);
            float s2 = fmaf(g0, xh[u].x, fmaf(g1, xh[u].y, fmaf(g2, xh[u].z, g3 * xh[u].w)));
            dg[0] = fmaf(dy[u].x, xh[u].x, dg[0]); dg[1] = fmaf(dy[u].y, xh[u].y, dg[1]);
            dg[2] = fmaf(dy[u].z, xh[u].z, dg[2]); dg[3] = fmaf(dy[u].w, xh[u].w, dg[3]);
            dbt[0] += dy[u].x; dbt[1] += dy[u].y; dbt[2] += dy[u].z; dbt[3] += dy[u].w;
            s1 = group_sum<LPR>(s1) * invE; s2 = group_sum<LPR>(s2) * invE;
            if (m < rows) {
                const float r = rs[u];
                *reinterpret_cast<float4*>(dZ + (size_t)m * E + cl * 4) =
                    make_float4(r * (g0 - s1 - xh[u].x * s2), r * (g1 - s1 - xh[u].y * s2), r * (g2 - s1 - xh[u].z * s2), r * (g3 - s1 - xh[u].w * s2));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[wid][lane][j] = dg[j]; red[wid][lane][4 + j] = dbt[j]; }
    __syncthreads();
    float* pp = partial + (size_t)blockIdx.x * pstride;
    for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) {
        const int sec = i / E, e = i % E;
        float s = 0.f;
        for (int w = 0; w < LNB_WARPS; ++w)
#pragma unroll
            for (int r = 0; r < RPW; ++r) s += red[w][r * LPR + (e >> 2)][sec * 4 + (e & 3)];
        pp[(sec == 0 ? goff : boff) + e] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// pooling: one CTA (128 threads) per sequence, thread e owns column e
__global__ void __launch_bounds__(128) pool_fwd_kernel(const float* __restrict__ X, const int32_t* __restrict__ cu,
                                                       const uint8_t* __restrict__ keyvalid, int T, int E, int agg,
                                                       float* __restrict__ pooled, int32_t* __restrict__ argmax) {
    const int b = blockIdx.x;
    const int r0 = cu[b], r1 = cu[b + 1];
    __shared__ int nvalid_s;
    if (threadIdx.x == 0) nvalid_s = 0;
    __syncthreads();
    int c = 0;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) c += keyvalid ? (keyvalid[r] != 0) : 1;
    if (c) atomicAdd(&nvalid_s, c);
    __syncthreads();
    const int nvalid = nvalid_s;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        if (agg == MVN_AGG_MEAN) {
            float s = 0.f;
            for (int r = r0; r < r1; ++r) if (!keyvalid || keyvalid[r]) s += X[(size_t)r * E + e];
            pooled[(size_t)b * E + e] = s / (float)nvalid;          // 0/0 -> NaN like the reference
        } else {
            // x*mask then max over all T positions: zeroed padded rows take part when nvalid < T
            float best = (nvalid < T) ? 0.f : -INFINITY;
            int arg = -1;
            for (int r = r0; r < r1; ++r) {
                if (keyvalid && !keyvalid[r]) continue;
                const float v = X[(size_t)r * E + e];
                if (v > best) { best = v; arg = r; }
            }
            pooled[(size_t)b * E + e] = best;
            if (argmax) argmax[(size_t)b * E + e] = arg;
        }
    }
}

__global__ void __launch_bounds__(128) pool_bwd_kernel(const float* __restrict__ dpooled, const int32_t* __restrict__ cu,
                                                       const uint8_t* __restrict__ keyvalid, const int32_t* __restrict__ argmax,
                                                       int E, int agg, float* __restrict__ dX) {
    const int b = blockIdx.x;
    const int r0 = cu[b], r1 = cu[b + 1];
    __shared__ int nvalid_s;
    if (threadIdx.x == 0) nvalid_s = 0;
    __syncthreads();
    int c = 0;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) c += keyvalid ? (keyvalid[r] != 0) : 1;
    if (c) atomicAdd(&nvalid_s, c);
    __syncthreads();
    const float inv = 1.0f / (float)nvalid_s;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const float g = dpooled[(size_t)b * E + e];
        if (agg == MVN_AGG_MEAN) {
            for (int r = r0; r < r1; ++r) dX[(size_t)r * E + e] = (!keyvalid || keyvalid[r]) ? g * inv : 0.f;
        } else {
            const int arg = argmax[(size_t)b * E + e];
            for (int r = r0; r < r1; ++r) dX[(size_t)r * E + e] = (r == arg) ? g : 0.f;
        }
    }
}

__global__ void unpack_rows_kernel(const float* __restrict__ X, const int32_t* __restrict__ tok_src, const uint8_t* __restrict__ keyvalid,
                                   const int32_t* n_rows_dev, int BT, int E, float* __restrict__ out) {
    const int rows = n_rows_dev ? min(*n_rows_dev, BT) : BT;
    const int lane = threadIdx.x & 31;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < rows; m += (gridDim.x * blockDim.x) >> 5) {
        const int src = tok_src[m];
        const bool v = !keyvalid || keyvalid[m];
        for (int e = lane; e < E; e += 32) out[(size_t)src * E + e] = v ? X[(size_t)m * E + e] : 0.f;
    }
}
__global__ void pack_rows_kernel(const float* __restrict__ dense, const int32_t* __restrict__ tok_src, const uint8_t* __restrict__ keyvalid,
                                 const int32_t* n_rows_dev, int BT, int E, float* __restrict__ X) {
    const int rows = n_rows_dev ? min(*n_rows_dev, BT) : BT;
    const int lane = threadIdx.x & 31;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < rows; m += (gridDim.x * blockDim.x) >> 5) {
        const int src = tok_src[m];
        const bool v = !keyvalid || keyvalid[m];
        for (int e = lane; e < E; e += 32) X[(size_t)m * E + e] = v ? dense[(size_t)src * E + e] : 0.f;
    }
}

// L2 norm, one warp per row
__global__ void l2norm_fwd_kernel(const float* __restrict__ X, float* __restrict__ Y, float* __restrict__ norm, int B, int D) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= B) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = X[(size_t)row * D + d]; s = fmaf(v, v, s); }
    const float n = sqrtf(warp_sum(s));
    for (int d = lane; d < D; d += 32) Y[(size_t)row * D + d] = X[(size_t)row * D + d] / n;
    if (lane == 0 && norm) norm[row] = n;
}
__global__ void l2norm_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, const float* __restrict__ norm,
                                  float* __restrict__ dX, int B, int D) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= B) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(dY[(size_t)row * D + d], Y[(size_t)row * D + d], s);
    s = warp_sum(s);
    const float inv = 1.0f / norm[row];
    for (int d = lane; d < D; d += 32) dX[(size_t)row * D + d] = (dY[(size_t)row * D + d] - Y[(size_t)row * D + d] * s) * inv;
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ h, size_t n, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = h[i] > 0.f ? dy[i] : 0.f;
}

}  // namespace

int launch_ln_bwd(const float* dY, const float* xhat, const float* rstd, const float* gamma, float* dZ,
                  const int32_t* n_rows_dev, int M_cap, int E, float* partial, size_t pstride, size_t goff, size_t boff,
                  cudaStream_t st, const DropCfg& drop) {
    MVN_CHECK_ARG(dY && xhat && rstd && gamma && dZ && partial, "layernorm_bwd: null pointer");
    MVN_UNSUPPORTED(E >= 1 && E <= 128, "layernorm_bwd: E=%d outside [1,128]", E);
    ProfScope prof(PROF_ROW, st);
    if (aligned16(dY) && aligned16(xhat) && aligned16(dZ) && aligned16(gamma) && (E == 16 || E == 32 || E == 64 || E == 128)) {
        switch (E) {
            case 16: MVN_CUDA(launch_dependent(ln_bwd_vec_kernel<4, 4>, dim3(kSlabs), dim3(LNB_WARPS * 32), 0, st, dY, xhat, rstd, gamma, dZ, n_rows_dev, M_cap, partial, pstride, goff, boff, drop)); break;
            case 32: MVN_CUDA(launch_dependent(ln_bwd_vec_kernel<8, 4>, dim3(kSlabs), dim3(LNB_WARPS * 32), 0, st, dY, xhat, rstd, gamma, dZ, n_rows_dev, M_cap, partial, pstride, goff, boff, drop)); break;
            case 64: MVN_CUDA(launch_dependent(ln_bwd_vec_kernel<16, 4>, dim3(kSlabs), dim3(LNB_WARPS * 32), 0, st, dY, xhat, rstd, gamma, dZ, n_rows_dev, M_cap, partial, pstride, goff, boff, drop)); break;
            default: MVN_CUDA(launch_dependent(ln_bwd_vec_kernel<32, 4>, dim3(kSlabs), dim3(LNB_WARPS * 32), 0, st, dY, xhat, rstd, gamma, dZ, n_rows_dev, M_cap, partial, pstride, goff, boff, drop)); break;
        }
        MVN_LAUNCH_CHECK();
        return 0;
    }
    const int per = (E + 31) / 32;
    switch (per) {
        case 1: ln_bwd_kernel<1><<<kSlabs, LNB_WARPS * 32, 0, st>>>(dY, xhat, rstd, gamma, dZ, n_rows_dev, M_cap, E, partial, pstride, goff, boff, drop); break;
        case 2: ln_bwd_kernel<2><<<kSlabs, LNB_WARPS * 32, 0, st>>>(dY, xhat, rstd, gamma, dZ, n_rows_dev, M_cap, E, partial, pstride, goff, boff, drop); break;
        default: ln_bwd_kernel<4><<<kSlabs, LNB_WARPS * 32, 0, st>>>(dY, xhat, rstd, gamma, dZ, n_rows_dev, M_cap, E, partial, pstride, goff, boff, drop); break;
    }
    MVN_LAUNCH_CHECK();
    return 0;
}

int launch_embed_bwd_partials(const float* x, const int32_t* tok_src, const float* dout, const int32_t* n_rows_dev,
                              int M_cap, int T, int E, int nband, float* partial, size_t pstride, size_t off, cudaStream_t st,
                              const DropCfg& drop) {
    MVN_CHECK_ARG(x && tok_src && dout && partial, "embed_bwd: null pointer");
    MVN_UNSUPPORTED(E >= 1 && E <= 128 && nband >= 1 && nband <= 4, "embed_bwd: E=%d nband=%d unsupported", E, nband);
    ProfScope prof(PROF_ROW, st);
    const int per = (E + 31) / 32;
    switch (per) {
        case 1: embed_bwd_kernel<1><<<kSlabs, EMB_WARPS * 32, 0, st>>>(x, tok_src, dout, n_rows_dev, M_cap, T, E, nband, partial, pstride, off, drop); break;
        case 2: embed_bwd_kernel<2><<<kSlabs, EMB_WARPS * 32, 0, st>>>(x, tok_src, dout, n_rows_dev, M_cap, T, E, nband, partial, pstride, off, drop); break;
        default: embed_bwd_kernel<4><<<kSlabs, EMB_WARPS * 32, 0, st>>>(x, tok_src, dout, n_rows_dev, M_cap, T, E, nband, partial, pstride, off, drop); break;
    }
    MVN_LAUNCH_CHECK();
    return 0;
}

int launch_embed_fwd(const float* x, const float* t, const int32_t* cu_seqlens, const int32_t* tok_src, const float* div_term,
                     const float* w, const float* b, const float* band_emb, int B, int T, int E, int nband, float* out,
                     cudaStream_t st, const DropCfg& drop) {
    MVN_CHECK_ARG(x && t && cu_seqlens && tok_src && div_term && w && b && out && B > 0 && T > 0 && E > 0, "embed_fwd: bad arguments");
    MVN_CHECK_ARG(E % 2 == 0, "embed_fwd: E must be even (sin/cos pairs), got %d", E);
    MVN_CHECK_ARG(nband >= 1 && (nband == 1 || (band_emb && T % nband == 0)), "embed_fwd: nband=%d needs band_emb and T%%nband==0", nband);
    const int blocks = min(cdiv(B * T, 8), num_sms() * 8);
    ProfScope prof(PROF_ROW, st);
    embed_fwd_kernel<<<blocks, 256, 0, st>>>(x, t, cu_seqlens, tok_src, div_term, w, b, band_emb, B, T, E, nband, out, drop);
    MVN_LAUNCH_CHECK();
    return 0;
}

__global__ void dropout_scale_kernel(const DropCfg drop, size_t n, int cols, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = drop.thresh ? drop_scale(drop, drop_rowkey(drop, (uint32_t)(i / cols)), (uint32_t)(i % cols)) : 1.0f;
}

// y = x * keep-factor (vectorised when cols % 4 == 0); the same call with dy -> dx is the backward
__global__ void dropout_apply_kernel(const DropCfg drop, const float* __restrict__ x, float* __restrict__ y, size_t n, int cols, int vec) {
    if (vec) {
        const size_t n4 = n / 4;
        const int c4 = cols / 4;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            float4 v = reinterpret_cast<const float4*>(x)[i];
            const uint32_t rk = drop_rowkey(drop, (uint32_t)(i / c4)), c = (uint32_t)(i % c4) * 4u;
            v.x *= drop_scale(drop, rk, c); v.y *= drop_scale(drop, rk, c + 1); v.z *= drop_scale(drop, rk, c + 2); v.w *= drop_scale(drop, rk, c + 3);
            reinterpret_cast<float4*>(y)[i] = v;
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = x[i] * drop_scale(drop, drop_rowkey(drop, (uint32_t)(i / cols)), (uint32_t)(i % cols));
}

}  // namespace mvn

using namespace mvn;

extern "C" int mvn_dropout_apply(const float* x, float* y, int rows, int cols, uint64_t seed, int site, float p, void* stream) {
    MVN_CHECK_ARG(x && y && rows > 0 && cols > 0 && p >= 0.f && p < 1.f && site >= 0, "dropout_apply: bad arguments");
    const size_t n = (size_t)rows * cols;
    if (!(p > 0.f)) {
        if (x != y) MVN_CUDA(cudaMemcpyAsync(y, x, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return 0;
    }
    const int vec = (cols % 4 == 0) && aligned16(x) && aligned16(y);
    const size_t work = vec ? n / 4 : n;
    const int blocks = (int)((work + 255) / 256 < (size_t)(8 * num_sms()) ? (work + 255) / 256 : (size_t)(8 * num_sms()));
    ProfScope prof(PROF_ROW, (cudaStream_t)stream);
    dropout_apply_kernel<<<blocks > 0 ? blocks : 1, 256, 0, (cudaStream_t)stream>>>(make_drop(p, seed, (uint32_t)site), x, y, n, cols, vec);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_dropout_scale(uint64_t seed, int site, float p, int rows, int cols, float* out, void* stream) {
    MVN_CHECK_ARG(out && rows > 0 && cols > 0 && p >= 0.f && p < 1.f && site >= 0, "dropout_scale: bad arguments");
    const size_t n = (size_t)rows * cols;
    dropout_scale_kernel<<<(int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, (cudaStream_t)stream>>>(make_drop(p, seed, (uint32_t)site), n, cols, out);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_pack_plan(const uint8_t* mask, int B, int T, int valid_only, int32_t* cu_seqlens, int32_t* tok_src,
                             uint8_t* keyvalid, void* stream) {
    MVN_CHECK_ARG(B > 0 && T > 0 && cu_seqlens && tok_src && keyvalid, "pack_plan: bad arguments");
    MVN_CHECK_ARG((long long)B * T < (1ll << 31), "pack_plan: B*T overflows int32");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ROW, st);
    // counts are staged in tok_src[0..B) (B <= B*T), consumed by the scan before the index pass overwrites them
    int32_t* counts = tok_src;
    count_valid_kernel<<<cdiv(B * 32, 256), 256, 0, st>>>(mask, B, T, valid_only, counts);
    MVN_LAUNCH_CHECK();
    scan_kernel<<<1, 1024, 0, st>>>(counts, B, cu_seqlens);
    MVN_LAUNCH_CHECK();
    write_index_kernel<<<cdiv(B * 32, 256), 256, 0, st>>>(mask, B, T, valid_only, cu_seqlens, tok_src, keyvalid);
    MVN_LAUNCH_CHECK();
    fill_tail_kernel<<<296, 256, 0, st>>>(cu_seqlens, B, B * T, tok_src, keyvalid);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_embed_fwd(const float* x, const float* t, const int32_t* cu_seqlens, const int32_t* tok_src,
                             const float* div_term, const float* w, const float* b, const float* band_emb, int B, int T,
                             int E, int nband, float* out, void* stream) {
    return launch_embed_fwd(x, t, cu_seqlens, tok_src, div_term, w, b, band_emb, B, T, E, nband, out, (cudaStream_t)stream, DropCfg());
}

extern "C" int mvn_embed_bwd(const float* x, const int32_t* cu_seqlens, const int32_t* tok_src, const float* dout, int B, int T,
                             int E, int nband, float* dw, float* db, float* dband, void* workspace, size_t workspace_bytes,
                             void* stream) {
    MVN_CHECK_ARG(cu_seqlens && dw && db && workspace && (nband == 1 || dband), "embed_bwd: bad arguments");
    const size_t pstride = (size_t)(2 + nband) * E;
    if (workspace_bytes < (size_t)kSlabs * pstride * sizeof(float)) { set_error("embed_bwd: workspace too small"); return MVN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)workspace;
    MVN_TRY(launch_embed_bwd_partials(x, tok_src, dout, cu_seqlens + B, B * T, T, E, nband, part, pstride, 0, st));
    MVN_TRY(launch_reduce_partials(part, pstride, E, dw, 0, st));
    MVN_TRY(launch_reduce_partials(part + E, pstride, E, db, 0, st));
    if (nband > 1) MVN_TRY(launch_reduce_partials(part + 2 * E, pstride, (size_t)nband * E, dband, 0, st));
    return 0;
}

extern "C" int mvn_layernorm_bwd(const float* dY, const float* xhat, const float* rstd, const float* gamma, float* dZ,
                                 float* dgamma, float* dbeta, const int32_t* n_rows_dev, int M_cap, int E, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    MVN_CHECK_ARG(dgamma && dbeta && workspace, "layernorm_bwd: null outputs");
    const size_t pstride = 2 * (size_t)E;
    if (workspace_bytes < (size_t)kSlabs * pstride * sizeof(float)) { set_error("layernorm_bwd: workspace too small"); return MVN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)workspace;
    MVN_TRY(launch_ln_bwd(dY, xhat, rstd, gamma, dZ, n_rows_dev, M_cap, E, part, pstride, 0, E, st));
    MVN_TRY(launch_reduce_partials(part, pstride, E, dgamma, 0, st));
    MVN_TRY(launch_reduce_partials(part + E, pstride, E, dbeta, 0, st));
    return 0;
}

extern "C" int mvn_pool_fwd(const float* X, const int32_t* cu_seqlens, const uint8_t* keyvalid, int B, int T, int E, int agg,
                            float* pooled, int32_t* argmax, void* stream) {
    MVN_CHECK_ARG(X && cu_seqlens && pooled && B > 0 && E > 0, "pool_fwd: bad arguments");
    MVN_CHECK_ARG(agg == MVN_AGG_MEAN || (agg == MVN_AGG_MAX && argmax), "pool_fwd: agg=%d unsupported or argmax missing", agg);
    ProfScope prof(PROF_ROW, (cudaStream_t)stream);
    pool_fwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(X, cu_seqlens, keyvalid, T, E, agg, pooled, argmax);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_pool_bwd(const float* dpooled, const int32_t* cu_seqlens, const uint8_t* keyvalid, const int32_t* argmax, int B,
                            int T, int E, int agg, float* dX, void* stream) {
    (void)T;
    MVN_CHECK_ARG(dpooled && cu_seqlens && dX && B > 0 && E > 0, "pool_bwd: bad arguments");
    MVN_CHECK_ARG(agg == MVN_AGG_MEAN || (agg == MVN_AGG_MAX && argmax), "pool_bwd: agg=%d unsupported or argmax missing", agg);
    ProfScope prof(PROF_ROW, (cudaStream_t)stream);
    pool_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(dpooled, cu_seqlens, keyvalid, argmax, E, agg, dX);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_unpack_rows(const float* X, const int32_t* tok_src, const uint8_t* keyvalid, const int32_t* n_rows_dev, int BT,
                               int E, float* out, void* stream) {
    MVN_CHECK_ARG(X && tok_src && out && BT > 0 && E > 0, "unpack_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    MVN_CUDA(cudaMemsetAsync(out, 0, (size_t)BT * E * sizeof(float), st));
    unpack_rows_kernel<<<min(cdiv(BT, 8), num_sms() * 8), 256, 0, st>>>(X, tok_src, keyvalid, n_rows_dev, BT, E, out);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_pack_rows(const float* dense, const int32_t* tok_src, const uint8_t* keyvalid, const int32_t* n_rows_dev, int BT,
                             int E, float* X, void* stream) {
    MVN_CHECK_ARG(dense && tok_src && X && BT > 0 && E > 0, "pack_rows: bad arguments");
    pack_rows_kernel<<<min(cdiv(BT, 8), num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(dense, tok_src, keyvalid, n_rows_dev, BT, E, X);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_l2norm_fwd(const float* X, float* Y, float* norm, int B, int D, void* stream) {
    MVN_CHECK_ARG(X && Y && B > 0 && D > 0, "l2norm_fwd: bad arguments");
    l2norm_fwd_kernel<<<cdiv(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(X, Y, norm, B, D);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_l2norm_bwd(const float* dY, const float* Y, const float* norm, float* dX, int B, int D, void* stream) {
    MVN_CHECK_ARG(dY && Y && norm && dX && B > 0 && D > 0, "l2norm_bwd: bad arguments");
    l2norm_bwd_kernel<<<cdiv(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(dY, Y, norm, dX, B, D);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_relu_bwd(const float* dY, const float* H, int64_t n, float* dPre, void* stream) {
    MVN_CHECK_ARG(dY && H && dPre && n > 0, "relu_bwd: bad arguments");
    const long long blocks = (n + 255) / 256;
    relu_bwd_kernel<<<(int)(blocks < 4096 ? blocks : 4096), 256, 0, (cudaStream_t)stream>>>(dY, H, (size_t)n, dPre);
    MVN_LAUNCH_CHECK();
    return 0;
}
