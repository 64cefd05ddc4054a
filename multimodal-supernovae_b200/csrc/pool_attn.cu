// A6, agg="attn": nn.MultiheadAttention(emb, heads) with ONE learnable query per sequence over the zero-padded token
// tensor and NO key mask (src/transformer_utils.py:241-247).  kv is [B*T, 2E] = k|v (in_proj already applied, so padded
// rows carry k = b_k, v = b_v exactly like the reference); q is the projected query [E], identical for every sequence.
// One CTA per (sequence, head): scores over the T keys, block softmax, weighted sum of values.  Saves the probabilities.
#include "common.cuh"

namespace mvn {
namespace {

constexpr int QP_THREADS = 128;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < QP_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

template <int HD>
__global__ void __launch_bounds__(QP_THREADS) query_pool_fwd_kernel(const float* __restrict__ q, const float* __restrict__ kv, int T, int E, int H,
                                                                    float scale, float* __restrict__ out, float* __restrict__ probs) {
    __shared__ float red[QP_THREADS / 32];
    __shared__ float acc_s[QP_THREADS / 32][HD];
    const int b = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x;
    float qh[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) qh[d] = q[h * HD + d] * scale;
    const float* base = kv + (size_t)b * T * 2 * E + h * HD;
    float* pr = probs + ((size_t)b * H + h) * T;
    float mx = -INFINITY;
    for (int t = tid; t < T; t += QP_THREADS) {
        const float* kp = base + (size_t)t * 2 * E;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(kp + d);
            s = fmaf(qh[d], k4.x, fmaf(qh[d + 1], k4.y, fmaf(qh[d + 2], k4.z, fmaf(qh[d + 3], k4.w, s))));
        }
        pr[t] = s;
        mx = fmaxf(mx, s);
    }
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int t = tid; t < T; t += QP_THREADS) { const float p = expf(pr[t] - mx); pr[t] = p; sum += p; }
    sum = block_reduce(sum, red, false);
    const float inv = 1.0f / sum;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    for (int t = tid; t < T; t += QP_THREADS) {
        const float p = pr[t] * inv;
        pr[t] = p;
        const float* vp = base + (size_t)t * 2 * E + E;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(vp + d);
            o[d] = fmaf(p, v4.x, o[d]); o[d + 1] = fmaf(p, v4.y, o[d + 1]); o[d + 2] = fmaf(p, v4.z, o[d + 2]); o[d + 3] = fmaf(p, v4.w, o[d + 3]);
        }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = warp_sum(o[d]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int d = 0; d < HD; ++d) acc_s[tid >> 5][d] = o[d];
    }
    __syncthreads();
    if (tid < HD) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < QP_THREADS / 32; ++w) r += acc_s[w][tid];
        out[(size_t)b * E + h * HD + tid] = r;
    }
}

// dkv rows: dk_t = scale * ds_t * q_h, dv_t = p_t * do_h with ds_t = p_t (do_h . v_t - D), D = sum_t p_t (do_h . v_t);
// dq_part[b, h*HD + d] = scale * sum_t ds_t k_t[d]   (summed over b by rows_sum_kernel)
template <int HD>
__global__ void __launch_bounds__(QP_THREADS) query_pool_bwd_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                                    const float* __restrict__ probs, const float* __restrict__ dout, int T, int E, int H,
                                                                    float scale, float* __restrict__ dkv, float* __restrict__ dq_part) {
    __shared__ float red[QP_THREADS / 32];
    __shared__ float acc_s[QP_THREADS / 32][HD];
    const int b = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x;
    float qh[HD], go[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { qh[d] = q[h * HD + d]; go[d] = dout[(size_t)b * E + h * HD + d]; }
    const float* base = kv + (size_t)b * T * 2 * E + h * HD;
    float* dbase = dkv + (size_t)b * T * 2 * E + h * HD;
    const float* pr = probs + ((size_t)b * H + h) * T;
    float Dp = 0.f;
    for (int t = tid; t < T; t += QP_THREADS) {
        const float* vp = base + (size_t)t * 2 * E + E;
        float dp = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) dp = fmaf(go[d], vp[d], dp);
        Dp = fmaf(pr[t], dp, Dp);
    }
    Dp = block_reduce(Dp, red, false);
    float dq[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = 0.f;
    for (int t = tid; t < T; t += QP_THREADS) {
        const float* kp = base + (size_t)t * 2 * E;
        const float p = pr[t];
        float dp = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) dp = fmaf(go[d], kp[E + d], dp);
        const float ds = p * (dp - Dp) * scale;
        float* dk = dbase + (size_t)t * 2 * E;
#pragma unroll
        for (int d = 0; d < HD; ++d) { dk[d] = ds * qh[d]; dk[E + d] = p * go[d]; dq[d] = fmaf(ds, kp[d], dq[d]); }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = warp_sum(dq[d]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int d = 0; d < HD; ++d) acc_s[tid >> 5][d] = dq[d];
    }
    __syncthreads();
    if (tid < HD) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < QP_THREADS / 32; ++w) r += acc_s[w][tid];
        dq_part[(size_t)b * E + h * HD + tid] = r;
    }
}

__global__ void rows_sum_kernel(const float* __restrict__ in, int B, int E, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += in[(size_t)b * E + e];
    out[e] = s;
}


// ---- closed form (N4) ---------------------------------------------------------------------------------------------------
// All zeroed padded rows of one sequence are identical keys (k = b_k, v = b_v), and the single query is the same for every
// sequence.  With q = W_q query + b_q, u_h = W_{k,h}^T q_h and c_h = q_h . b_{k,h} the score of token t in head h is
// s_t = scale (u_h . x_t + c_h), every padded row scores scale c_h, and
//     o_h = sum_t p_t (W_{v,h} x_t + b_{v,h}) = W_{v,h} (sum_{valid t} p_t x_t) + b_{v,h}        (sum over ALL p is 1)
// with the softmax normalised over the valid scores plus (T - n) copies of the padded one.  One CTA per sequence reads its n valid
// rows of the token tensor and nothing else: no k|v projection of B*T rows, no probabilities in HBM.
constexpr int AP_THREADS = 128, AP_MAXE = 128, AP_MAXH = 8;

struct ApSaved {                 // carved from the caller's `saved` buffer
    float *q, *u, *c, *lse, *xbar, *o;
};
__host__ __device__ inline ApSaved ap_carve(float* base, int B, int E, int H) {
    ApSaved s;
    s.q = base; base += E;
    s.u = base; base += (size_t)H * E;
    s.c = base; base += AP_MAXH;
    s.lse = base; base += (size_t)B * H;
    s.xbar = base; base += (size_t)H * B * E;         // [H][B][E]
    s.o = base;                                       // [B][E]
    return s;
}

// q, u_h, c_h into shared memory (every CTA recomputes them: 2 E^2 MACs)
__device__ __forceinline__ void ap_prep(const float* __restrict__ query, const float* __restrict__ in_w, const float* __restrict__ in_b, int E, int H,
                                        float* qs, float* us, float* cs) {
    const int hd = E / H;
    for (int j = threadIdx.x; j < E; j += AP_THREADS) {
        float a = in_b[j];
        for (int e = 0; e < E; ++e) a = fmaf(in_w[(size_t)j * E + e], query[e], a);
        qs[j] = a;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += AP_THREADS)
        for (int h = 0; h < H; ++h) {
            float a = 0.f;
            for (int j = h * hd; j < (h + 1) * hd; ++j) a = fmaf(qs[j], in_w[(size_t)(E + j) * E + e], a);
            us[h * E + e] = a;
        }
    if (threadIdx.x < H) {
        float a = 0.f;
        for (int j = threadIdx.x * hd; j < (threadIdx.x + 1) * hd; ++j) a = fmaf(qs[j], in_b[E + j], a);
        cs[threadIdx.x] = a;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(AP_THREADS) attn_pool_fwd_kernel(const float* __restrict__ x, const unsigned char* __restrict__ mask,
                                                                   const float* __restrict__ query, const float* __restrict__ in_w,
                                                                   const float* __restrict__ in_b, const float* __restrict__ out_w,
                                                                   const float* __restrict__ out_b, int B, int T, int E, int H, float scale,
                                                                   float* __restrict__ out, float* __restrict__ saved) {
    extern __shared__ float ap_sm[];                 // scores [H][T]
    __shared__ float qs[AP_MAXE], us[AP_MAXH * AP_MAXE], cs[AP_MAXH], xb[AP_MAXH * AP_MAXE], os[AP_MAXE];
    __shared__ float red[AP_THREADS / 32][AP_MAXH], mh[AP_MAXH], zh[AP_MAXH];
    __shared__ int nvalid_s[AP_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, hd = E / H;
    const ApSaved sv = ap_carve(saved, B, E, H);
    ap_prep(query, in_w, in_b, E, H, qs, us, cs);
    if (b == 0) {
        for (int i = tid; i < E; i += AP_THREADS) sv.q[i] = qs[i];
        for (int i = tid; i < H * E; i += AP_THREADS) sv.u[i] = us[i];
        if (tid < H) sv.c[tid] = cs[tid];
    }
    const float* xb_ = x + (size_t)b * T * E;
    const unsigned char* mb = mask + (size_t)b * T;
    // scores of the valid tokens (a warp per token), running max per head
    float mx[AP_MAXH];
    for (int h = 0; h < H; ++h) mx[h] = -INFINITY;
    int nv = 0;
    for (int t = warp; t < T; t += AP_THREADS / 32) {
        if (!mb[t]) { if (lane == 0) for (int h = 0; h < H; ++h) ap_sm[h * T + t] = -INFINITY; continue; }
        ++nv;
        float d[AP_MAXH];
        for (int h = 0; h < H; ++h) d[h] = 0.f;
        for (int e = lane; e < E; e += 32) {
            const float xv = xb_[(size_t)t * E + e];
            for (int h = 0; h < H; ++h) d[h] = fmaf(xv, us[h * E + e], d[h]);
        }
        for (int h = 0; h < H; ++h) {
            const float sc = (warp_sum(d[h]) + cs[h]) * scale;
            mx[h] = fmaxf(mx[h], sc);
            if (lane == 0) ap_sm[h * T + t] = sc;
        }
    }
    if (lane == 0) { for (int h = 0; h < H; ++h) red[warp][h] = mx[h]; nvalid_s[warp] = nv; }
    __syncthreads();
    int n = 0;
    for (int w = 0; w < AP_THREADS / 32; ++w) n += nvalid_s[w];
    const int npad = T - n;
    if (tid < H) {
        float m = npad > 0 ? cs[tid] * scale : -INFINITY;            // the padded rows' common score
        for (int w = 0; w < AP_THREADS / 32; ++w) m = fmaxf(m, red[w][tid]);
        mh[tid] = m;
    }
    __syncthreads();
    // denominators (fixed order over t per head: one warp per head pair is enough at H = 2, T <= 1024)
    if (warp < H) {
        const int h = warp;
        float z = 0.f;
        for (int t = lane; t < T; t += 32) z += expf(ap_sm[h * T + t] - mh[h]);      // exp(-inf) = 0 on padding
        z = warp_sum(z);
        if (npad > 0) z += (float)npad * expf(cs[h] * scale - mh[h]);
        if (lane == 0) zh[h] = z;
    }
    __syncthreads();
    // xbar_h[e] = sum_t p_t x_t[e]: thread (e, token group)
    {
        const int ngrp = AP_THREADS / E > 0 ? AP_THREADS / E : 1;
        float* part = us;                                // u is no longer needed in this kernel: [ngrp][H][E] partials fit (ngrp*E <= 128)
        for (int e0 = 0; e0 < E; e0 += AP_THREADS) {
            const int e = e0 + (tid % (E < AP_THREADS ? E : AP_THREADS)), grp = E < AP_THREADS ? tid / E : 0;
            float acc[AP_MAXH];
            for (int h = 0; h < H; ++h) acc[h] = 0.f;
            if (e < E && grp < ngrp)
                for (int t = grp; t < T; t += ngrp) {
                    if (!mb[t]) continue;
                    const float xv = xb_[(size_t)t * E + e];
                    for (int h = 0; h < H; ++h) acc[h] = fmaf(expf(ap_sm[h * T + t] - mh[h]), xv, acc[h]);
                }
            __syncthreads();
            if (e < E && grp < ngrp) for (int h = 0; h < H; ++h) part[(grp * H + h) * E + e] = acc[h];
            __syncthreads();
            if (e < E && grp == 0)
                for (int h = 0; h < H; ++h) {
                    float a = 0.f;
                    for (int g2 = 0; g2 < ngrp; ++g2) a += part[(g2 * H + h) * E + e];
                    a /= zh[h];
                    xb[h * E + e] = a;
                    sv.xbar[((size_t)h * B + b) * E + e] = a;
                }
        }
    }
    __syncthreads();
    for (int j = tid; j < E; j += AP_THREADS) {         // o = W_v xbar_h + b_v
        const int h = j / hd;
        float a = in_b[2 * E + j];
        for (int e = 0; e < E; ++e) a = fmaf(in_w[(size_t)(2 * E + j) * E + e], xb[h * E + e], a);
        os[j] = a;
        sv.o[(size_t)b * E + j] = a;
    }
    if (tid < H) sv.lse[(size_t)b * H + tid] = mh[tid] + logf(zh[tid]);
    __syncthreads();
    for (int i = tid; i < E; i += AP_THREADS) {         // out = W_o o + b_o
        float a = out_b[i];
        for (int j = 0; j < E; ++j) a = fmaf(out_w[(size_t)i * E + j], os[j], a);
        out[(size_t)b * E + i] = a;
    }
}

// per sequence: do = W_o^T dout, dxbar_h = W_{v,h}^T do_h, then per valid token p_t = exp(s_t - lse), ds_t = p_t (dxbar_h . x_t - D_h)
// with D_h = dxbar_h . xbar_h, dx_t = sum_h p_t dxbar_h + scale ds_t u_h; partials of du_h, dc_h per sequence
__global__ void __launch_bounds__(AP_THREADS) attn_pool_bwd_kernel(const float* __restrict__ x, const unsigned char* __restrict__ mask,
                                                                   const float* __restrict__ in_w, const float* __restrict__ out_w,
                                                                   const float* __restrict__ saved, const float* __restrict__ dout, int B, int T,
                                                                   int E, int H, float scale, float* __restrict__ dx, float* __restrict__ dO,
                                                                   float* __restrict__ dxbar, float* __restrict__ part) {
    __shared__ float us[AP_MAXH * AP_MAXE], cs[AP_MAXH], dos[AP_MAXE], dxb[AP_MAXH * AP_MAXE], Dh[AP_MAXH], lse[AP_MAXH];
    __shared__ float duw[AP_THREADS / 32][AP_MAXH * AP_MAXE], dcw[AP_THREADS / 32][AP_MAXH];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, hd = E / H;
    const ApSaved sv = ap_carve(const_cast<float*>(saved), B, E, H);
    for (int i = tid; i < H * E; i += AP_THREADS) us[i] = sv.u[i];
    if (tid < H) { cs[tid] = sv.c[tid]; lse[tid] = sv.lse[(size_t)b * H + tid]; }
    for (int j = tid; j < E; j += AP_THREADS) {
        float a = 0.f;
        for (int i = 0; i < E; ++i) a = fmaf(out_w[(size_t)i * E + j], dout[(size_t)b * E + i], a);
        dos[j] = a;
        dO[(size_t)b * E + j] = a;
    }
    __syncthreads();
    for (int e = tid; e < E; e += AP_THREADS)
        for (int h = 0; h < H; ++h) {
            float a = 0.f;
            for (int j = h * hd; j < (h + 1) * hd; ++j) a = fmaf(in_w[(size_t)(2 * E + j) * E + e], dos[j], a);
            dxb[h * E + e] = a;
            dxbar[((size_t)h * B + b) * E + e] = a;
        }
    __syncthreads();
    if (warp < H) {
        float d = 0.f;
        for (int e = lane; e < E; e += 32) d = fmaf(dxb[warp * E + e], sv.xbar[((size_t)warp * B + b) * E + e], d);
        d = warp_sum(d);
        if (lane == 0) Dh[warp] = d;
    }
    for (int i = lane; i < H * E; i += 32) duw[warp][i] = 0.f;
    __syncthreads();
    const float* xb_ = x + (size_t)b * T * E;
    float* dxb_ = dx + (size_t)b * T * E;
    const unsigned char* mb = mask + (size_t)b * T;
    float dcl[AP_MAXH];
    for (int h = 0; h < H; ++h) dcl[h] = 0.f;
    int nv = 0;
    for (int t = warp; t < T; t += AP_THREADS / 32) {
        if (!mb[t]) { for (int e = lane; e < E; e += 32) dxb_[(size_t)t * E + e] = 0.f; continue; }
        ++nv;
        float ds[AP_MAXH], pp[AP_MAXH];
        for (int h = 0; h < H; ++h) {
            float d1 = 0.f, d2 = 0.f;
            for (int e = lane; e < E; e += 32) {
                const float xv = xb_[(size_t)t * E + e];
                d1 = fmaf(xv, us[h * E + e], d1);
                d2 = fmaf(xv, dxb[h * E + e], d2);
            }
            d1 = warp_sum(d1); d2 = warp_sum(d2);
            const float p = expf((d1 + cs[h]) * scale - lse[h]);
            pp[h] = p;
            ds[h] = p * (d2 - Dh[h]) * scale;
            dcl[h] += ds[h];
        }
        for (int e = lane; e < E; e += 32) {
            const float xv = xb_[(size_t)t * E + e];
            float a = 0.f;
            for (int h = 0; h < H; ++h) {
                a = fmaf(pp[h], dxb[h * E + e], a);
                a = fmaf(ds[h], us[h * E + e], a);
                duw[warp][h * E + e] = fmaf(ds[h], xv, duw[warp][h * E + e]);      // each lane owns its e's: no race
            }
            dxb_[(size_t)t * E + e] = a;
        }
    }
    if (lane == 0) for (int h = 0; h < H; ++h) dcw[warp][h] = dcl[h];
    __shared__ int nvs[AP_THREADS / 32];
    if (lane == 0) nvs[warp] = nv;
    __syncthreads();
    float* pr = part + (size_t)b * (H * E + H);
    for (int i = tid; i < H * E; i += AP_THREADS) {
        float a = 0.f;
        for (int w = 0; w < AP_THREADS / 32; ++w) a += duw[w][i];
        pr[i] = a;
    }
    if (tid < H) {
        int n = 0;
        float a = 0.f;
        for (int w = 0; w < AP_THREADS / 32; ++w) { a += dcw[w][tid]; n += nvs[w]; }
        const int npad = T - n;
        if (npad > 0) a += (float)npad * expf(cs[tid] * scale - lse[tid]) * (0.f - Dh[tid]) * scale;      // the padded rows: x = 0
        pr[H * E + tid] = a;
    }
}

// out[i] = sum_b in[b][i]  (rows_sum_kernel), then the parameter gradients that only depend on batch sums
// C[i][j] (+ bias column) = sum_b A[b][i] * Bm[b][j]:  dW_o = dout^T o,  dW_{v,h} = do_h^T xbar_h
__global__ void __launch_bounds__(128) batch_outer_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int B, int N,
                                                          float* __restrict__ C, int ldc, float* __restrict__ bias) {
    const int i = blockIdx.x;
    for (int j = threadIdx.x; j < N; j += 128) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a = fmaf(A[(size_t)b * lda + i], Bm[(size_t)b * ldb + j], a);
        C[(size_t)i * ldc + j] = a;
    }
    if (bias && threadIdx.x == 0) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += A[(size_t)b * lda + i];
        bias[i] = a;
    }
}
// gradients of the query path from du_h, dc_h (already scaled): k rows of in_proj, q rows of in_proj, the query
__global__ void __launch_bounds__(AP_THREADS) attn_pool_prep_bwd_kernel(const float* __restrict__ query, const float* __restrict__ in_w,
                                                                        const float* __restrict__ in_b, const float* __restrict__ saved,
                                                                        const float* __restrict__ duc, int B, int E, int H,
                                                                        float* __restrict__ d_in_w, float* __restrict__ d_in_b, float* __restrict__ dquery) {
    __shared__ float dq[AP_MAXE];
    const ApSaved sv = ap_carve(const_cast<float*>(saved), B, E, H);
    const int hd = E / H, tid = threadIdx.x;
    for (int j = tid; j < E; j += AP_THREADS) {
        const int h = j / hd;
        const float qj = sv.q[j], dc = duc[H * E + h];
        float a = in_b[E + j] * dc;
        for (int e = 0; e < E; ++e) {
            const float du = duc[h * E + e];
            a = fmaf(in_w[(size_t)(E + j) * E + e], du, a);
            d_in_w[(size_t)(E + j) * E + e] = qj * du;                         // dW_k
        }
        d_in_b[E + j] = qj * dc;                                                // db_k
        dq[j] = a;
        d_in_b[j] = a;                                                          // db_q
        for (int e = 0; e < E; ++e) d_in_w[(size_t)j * E + e] = a * query[e];   // dW_q
    }
    __syncthreads();
    for (int e = tid; e < E; e += AP_THREADS) {
        float a = 0.f;
        for (int j = 0; j < E; ++j) a = fmaf(in_w[(size_t)j * E + e], dq[j], a);
        dquery[e] = a;
    }
}

}  // namespace
}  // namespace mvn

using namespace mvn;

extern "C" int mvn_query_pool_fwd(const float* q, const float* kv, int B, int T, int E, int H, float* out, float* probs, void* stream) {
    MVN_CHECK_ARG(q && kv && out && probs && B > 0 && T > 0 && E > 0 && H > 0 && E % H == 0, "query_pool_fwd: bad arguments");
    MVN_CHECK_ARG(aligned16(kv), "query_pool_fwd: kv must be 16-byte aligned");
    const int hd = E / H;
    const float scale = 1.0f / sqrtf((float)hd);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ROW, st);
    switch (hd) {
        case 8: query_pool_fwd_kernel<8><<<B * H, QP_THREADS, 0, st>>>(q, kv, T, E, H, scale, out, probs); break;
        case 16: query_pool_fwd_kernel<16><<<B * H, QP_THREADS, 0, st>>>(q, kv, T, E, H, scale, out, probs); break;
        case 32: query_pool_fwd_kernel<32><<<B * H, QP_THREADS, 0, st>>>(q, kv, T, E, H, scale, out, probs); break;
        case 64: query_pool_fwd_kernel<64><<<B * H, QP_THREADS, 0, st>>>(q, kv, T, E, H, scale, out, probs); break;
        default: MVN_UNSUPPORTED(false, "query_pool: head dim %d not in {8,16,32,64}", hd);
    }
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_query_pool_bwd(const float* q, const float* kv, const float* probs, const float* dout, int B, int T, int E, int H,
                                  float* dkv, float* dq, void* workspace, size_t workspace_bytes, void* stream) {
    MVN_CHECK_ARG(q && kv && probs && dout && dkv && dq && workspace && B > 0 && T > 0 && E > 0 && H > 0 && E % H == 0, "query_pool_bwd: bad arguments");
    if (workspace_bytes < (size_t)B * E * sizeof(float)) { set_error("query_pool_bwd: workspace %zu < %zu", workspace_bytes, (size_t)B * E * 4); return MVN_E_WORKSPACE; }
    const int hd = E / H;
    const float scale = 1.0f / sqrtf((float)hd);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ROW, st);
    float* part = (float*)workspace;
    switch (hd) {
        case 8: query_pool_bwd_kernel<8><<<B * H, QP_THREADS, 0, st>>>(q, kv, probs, dout, T, E, H, scale, dkv, part); break;
        case 16: query_pool_bwd_kernel<16><<<B * H, QP_THREADS, 0, st>>>(q, kv, probs, dout, T, E, H, scale, dkv, part); break;
        case 32: query_pool_bwd_kernel<32><<<B * H, QP_THREADS, 0, st>>>(q, kv, probs, dout, T, E, H, scale, dkv, part); break;
        case 64: query_pool_bwd_kernel<64><<<B * H, QP_THREADS, 0, st>>>(q, kv, probs, dout, T, E, H, scale, dkv, part); break;
        default: MVN_UNSUPPORTED(false, "query_pool: head dim %d not in {8,16,32,64}", hd);
    }
    MVN_LAUNCH_CHECK();
    rows_sum_kernel<<<cdiv(E, 128), 128, 0, st>>>(part, B, E, dq);
    MVN_LAUNCH_CHECK();
    return 0;
}


// ---- closed-form agg="attn" pooling (N4): tokens [B,T,E] (zero on padding) + bool mask -> pooled [B,E] -----------------------
extern "C" size_t mvn_attn_pool_saved_bytes(int B, int E, int H) {
    return ((size_t)E + (size_t)H * E + AP_MAXH + (size_t)B * H + (size_t)H * B * E + (size_t)B * E) * sizeof(float);
}
extern "C" size_t mvn_attn_pool_bwd_workspace_bytes(int B, int E, int H) {
    return ((size_t)B * E + (size_t)H * B * E + (size_t)B * (H * E + H) + (size_t)H * E + H) * sizeof(float);
}
static int attn_pool_check(int B, int T, int E, int H) {
    MVN_CHECK_ARG(B > 0 && T > 0 && E > 0 && H > 0 && E % H == 0, "attn_pool: bad dims");
    MVN_UNSUPPORTED(E <= AP_MAXE && H <= AP_MAXH && (E >= AP_THREADS || AP_THREADS % E == 0), "attn_pool: emb %d (<= 128, a divisor of 128) / heads %d (<= 8)", E, H);
    MVN_UNSUPPORTED((size_t)H * T * sizeof(float) <= 96 * 1024, "attn_pool: T=%d too long for the score buffer", T);
    return 0;
}
extern "C" int mvn_attn_pool_fwd(const float* x, const unsigned char* mask, const float* query, const float* in_w, const float* in_b,
                                 const float* out_w, const float* out_b, int B, int T, int E, int H, float* out, float* saved, size_t saved_bytes,
                                 void* stream) {
    MVN_CHECK_ARG(x && mask && query && in_w && in_b && out_w && out_b && out && saved, "attn_pool_fwd: null pointer");
    MVN_TRY(attn_pool_check(B, T, E, H));
    if (saved_bytes < mvn_attn_pool_saved_bytes(B, E, H)) { set_error("attn_pool_fwd: saved buffer %zu < %zu", saved_bytes, mvn_attn_pool_saved_bytes(B, E, H)); return MVN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ROW, st);
    const size_t sm = (size_t)H * T * sizeof(float);
    static bool configured = false;
    if (!configured) { MVN_CUDA(cudaFuncSetAttribute(attn_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); configured = true; }
    attn_pool_fwd_kernel<<<B, AP_THREADS, sm, st>>>(x, mask, query, in_w, in_b, out_w, out_b, B, T, E, H, 1.0f / sqrtf((float)(E / H)), out, saved);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_attn_pool_bwd(const float* x, const unsigned char* mask, const float* query, const float* in_w, const float* in_b,
                                 const float* out_w, const float* saved, const float* dout, int B, int T, int E, int H, float* dx, float* dquery,
                                 float* d_in_w, float* d_in_b, float* d_out_w, float* d_out_b, void* workspace, size_t workspace_bytes, void* stream) {
    MVN_CHECK_ARG(x && mask && query && in_w && in_b && out_w && saved && dout && dx && dquery && d_in_w && d_in_b && d_out_w && d_out_b && workspace,
                  "attn_pool_bwd: null pointer");
    MVN_TRY(attn_pool_check(B, T, E, H));
    if (workspace_bytes < mvn_attn_pool_bwd_workspace_bytes(B, E, H)) { set_error("attn_pool_bwd: workspace too small"); return MVN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ROW, st);
    float* dO = (float*)workspace;                       // [B][E]
    float* dxbar = dO + (size_t)B * E;                   // [H][B][E]
    float* part = dxbar + (size_t)H * B * E;             // [B][H*E + H]
    float* duc = part + (size_t)B * (H * E + H);         // [H*E + H]
    const float scale = 1.0f / sqrtf((float)(E / H));
    const ApSaved sv = ap_carve(const_cast<float*>(saved), B, E, H);
    const int hd = E / H;
    attn_pool_bwd_kernel<<<B, AP_THREADS, 0, st>>>(x, mask, in_w, out_w, saved, dout, B, T, E, H, scale, dx, dO, dxbar, part);
    MVN_LAUNCH_CHECK();
    rows_sum_kernel<<<cdiv(H * E + H, 128), 128, 0, st>>>(part, B, H * E + H, duc);
    MVN_LAUNCH_CHECK();
    attn_pool_prep_bwd_kernel<<<1, AP_THREADS, 0, st>>>(query, in_w, in_b, saved, duc, B, E, H, d_in_w, d_in_b, dquery);
    MVN_LAUNCH_CHECK();
    batch_outer_kernel<<<E, 128, 0, st>>>(dout, E, sv.o, E, B, E, d_out_w, E, d_out_b);                                      // dW_o, db_o
    MVN_LAUNCH_CHECK();
    for (int h = 0; h < H; ++h) {                                                                                            // dW_v, db_v
        batch_outer_kernel<<<hd, 128, 0, st>>>(dO + h * hd, E, sv.xbar + (size_t)h * B * E, E, B, E, d_in_w + (size_t)(2 * E + h * hd) * E, E,
                                               d_in_b + 2 * E + h * hd);
        MVN_LAUNCH_CHECK();
    }
    return 0;
}
