// fp32 FFMA GEMM family (prec==0 tier): Y = X W^T with fused bias/activation/residual/LayerNorm epilogues,
// input-gradient GEMMs, and slab-partial weight-gradient GEMMs.  Tensor-core (tcgen05) variants live in
// gemm_tc.cu and are selected by prec==1 inside launch_gemm / launch_wgrad_partials.
#include "common.cuh"

namespace mvn {

int launch_gemm_tc(const float* A, const float* Bm, float* C, const int32_t* n_rows_dev, int M_cap, int N, int K,
                   bool b_is_nk, const GemmEpilogue& ep, cudaStream_t st);   // gemm_tc.cu; returns MVN_E_UNSUPPORTED if shape not covered

int launch_wgrad_tc(const float* dY, const float* X, const int32_t* n_rows_dev, int M_cap, int N, int K, float* partial, size_t pstride,
                    size_t woff, long long boff, cudaStream_t st);             // gemm_tc.cu

namespace {

constexpr int BK = 16;

struct GemmArgs {
    const float* A; const float* B; float* C;
    const int32_t* n_rows_dev;
    int M_cap, N, K;
    int b_is_nk;
    int vecA, vecB, vecC;
    GemmEpilogue ep;
};

__device__ __forceinline__ float4 ld4_guard(const float* __restrict__ p, int row, int col, int ld, int nrows, int ncols, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nrows) {
        const float* q = p + (size_t)row * ld + col;
        if (vec && col + 3 < ncols) {
            v = *reinterpret_cast<const float4*>(q);
        } else {
            if (col + 0 < ncols) v.x = q[0];
            if (col + 1 < ncols) v.y = q[1];
            if (col + 2 < ncols) v.z = q[2];
            if (col + 3 < ncols) v.w = q[3];
        }
    }
    return v;
}

// KS > 1: in-CTA split-K for the per-sample heads (M = batch <= 4096 rows, K up to 1024): KS thread groups of NT threads each
// own a K range with their own staging tiles, and group 0 adds the partial accumulators in group order (deterministic) before the
// epilogue.  Without it a K = 1024, N = 32 head is 64 single-warp CTAs walking 64 K-steps each (80 - 190 us per launch).
template <int BM, int BN, bool LN, int KS = 1>
__global__ void __launch_bounds__((BM / 4) * (BN / 4) * KS) gemm_kernel(const GemmArgs a) {
    constexpr int NT = (BM / 4) * (BN / 4);
    constexpr int TX = BN / 4;
    __shared__ __align__(16) float As_[KS][BK][BM + 4];
    __shared__ __align__(16) float Bs_[KS][BK][BN + 4];
    __shared__ __align__(16) float red_[KS > 1 ? (KS - 1) * NT * 16 : 4];

    const int rows = a.n_rows_dev ? min(*a.n_rows_dev, a.M_cap) : a.M_cap;
    const int m0 = blockIdx.x * BM;
    if (m0 >= rows) return;
    const int n0 = blockIdx.y * BN;
    const int ks = threadIdx.x / NT;
    const int tid = threadIdx.x % NT;
    const int tx = tid % TX, ty = tid / TX;
    const int N = a.N, K = a.K;
    float (*As)[BM + 4] = As_[ks];
    float (*Bs)[BN + 4] = Bs_[ks];

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int kper = KS == 1 ? K : ((K + KS * BK - 1) / (KS * BK)) * BK;          // K range of a group, whole BK steps
    const int kbeg = ks * kper;
    for (int it = 0; it < kper; it += BK) {                                      // same trip count in every group: the barriers are CTA-wide
        const int k0 = kbeg + it;
        // A tile: BM rows x BK cols, K contiguous
        for (int idx = tid; idx < BM * (BK / 4); idx += NT) {
            const int r = idx / (BK / 4), kc = (idx % (BK / 4)) * 4;
            const float4 v = ld4_guard(a.A, m0 + r, k0 + kc, K, rows, K, a.vecA);
            As[kc + 0][r] = v.x; As[kc + 1][r] = v.y; As[kc + 2][r] = v.z; As[kc + 3][r] = v.w;
        }
        if (a.b_is_nk) {    // B[n][k]
            for (int idx = tid; idx < BN * (BK / 4); idx += NT) {
                const int r = idx / (BK / 4), kc = (idx % (BK / 4)) * 4;
                const float4 v = ld4_guard(a.B, n0 + r, k0 + kc, K, N, K, a.vecB);
                Bs[kc + 0][r] = v.x; Bs[kc + 1][r] = v.y; Bs[kc + 2][r] = v.z; Bs[kc + 3][r] = v.w;
            }
        } else {            // B[k][n]
            for (int idx = tid; idx < BK * (BN / 4); idx += NT) {
                const int kr = idx / (BN / 4), nc = (idx % (BN / 4)) * 4;
                const float4 v = ld4_guard(a.B, k0 + kr, n0 + nc, N, K, N, a.vecB);
                *reinterpret_cast<float4*>(&Bs[kr][nc]) = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
    if constexpr (KS > 1) {
        if (ks > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(&red_[((ks - 1) * NT + tid) * 16 + 4 * i]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        if (ks > 0) return;
#pragma unroll
        for (int g = 0; g < KS - 1; ++g)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(&red_[(g * NT + tid) * 16 + 4 * i]);
                acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
            }
    }

    const GemmEpilogue& ep = a.ep;
    const int nb = n0 + tx * 4;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (nb + j < N) bias[j] = ep.bias[nb + j];
    }

    if constexpr (LN) {
        float g[4], be[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { g[j] = ep.gamma[nb + j]; be[j] = ep.beta[nb + j]; }
        const float invn = 1.0f / (float)N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            const bool live = m < rows;
            float v[4];
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && ep.addend) r4 = *reinterpret_cast<const float4*>(ep.addend + (size_t)m * N + nb);
            v[0] = acc[i][0] + bias[0] + r4.x; v[1] = acc[i][1] + bias[1] + r4.y;
            v[2] = acc[i][2] + bias[2] + r4.z; v[3] = acc[i][3] + bias[3] + r4.w;
            const float mean = group_sum<TX>(v[0] + v[1] + v[2] + v[3]) * invn;
            float d[4], sq = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { d[j] = v[j] - mean; sq = fmaf(d[j], d[j], sq); }
            const float var = group_sum<TX>(sq) * invn;
            const float rs = rsqrtf(var + ep.eps);
            if (live) {
                float4 xh = make_float4(d[0] * rs, d[1] * rs, d[2] * rs, d[3] * rs);
                float4 y = make_float4(fmaf(xh.x, g[0], be[0]), fmaf(xh.y, g[1], be[1]), fmaf(xh.z, g[2], be[2]), fmaf(xh.w, g[3], be[3]));
                if (ep.drop.thresh) {
                    const uint32_t rk = drop_rowkey(ep.drop, (uint32_t)m);
                    y.x *= drop_scale(ep.drop, rk, nb); y.y *= drop_scale(ep.drop, rk, nb + 1); y.z *= drop_scale(ep.drop, rk, nb + 2); y.w *= drop_scale(ep.drop, rk, nb + 3);
                }
                *reinterpret_cast<float4*>(a.C + (size_t)m * N + nb) = y;
                if (ep.xhat) *reinterpret_cast<float4*>(ep.xhat + (size_t)m * N + nb) = xh;
                if (ep.rstd && tx == 0) ep.rstd[m] = rs;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m >= rows) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = acc[i][j] + bias[j];
                const int n = nb + j;
                if (n < N) {
                    if (ep.addend) v[j] += ep.addend[(size_t)m * N + n];
                    if (ep.act == MVN_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
                    else if (ep.act == MVN_ACT_GELU) v[j] = gelu_erf(v[j]);
                    if (ep.dact == 1) v[j] = ep.act_src[(size_t)m * N + n] > 0.f ? v[j] : 0.f;
                    else if (ep.dact == 2) v[j] *= gelu_erf_grad(ep.act_src[(size_t)m * N + n]);
                }
            }
            if (ep.drop.thresh) {       // dropout after the activation (forward) / mask of the forward's dropout (dact: backward)
                const uint32_t rk = drop_rowkey(ep.drop, (uint32_t)m);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] *= drop_scale(ep.drop, rk, (uint32_t)(nb + j));
            }
            float* c = a.C + (size_t)m * N + nb;
            if (a.vecC && nb + 3 < N) {
                *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (nb + j < N) c[j] = v[j];
            }
        }
    }
}

template <int BM, int BN, bool LN, int KS = 1>
int launch_t(const GemmArgs& a, cudaStream_t st) {
    dim3 grid(cdiv(a.M_cap, BM), cdiv(a.N, BN));
    gemm_kernel<BM, BN, LN, KS><<<grid, (BM / 4) * (BN / 4) * KS, 0, st>>>(a);
    MVN_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// weight gradient partials: tile 64(n) x 64(k), row-chunks of 16 interleaved over kSlabs slabs
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                                                    const int32_t* n_rows_dev, int M_cap, int N, int K,
                                                    float* __restrict__ partial, size_t pstride, size_t woff, long long boff,
                                                    int vecY, int vecX) {
    __shared__ __align__(16) float Ys[16][64 + 4];
    __shared__ __align__(16) float Xs[16][64 + 4];
    const int rows = n_rows_dev ? min(*n_rows_dev, M_cap) : M_cap;
    const int s = blockIdx.x;
    const int n0 = blockIdx.y * 64, k0 = blockIdx.z * 64;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    const bool do_bias = (boff >= 0) && blockIdx.z == 0;
    const int nchunks = (rows + 15) / 16;
    const int lr = tid / 16, lc = (tid % 16) * 4;
    for (int c = s; c < nchunks; c += gridDim.x) {
        const int r = c * 16 + lr;
        *reinterpret_cast<float4*>(&Ys[lr][lc]) = ld4_guard(dY, r, n0 + lc, N, rows, N, vecY);
        *reinterpret_cast<float4*>(&Xs[lr][lc]) = ld4_guard(X, r, k0 + lc, K, rows, K, vecX);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const float4 av = *reinterpret_cast<const float4*>(&Ys[rr][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Xs[rr][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        if (do_bias && tid < 64) {
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) bsum += Ys[rr][tid];
        }
        __syncthreads();
    }
    float* p = partial + (size_t)s * pstride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) p[woff + (size_t)n * K + k] = acc[i][j];
        }
    }
    if (do_bias && tid < 64 && n0 + tid < N) p[boff + n0 + tid] = bsum;
}

// out[i] = sum over slabs of partial[s][i]: a block owns 32 consecutive columns, its 8 warps split the slabs (16 independent
// loads in flight per thread), fixed summation order -> deterministic.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, size_t pstride, size_t n,
                                                              int nslabs, float* __restrict__ out, int accumulate, size_t in_lstride, size_t out_lstride) {
    __shared__ float red[8][33];
    pdl_trigger();
    pdl_wait();
    partial += (size_t)blockIdx.y * in_lstride;              // blockIdx.y: independent range (a layer)
    out += (size_t)blockIdx.y * out_lstride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t i = (size_t)blockIdx.x * 32 + lane;
    float acc = 0.f;
    if (i < n) {
        const int per = (nslabs + 7) / 8;
        const int s0 = warp * per, s1 = min(nslabs, s0 + per);
        const float* p = partial + i;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int s = s0;
        for (; s + 3 < s1; s += 4) {
            a0 += p[(size_t)(s + 0) * pstride]; a1 += p[(size_t)(s + 1) * pstride];
            a2 += p[(size_t)(s + 2) * pstride]; a3 += p[(size_t)(s + 3) * pstride];
        }
        for (; s < s1; ++s) a0 += p[(size_t)s * pstride];
        acc = (a0 + a1) + (a2 + a3);
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && i < n) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][lane];
        out[i] = accumulate ? out[i] + v : v;
    }
}

}  // namespace

int launch_gemm(const float* A, const float* Bm, float* C, const int32_t* n_rows_dev, int M_cap, int N, int K,
                bool b_is_nk, const GemmEpilogue& ep, int prec, cudaStream_t st) {
    MVN_CHECK_ARG(A && Bm && C && M_cap > 0 && N > 0 && K > 0, "gemm: null pointer or non-positive size (M=%d N=%d K=%d)", M_cap, N, K);
    ProfScope prof(PROF_GEMM, st);
    if (prec >= 1 && !(ep.drop.thresh && !ep.gamma)) {     // the tensor-core epilogue applies dropout only after LayerNorm
        int r = launch_gemm_tc(A, Bm, C, n_rows_dev, M_cap, N, K, b_is_nk, ep, st);
        if (r != MVN_E_UNSUPPORTED) { if (r == 0) count_tier(TIER_TC); return r; }     // shapes the tensor-core kernel does not cover use the FFMA kernel
    }
    count_tier(TIER_FFMA);
    GemmArgs a;
    a.A = A; a.B = Bm; a.C = C; a.n_rows_dev = n_rows_dev; a.M_cap = M_cap; a.N = N; a.K = K; a.b_is_nk = b_is_nk ? 1 : 0;
    a.vecA = (K % 4 == 0) && aligned16(A);
    a.vecB = b_is_nk ? ((K % 4 == 0) && aligned16(Bm)) : ((N % 4 == 0) && aligned16(Bm));
    a.vecC = (N % 4 == 0) && aligned16(C);
    a.ep = ep;
    const bool ln = ep.gamma != nullptr;
    if (ln) {
        MVN_CHECK_ARG(ep.beta && aligned16(C) && (!ep.addend || aligned16(ep.addend)) && (!ep.xhat || aligned16(ep.xhat)),
                      "gemm+LN: beta missing or misaligned buffers");
        switch (N) {
            case 16: return launch_t<64, 16, true>(a, st);
            case 32: return launch_t<64, 32, true>(a, st);
            case 64: return launch_t<64, 64, true>(a, st);
            case 128: return launch_t<32, 128, true>(a, st);
            default: MVN_UNSUPPORTED(false, "linear+residual+LayerNorm needs N in {16,32,64,128}, got %d", N);
        }
    }
    if (M_cap <= 4096) {      // per-sample heads (M = batch): 16-row tiles give 4x the CTAs of the 64-row ones (a K = 1024 head ran on 16 SMs)
        if (N <= 32) return K >= 256 ? launch_t<16, 32, false, 8>(a, st) : launch_t<16, 32, false>(a, st);
        return K >= 256 ? launch_t<16, 64, false, 4>(a, st) : launch_t<16, 64, false>(a, st);
    }
    if (N <= 16) return launch_t<64, 16, false>(a, st);
    if (N <= 32) return launch_t<64, 32, false>(a, st);
    return launch_t<64, 64, false>(a, st);
}

int launch_wgrad_partials(const float* dY, const float* X, const int32_t* n_rows_dev, int M_cap, int N, int K,
                          float* partial, size_t pstride, size_t woff, long long boff, int prec, cudaStream_t st) {
    MVN_CHECK_ARG(dY && X && partial && M_cap > 0 && N > 0 && K > 0, "wgrad: null pointer or non-positive size");
    ProfScope prof(PROF_WGRAD, st);
    if (prec >= 1) {
        const int r = launch_wgrad_tc(dY, X, n_rows_dev, M_cap, N, K, partial, pstride, woff, boff, st);
        if (r != MVN_E_UNSUPPORTED) { if (r == 0) count_tier(TIER_TC); return r; }
    }
    count_tier(TIER_FFMA);
    dim3 grid(kSlabs, cdiv(N, 64), cdiv(K, 64));
    wgrad_kernel<<<grid, 256, 0, st>>>(dY, X, n_rows_dev, M_cap, N, K, partial, pstride, woff, boff,
                                       (N % 4 == 0) && aligned16(dY), (K % 4 == 0) && aligned16(X));
    MVN_LAUNCH_CHECK();
    return 0;
}

int launch_reduce_partials_n(const float* partial, size_t pstride, size_t n, int nslabs, float* out, int accumulate, cudaStream_t st) {
    if (n == 0) return 0;
    ProfScope prof(PROF_ROW, st);
    const int blocks = (int)((n + 31) / 32);
    MVN_CUDA(launch_dependent(reduce_partials_kernel, dim3(blocks), dim3(256), 0, st, partial, pstride, n, nslabs, out, accumulate, (size_t)0, (size_t)0));
    MVN_LAUNCH_CHECK();
    return 0;
}
int launch_reduce_partials_2d(const float* partial, size_t pstride, size_t n, int nslabs, float* out, int accumulate, int nl, size_t in_lstride,
                              size_t out_lstride, cudaStream_t st) {
    if (n == 0 || nl <= 0) return 0;
    ProfScope prof(PROF_ROW, st);
    const int blocks = (int)((n + 31) / 32);
    MVN_CUDA(launch_dependent(reduce_partials_kernel, dim3(blocks, nl), dim3(256), 0, st, partial, pstride, n, nslabs, out, accumulate, in_lstride, out_lstride));
    MVN_LAUNCH_CHECK();
    return 0;
}
int launch_reduce_partials(const float* partial, size_t pstride, size_t n, float* out, int accumulate, cudaStream_t st) {
    return launch_reduce_partials_n(partial, pstride, n, kSlabs, out, accumulate, st);
}

}  // namespace mvn

using namespace mvn;

extern "C" int mvn_linear_fwd(const float* X, const float* W, const float* bias, float* Y, const int32_t* n_rows_dev,
                              int M_cap, int N, int K, int act, int prec, void* stream) {
    GemmEpilogue ep;
    ep.bias = bias; ep.act = act;
    return launch_gemm(X, W, Y, n_rows_dev, M_cap, N, K, true, ep, prec, (cudaStream_t)stream);
}

extern "C" int mvn_linear_res_ln_fwd(const float* X, const float* W, const float* bias, const float* R, const float* gamma,
                                     const float* beta, float* Y, float* xhat, float* rstd, const int32_t* n_rows_dev,
                                     int M_cap, int N, int K, float eps, int prec, void* stream) {
    MVN_CHECK_ARG(gamma && beta, "linear_res_ln: gamma/beta required");
    GemmEpilogue ep;
    ep.bias = bias; ep.addend = R; ep.gamma = gamma; ep.beta = beta; ep.xhat = xhat; ep.rstd = rstd; ep.eps = eps;
    return launch_gemm(X, W, Y, n_rows_dev, M_cap, N, K, true, ep, prec, (cudaStream_t)stream);
}

extern "C" int mvn_linear_bwd_input(const float* dY, const float* W, float* dX, const float* addend, const float* act_src,
                                    int dact, const int32_t* n_rows_dev, int M_cap, int N, int K, int prec, void* stream) {
    MVN_CHECK_ARG(dact == 0 || act_src, "linear_bwd_input: dact needs act_src");
    GemmEpilogue ep;
    ep.addend = addend; ep.act_src = act_src; ep.dact = dact;
    // dX[M,K] = dY[M,N] * W[N,K]: contraction over N, W is stored [contraction][output]
    return launch_gemm(dY, W, dX, n_rows_dev, M_cap, K, N, false, ep, prec, (cudaStream_t)stream);
}

extern "C" size_t mvn_linear_bwd_weight_workspace_bytes(int M_cap, int N, int K) {
    (void)M_cap;
    return (size_t)kSlabs * ((size_t)N * K + N) * sizeof(float);
}

extern "C" int mvn_linear_bwd_weight(const float* dY, const float* X, float* dW, float* db, const int32_t* n_rows_dev,
                                     int M_cap, int N, int K, int accumulate, void* workspace, size_t workspace_bytes,
                                     int prec, void* stream) {
    MVN_CHECK_ARG(dW && workspace, "linear_bwd_weight: null dW/workspace");
    if (workspace_bytes < mvn_linear_bwd_weight_workspace_bytes(M_cap, N, K)) {
        set_error("linear_bwd_weight: workspace %zu < %zu", workspace_bytes, mvn_linear_bwd_weight_workspace_bytes(M_cap, N, K));
        return MVN_E_WORKSPACE;
    }
    const size_t pstride = (size_t)N * K + N;
    cudaStream_t st = (cudaStream_t)stream;
    MVN_TRY(launch_wgrad_partials(dY, X, n_rows_dev, M_cap, N, K, (float*)workspace, pstride, 0, db ? (long long)N * K : -1, prec, st));
    MVN_TRY(launch_reduce_partials((const float*)workspace, pstride, (size_t)N * K, dW, accumulate, st));
    if (db) MVN_TRY(launch_reduce_partials((const float*)workspace + (size_t)N * K, pstride, (size_t)N, db, accumulate, st));
    return 0;
}
