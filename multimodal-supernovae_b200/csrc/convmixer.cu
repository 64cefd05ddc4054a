// A8/A9: ConvMixer image encoder (src/models_multimodal.py:38-95), forward and backward, channels-last feature
// maps [R = B*Hp*Wp, dim].  HBM-bound on the 43 KB/sample input; everything after the patch embedding works on
// 4.6 KB/sample.  The work is split into stages at every BatchNorm so a data-parallel caller can all-reduce the
// per-channel sums (SyncBN semantics) between stages; single-GPU callers run the stages back to back.
//   BN 0 = patch BN;  mixer layer d: BN 2d+1 = depthwise ("A", inside the Residual), BN 2d+2 = pointwise ("B").
#include "common.cuh"

namespace mvn {
namespace {

struct ConvOff {           // flat parameter offsets (floats)
    size_t patch_w, bn0_g, bn0_b, layer0, layer_stride, fc1_w, fc1_b, fc2_w, fc2_b, ip_w, ip_b, total;
    size_t dw_w, dw_b, bnA_g, bnA_b, pw_w, pw_b, bnB_g, bnB_b;      // within a layer
};
ConvOff conv_offsets(const mvn_conv_cfg& c) {
    ConvOff o;
    const size_t dim = c.dim, kk = (size_t)c.kernel_size * c.kernel_size;
    size_t p = 0;
    o.patch_w = p; p += dim * c.C * c.patch_size * c.patch_size;
    o.bn0_g = p; p += dim;
    o.bn0_b = p; p += dim;
    o.layer0 = p;
    size_t q = 0;
    o.dw_w = q; q += dim * kk;
    o.dw_b = q; q += dim;
    o.bnA_g = q; q += dim;
    o.bnA_b = q; q += dim;
    o.pw_w = q; q += dim * dim;
    o.pw_b = q; q += dim;
    o.bnB_g = q; q += dim;
    o.bnB_b = q; q += dim;
    o.layer_stride = q;
    p += q * (size_t)c.depth;
    o.fc1_w = p; p += (size_t)c.hidden * dim;
    o.fc1_b = p; p += c.hidden;
    o.fc2_w = p; p += (size_t)c.n_out * c.hidden;
    o.fc2_b = p; p += c.n_out;
    o.ip_w = p; p += (size_t)c.enc_dim * c.n_out;
    o.ip_b = p; p += c.enc_dim;
    o.total = p;
    return o;
}
// parameter offsets of BN s: gamma, beta
void bn_param(const ConvOff& o, int s, size_t* g, size_t* b) {
    if (s == 0) { *g = o.bn0_g; *b = o.bn0_b; return; }
    const int d = (s - 1) / 2;
    const size_t base = o.layer0 + (size_t)d * o.layer_stride;
    if (s & 1) { *g = base + o.bnA_g; *b = base + o.bnA_b; } else { *g = base + o.bnB_g; *b = base + o.bnB_b; }
}

struct ConvWs {
    float *col, *pooled, *u1, *h1, *f, *p, *ynorm, *norm;
    float *dZ, *dU, *dYres, *dpooled, *du1, *df, *dp;
    double* stat_part;     // [kSlabs][2*dim]
    float* partial; size_t pstride;
    char* bn_base; size_t bn_bytes;
    size_t R, dim, bytes;
    struct Bn { float *u, *a, *z, *mean, *rstd, *scale, *shift; };
    Bn bn(int s) const {
        char* p = bn_base + (size_t)s * bn_bytes;
        Bn b;
        auto take = [&](size_t n) { float* r = (float*)p; p += align_up(n * 4, 256); return r; };
        b.u = take(R * dim); b.a = take(R * dim); b.z = take(R * dim);
        b.mean = take(dim); b.rstd = take(dim); b.scale = take(dim); b.shift = take(dim);
        return b;
    }
};
ConvWs conv_carve(const mvn_conv_cfg& c, void* base) {
    ConvWs w;
    const size_t B = c.B, P = (size_t)(c.H / c.patch_size) * (c.W / c.patch_size), R = B * P, dim = c.dim;
    const size_t Kp = (size_t)c.C * c.patch_size * c.patch_size;
    const size_t Dm = (size_t)(c.enc_dim > c.n_out ? c.enc_dim : c.n_out);
    w.R = R; w.dim = dim;
    char* p = (char*)base;
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
    w.col = (float*)take(R * Kp * 4);
    w.pooled = (float*)take(B * dim * 4);
    w.u1 = (float*)take(B * (size_t)c.hidden * 4);
    w.h1 = (float*)take(B * (size_t)c.hidden * 4);
    w.f = (float*)take(B * (size_t)c.n_out * 4);
    w.p = (float*)take(B * Dm * 4);
    w.ynorm = (float*)take(B * Dm * 4);
    w.norm = (float*)take(B * 4);
    w.dZ = (float*)take(R * dim * 4);
    w.dU = (float*)take(R * dim * 4);
    w.dYres = (float*)take(R * dim * 4);
    w.dpooled = (float*)take(B * dim * 4);
    w.du1 = (float*)take(B * (size_t)c.hidden * 4);
    w.df = (float*)take(B * (size_t)c.n_out * 4);
    w.dp = (float*)take(B * Dm * 4);
    w.stat_part = (double*)take((size_t)kSlabs * 2 * dim * 8);
    w.pstride = conv_offsets(c).total;
    w.partial = (float*)take((size_t)kSlabs * w.pstride * 4);
    size_t bb = 0;
    auto add = [&](size_t n) { bb += align_up(n * 4, 256); };
    add(R * dim); add(R * dim); add(R * dim); add(dim); add(dim); add(dim); add(dim);
    w.bn_bytes = bb;
    w.bn_base = p;
    p += bb * (size_t)(1 + 2 * c.depth);
    w.bytes = (size_t)(p - (char*)base);
    return w;
}

// Dropout sites: s = 1..2*depth follows BatchNorm s of the mixer layers (the patch BN, s = 0, has none); s = 1+2*depth is
// the projection head's dropout after its GELU.  Element index = (row of the [R,dim] / [B,hidden] matrix, column).
DropCfg conv_drop(const mvn_conv_cfg& c, int site) {
    if (!c.training || site == 0) return DropCfg();
    return make_drop(c.dropout_p, c.seed, (uint32_t)site);
}

int conv_check(const mvn_conv_cfg* c) {
    MVN_CHECK_ARG(c != nullptr, "convmixer: null cfg");
    MVN_CHECK_ARG(c->B > 0 && c->C > 0 && c->H > 0 && c->W > 0 && c->dim > 0 && c->depth >= 0 && c->n_out > 0 && c->hidden > 0 && c->enc_dim >= 0,
                  "convmixer: non-positive dims");
    MVN_UNSUPPORTED(c->dim % 4 == 0 && c->dim <= 128, "convmixer: dim=%d must be a multiple of 4 and <= 128", c->dim);
    MVN_UNSUPPORTED(c->patch_size > 0 && c->H % c->patch_size == 0 && c->W % c->patch_size == 0, "convmixer: patch %d must divide %dx%d", c->patch_size, c->H, c->W);
    MVN_UNSUPPORTED(c->kernel_size % 2 == 1 && c->kernel_size <= 9, "convmixer: kernel_size=%d must be odd and <= 9", c->kernel_size);
    MVN_UNSUPPORTED((c->C * c->patch_size * c->patch_size) % 4 == 0, "convmixer: C*p*p must be a multiple of 4");
    MVN_CHECK_ARG((long long)c->B * (c->H / c->patch_size) * (c->W / c->patch_size) < (1ll << 30), "convmixer: too many patches");
    MVN_CHECK_ARG(c->dropout_p >= 0.0f && c->dropout_p < 1.0f, "convmixer: dropout_p=%g outside [0,1)", (double)c->dropout_p);
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// im2col: col[(b,py,px), (c,i,j)] = img[b,c,py*p+i,px*p+j]   (conv weight layout (dim,C,p,p) -> [dim, C*p*p])
__global__ void im2col_kernel(const float* __restrict__ img, float* __restrict__ col, int B, int C, int H, int W, int p) {
    const int Hp = H / p, Wp = W / p, Kp = C * p * p;
    const size_t total = (size_t)B * Hp * Wp * Kp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kp);
        const size_t r = i / Kp;
        const int px = (int)(r % Wp), py = (int)((r / Wp) % Hp), b = (int)(r / ((size_t)Wp * Hp));
        const int j = k % p, ii = (k / p) % p, c = k / (p * p);
        col[i] = img[(((size_t)b * C + c) * H + py * p + ii) * W + px * p + j];
    }
}

// a = gelu(u); per-CTA partial sums of a and a^2 per channel (double) when stat_part != nullptr
__global__ void __launch_bounds__(256) gelu_stats_kernel(const float* __restrict__ u, float* __restrict__ a, int R, int dim, double* __restrict__ stat_part) {
    __shared__ double red[8][2][128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = blockIdx.x * 8 + wid; r < R; r += gridDim.x * 8) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            if (c < dim) {
                const float v = gelu_erf(u[(size_t)r * dim + c]);
                a[(size_t)r * dim + c] = v;
                s1[q] += v; s2[q] = fmaf(v, v, s2[q]);
            }
        }
    }
    if (!stat_part) return;
#pragma unroll
    for (int q = 0; q < 4; ++q) { red[wid][0][lane + 32 * q] = (double)s1[q]; red[wid][1][lane + 32 * q] = (double)s2[q]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * dim; i += blockDim.x) {
        const int sec = i / dim, c = i % dim;
        double s = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) s += red[w8][sec][c];
        stat_part[(size_t)blockIdx.x * 2 * dim + i] = s;
    }
}
// Same pass, vectorised: TPR = dim/4 threads per row (one float4 each), 256/TPR rows per CTA pass, 4 passes in flight per thread
// (the scalar version above keeps ONE 128-byte load in flight per warp: ~20 us of pure latency for a 4.7 MB map).
template <int TPR>
__global__ void __launch_bounds__(256) gelu_stats_vec_kernel(const float* __restrict__ u, float* __restrict__ a, int R, double* __restrict__ stat_part) {
    constexpr int DIM = 4 * TPR, RPP = 256 / TPR, U = 4;
    __shared__ double red[2][RPP][DIM];
    const int c4 = (threadIdx.x % TPR) * 4, rl = threadIdx.x / TPR;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r0 = blockIdx.x * RPP + rl; r0 < R; r0 += gridDim.x * RPP * U) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int r = r0 + k * gridDim.x * RPP;
            v[k] = r < R ? __ldg(reinterpret_cast<const float4*>(u + (size_t)r * DIM + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int r = r0 + k * gridDim.x * RPP;
            if (r >= R) break;
            const float4 g = make_float4(gelu_erf(v[k].x), gelu_erf(v[k].y), gelu_erf(v[k].z), gelu_erf(v[k].w));
            *reinterpret_cast<float4*>(a + (size_t)r * DIM + c4) = g;
            s1[0] += g.x; s1[1] += g.y; s1[2] += g.z; s1[3] += g.w;
            s2[0] = fmaf(g.x, g.x, s2[0]); s2[1] = fmaf(g.y, g.y, s2[1]); s2[2] = fmaf(g.z, g.z, s2[2]); s2[3] = fmaf(g.w, g.w, s2[3]);
        }
    }
    if (!stat_part) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[0][rl][c4 + j] = (double)s1[j]; red[1][rl][c4 + j] = (double)s2[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * DIM; i += 256) {
        const int sec = i / DIM, c = i % DIM;
        double s = 0.0;
#pragma unroll 8
        for (int w = 0; w < RPP; ++w) s += red[sec][w][c];
        stat_part[(size_t)blockIdx.x * 2 * DIM + i] = s;
    }
}
template <int TPR>
__global__ void __launch_bounds__(256) bn_bwd_stats_vec_kernel(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, int R, double* __restrict__ stat_part, const DropCfg drop) {
    constexpr int DIM = 4 * TPR, RPP = 256 / TPR, U = 4;
    __shared__ double red[2][RPP][DIM];
    const int c4 = (threadIdx.x % TPR) * 4, rl = threadIdx.x / TPR;
    const float4 mu = *reinterpret_cast<const float4*>(mean + c4), rs = *reinterpret_cast<const float4*>(rstd + c4);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r0 = blockIdx.x * RPP + rl; r0 < R; r0 += gridDim.x * RPP * U) {
        float4 g[U], av[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int r = r0 + k * gridDim.x * RPP;
            g[k] = av[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < R) {
                g[k] = __ldg(reinterpret_cast<const float4*>(dout + (size_t)r * DIM + c4));
                av[k] = __ldg(reinterpret_cast<const float4*>(a + (size_t)r * DIM + c4));
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int r = r0 + k * gridDim.x * RPP;
            if (r >= R) break;
            if (drop.thresh) {                                                     // dout is the gradient AFTER this BN's dropout
                const uint32_t rk = drop_rowkey(drop, (uint32_t)r);
                g[k].x *= drop_scale(drop, rk, (uint32_t)c4); g[k].y *= drop_scale(drop, rk, (uint32_t)c4 + 1);
                g[k].z *= drop_scale(drop, rk, (uint32_t)c4 + 2); g[k].w *= drop_scale(drop, rk, (uint32_t)c4 + 3);
            }
            s1[0] += g[k].x; s1[1] += g[k].y; s1[2] += g[k].z; s1[3] += g[k].w;
            s2[0] = fmaf(g[k].x, (av[k].x - mu.x) * rs.x, s2[0]); s2[1] = fmaf(g[k].y, (av[k].y - mu.y) * rs.y, s2[1]);
            s2[2] = fmaf(g[k].z, (av[k].z - mu.z) * rs.z, s2[2]); s2[3] = fmaf(g[k].w, (av[k].w - mu.w) * rs.w, s2[3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[0][rl][c4 + j] = (double)s1[j]; red[1][rl][c4 + j] = (double)s2[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * DIM; i += 256) {
        const int sec = i / DIM, c = i % DIM;
        double s = 0.0;
#pragma unroll 8
        for (int w = 0; w < RPP; ++w) s += red[sec][w][c];
        stat_part[(size_t)blockIdx.x * 2 * DIM + i] = s;
    }
}
// stats[i] = sum over CTAs (one warp per output, fixed lane/shuffle order -> deterministic)
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void stat_reduce_kernel(const double* __restrict__ part, int nblk, int n, double* __restrict__ out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    double s = 0.0;
    for (int b = lane; b < nblk; b += 32) s += part[(size_t)b * n + i];
    s = warp_sum_f64(s);
    if (lane == 0) out[i] = s;
}
// batch statistics -> normalisation coefficients (+ running-stat update); stats = [sum a | sum a^2] over `count` values
__global__ void bn_coeffs_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps, float momentum, int training, float* __restrict__ running, int dim,
                                 float* __restrict__ mean_o, float* __restrict__ rstd_o, float* __restrict__ scale_o, float* __restrict__ shift_o) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= dim) return;
    float mean, var;
    if (training) {
        const double m = stats[c] / count;
        double v = stats[dim + c] / count - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m; var = (float)v;
        if (running) {
            const double unb = count > 1.0 ? v * count / (count - 1.0) : v;
            running[c] = (1.0f - momentum) * running[c] + momentum * mean;
            running[dim + c] = (1.0f - momentum) * running[dim + c] + momentum * (float)unb;
        }
    } else {
        mean = running[c]; var = running[dim + c];
    }
    const float rs = 1.0f / sqrtf(var + eps);
    mean_o[c] = mean; rstd_o[c] = rs;
    const float sc = gamma[c] * rs;
    scale_o[c] = sc; shift_o[c] = beta[c] - mean * sc;
}
// z = dropout(a*scale + shift) (+ res)      (nn.Dropout follows every mixer BatchNorm, src/models_multimodal.py:62-77)
__global__ void bn_apply_kernel(const float* __restrict__ a, const float* __restrict__ scale, const float* __restrict__ shift,
                                const float* __restrict__ res, float* __restrict__ z, size_t n, int dim, const DropCfg drop) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % dim);
        float v = fmaf(a[i], scale[c], shift[c]);
        if (drop.thresh) v *= drop_scale(drop, drop_rowkey(drop, (uint32_t)(i / dim)), (uint32_t)c);
        if (res) v += res[i];
        z[i] = v;
    }
}
// float4 variants of the two elementwise BatchNorm passes (dim % 4 == 0, 16-byte aligned maps): a quarter of the threads and memory
// instructions of the scalar kernels
__global__ void bn_apply_vec_kernel(const float* __restrict__ a, const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ res, float* __restrict__ z, size_t n4, int dim, const DropCfg drop) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i * 4;
        const int c = (int)(e % dim);
        const float4 av = __ldg(reinterpret_cast<const float4*>(a) + i);
        const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
        float4 v = make_float4(fmaf(av.x, sc.x, sh.x), fmaf(av.y, sc.y, sh.y), fmaf(av.z, sc.z, sh.z), fmaf(av.w, sc.w, sh.w));
        if (drop.thresh) {
            const uint32_t rk = drop_rowkey(drop, (uint32_t)(e / dim));
            v.x *= drop_scale(drop, rk, (uint32_t)c); v.y *= drop_scale(drop, rk, (uint32_t)c + 1);
            v.z *= drop_scale(drop, rk, (uint32_t)c + 2); v.w *= drop_scale(drop, rk, (uint32_t)c + 3);
        }
        if (res) { const float4 r = __ldg(reinterpret_cast<const float4*>(res) + i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        reinterpret_cast<float4*>(z)[i] = v;
    }
}
__global__ void bn_bwd_apply_vec_kernel(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ u,
                                        const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
                                        const double* __restrict__ stats, double count, size_t n4, int dim, float* __restrict__ du, const DropCfg drop) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i * 4;
        const int c = (int)(e % dim);
        const float4 av = __ldg(reinterpret_cast<const float4*>(a) + i), uv = __ldg(reinterpret_cast<const float4*>(u) + i);
        float4 g = __ldg(reinterpret_cast<const float4*>(dout) + i);
        if (drop.thresh) {
            const uint32_t rk = drop_rowkey(drop, (uint32_t)(e / dim));
            g.x *= drop_scale(drop, rk, (uint32_t)c); g.y *= drop_scale(drop, rk, (uint32_t)c + 1);
            g.z *= drop_scale(drop, rk, (uint32_t)c + 2); g.w *= drop_scale(drop, rk, (uint32_t)c + 3);
        }
        const float avv[4] = {av.x, av.y, av.z, av.w}, uvv[4] = {uv.x, uv.y, uv.z, uv.w}, gv[4] = {g.x, g.y, g.z, g.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float xh = (avv[j] - mean[c + j]) * rstd[c + j];
            const float m1 = (float)(stats[c + j] / count), m2 = (float)(stats[dim + c + j] / count);
            o[j] = scale[c + j] * (gv[j] - m1 - xh * m2) * gelu_erf_grad(uvv[j]);
        }
        reinterpret_cast<float4*>(du)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}
// depthwise k x k 'same' convolution, channels-last.  transposed=1 gives the input-gradient (flipped taps) + addend
__global__ void dwconv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                              const float* __restrict__ addend, float* __restrict__ y, int B, int Hp, int Wp, int dim, int k, int transposed) {
    extern __shared__ float wt[];                  // [k*k][dim]
    for (int i = threadIdx.x; i < k * k * dim; i += blockDim.x) { const int tap = i / dim, c = i % dim; wt[i] = w[c * k * k + tap]; }
    __syncthreads();
    const int h = k / 2;
    const size_t total = (size_t)B * Hp * Wp * dim;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % dim);
        const size_t r = i / dim;
        const int px = (int)(r % Wp), py = (int)((r / Wp) % Hp);
        const size_t b0 = r - (size_t)py * Wp - px;
        float acc = bias ? bias[c] : 0.f;
        for (int ii = 0; ii < k; ++ii) {
            const int yy = transposed ? py - (ii - h) : py + (ii - h);
            if (yy < 0 || yy >= Hp) continue;
            for (int jj = 0; jj < k; ++jj) {
                const int xx = transposed ? px - (jj - h) : px + (jj - h);
                if (xx < 0 || xx >= Wp) continue;
                acc = fmaf(x[(b0 + (size_t)yy * Wp + xx) * dim + c], wt[(ii * k + jj) * dim + c], acc);
            }
        }
        if (addend) acc += addend[i];
        y[i] = acc;
    }
}
// Same convolution, one sample at a time per CTA: the [P, dim] map (4.6 KB at 6x6x32) is staged in shared memory with coalesced
// float4 loads and the thread's k*k taps of its channel live in registers, so the 25 reads per output hit shared memory instead
// of 25 scattered global loads.  dim == 32, k <= 5, P <= 64.
template <int K>
__global__ void __launch_bounds__(256) dwconv_smem_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ addend, float* __restrict__ y, int B, int Hp, int Wp,
                                                          int transposed) {
    constexpr int DIM = 32, MAXP = 64, MAXKK = K * K, k = K, h = K / 2, kk = K * K;
    __shared__ __align__(16) float sx[2][MAXP * DIM];
    const int P = Hp * Wp, n = P * DIM;
    const int c = threadIdx.x % DIM, pg = threadIdx.x / DIM;          // channel, position group (8 groups)
    float wt[MAXKK];
#pragma unroll
    for (int i = 0; i < MAXKK; ++i) {
        // transposed (input gradient): tap (ii, jj) reads x[p - (ii-h, jj-h)] = correlation with the flipped kernel
        const int src = transposed ? (kk - 1 - i) : i;
        wt[i] = i < kk ? __ldg(w + c * kk + src) : 0.f;
    }
    const float bv = bias ? __ldg(bias + c) : 0.f;
    int buf = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, buf ^= 1) {
        const float* xb = x + (size_t)b * n;
        for (int i = threadIdx.x * 4; i < n; i += 256 * 4) *reinterpret_cast<float4*>(&sx[buf][i]) = __ldg(reinterpret_cast<const float4*>(xb + i));
        __syncthreads();                                             // double buffered: one barrier per sample
        for (int p = pg; p < P; p += 256 / DIM) {
            const int py = p / Wp, px = p % Wp;
            float acc = bv;
#pragma unroll
            for (int ii = 0; ii < K; ++ii) {
                const int yy = py + ii - h;
                if (yy < 0 || yy >= Hp) continue;
#pragma unroll
                for (int jj = 0; jj < K; ++jj) {
                    const int xx = px + jj - h;
                    if (xx < 0 || xx >= Wp) continue;
                    acc = fmaf(sx[buf][(yy * Wp + xx) * DIM + c], wt[ii * k + jj], acc);
                }
            }
            const size_t o = (size_t)b * n + (size_t)p * DIM + c;
            if (addend) acc += __ldg(addend + o);
            y[o] = acc;
        }
    }
}
// depthwise weight/bias gradient partials: partial[cta][woff + c*k*k + tap] = sum du[r,c]*x[r+tap,c], [boff + c] = sum du.
// One CTA per slab of samples: the sample's two [P, dim] maps are staged in shared memory (coalesced float4 loads), then
// thread (c = tid % dim, tg = tid / dim) accumulates its taps {tg, tg+ntg, ...} over the P positions from shared memory.
// Tap index k*k is the bias column.  Fixed summation order per CTA -> deterministic.
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const float* __restrict__ du, const float* __restrict__ x, int B, int Hp, int Wp, int dim, int k,
                                                           float* __restrict__ partial, size_t pstride, size_t woff, size_t boff) {
    extern __shared__ __align__(16) float dw_sm[];
    const int P = Hp * Wp, n = P * dim;
    float* sdu = dw_sm;
    float* sx = dw_sm + n;
    const int c = threadIdx.x % dim, tg = threadIdx.x / dim, ntg = blockDim.x / dim;
    const int h = k / 2, kk = k * k, ntaps = kk + 1;
    constexpr int TPT = 8;                                   // taps per thread per pass
    float* pp = partial + (size_t)blockIdx.x * pstride;
    for (int t0 = 0; t0 < ntaps; t0 += TPT * ntg) {
        float acc[TPT];
#pragma unroll
        for (int j = 0; j < TPT; ++j) acc[j] = 0.f;
        for (int b = blockIdx.x; b < B; b += gridDim.x) {
            __syncthreads();
            const float4* gdu = reinterpret_cast<const float4*>(du + (size_t)b * n);
            const float4* gx = reinterpret_cast<const float4*>(x + (size_t)b * n);
            for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
                reinterpret_cast<float4*>(sdu)[i] = gdu[i];
                reinterpret_cast<float4*>(sx)[i] = gx[i];
            }
            __syncthreads();
            if (tg < ntg) {
#pragma unroll
                for (int j = 0; j < TPT; ++j) {
                    const int tap = t0 + tg + j * ntg;
                    if (tap >= ntaps) continue;
                    float a = 0.f;
                    if (tap == kk) {
                        for (int q = 0; q < P; ++q) a += sdu[q * dim + c];
                    } else {
                        const int dy = tap / k - h, dx = tap % k - h;
                        const int y0 = dy < 0 ? -dy : 0, y1 = dy > 0 ? Hp - dy : Hp;
                        const int x0 = dx < 0 ? -dx : 0, x1 = dx > 0 ? Wp - dx : Wp;
                        for (int py = y0; py < y1; ++py)
                            for (int px = x0; px < x1; ++px)
                                a = fmaf(sdu[(py * Wp + px) * dim + c], sx[((py + dy) * Wp + px + dx) * dim + c], a);
                    }
                    acc[j] += a;
                }
            }
        }
        if (tg < ntg) {
#pragma unroll
            for (int j = 0; j < TPT; ++j) {
                const int tap = t0 + tg + j * ntg;
                if (tap < kk) pp[woff + (size_t)c * kk + tap] = acc[j];
                else if (tap == kk) pp[boff + c] = acc[j];
            }
        }
    }
}
__global__ void avgpool_fwd_kernel(const float* __restrict__ z, int B, int P, int dim, float* __restrict__ pooled) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * dim) return;
    const int b = i / dim, c = i % dim;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += z[((size_t)b * P + p) * dim + c];
    pooled[i] = s / (float)P;
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ dpooled, int B, int P, int dim, float* __restrict__ dz) {
    const size_t total = (size_t)B * P * dim;
    const float inv = 1.0f / (float)P;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % dim);
        const size_t b = i / ((size_t)P * dim);
        dz[i] = dpooled[b * dim + c] * inv;
    }
}
// BN backward statistics: per channel sum(dout), sum(dout*xhat), xhat = (a-mean)*rstd
__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, int R, int dim, double* __restrict__ stat_part,
                                                           const DropCfg drop) {
    __shared__ double red[8][2][128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f}, mu[4], rs[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int c = lane + 32 * q; mu[q] = c < dim ? mean[c] : 0.f; rs[q] = c < dim ? rstd[c] : 0.f; }
    for (int r = blockIdx.x * 8 + wid; r < R; r += gridDim.x * 8) {
        const uint32_t rk = drop.thresh ? drop_rowkey(drop, (uint32_t)r) : 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            if (c < dim) {
                float g = dout[(size_t)r * dim + c];
                if (drop.thresh) g *= drop_scale(drop, rk, (uint32_t)c);     // dout is the gradient AFTER this BN's dropout
                const float xh = (a[(size_t)r * dim + c] - mu[q]) * rs[q];
                s1[q] += g; s2[q] = fmaf(g, xh, s2[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { red[wid][0][lane + 32 * q] = (double)s1[q]; red[wid][1][lane + 32 * q] = (double)s2[q]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * dim; i += blockDim.x) {
        const int sec = i / dim, c = i % dim;
        double s = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) s += red[w8][sec][c];
        stat_part[(size_t)blockIdx.x * 2 * dim + i] = s;
    }
}
// local sums -> stats_out (for the all-reduce) and the parameter gradients dgamma = sum dout*xhat, dbeta = sum dout
__global__ void bn_bwd_reduce_kernel(const double* __restrict__ part, int nblk, int dim, double* __restrict__ out, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= 2 * dim) return;
    double s = 0.0;
    for (int b = lane; b < nblk; b += 32) s += part[(size_t)b * 2 * dim + i];
    s = warp_sum_f64(s);
    if (lane == 0) {
        out[i] = s;
        if (i < dim) dbeta[i] = (float)s; else dgamma[i - dim] = (float)s;
    }
}
// du = gelu'(u) * scale * (dout - S1/N - xhat*S2/N)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ u,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
                                    const double* __restrict__ stats, double count, size_t n, int dim, float* __restrict__ du,
                                    const DropCfg drop) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % dim);
        const float xh = (a[i] - mean[c]) * rstd[c];
        const float m1 = (float)(stats[c] / count), m2 = (float)(stats[dim + c] / count);
        float g = dout[i];
        if (drop.thresh) g *= drop_scale(drop, drop_rowkey(drop, (uint32_t)(i / dim)), (uint32_t)c);
        const float da = scale[c] * (g - m1 - xh * m2);
        du[i] = da * gelu_erf_grad(u[i]);
    }
}
// eval-mode BN backward (running statistics are constants): du = gelu'(u) * scale * dout
__global__ void bn_bwd_eval_kernel(const float* __restrict__ dout, const float* __restrict__ u, const float* __restrict__ scale, size_t n, int dim,
                                   float* __restrict__ du) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        du[i] = dout[i] * scale[(int)(i % dim)] * gelu_erf_grad(u[i]);
}
// h = dropout(gelu(u)), rows of `cols` elements   (projection head, src/models_multimodal.py:85-87)
__global__ void gelu_pair_kernel(const float* __restrict__ u, float* __restrict__ h, size_t n, int cols, const DropCfg drop) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = gelu_erf(u[i]);
        if (drop.thresh) v *= drop_scale(drop, drop_rowkey(drop, (uint32_t)(i / cols)), (uint32_t)(i % cols));
        h[i] = v;
    }
}

// dispatchers: vectorised / shared-memory variants for the shapes of the shipped configurations, general kernels otherwise
inline void launch_gelu_stats(const float* u, float* a, int R, int dim, double* stat_part, cudaStream_t st) {
    if (dim == 32 && aligned16(u) && aligned16(a)) gelu_stats_vec_kernel<8><<<kSlabs, 256, 0, st>>>(u, a, R, stat_part);
    else if (dim == 64 && aligned16(u) && aligned16(a)) gelu_stats_vec_kernel<16><<<kSlabs, 256, 0, st>>>(u, a, R, stat_part);
    else gelu_stats_kernel<<<kSlabs, 256, 0, st>>>(u, a, R, dim, stat_part);
}
inline void launch_bn_bwd_stats(const float* dout, const float* a, const float* mean, const float* rstd, int R, int dim, double* stat_part,
                                const DropCfg& drop, cudaStream_t st) {
    if (dim == 32 && aligned16(dout) && aligned16(a)) bn_bwd_stats_vec_kernel<8><<<kSlabs, 256, 0, st>>>(dout, a, mean, rstd, R, stat_part, drop);
    else if (dim == 64 && aligned16(dout) && aligned16(a)) bn_bwd_stats_vec_kernel<16><<<kSlabs, 256, 0, st>>>(dout, a, mean, rstd, R, stat_part, drop);
    else bn_bwd_stats_kernel<<<kSlabs, 256, 0, st>>>(dout, a, mean, rstd, R, dim, stat_part, drop);
}
inline int ew_blocks(size_t n);
inline void launch_dwconv(const float* x, const float* w, const float* bias, const float* addend, float* y, int B, int Hp, int Wp, int dim, int k,
                          int transposed, cudaStream_t st) {
    const int grid = B < 4 * kSlabs ? B : 4 * kSlabs;
    if (dim == 32 && Hp * Wp <= 64 && k == 5 && aligned16(x)) dwconv_smem_kernel<5><<<grid, 256, 0, st>>>(x, w, bias, addend, y, B, Hp, Wp, transposed);
    else if (dim == 32 && Hp * Wp <= 64 && k == 3 && aligned16(x)) dwconv_smem_kernel<3><<<grid, 256, 0, st>>>(x, w, bias, addend, y, B, Hp, Wp, transposed);
    else dwconv_kernel<<<ew_blocks((size_t)B * Hp * Wp * dim), 256, (size_t)k * k * dim * 4, st>>>(x, w, bias, addend, y, B, Hp, Wp, dim, k, transposed);
}
inline int ew_blocks(size_t n) { size_t b = (n + 255) / 256; return (int)(b < 4096 ? (b ? b : 1) : 4096); }

}  // namespace
// convmixer_tc.cu: patch embedding as an implicit GEMM on tensor cores (prec >= 1)
bool patch_conv_tc_supported(const mvn_conv_cfg& c);
int launch_patch_conv_fwd_tc(const mvn_conv_cfg& c, const float* img, const float* Wt, float* u, float* a, double* stat_part, cudaStream_t st);
int launch_patch_conv_wgrad_tc(const mvn_conv_cfg& c, const float* img, const float* dU, float* partial, size_t pstride, size_t woff, cudaStream_t st);
// convmixer_fused.cu: one kernel per BatchNorm stage of the mixer layers (dim 32, training mode)
bool mixer_fused_supported(int dim, int P, int k, int training);
size_t mixer_scratch_bytes(int B);
int launch_mixer_fwd(int kind, int k, const double* stats_prev, double count, const float* gamma, const float* beta, float eps, float momentum,
                     float* running, float* mean_o, float* rstd_o, float* scale_o, float* shift_o, const float* a_prev, const float* res,
                     float* z_prev, const DropCfg& drop, const float* w, const float* bias, float* u, float* a_out, void* scratch,
                     double* stats_out, int stage, float* pooled, int B, int Hp, int Wp, cudaStream_t st);
int launch_mixer_bwd(int kind, int k, float* dZ, const float* a_s, const float* u_s, const float* mean_s, const float* rstd_s, const float* scale_s,
                     const double* stats_s, double count, const DropCfg& drop_s, const float* w, const float* x, const float* a_p,
                     const float* mean_p, const float* rstd_p, const DropCfg& drop_p, void* scratch, double* stats_out, float* dgamma,
                     float* dbeta, int stage, float* dW_out, int B, int Hp, int Wp, cudaStream_t st);
int launch_pool_bwd_stats(const float* dpooled, float* dZ, const float* a_p, const float* mean_p, const float* rstd_p, const DropCfg& drop_p,
                          void* scratch, double* stats_out, float* dgamma, float* dbeta, int stage, int B, int P, cudaStream_t st);
}  // namespace mvn

using namespace mvn;

extern "C" size_t mvn_conv_param_count(const mvn_conv_cfg* cfg) { return conv_check(cfg) ? 0 : conv_offsets(*cfg).total; }
extern "C" size_t mvn_conv_workspace_bytes(const mvn_conv_cfg* cfg) { return conv_check(cfg) ? 0 : conv_carve(*cfg, nullptr).bytes + 256; }
extern "C" int mvn_conv_num_bn(const mvn_conv_cfg* c) { return c ? 1 + 2 * c->depth : 0; }

extern "C" int mvn_convmixer_fwd_stage(const mvn_conv_cfg* cfg, int stage, const float* params, const float* img, float* running_stats,
                                       double* bn_stats, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    MVN_TRY(conv_check(cfg));
    const mvn_conv_cfg& c = *cfg;
    const int nbn = 1 + 2 * c.depth;
    MVN_CHECK_ARG(stage >= 0 && stage <= nbn && params && img && bn_stats && workspace && running_stats, "convmixer_fwd_stage: bad arguments");
    void* base = (void*)align_up((size_t)workspace, 256);
    const ConvWs w = conv_carve(c, base);
    if ((char*)base + w.bytes > (char*)workspace + workspace_bytes) { set_error("convmixer: workspace too small"); return MVN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_CONV, st);
    const ConvOff o = conv_offsets(c);
    const int Hp = c.H / c.patch_size, Wp = c.W / c.patch_size, P = Hp * Wp, R = c.B * P, dim = c.dim;
    const int Kp = c.C * c.patch_size * c.patch_size;
    const size_t n = (size_t)R * dim;
    const double count = (double)(c.global_count > 0 ? c.global_count : (long long)R);

    auto finish_bn = [&](int s) -> int {          // coefficients of BN s from its (already reduced) sums, then z_s = BN(a_s) (+ residual)
        size_t g, b;
        bn_param(o, s, &g, &b);
        const ConvWs::Bn bn = w.bn(s);
        bn_coeffs_kernel<<<cdiv(dim, 128), 128, 0, st>>>(bn_stats + (size_t)s * 2 * dim, count, params + g, params + b, c.bn_eps, c.bn_momentum,
                                                         c.training, running_stats + (size_t)s * 2 * dim, dim, bn.mean, bn.rstd, bn.scale, bn.shift);
        MVN_LAUNCH_CHECK();
        const float* res = (s & 1) ? w.bn(s - 1).z : nullptr;      // "A" BNs sit inside the Residual: add the layer input
        if (dim % 4 == 0 && aligned16(bn.a) && aligned16(bn.z) && (!res || aligned16(res)))
            bn_apply_vec_kernel<<<ew_blocks(n / 4), 256, 0, st>>>(bn.a, bn.scale, bn.shift, res, bn.z, n / 4, dim, conv_drop(c, s));
        else
            bn_apply_kernel<<<ew_blocks(n), 256, 0, st>>>(bn.a, bn.scale, bn.shift, res, bn.z, n, dim, conv_drop(c, s));
        MVN_LAUNCH_CHECK();
        return 0;
    };
    auto gelu_stats = [&](int s) -> int {
        const ConvWs::Bn bn = w.bn(s);
        launch_gelu_stats(bn.u, bn.a, R, dim, c.training ? w.stat_part : nullptr, st);
        MVN_LAUNCH_CHECK();
        if (c.training) {
            stat_reduce_kernel<<<cdiv(2 * dim, 8), 256, 0, st>>>(w.stat_part, kSlabs, 2 * dim, bn_stats + (size_t)s * 2 * dim);
            MVN_LAUNCH_CHECK();
        }
        return 0;
    };

    if (stage == 0 && c.prec >= 1 && patch_conv_tc_supported(c)) {
        // implicit GEMM from the image, GELU and the BatchNorm partial sums in the epilogue: one pass over 43.2 KB per sample
        const ConvWs::Bn bn = w.bn(0);
        MVN_TRY(launch_patch_conv_fwd_tc(c, img, params + o.patch_w, bn.u, bn.a, c.training ? w.stat_part : nullptr, st));
        if (c.training) {
            stat_reduce_kernel<<<cdiv(2 * dim, 8), 256, 0, st>>>(w.stat_part, kSlabs, 2 * dim, bn_stats);
            MVN_LAUNCH_CHECK();
        }
        return 0;
    }
    if (stage == 0) {
        im2col_kernel<<<ew_blocks((size_t)R * Kp), 256, 0, st>>>(img, w.col, c.B, c.C, c.H, c.W, c.patch_size);
        MVN_LAUNCH_CHECK();
        GemmEpilogue e;
        MVN_TRY(launch_gemm(w.col, params + o.patch_w, w.bn(0).u, nullptr, R, dim, Kp, true, e, 0, st));
        return gelu_stats(0);
    }
    const bool fused = mixer_fused_supported(dim, P, c.kernel_size, c.training) && mixer_scratch_bytes(c.B) <= (size_t)kSlabs * w.pstride * sizeof(float);      // per-CTA partials live in the weight-gradient slab area
    if (fused) {
        // one kernel: BN stage-1 (coefficients, apply, dropout, residual) + this stage's convolution + GELU + the sums of BN `stage`
        // (or, at the head, + the average pool)
        const int sp = stage - 1;
        size_t g, b;
        bn_param(o, sp, &g, &b);
        const ConvWs::Bn bp = w.bn(sp);
        const float* res = (sp & 1) ? w.bn(sp - 1).z : nullptr;
        const float* cw = nullptr; const float* cb = nullptr;
        float *uo = nullptr, *ao = nullptr;
        int kind = 2;
        if (stage < nbn) {
            const float* LP = params + o.layer0 + (size_t)((stage - 1) / 2) * o.layer_stride;
            kind = (stage & 1) ? 0 : 1;
            cw = LP + (kind == 0 ? o.dw_w : o.pw_w); cb = LP + (kind == 0 ? o.dw_b : o.pw_b);
            uo = w.bn(stage).u; ao = w.bn(stage).a;
        }
        MVN_TRY(launch_mixer_fwd(kind, c.kernel_size, bn_stats + (size_t)sp * 2 * dim, count, params + g, params + b, c.bn_eps, c.bn_momentum,
                                 running_stats + (size_t)sp * 2 * dim, bp.mean, bp.rstd, bp.scale, bp.shift, bp.a, res, bp.z, conv_drop(c, sp), cw, cb,
                                 uo, ao, w.partial, bn_stats + (size_t)stage * 2 * dim, stage, w.pooled, c.B, Hp, Wp, st));
        if (stage < nbn) return 0;
    } else {
        MVN_TRY(finish_bn(stage - 1));
    }
    if (stage < nbn) {
        const int d = (stage - 1) / 2;
        const float* LP = params + o.layer0 + (size_t)d * o.layer_stride;
        const float* xin = w.bn(stage - 1).z;
        if (stage & 1) {        // depthwise conv of layer d on z_{2d}
            launch_dwconv(xin, LP + o.dw_w, LP + o.dw_b, nullptr, w.bn(stage).u, c.B, Hp, Wp, dim, c.kernel_size, 0, st);
            MVN_LAUNCH_CHECK();
        } else {                // pointwise conv on y = BN_A(...) + x
            GemmEpilogue e;
            e.bias = LP + o.pw_b;
            MVN_TRY(launch_gemm(xin, LP + o.pw_w, w.bn(stage).u, nullptr, R, dim, dim, true, e, c.prec >= 1 ? 1 : 0, st));
        }
        return gelu_stats(stage);
    }
    // head
    MVN_CHECK_ARG(out != nullptr, "convmixer_fwd_stage: out is null at the last stage");
    if (!fused) {
        avgpool_fwd_kernel<<<cdiv(c.B * dim, 256), 256, 0, st>>>(w.bn(nbn - 1).z, c.B, P, dim, w.pooled);
        MVN_LAUNCH_CHECK();
    }
    GemmEpilogue e1;
    e1.bias = params + o.fc1_b;
    MVN_TRY(launch_gemm(w.pooled, params + o.fc1_w, w.u1, nullptr, c.B, c.hidden, dim, true, e1, 0, st));
    gelu_pair_kernel<<<ew_blocks((size_t)c.B * c.hidden), 256, 0, st>>>(w.u1, w.h1, (size_t)c.B * c.hidden, c.hidden, conv_drop(c, nbn));
    MVN_LAUNCH_CHECK();
    GemmEpilogue e2;
    e2.bias = params + o.fc2_b;
    MVN_TRY(launch_gemm(w.h1, params + o.fc2_w, w.f, nullptr, c.B, c.n_out, c.hidden, true, e2, 0, st));
    const float* feat = w.f;
    const int D = c.enc_dim > 0 ? c.enc_dim : c.n_out;
    if (c.enc_dim > 0) {
        GemmEpilogue e3;
        e3.bias = params + o.ip_b;
        MVN_TRY(launch_gemm(w.f, params + o.ip_w, w.p, nullptr, c.B, c.enc_dim, c.n_out, true, e3, 0, st));
        feat = w.p;
    }
    if (c.normalize) {
        MVN_TRY(mvn_l2norm_fwd(feat, w.ynorm, w.norm, c.B, D, st));
        feat = w.ynorm;
    }
    MVN_CUDA(cudaMemcpyAsync(out, feat, (size_t)c.B * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int mvn_convmixer_bwd_stage(const mvn_conv_cfg* cfg, int stage, const float* params, const float* img, const double* bn_stats,
                                       double* bn_stats_bwd, const float* dout, float* grads, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    (void)bn_stats;
    MVN_TRY(conv_check(cfg));
    const mvn_conv_cfg& c = *cfg;
    const int nbn = 1 + 2 * c.depth;
    MVN_CHECK_ARG(stage >= 0 && stage <= nbn && params && bn_stats_bwd && grads && workspace, "convmixer_bwd_stage: bad arguments");
    void* base = (void*)align_up((size_t)workspace, 256);
    const ConvWs w = conv_carve(c, base);
    if ((char*)base + w.bytes > (char*)workspace + workspace_bytes) { set_error("convmixer: workspace too small"); return MVN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_CONV, st);
    const ConvOff o = conv_offsets(c);
    const int Hp = c.H / c.patch_size, Wp = c.W / c.patch_size, P = Hp * Wp, R = c.B * P, dim = c.dim;
    const int Kp = c.C * c.patch_size * c.patch_size;
    const size_t n = (size_t)R * dim;
    const double count = (double)(c.global_count > 0 ? c.global_count : (long long)R);
    float* part = w.partial;
    const size_t ps = w.pstride;

    // local statistics of BN s from dZ (grad of its output), plus its dgamma/dbeta (from LOCAL sums: the caller's
    // gradient all-reduce adds the other ranks' shares)
    auto bwd_stats = [&](int s) -> int {
        size_t g, b;
        bn_param(o, s, &g, &b);
        const ConvWs::Bn bn = w.bn(s);
        launch_bn_bwd_stats(w.dZ, bn.a, bn.mean, bn.rstd, R, dim, w.stat_part, conv_drop(c, s), st);
        MVN_LAUNCH_CHECK();
        bn_bwd_reduce_kernel<<<cdiv(2 * dim, 8), 256, 0, st>>>(w.stat_part, kSlabs, dim, bn_stats_bwd + (size_t)s * 2 * dim, grads + g, grads + b);
        MVN_LAUNCH_CHECK();
        return 0;
    };

    const bool fused = mixer_fused_supported(dim, P, c.kernel_size, c.training) && mixer_scratch_bytes(c.B) <= (size_t)kSlabs * w.pstride * sizeof(float);      // per-CTA partials live in the weight-gradient slab area
    if (stage == nbn) {
        MVN_CHECK_ARG(dout != nullptr, "convmixer_bwd_stage: dout is null at the head stage");
        const int D = c.enc_dim > 0 ? c.enc_dim : c.n_out;
        const float* dfeat = dout;
        if (c.normalize) {
            MVN_TRY(mvn_l2norm_bwd(dout, w.ynorm, w.norm, w.dp, c.B, D, st));
            dfeat = w.dp;
        }
        GemmEpilogue e0;
        const float* df = dfeat;
        if (c.enc_dim > 0) {
            MVN_TRY(launch_wgrad_partials(dfeat, w.f, nullptr, c.B, c.enc_dim, c.n_out, part, ps, o.ip_w, (long long)o.ip_b, 0, st));
            MVN_TRY(launch_gemm(dfeat, params + o.ip_w, w.df, nullptr, c.B, c.n_out, c.enc_dim, false, e0, 0, st));
            df = w.df;
        }
        MVN_TRY(launch_wgrad_partials(df, w.h1, nullptr, c.B, c.n_out, c.hidden, part, ps, o.fc2_w, (long long)o.fc2_b, 0, st));
        GemmEpilogue eg;
        eg.act_src = w.u1; eg.dact = 2;
        eg.drop = conv_drop(c, nbn);
        MVN_TRY(launch_gemm(df, params + o.fc2_w, w.du1, nullptr, c.B, c.hidden, c.n_out, false, eg, 0, st));
        MVN_TRY(launch_wgrad_partials(w.du1, w.pooled, nullptr, c.B, c.hidden, dim, part, ps, o.fc1_w, (long long)o.fc1_b, 0, st));
        MVN_TRY(launch_gemm(w.du1, params + o.fc1_w, w.dpooled, nullptr, c.B, dim, c.hidden, false, e0, 0, st));
        MVN_TRY(launch_reduce_partials(part + o.fc1_w, ps, o.total - o.fc1_w, grads + o.fc1_w, 0, st));
        if (fused) {
            size_t g, b;
            bn_param(o, nbn - 1, &g, &b);
            const ConvWs::Bn bl = w.bn(nbn - 1);
            return launch_pool_bwd_stats(w.dpooled, w.dZ, bl.a, bl.mean, bl.rstd, conv_drop(c, nbn - 1), w.partial,
                                         bn_stats_bwd + (size_t)(nbn - 1) * 2 * dim, grads + g, grads + b, nbn, c.B, P, st);
        }
        avgpool_bwd_kernel<<<ew_blocks(n), 256, 0, st>>>(w.dpooled, c.B, P, dim, w.dZ);
        MVN_LAUNCH_CHECK();
        return bwd_stats(nbn - 1);
    }

    // stage s in [0, nbn): BN s backward (needs the reduced sums), then the convolution that feeds BN s
    const int s = stage;
    const ConvWs::Bn bn = w.bn(s);
    if (fused && s >= 1) {
        // one kernel: BN s backward * GELU' -> input gradient of the convolution (+ residual path) = gradient of z_{s-1} (in place in
        // dZ), the sums of BN s-1 backward (+ dgamma / dbeta), and the convolution's weight-gradient partials
        const int kind = (s & 1) ? 0 : 1;
        const size_t lb = o.layer0 + (size_t)((s - 1) / 2) * o.layer_stride;
        const size_t woff = lb + (kind == 0 ? o.dw_w : o.pw_w);          // the bias follows its weight in the flat layout
        size_t g, b;
        bn_param(o, s - 1, &g, &b);
        const ConvWs::Bn bp = w.bn(s - 1);
        return launch_mixer_bwd(kind, c.kernel_size, w.dZ, bn.a, bn.u, bn.mean, bn.rstd, bn.scale, bn_stats_bwd + (size_t)s * 2 * dim, count,
                                conv_drop(c, s), params + woff, bp.z, bp.a, bp.mean, bp.rstd, conv_drop(c, s - 1), w.partial,
                                bn_stats_bwd + (size_t)(s - 1) * 2 * dim, grads + g, grads + b, s, grads + woff, c.B, Hp, Wp, st);
    }
    if (c.training) {
        if (dim % 4 == 0 && aligned16(w.dZ) && aligned16(bn.a) && aligned16(bn.u) && aligned16(w.dU))
            bn_bwd_apply_vec_kernel<<<ew_blocks(n / 4), 256, 0, st>>>(w.dZ, bn.a, bn.u, bn.mean, bn.rstd, bn.scale, bn_stats_bwd + (size_t)s * 2 * dim, count,
                                                                      n / 4, dim, w.dU, conv_drop(c, s));
        else
            bn_bwd_apply_kernel<<<ew_blocks(n), 256, 0, st>>>(w.dZ, bn.a, bn.u, bn.mean, bn.rstd, bn.scale, bn_stats_bwd + (size_t)s * 2 * dim, count, n, dim, w.dU,
                                                              conv_drop(c, s));
    } else {
        bn_bwd_eval_kernel<<<ew_blocks(n), 256, 0, st>>>(w.dZ, bn.u, bn.scale, n, dim, w.dU);
    }
    MVN_LAUNCH_CHECK();
    if (s == 0 && c.prec >= 1 && patch_conv_tc_supported(c)) {
        MVN_CHECK_ARG(img != nullptr, "convmixer_bwd_stage: img is null at stage 0");
        MVN_TRY(launch_patch_conv_wgrad_tc(c, img, w.dU, part, ps, o.patch_w, st));
        return launch_reduce_partials(part + o.patch_w, ps, (size_t)dim * Kp, grads + o.patch_w, 0, st);
    }
    if (s == 0) {
        MVN_TRY(launch_wgrad_partials(w.dU, w.col, nullptr, R, dim, Kp, part, ps, o.patch_w, -1, 0, st));
        return launch_reduce_partials(part + o.patch_w, ps, (size_t)dim * Kp, grads + o.patch_w, 0, st);
    }
    const int d = (s - 1) / 2;
    const size_t lbase = o.layer0 + (size_t)d * o.layer_stride;
    const float* LP = params + lbase;
    if (s & 1) {    // depthwise conv: input x = z_{s-1}; total grad of x = dYres (residual path) + dwconv^T(dU)
        dwconv_wgrad_kernel<<<kSlabs, 256, (size_t)2 * P * dim * sizeof(float), st>>>(w.dU, w.bn(s - 1).z, c.B, Hp, Wp, dim, c.kernel_size, part, ps, lbase + o.dw_w, lbase + o.dw_b);
        MVN_LAUNCH_CHECK();
        launch_dwconv(w.dU, LP + o.dw_w, nullptr, w.dYres, w.dZ, c.B, Hp, Wp, dim, c.kernel_size, 1, st);
        MVN_LAUNCH_CHECK();
        MVN_TRY(launch_reduce_partials(part + lbase + o.dw_w, ps, (size_t)dim * c.kernel_size * c.kernel_size + dim, grads + lbase + o.dw_w, 0, st));
    } else {        // pointwise conv: input y = z_{s-1} (BN_A output + residual); dy feeds BN_A and the residual path
        const int gp = c.prec >= 1 ? 1 : 0;
        MVN_TRY(launch_wgrad_partials(w.dU, w.bn(s - 1).z, nullptr, R, dim, dim, part, ps, lbase + o.pw_w, (long long)(lbase + o.pw_b), gp, st));
        GemmEpilogue e0;
        MVN_TRY(launch_gemm(w.dU, LP + o.pw_w, w.dZ, nullptr, R, dim, dim, false, e0, gp, st));
        MVN_CUDA(cudaMemcpyAsync(w.dYres, w.dZ, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
        MVN_TRY(launch_reduce_partials(part + lbase + o.pw_w, ps, (size_t)dim * dim + dim, grads + lbase + o.pw_w, 0, st));
    }
    return bwd_stats(s - 1);      // also writes dgamma/dbeta of BN s-1 (eval mode: same sums against the running statistics)
}
