// A8: ConvMixer mixer stages as ONE kernel per BatchNorm stage (src/models_multimodal.py:52-95).
//
// Everything after the patch embedding works on a [P, dim] map per sample (6 x 6 x 32 floats = 4.6 KB): the stage-by-stage
// path of convmixer.cu spends its time in launches (5 per forward stage, 6-7 per backward stage), not in bytes.  Here a CTA
// walks whole samples with the map in shared memory:
//   forward  stage s: coefficients of BN s-1 from its (already reduced) sums -> z_{s-1} = dropout(BN(a_{s-1})) (+ residual)
//                     -> depthwise k x k / pointwise 1 x 1 convolution + bias -> u_s -> a_s = GELU(u_s) -> per-channel sums of
//                     BN s (double), reduced by the last CTA to finish (fixed order: deterministic);
//   head     stage  : the same BN step fused with the average pool;
//   backward stage s: dU = BN_s backward (given its reduced sums) * GELU'(u_s) -> transposed depthwise conv (+ residual path)
//                     / pointwise input gradient -> gradient of z_{s-1} (in place) -> sums of BN s-1 backward (+ its dgamma,
//                     dbeta), and this CTA's share of the convolution's weight / bias gradient (slab partials).
// The stage boundaries are unchanged (one C-ABI call per BatchNorm), so a data-parallel caller still all-reduces the
// per-channel sums between calls (SyncBN).  dim == 32, P <= 64, k in {3, 5}, training mode; anything else runs convmixer.cu's kernels.
#include <stdlib.h>

#include "common.cuh"

namespace mvn {
namespace {

constexpr int MX_DIM = 32, MX_MAXP = 64, MX_THREADS = 256, MX_PG = MX_THREADS / MX_DIM, MX_PT = MX_MAXP / MX_PG;
constexpr int MX_WSTR = MX_DIM * MX_DIM + MX_DIM;      // floats per CTA of weight-gradient partials ([32][25] + 32 or [32][32] + 32)
constexpr int MX_CTAS_PER_SM = 4;

// arrival counters of the last-CTA reductions, one per (direction, stage): zero at module load, reset by the CTA that consumes them.
// (One ConvMixer runs its stages in stream order; two models interleaving the same stage on different streams of one device would
// share a counter -- the drop-in runs one training thread per device)
__device__ unsigned g_mixer_counters[2][32];

struct MixFwdArgs {
    // BatchNorm s-1
    const double* stats_prev; double count; const float* gamma; const float* beta; float eps, momentum; float* running;
    float *mean_o, *rstd_o, *scale_o, *shift_o;
    const float* a_prev; const float* res; float* z_prev; DropCfg drop;
    // convolution of stage s and BatchNorm s statistics
    const float* w; const float* bias; float* u; float* a_out; double* stat_part; double* stats_out; unsigned* counter;
    float* pooled;
    int B, Hp, Wp;
};

// per-channel sums of this CTA -> stat_part[cta]; the last CTA to arrive adds the CTAs' partials in a fixed order (four contiguous
// ranges of CTAs summed in parallel, then combined in range order): deterministic
__device__ __forceinline__ void reduce_channel_sums(float s1, float s2, double (*red)[2][MX_DIM], double* __restrict__ stat_part, double* __restrict__ out,
                                                    unsigned* counter, float* __restrict__ o1, float* __restrict__ o2) {
    __shared__ bool last;
    const int c = threadIdx.x % MX_DIM, pg = threadIdx.x / MX_DIM;
    red[pg][0][c] = (double)s1; red[pg][1][c] = (double)s2;
    __syncthreads();
    if (threadIdx.x < 2 * MX_DIM) {
        const int sec = threadIdx.x / MX_DIM;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < MX_PG; ++w) s += red[w][sec][c];
        stat_part[(size_t)blockIdx.x * 2 * MX_DIM + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    {
        const int i = threadIdx.x % (2 * MX_DIM), q = threadIdx.x / (2 * MX_DIM);      // 4 ranges of CTAs
        const unsigned per = (gridDim.x + 3) / 4, b0 = q * per, b1 = min(gridDim.x, b0 + per);
        double s = 0.0;
        unsigned b = b0;
        for (; b + 8 <= b1; b += 8) {                       // 8 loads in flight, added in index order
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldcg(stat_part + (size_t)(b + k) * 2 * MX_DIM + i);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[k];
        }
        for (; b < b1; ++b) s += __ldcg(stat_part + (size_t)b * 2 * MX_DIM + i);
        double* r4 = &red[0][0][0];                                                    // 8*2*32 doubles: room for 4*64
        r4[q * 2 * MX_DIM + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 2 * MX_DIM) {
        const double* r4 = &red[0][0][0];
        const double s = ((r4[threadIdx.x] + r4[2 * MX_DIM + threadIdx.x]) + r4[4 * MX_DIM + threadIdx.x]) + r4[6 * MX_DIM + threadIdx.x];
        out[threadIdx.x] = s;
        if (o1 && threadIdx.x < MX_DIM) o1[threadIdx.x] = (float)s;
        if (o2 && threadIdx.x >= MX_DIM) o2[threadIdx.x - MX_DIM] = (float)s;
    }
    if (threadIdx.x == 0) *counter = 0;                     // ready for the next launch / graph replay
}

// KIND 0: depthwise conv follows, 1: pointwise conv follows, 2: average pool follows (head)
template <int KIND, int K>
__global__ void __launch_bounds__(MX_THREADS, 3) mixer_fwd_kernel(const MixFwdArgs a) {
    constexpr int DIM = MX_DIM, kk = K * K, h = K / 2;
    __shared__ float coef[2][DIM];
    __shared__ __align__(16) float sz[MX_MAXP * DIM];
    __shared__ double red[MX_PG][2][DIM];
    const int c = threadIdx.x % DIM, pg = threadIdx.x / DIM;
    const int Hp = a.Hp, Wp = a.Wp, P = Hp * Wp;
    if (threadIdx.x < DIM) {                                // BatchNorm s-1 coefficients (training mode), as bn_coeffs_kernel
        const double m = a.stats_prev[c] / a.count;
        double v = a.stats_prev[DIM + c] / a.count - m * m;
        if (v < 0.0) v = 0.0;
        const float mean = (float)m, var = (float)v;
        const float rs = 1.0f / sqrtf(var + a.eps);
        const float sc = a.gamma[c] * rs, sh = a.beta[c] - mean * sc;
        coef[0][c] = sc; coef[1][c] = sh;
        if (blockIdx.x == 0) {
            a.mean_o[c] = mean; a.rstd_o[c] = rs; a.scale_o[c] = sc; a.shift_o[c] = sh;
            if (a.running) {
                const double unb = a.count > 1.0 ? v * a.count / (a.count - 1.0) : v;
                a.running[c] = (1.0f - a.momentum) * a.running[c] + a.momentum * mean;
                a.running[DIM + c] = (1.0f - a.momentum) * a.running[DIM + c] + a.momentum * (float)unb;
            }
        }
    }
    // the thread's weights (channel c's taps / row c of W[out][in]) go through shared memory: read straight from global they are
    // 25- / 32-line gathers per warp instruction, which serialise in the L1 (measured: 17 us of a 41 us kernel)
    float wt[KIND == 0 ? kk : DIM];
    float bv = 0.f;
    if (KIND != 2) {
        constexpr int NW = KIND == 0 ? DIM * kk : DIM * DIM, LDW = KIND == 0 ? kk : DIM + 1;
        for (int i = threadIdx.x; i < NW; i += MX_THREADS) sz[(i / (KIND == 0 ? kk : DIM)) * LDW + i % (KIND == 0 ? kk : DIM)] = __ldg(a.w + i);
        bv = __ldg(a.bias + c);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < (KIND == 0 ? kk : DIM); ++i) wt[i] = sz[c * LDW + i];
    }
    __syncthreads();
    const float sc = coef[0][c], sh = coef[1][c];
    float s1 = 0.f, s2 = 0.f;
    const float invP = 1.0f / (float)P;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const size_t base = (size_t)b * P * DIM + c;
        float pool = 0.f;
        float av[MX_PT], rv[MX_PT];
#pragma unroll
        for (int i = 0; i < MX_PT; ++i) {                   // every load of the sample in flight before the first use
            const int p = pg + MX_PG * i;
            av[i] = rv[i] = 0.f;
            if (p < P) {
                av[i] = __ldg(a.a_prev + base + (size_t)p * DIM);
                if (a.res) rv[i] = __ldg(a.res + base + (size_t)p * DIM);
            }
        }
#pragma unroll
        for (int i = 0; i < MX_PT; ++i) {
            const int p = pg + MX_PG * i;
            if (p >= P) break;
            float v = fmaf(av[i], sc, sh);
            if (a.drop.thresh) v *= drop_scale(a.drop, drop_rowkey(a.drop, (uint32_t)(b * P + p)), (uint32_t)c);
            v += rv[i];
            if (KIND == 2) { pool += v; } else { a.z_prev[base + (size_t)p * DIM] = v; sz[p * DIM + c] = v; }
        }
        if (KIND == 2) {
            __syncthreads();
            sz[pg * DIM + c] = pool;
            __syncthreads();
            if (pg == 0) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < MX_PG; ++w) s += sz[w * DIM + c];
                a.pooled[(size_t)b * DIM + c] = s * invP;
            }
            continue;
        }
        __syncthreads();
        for (int p = pg; p < P; p += MX_PG) {
            float acc = bv;
            if (KIND == 0) {
                const int py = p / Wp, px = p % Wp;
#pragma unroll
                for (int ii = 0; ii < K; ++ii) {
                    const int yy = py + ii - h;
                    if (yy < 0 || yy >= Hp) continue;
#pragma unroll
                    for (int jj = 0; jj < K; ++jj) {
                        const int xx = px + jj - h;
                        if (xx < 0 || xx >= Wp) continue;
                        acc = fmaf(sz[(yy * Wp + xx) * DIM + c], wt[ii * K + jj], acc);
                    }
                }
            } else {
                const float4* zr = reinterpret_cast<const float4*>(sz + p * DIM);          // broadcast reads
#pragma unroll
                for (int i = 0; i < DIM / 4; ++i) {
                    const float4 z4 = zr[i];
                    acc = fmaf(z4.x, wt[4 * i], acc); acc = fmaf(z4.y, wt[4 * i + 1], acc);
                    acc = fmaf(z4.z, wt[4 * i + 2], acc); acc = fmaf(z4.w, wt[4 * i + 3], acc);
                }
            }
            const float g = gelu_erf(acc);
            a.u[base + (size_t)p * DIM] = acc; a.a_out[base + (size_t)p * DIM] = g;
            s1 += g; s2 = fmaf(g, g, s2);
        }
        __syncthreads();
    }
    if (KIND != 2) reduce_channel_sums(s1, s2, red, a.stat_part, a.stats_out, a.counter, nullptr, nullptr);
}

struct MixBwdArgs {
    float* dZ;                                                                              // in: grad of z_s; out: grad of z_{s-1}
    const float *a_s, *u_s, *mean_s, *rstd_s, *scale_s; const double* stats_s; double count; DropCfg drop_s;
    const float* w; const float* x;                                                          // conv weights of stage s, its input z_{s-1}
    const float *a_p, *mean_p, *rstd_p; DropCfg drop_p;                                      // BatchNorm s-1
    double* stat_part; double* stats_out; float* dgamma; float* dbeta; unsigned* counter;
    float* wpart;                                                                            // [grid][MX_WSTR]: weights then bias
    int B, Hp, Wp;
};

// KIND 0: stage s is a depthwise conv inside the Residual (odd s), 1: pointwise conv (even s)
template <int KIND, int K>
__global__ void __launch_bounds__(MX_THREADS, 2) mixer_bwd_kernel(const MixBwdArgs a) {
    constexpr int DIM = MX_DIM, kk = K * K, h = K / 2;
    __shared__ __align__(16) float sdu[MX_MAXP * DIM];
    __shared__ __align__(16) float sx[MX_MAXP * DIM];
    __shared__ __align__(16) float sxh[MX_MAXP * DIM];                                       // xhat of BatchNorm s-1
    __shared__ __align__(16) float sres[KIND == 0 ? MX_MAXP * DIM : 4];
    __shared__ double red[MX_PG][2][DIM];
    const int c = threadIdx.x % DIM, pg = threadIdx.x / DIM;
    const int Hp = a.Hp, Wp = a.Wp, P = Hp * Wp;
    const float mean_s = a.mean_s[c], rstd_s = a.rstd_s[c], scale_s = a.scale_s[c];
    const float m1 = (float)(a.stats_s[c] / a.count), m2 = (float)(a.stats_s[DIM + c] / a.count);
    const float mean_p = a.mean_p[c], rstd_p = a.rstd_p[c];
    float wt[KIND == 0 ? kk : DIM];
    if (KIND == 0) {
        for (int i = threadIdx.x; i < DIM * kk; i += MX_THREADS) sdu[i] = __ldg(a.w + i);     // coalesced, then odd-stride rows: conflict-free
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kk; ++i) wt[i] = sdu[c * kk + (kk - 1 - i)];                      // flipped taps: input gradient
        __syncthreads();
    } else {
#pragma unroll
        for (int o = 0; o < DIM; ++o) wt[o] = __ldg(a.w + o * DIM + c);                       // column c of W[out][in]
    }
    // weight-gradient share of (channel c, this thread's positions): KIND 0: dW[c][tap] over the forward's tap order; KIND 1: dW[o][c]
    float acc[KIND == 0 ? kk : DIM];
#pragma unroll
    for (int i = 0; i < (KIND == 0 ? kk : DIM); ++i) acc[i] = 0.f;
    float accb = 0.f;                             // bias gradient share
    float s1 = 0.f, s2 = 0.f;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const size_t base = (size_t)b * P * DIM + c;
        float gv[MX_PT], asv[MX_PT], usv[MX_PT], xv[MX_PT], apv[MX_PT];
#pragma unroll
        for (int i = 0; i < MX_PT; ++i) {                   // every load of the sample in flight before the first use
            const int p = pg + MX_PG * i;
            gv[i] = asv[i] = usv[i] = xv[i] = apv[i] = 0.f;
            if (p < P) {
                const size_t o = base + (size_t)p * DIM;
                gv[i] = a.dZ[o]; asv[i] = __ldg(a.a_s + o); usv[i] = __ldg(a.u_s + o); xv[i] = __ldg(a.x + o); apv[i] = __ldg(a.a_p + o);
            }
        }
#pragma unroll
        for (int i = 0; i < MX_PT; ++i) {
            const int p = pg + MX_PG * i;
            if (p >= P) break;
            float g = gv[i];
            if (a.drop_s.thresh) g *= drop_scale(a.drop_s, drop_rowkey(a.drop_s, (uint32_t)(b * P + p)), (uint32_t)c);
            const float xh = (asv[i] - mean_s) * rstd_s;
            const float da = scale_s * (g - m1 - xh * m2);
            const float du = da * gelu_erf_grad(usv[i]);
            sdu[p * DIM + c] = du;
            sx[p * DIM + c] = xv[i];
            sxh[p * DIM + c] = (apv[i] - mean_p) * rstd_p;
            if (KIND == 0) sres[p * DIM + c] = gv[i];
            accb += du;
        }
        __syncthreads();
        // one pass over the thread's positions: gradient of the convolution's input (= gradient of z_{s-1}), the sums of BatchNorm
        // s-1 backward, and the weight-gradient products -- both convolution gradients read the SAME neighbour (yy, xx)
        for (int p = pg; p < P; p += MX_PG) {
            float v;
            if (KIND == 0) {
                const int py = p / Wp, px = p % Wp;
                v = sres[p * DIM + c];                                                      // residual path: z_s = dropout(BN(.)) + x
                const float dup = sdu[p * DIM + c];
#pragma unroll
                for (int ii = 0; ii < K; ++ii) {
                    const int yy = py + ii - h;
                    if (yy < 0 || yy >= Hp) continue;
#pragma unroll
                    for (int jj = 0; jj < K; ++jj) {
                        const int xx = px + jj - h;
                        if (xx < 0 || xx >= Wp) continue;
                        const int q = (yy * Wp + xx) * DIM + c;
                        v = fmaf(sdu[q], wt[ii * K + jj], v);                               // flipped taps (wt is reversed)
                        acc[ii * K + jj] = fmaf(dup, sx[q], acc[ii * K + jj]);              // dW[c][tap] += du[p] x[p + off(tap)]
                    }
                }
            } else {
                v = 0.f;
                const float xq = sx[p * DIM + c];
                const float4* dr = reinterpret_cast<const float4*>(sdu + p * DIM);           // broadcast reads
#pragma unroll
                for (int i = 0; i < DIM / 4; ++i) {
                    const float4 d4 = dr[i];
                    v = fmaf(d4.x, wt[4 * i], v); v = fmaf(d4.y, wt[4 * i + 1], v);
                    v = fmaf(d4.z, wt[4 * i + 2], v); v = fmaf(d4.w, wt[4 * i + 3], v);
                    acc[4 * i] = fmaf(d4.x, xq, acc[4 * i]); acc[4 * i + 1] = fmaf(d4.y, xq, acc[4 * i + 1]);      // dW[o][c] += dU[p][o] y[p][c]
                    acc[4 * i + 2] = fmaf(d4.z, xq, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(d4.w, xq, acc[4 * i + 3]);
                }
            }
            a.dZ[base + (size_t)p * DIM] = v;
            float gs = v;
            if (a.drop_p.thresh) gs *= drop_scale(a.drop_p, drop_rowkey(a.drop_p, (uint32_t)(b * P + p)), (uint32_t)c);
            s1 += gs;
            s2 = fmaf(gs, sxh[p * DIM + c], s2);
        }
        __syncthreads();
    }
    // this CTA's slab of weight / bias gradient partials ([weights | bias]): the eight position groups add their shares through shared
    // memory one after the other (fixed order)
    {
        constexpr int NA = KIND == 0 ? kk : DIM;
        float* sacc = sdu;                                   // [NA + 1][DIM] <= 33 * 32 floats
        for (int k = 0; k < MX_PG; ++k) {
            if (pg == k) {
#pragma unroll
                for (int i = 0; i < NA; ++i) sacc[i * DIM + c] = (k == 0 ? 0.f : sacc[i * DIM + c]) + acc[i];
                sacc[NA * DIM + c] = (k == 0 ? 0.f : sacc[NA * DIM + c]) + accb;
            }
            __syncthreads();
        }
        float* pp = a.wpart + (size_t)blockIdx.x * MX_WSTR;
        for (int i = threadIdx.x; i < (NA + 1) * DIM; i += MX_THREADS) {
            const int r = i / DIM, cc = i % DIM;             // sacc[r][cc]: KIND 0: tap r of channel cc (r == kk: bias); KIND 1: dW[r][cc] (r == DIM: bias)
            if (KIND == 0) pp[r < kk ? cc * kk + r : DIM * kk + cc] = sacc[i];
            else pp[i] = sacc[i];                            // [o][c] row-major, bias row last = offset DIM*DIM
        }
    }
    reduce_channel_sums(s1, s2, red, a.stat_part, a.stats_out, a.counter, a.dbeta, a.dgamma);
}

// head backward: dZ of the last BatchNorm = dpooled / P broadcast over the positions, with the sums of that BatchNorm's backward
struct PoolBwdArgs {
    const float* dpooled; float* dZ; const float *a_p, *mean_p, *rstd_p; DropCfg drop_p;
    double* stat_part; double* stats_out; float* dgamma; float* dbeta; unsigned* counter;
    int B, P;
};
__global__ void __launch_bounds__(MX_THREADS) pool_bwd_stats_kernel(const PoolBwdArgs a) {
    constexpr int DIM = MX_DIM;
    __shared__ double red[MX_PG][2][DIM];
    const int c = threadIdx.x % DIM, pg = threadIdx.x / DIM;
    const float mean_p = a.mean_p[c], rstd_p = a.rstd_p[c], inv = 1.0f / (float)a.P;
    float s1 = 0.f, s2 = 0.f;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const float v = __ldg(a.dpooled + (size_t)b * DIM + c) * inv;
        float apv[MX_PT];
#pragma unroll
        for (int i = 0; i < MX_PT; ++i) {
            const int p = pg + MX_PG * i;
            apv[i] = p < a.P ? __ldg(a.a_p + ((size_t)b * a.P + p) * DIM + c) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < MX_PT; ++i) {
            const int p = pg + MX_PG * i;
            if (p >= a.P) break;
            a.dZ[((size_t)b * a.P + p) * DIM + c] = v;
            float gs = v;
            if (a.drop_p.thresh) gs *= drop_scale(a.drop_p, drop_rowkey(a.drop_p, (uint32_t)(b * a.P + p)), (uint32_t)c);
            s1 += gs;
            s2 = fmaf(gs, (apv[i] - mean_p) * rstd_p, s2);
        }
    }
    reduce_channel_sums(s1, s2, red, a.stat_part, a.stats_out, a.counter, a.dbeta, a.dgamma);
}

}  // namespace

static unsigned* mixer_counters() {          // resolved once, outside any stream capture
    static unsigned* p = nullptr;
    if (!p) cudaGetSymbolAddress((void**)&p, g_mixer_counters);
    return p;
}

bool mixer_fused_supported(int dim, int P, int k, int training) {
    static const int off = getenv("MVN_CONV_FUSED") ? atoi(getenv("MVN_CONV_FUSED")) == 0 : 0;      // MVN_CONV_FUSED=0: stage-by-stage kernels (A/B)
    return !off && training && dim == MX_DIM && P <= MX_MAXP && (k == 3 || k == 5);
}

// scratch the fused stages need for their per-CTA partials (statistics in double, then weight gradients)
static int mixer_grid(int B) { return B < MX_CTAS_PER_SM * kSlabs ? B : MX_CTAS_PER_SM * kSlabs; }
size_t mixer_scratch_bytes(int B) {
    const size_t G = (size_t)mixer_grid(B);
    return align_up(G * 2 * MX_DIM * sizeof(double), 256) + align_up(G * MX_WSTR * sizeof(float), 256);
}

int launch_mixer_fwd(int kind, int k, const double* stats_prev, double count, const float* gamma, const float* beta, float eps, float momentum,
                     float* running, float* mean_o, float* rstd_o, float* scale_o, float* shift_o, const float* a_prev, const float* res,
                     float* z_prev, const DropCfg& drop, const float* w, const float* bias, float* u, float* a_out, void* scratch,
                     double* stats_out, int stage, float* pooled, int B, int Hp, int Wp, cudaStream_t st) {
    unsigned* counter = mixer_counters();
    MVN_CHECK_ARG(counter != nullptr, "convmixer: arrival counters unavailable");
    counter += stage & 31;
    MixFwdArgs a;
    a.stats_prev = stats_prev; a.count = count; a.gamma = gamma; a.beta = beta; a.eps = eps; a.momentum = momentum; a.running = running;
    a.mean_o = mean_o; a.rstd_o = rstd_o; a.scale_o = scale_o; a.shift_o = shift_o;
    a.a_prev = a_prev; a.res = res; a.z_prev = z_prev; a.drop = drop;
    a.w = w; a.bias = bias; a.u = u; a.a_out = a_out; a.stat_part = (double*)scratch; a.stats_out = stats_out; a.counter = counter; a.pooled = pooled;
    a.B = B; a.Hp = Hp; a.Wp = Wp;
    const int grid = mixer_grid(B);
    if (kind == 2) mixer_fwd_kernel<2, 3><<<grid, MX_THREADS, 0, st>>>(a);
    else if (kind == 0 && k == 5) mixer_fwd_kernel<0, 5><<<grid, MX_THREADS, 0, st>>>(a);
    else if (kind == 0 && k == 3) mixer_fwd_kernel<0, 3><<<grid, MX_THREADS, 0, st>>>(a);
    else if (kind == 1) mixer_fwd_kernel<1, 3><<<grid, MX_THREADS, 0, st>>>(a);
    else return MVN_E_UNSUPPORTED;
    MVN_LAUNCH_CHECK();
    return 0;
}

// dW / db of the stage's convolution land in dW_out ([dim][k*k] or [dim][dim]) and db_out (contiguous after it in the flat layout)
int launch_mixer_bwd(int kind, int k, float* dZ, const float* a_s, const float* u_s, const float* mean_s, const float* rstd_s, const float* scale_s,
                     const double* stats_s, double count, const DropCfg& drop_s, const float* w, const float* x, const float* a_p,
                     const float* mean_p, const float* rstd_p, const DropCfg& drop_p, void* scratch, double* stats_out, float* dgamma,
                     float* dbeta, int stage, float* dW_out, int B, int Hp, int Wp, cudaStream_t st) {
    unsigned* counter = mixer_counters();
    MVN_CHECK_ARG(counter != nullptr, "convmixer: arrival counters unavailable");
    counter += 32 + (stage & 31);
    const int grid = mixer_grid(B);
    MixBwdArgs a;
    a.dZ = dZ; a.a_s = a_s; a.u_s = u_s; a.mean_s = mean_s; a.rstd_s = rstd_s; a.scale_s = scale_s; a.stats_s = stats_s; a.count = count; a.drop_s = drop_s;
    a.w = w; a.x = x; a.a_p = a_p; a.mean_p = mean_p; a.rstd_p = rstd_p; a.drop_p = drop_p;
    a.stat_part = (double*)scratch; a.stats_out = stats_out; a.dgamma = dgamma; a.dbeta = dbeta; a.counter = counter;
    a.wpart = (float*)((char*)scratch + align_up((size_t)grid * 2 * MX_DIM * sizeof(double), 256));
    a.B = B; a.Hp = Hp; a.Wp = Wp;
    if (kind == 0 && k == 5) mixer_bwd_kernel<0, 5><<<grid, MX_THREADS, 0, st>>>(a);
    else if (kind == 0 && k == 3) mixer_bwd_kernel<0, 3><<<grid, MX_THREADS, 0, st>>>(a);
    else if (kind == 1) mixer_bwd_kernel<1, 3><<<grid, MX_THREADS, 0, st>>>(a);
    else return MVN_E_UNSUPPORTED;
    MVN_LAUNCH_CHECK();
    const size_t wn = kind == 0 ? (size_t)MX_DIM * k * k : (size_t)MX_DIM * MX_DIM;
    return launch_reduce_partials_n(a.wpart, MX_WSTR, wn + MX_DIM, grid, dW_out, 0, st);      // every CTA owns >= 1 sample: all slabs written
}

int launch_pool_bwd_stats(const float* dpooled, float* dZ, const float* a_p, const float* mean_p, const float* rstd_p, const DropCfg& drop_p,
                          void* scratch, double* stats_out, float* dgamma, float* dbeta, int stage, int B, int P, cudaStream_t st) {
    unsigned* counter = mixer_counters();
    MVN_CHECK_ARG(counter != nullptr, "convmixer: arrival counters unavailable");
    counter += 32 + (stage & 31);
    PoolBwdArgs a;
    a.dpooled = dpooled; a.dZ = dZ; a.a_p = a_p; a.mean_p = mean_p; a.rstd_p = rstd_p; a.drop_p = drop_p;
    a.stat_part = (double*)scratch; a.stats_out = stats_out; a.dgamma = dgamma; a.dbeta = dbeta; a.counter = counter; a.B = B; a.P = P;
    pool_bwd_stats_kernel<<<mixer_grid(B), MX_THREADS, 0, st>>>(a);
    MVN_LAUNCH_CHECK();
    return 0;
}

}  // namespace mvn
