// A10/A11: symmetric InfoNCE (src/loss.py:14-38), streamed.  One kernel per direction: a CTA owns 64 rows and a
// range of column tiles, recomputes 64x64 logit tiles on the fly and folds them into an online log-sum-exp
// (forward) or into G = (P_row + P_col - 2I)/(2N) and dR += G * C (backward).  The N x N matrix is never stored.
// Deterministic: column splits write partials that a second kernel combines in a fixed order.
#include "common.cuh"

namespace mvn {
namespace {

constexpr int TB = 64;     // tile rows / cols
constexpr int KB = 16;

struct Split { int row_tiles, col_tiles, nsplit, tiles_per_split; };
Split make_split(int n, int N) {
    Split s;
    s.row_tiles = cdiv(n, TB);
    s.col_tiles = cdiv(N, TB);
    int want = cdiv(2 * num_sms(), s.row_tiles);
    if (want < 1) want = 1;
    if (want > s.col_tiles) want = s.col_tiles;
    s.tiles_per_split = cdiv(s.col_tiles, want);
    s.nsplit = cdiv(s.col_tiles, s.tiles_per_split);
    return s;
}

// computes the raw dot-product tile acc[i][j] = R[m0+ty*4+i] . C[c0+tx*4+j]
__device__ __forceinline__ void dot_tile(const float* __restrict__ R, const float* __restrict__ C, int nr, int N, int D, int m0, int c0,
                                         float (*Rs)[TB + 4], float (*Cs)[TB + 4], float acc[4][4]) {
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid / 4, lk = (tid % 4) * 4;
    for (int k0 = 0; k0 < D; k0 += KB) {
        float4 rv = make_float4(0.f, 0.f, 0.f, 0.f), cv = rv;
        if (m0 + lr < nr) rv = *reinterpret_cast<const float4*>(R + (size_t)(m0 + lr) * D + k0 + lk);
        if (c0 + lr < N) cv = *reinterpret_cast<const float4*>(C + (size_t)(c0 + lr) * D + k0 + lk);
        __syncthreads();
        Rs[lk + 0][lr] = rv.x; Rs[lk + 1][lr] = rv.y; Rs[lk + 2][lr] = rv.z; Rs[lk + 3][lr] = rv.w;
        Cs[lk + 0][lr] = cv.x; Cs[lk + 1][lr] = cv.y; Cs[lk + 2][lr] = cv.z; Cs[lk + 3][lr] = cv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&Rs[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Cs[kk][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
}

// forward: partial online LSE of rows R against the CTA's column range.  pm/pl are [nsplit][nr].
__global__ void __launch_bounds__(256) lse_dir_kernel(const float* __restrict__ R, const float* __restrict__ C, int nr, int N, int D,
                                                      const float* __restrict__ logit_scale, const float* __restrict__ logit_bias,
                                                      int tiles_per_split, float* __restrict__ pm, float* __restrict__ pl) {
    __shared__ __align__(16) float Rs[KB][TB + 4];
    __shared__ __align__(16) float Cs[KB][TB + 4];
    const float s = expf(*logit_scale), bz = *logit_bias;
    const int m0 = blockIdx.x * TB, split = blockIdx.y;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int ct0 = split * tiles_per_split;
    const int ct1 = min(ct0 + tiles_per_split, (N + TB - 1) / TB);
    float mrun[4], lrun[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { mrun[i] = -INFINITY; lrun[i] = 0.f; }
    for (int ct = ct0; ct < ct1; ++ct) {
        const int c0 = ct * TB;
        float acc[4][4];
        dot_tile(R, C, nr, N, D, m0, c0, Rs, Cs, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float z[4], tm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                z[j] = (c0 + tx * 4 + j < N) ? fmaf(acc[i][j], s, bz) : -INFINITY;
                tm = fmaxf(tm, z[j]);
            }
            tm = group_max<16>(tm);
            const float mnew = fmaxf(mrun[i], tm);
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) ps += expf(z[j] - mnew);       // exp(-inf) = 0 for out-of-range columns
            ps = group_sum<16>(ps);
            lrun[i] = lrun[i] * expf(mrun[i] - mnew) + ps;
            mrun[i] = mnew;
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m < nr) { pm[(size_t)split * nr + m] = mrun[i]; pl[(size_t)split * nr + m] = lrun[i]; }
        }
    }
}

// merge splits -> lse ; one warp per local row; also the per-row loss term lse_row + lse_col - 2 z_ii
__global__ void __launch_bounds__(256) lse_finish_kernel(const float* __restrict__ pm_r, const float* __restrict__ pl_r,
                                                         const float* __restrict__ pm_c, const float* __restrict__ pl_c, int nsplit,
                                                         const float* __restrict__ e1_local, const float* __restrict__ e2_local, int n, int D,
                                                         const float* __restrict__ logit_scale, const float* __restrict__ logit_bias,
                                                         float* __restrict__ lse_row, float* __restrict__ lse_col, float* __restrict__ terms) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    float out[2];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        const float* pm = dir ? pm_c : pm_r;
        const float* pl = dir ? pl_c : pl_r;
        float m = -INFINITY;
        for (int sp = lane; sp < nsplit; sp += 32) m = fmaxf(m, pm[(size_t)sp * n + row]);
        m = warp_max(m);
        float l = 0.f;
        for (int sp = lane; sp < nsplit; sp += 32) l += pl[(size_t)sp * n + row] * expf(pm[(size_t)sp * n + row] - m);
        l = warp_sum(l);
        out[dir] = m + logf(l);
    }
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) dot = fmaf(e2_local[(size_t)row * D + d], e1_local[(size_t)row * D + d], dot);
    dot = warp_sum(dot);
    if (lane == 0) {
        const float zii = fmaf(dot, expf(*logit_scale), *logit_bias);
        lse_row[row] = out[0];
        lse_col[row] = out[1];
        terms[row] = (out[0] - zii) + (out[1] - zii);
    }
}

// deterministic single-CTA sum: out[0] = factor * sum(v[0..n)) (* *gmul if given)
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ v, int n, float factor, const float* __restrict__ gmul, float* __restrict__ out) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = warp_sum(red[threadIdx.x]);
        if (threadIdx.x == 0) out[0] = s * factor * (gmul ? *gmul : 1.0f);
    }
}

// backward: dR partial [nsplit][nr][D] (unscaled) and per-CTA partial of sum G (z-b)
__global__ void __launch_bounds__(256) grad_dir_kernel(const float* __restrict__ R, const float* __restrict__ C, int nr, int N, int D,
                                                       const float* __restrict__ logit_scale, const float* __restrict__ logit_bias,
                                                       const float* __restrict__ lse_R, const float* __restrict__ lse_C, int row_offset,
                                                       int tiles_per_split, float* __restrict__ dR_part, float* __restrict__ dls_part) {
    __shared__ __align__(16) float Rs[KB][TB + 4];
    __shared__ __align__(16) float Cs[KB][TB + 4];
    __shared__ float Gs[TB][TB + 1];
    __shared__ float red[8];
    const float s = expf(*logit_scale), bz = *logit_bias;
    const float inv2n = 0.5f / (float)N;
    const int m0 = blockIdx.x * TB, split = blockIdx.y;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int ct0 = split * tiles_per_split;
    const int ct1 = min(ct0 + tiles_per_split, (N + TB - 1) / TB);
    float lr_[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int m = m0 + ty * 4 + i; lr_[i] = m < nr ? lse_R[m] : 0.f; }
    float dls = 0.f;
    // second-GEMM micro tile: rows ty*4..+4, columns tx*8..+8 of each 128-wide slice of D
    const int nslice = D / 128;       // D is a multiple of 128 here (checked on the host)
    float dacc[2][4][8];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dacc[sl][i][j] = 0.f;

    for (int ct = ct0; ct < ct1; ++ct) {
        const int c0 = ct * TB;
        float acc[4][4];
        dot_tile(R, C, nr, N, D, m0, c0, Rs, Cs, acc);
        float lc_[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int c = c0 + tx * 4 + j; lc_[j] = c < N ? lse_C[c] : 0.f; }
        __syncthreads();          // previous tile's Gs readers are done
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + tx * 4 + j;
                float g = 0.f;
                if (m < nr && c < N) {
                    const float z = fmaf(acc[i][j], s, bz);
                    g = (expf(z - lr_[i]) + expf(z - lc_[j]) - ((m + row_offset == c) ? 2.0f : 0.0f)) * inv2n;
                    dls = fmaf(g, z - bz, dls);
                }
                Gs[ty * 4 + i][tx * 4 + j] = g;
            }
        }
        __syncthreads();
        const int jn = min(TB, N - c0);
        for (int sl = 0; sl < nslice && sl < 2; ++sl) {
            for (int j = 0; j < jn; ++j) {
                const float* cp = C + (size_t)(c0 + j) * D + sl * 128 + tx * 8;
                const float4 c0v = *reinterpret_cast<const float4*>(cp);
                const float4 c1v = *reinterpret_cast<const float4*>(cp + 4);
                const float cr[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float g = Gs[ty * 4 + i][j];
#pragma unroll
                    for (int q = 0; q < 8; ++q) dacc[sl][i][q] = fmaf(g, cr[q], dacc[sl][i][q]);
                }
            }
        }
    }
    for (int sl = 0; sl < nslice && sl < 2; ++sl) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m >= nr) continue;
            float* o = dR_part + ((size_t)split * nr + m) * D + sl * 128 + tx * 8;
            *reinterpret_cast<float4*>(o) = make_float4(dacc[sl][i][0], dacc[sl][i][1], dacc[sl][i][2], dacc[sl][i][3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(dacc[sl][i][4], dacc[sl][i][5], dacc[sl][i][6], dacc[sl][i][7]);
        }
    }
    if (dls_part) {
        dls = warp_sum(dls);
        if ((tid & 31) == 0) red[tid >> 5] = dls;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w];
            dls_part[blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
}

// out[i] = g * s * sum_split part[split][i]
__global__ void __launch_bounds__(256) grad_reduce_kernel(const float* __restrict__ part, int nsplit, size_t n, const float* __restrict__ logit_scale,
                                                          const float* __restrict__ grad_out, float* __restrict__ out) {
    const float f = expf(*logit_scale) * (grad_out ? *grad_out : 1.0f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) s += part[(size_t)sp * n + i];
        out[i] = s * f;
    }
}

}  // namespace

// clip_loss_tc.cu: the same two passes on tcgen05 (prec >= 1, D == 128)
bool tc_loss_supported(int n, int N, int D);
void tc_loss_split(int n, int N, int nc, int* row_tiles, int* nsplit, int* tiles_per_split);
int launch_lse_tc(const float* R, const float* C, int nr, int N, const float* logit_scale, const float* logit_bias, int tiles_per_split, int row_tiles,
                  int nsplit, float* pm, float* pl, cudaStream_t st);
int launch_grad_tc(const float* R, const float* C, int nr, int N, const float* logit_scale, const float* logit_bias, const float* lse_R,
                   const float* lse_C, int row_offset, int tiles_per_split, int row_tiles, int nsplit, float* dR_part, float* dls_part, cudaStream_t st);
}  // namespace mvn

using namespace mvn;

extern "C" size_t mvn_clip_loss_workspace_bytes(int n, int N, int D) {
    if (n <= 0 || N <= 0 || D <= 0) return 0;
    const Split sp = make_split(n, N);
    size_t fwd = (size_t)4 * sp.nsplit * n + n;                                  // pm/pl x 2 directions + terms
    size_t bwd = (size_t)2 * sp.nsplit * n * D + (size_t)sp.nsplit * sp.row_tiles;   // dR partials x 2 + dls partials
    if (tc_loss_supported(n, N, D)) {                                            // the tensor-core passes split the columns differently
        int rt, ns, tps;
        tc_loss_split(n, N, 128, &rt, &ns, &tps);
        const size_t f2 = (size_t)8 * ns * n + n;                                // two partial (max, sum) slots per column split: one per column half
        tc_loss_split(n, N, 64, &rt, &ns, &tps);
        const size_t b2 = (size_t)2 * ns * n * D + (size_t)ns * rt;
        fwd = fwd > f2 ? fwd : f2;
        bwd = bwd > b2 ? bwd : b2;
    }
    return (fwd > bwd ? fwd : bwd) * sizeof(float) + 256;
}

static int check_loss_args(const float* a, const float* b, const float* c, const float* d, int n, int N, int D, int row_offset,
                           const float* ls, const float* lb, void* ws, size_t ws_bytes) {
    MVN_CHECK_ARG(a && b && c && d && ls && lb && ws, "clip_loss: null pointer");
    MVN_CHECK_ARG(n > 0 && N >= n && row_offset >= 0 && row_offset + n <= N, "clip_loss: bad sizes n=%d N=%d offset=%d", n, N, row_offset);
    MVN_UNSUPPORTED(D % 128 == 0 && D <= 256, "clip_loss: embedding dim %d must be 128 or 256", D);
    MVN_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d), "clip_loss: embeddings must be 16-byte aligned");
    if (ws_bytes < mvn_clip_loss_workspace_bytes(n, N, D)) { set_error("clip_loss: workspace %zu < %zu", ws_bytes, mvn_clip_loss_workspace_bytes(n, N, D)); return MVN_E_WORKSPACE; }
    return 0;
}

extern "C" int mvn_clip_loss_fwd(const float* e1_local, const float* e2_local, const float* e1_all, const float* e2_all, int n, int N,
                                 int D, int row_offset, const float* logit_scale, const float* logit_bias, float* loss_out,
                                 float* lse_row, float* lse_col, void* workspace, size_t workspace_bytes, int prec, void* stream) {
    MVN_TRY(check_loss_args(e1_local, e2_local, e1_all, e2_all, n, N, D, row_offset, logit_scale, logit_bias, workspace, workspace_bytes));
    MVN_CHECK_ARG(loss_out && lse_row && lse_col, "clip_loss_fwd: null outputs");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_LOSS, st);
    float* ws = (float*)workspace;
    if (prec >= 1 && tc_loss_supported(n, N, D)) {
        int rt, ns, tps;
        tc_loss_split(n, N, 128, &rt, &ns, &tps);
        const size_t np = (size_t)2 * ns * n;                                    // the kernel writes two partial slots per split (column halves)
        float* pm_r = ws;              float* pl_r = pm_r + np;
        float* pm_c = pl_r + np;       float* pl_c = pm_c + np;
        float* terms = pl_c + np;
        MVN_TRY(launch_lse_tc(e2_local, e1_all, n, N, logit_scale, logit_bias, tps, rt, ns, pm_r, pl_r, st));
        MVN_TRY(launch_lse_tc(e1_local, e2_all, n, N, logit_scale, logit_bias, tps, rt, ns, pm_c, pl_c, st));
        count_tier(TIER_TC);
        lse_finish_kernel<<<cdiv(n * 32, 256), 256, 0, st>>>(pm_r, pl_r, pm_c, pl_c, 2 * ns, e1_local, e2_local, n, D, logit_scale, logit_bias,
                                                             lse_row, lse_col, terms);
        MVN_LAUNCH_CHECK();
        sum_kernel<<<1, 1024, 0, st>>>(terms, n, 0.5f / (float)N, nullptr, loss_out);
        MVN_LAUNCH_CHECK();
        return 0;
    }
    count_tier(TIER_FFMA);
    const Split sp = make_split(n, N);
    float* pm_r = ws;                       float* pl_r = pm_r + (size_t)sp.nsplit * n;
    float* pm_c = pl_r + (size_t)sp.nsplit * n; float* pl_c = pm_c + (size_t)sp.nsplit * n;
    float* terms = pl_c + (size_t)sp.nsplit * n;
    dim3 grid(sp.row_tiles, sp.nsplit);
    // dim=1: rows e2_local against all e1 ; dim=0: columns j <-> rows e1_local against all e2
    lse_dir_kernel<<<grid, 256, 0, st>>>(e2_local, e1_all, n, N, D, logit_scale, logit_bias, sp.tiles_per_split, pm_r, pl_r);
    MVN_LAUNCH_CHECK();
    lse_dir_kernel<<<grid, 256, 0, st>>>(e1_local, e2_all, n, N, D, logit_scale, logit_bias, sp.tiles_per_split, pm_c, pl_c);
    MVN_LAUNCH_CHECK();
    lse_finish_kernel<<<cdiv(n * 32, 256), 256, 0, st>>>(pm_r, pl_r, pm_c, pl_c, sp.nsplit, e1_local, e2_local, n, D, logit_scale, logit_bias,
                                                         lse_row, lse_col, terms);
    MVN_LAUNCH_CHECK();
    sum_kernel<<<1, 1024, 0, st>>>(terms, n, 0.5f / (float)N, nullptr, loss_out);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_clip_loss_bwd(const float* e1_local, const float* e2_local, const float* e1_all, const float* e2_all, int n, int N,
                                 int D, int row_offset, const float* logit_scale, const float* logit_bias, const float* lse_row_all,
                                 const float* lse_col_all, const float* grad_out, float* d_e1_local, float* d_e2_local,
                                 float* d_logit_scale, void* workspace, size_t workspace_bytes, int prec, void* stream) {
    MVN_TRY(check_loss_args(e1_local, e2_local, e1_all, e2_all, n, N, D, row_offset, logit_scale, logit_bias, workspace, workspace_bytes));
    MVN_CHECK_ARG(lse_row_all && lse_col_all && d_e1_local && d_e2_local && d_logit_scale, "clip_loss_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_LOSS, st);
    float* ws = (float*)workspace;
    if (prec >= 1 && tc_loss_supported(n, N, D)) {
        int rt, ns, tps;
        tc_loss_split(n, N, 64, &rt, &ns, &tps);
        float* part2 = ws;
        float* part1 = part2 + (size_t)ns * n * D;
        float* dls_part = part1 + (size_t)ns * n * D;
        MVN_TRY(launch_grad_tc(e2_local, e1_all, n, N, logit_scale, logit_bias, lse_row_all + row_offset, lse_col_all, row_offset, tps, rt, ns, part2,
                               dls_part, st));
        MVN_TRY(launch_grad_tc(e1_local, e2_all, n, N, logit_scale, logit_bias, lse_col_all + row_offset, lse_row_all, row_offset, tps, rt, ns, part1,
                               nullptr, st));
        count_tier(TIER_TC);
        const size_t ne = (size_t)n * D;
        const int blocks = (int)((ne + 255) / 256 < 2048 ? (ne + 255) / 256 : 2048);
        grad_reduce_kernel<<<blocks, 256, 0, st>>>(part2, ns, ne, logit_scale, grad_out, d_e2_local);
        MVN_LAUNCH_CHECK();
        grad_reduce_kernel<<<blocks, 256, 0, st>>>(part1, ns, ne, logit_scale, grad_out, d_e1_local);
        MVN_LAUNCH_CHECK();
        sum_kernel<<<1, 1024, 0, st>>>(dls_part, ns * rt, 1.0f, grad_out, d_logit_scale);
        MVN_LAUNCH_CHECK();
        return 0;
    }
    count_tier(TIER_FFMA);
    const Split sp = make_split(n, N);
    float* part2 = ws;                                        // d_e2 partials  [nsplit][n][D]
    float* part1 = part2 + (size_t)sp.nsplit * n * D;         // d_e1 partials
    float* dls_part = part1 + (size_t)sp.nsplit * n * D;      // [nsplit*row_tiles]
    dim3 grid(sp.row_tiles, sp.nsplit);
    // rows i = e2_local: own LSE = lse_row (local slice), other = lse_col for all columns j
    grad_dir_kernel<<<grid, 256, 0, st>>>(e2_local, e1_all, n, N, D, logit_scale, logit_bias, lse_row_all + row_offset, lse_col_all, row_offset,
                                          sp.tiles_per_split, part2, dls_part);
    MVN_LAUNCH_CHECK();
    grad_dir_kernel<<<grid, 256, 0, st>>>(e1_local, e2_all, n, N, D, logit_scale, logit_bias, lse_col_all + row_offset, lse_row_all, row_offset,
                                          sp.tiles_per_split, part1, nullptr);
    MVN_LAUNCH_CHECK();
    const size_t ne = (size_t)n * D;
    const int blocks = (int)((ne + 255) / 256 < 2048 ? (ne + 255) / 256 : 2048);
    grad_reduce_kernel<<<blocks, 256, 0, st>>>(part2, sp.nsplit, ne, logit_scale, grad_out, d_e2_local);
    MVN_LAUNCH_CHECK();
    grad_reduce_kernel<<<blocks, 256, 0, st>>>(part1, sp.nsplit, ne, logit_scale, grad_out, d_e1_local);
    MVN_LAUNCH_CHECK();
    sum_kernel<<<1, 1024, 0, st>>>(dls_part, sp.nsplit * sp.row_tiles, 1.0f, grad_out, d_logit_scale);
    MVN_LAUNCH_CHECK();
    return 0;
}
