// N1: GPU-side batch augmentation of NoisyDataLoader.__iter__ (src/dataloader.py:88-287) --
//   light curves / spectra:  x + randn * err * noise_level                                   (:125, :136, ...)
//   host images:             rot90^k( img + (2*rand - 1) * max_noise_intensity * std(imgs) )  (:93-112)
// plus the 8-bit image upload: the PNG cut-outs are 8-bit, the reference stores them as fp32/255 (4x the H2D and HBM bytes).
// All three are one pass over the data (HBM-bound).  The random numbers are either GIVEN (parity: the result is then
// bit-identical to the torch expression, every product and sum rounded separately like eager torch) or drawn in-kernel from
// the counter-based generator the dropout sites use.
#include "common.cuh"

namespace mvn {
namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }        // [0, 1)
__device__ __forceinline__ float gauss(uint64_t seed, uint64_t i) {                                        // Box-Muller
    const uint32_t a = mix32((uint32_t)i ^ (uint32_t)seed), b = mix32((uint32_t)(i >> 32) ^ (uint32_t)(seed >> 32) ^ (a * 0x9E3779B9u));
    const float u1 = ((float)(mix32(a ^ b) >> 8) + 1.0f) * (1.0f / 16777216.0f);                           // (0, 1]
    const float u2 = u01(mix32(b + 0x85EBCA77u));
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__global__ void augment_seq_kernel(const float* __restrict__ x, const float* __restrict__ err, const float* __restrict__ noise, float level,
                                   size_t n, uint64_t seed, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float z = noise ? noise[i] : gauss(seed, i);
        out[i] = __fadd_rn(x[i], __fmul_rn(__fmul_rn(z, err[i]), level));           // mag + randn * magerr * level, torch's rounding order
    }
}

// per-CTA partial sums (double) of v and v^2, v = img (fp32) or img/255 (uint8)
__global__ void __launch_bounds__(256) image_moments_kernel(const void* __restrict__ img, int is_u8, size_t n, double* __restrict__ part) {
    __shared__ double r1[8], r2[8];
    double s1 = 0.0, s2 = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = is_u8 ? __fdiv_rn((float)reinterpret_cast<const uint8_t*>(img)[i], 255.0f) : reinterpret_cast<const float*>(img)[i];
        s1 += (double)v; s2 += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += r1[w]; b += r2[w]; }
        part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    }
}
// range = intensity * std (unbiased, like torch.std)
__global__ void image_range_kernel(const double* __restrict__ part, int nblk, double n, float intensity, float* __restrict__ range) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nblk; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
    const double mean = a / n;
    double var = (b - n * mean * mean) / (n > 1.0 ? n - 1.0 : 1.0);
    if (var < 0.0) var = 0.0;
    range[0] = __fmul_rn(intensity, (float)sqrt(var));
}
// out[b,c,y,x] = noisy[b,c,sy,sx], (sy,sx) = source of (y,x) under a counter-clockwise rotation by 90*k degrees (square images)
__global__ void augment_images_kernel(const void* __restrict__ img, int is_u8, const float* __restrict__ noise_u, const int32_t* __restrict__ rot_k,
                                      const float* __restrict__ range, uint64_t seed, int B, int C, int H, int W, float* __restrict__ out) {
    const size_t per = (size_t)C * H * W, total = (size_t)B * per;
    const float rg = range ? range[0] : 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const size_t bc = i / ((size_t)H * W);
        const int b = (int)(bc / C);
        const int k = rot_k ? (rot_k[b] & 3) : 0;
        int sy = y, sx = x;
        if (k == 1) { sy = x; sx = W - 1 - y; }
        else if (k == 2) { sy = H - 1 - y; sx = W - 1 - x; }
        else if (k == 3) { sy = H - 1 - x; sx = y; }
        const size_t src = bc * (size_t)H * W + (size_t)sy * W + sx;                 // index in the planar [B,C,H,W] frame (noise_u lives there)
        float v;
        if (is_u8 == 2) {                                                             // raw decoder layout [B,H,W,C] ("b h w c -> b c h w", :327)
            const int c = (int)(bc % C);
            v = __fdiv_rn((float)reinterpret_cast<const uint8_t*>(img)[(((size_t)b * H + sy) * W + sx) * C + c], 255.0f);
        } else if (is_u8 == 1) {
            v = __fdiv_rn((float)reinterpret_cast<const uint8_t*>(img)[src], 255.0f);
        } else {
            v = reinterpret_cast<const float*>(img)[src];
        }
        const float u = noise_u ? noise_u[src] : u01(mix32((uint32_t)src ^ (uint32_t)seed) ^ mix32((uint32_t)(src >> 32) + (uint32_t)(seed >> 32)));
        out[i] = __fadd_rn(v, __fmul_rn(__fadd_rn(__fmul_rn(2.0f, u), -1.0f), rg));   // img + (2*rand - 1) * range
    }
}

inline int grid_for(size_t n) { size_t b = (n + 255) / 256; const size_t cap = (size_t)num_sms() * 8; return (int)(b < cap ? (b ? b : 1) : cap); }

}  // namespace
}  // namespace mvn

using namespace mvn;

extern "C" int mvn_augment_seq(const float* x, const float* err, const float* noise, float level, int64_t n, uint64_t seed, float* out, void* stream) {
    MVN_CHECK_ARG(x && err && out && n > 0, "augment_seq: bad arguments");
    ProfScope prof(PROF_ROW, (cudaStream_t)stream);
    augment_seq_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>(x, err, noise, level, (size_t)n, seed, out);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" size_t mvn_image_noise_range_workspace_bytes(void) { return (size_t)kSlabs * 2 * sizeof(double); }
extern "C" int mvn_image_noise_range(const void* img, int is_u8, int64_t n, float max_noise_intensity, float* range_out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    MVN_CHECK_ARG(img && range_out && workspace && n > 0, "image_noise_range: bad arguments");
    if (workspace_bytes < mvn_image_noise_range_workspace_bytes()) { set_error("image_noise_range: workspace too small"); return MVN_E_WORKSPACE; }
    ProfScope prof(PROF_ROW, (cudaStream_t)stream);
    image_moments_kernel<<<kSlabs, 256, 0, (cudaStream_t)stream>>>(img, is_u8, (size_t)n, (double*)workspace);
    MVN_LAUNCH_CHECK();
    image_range_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const double*)workspace, kSlabs, (double)n, max_noise_intensity, range_out);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_augment_images(const void* img, int is_u8, const float* noise_u, const int32_t* rot_k, const float* range_dev, uint64_t seed,
                                  int B, int C, int H, int W, float* out, void* stream) {
    MVN_CHECK_ARG(img && out && B > 0 && C > 0 && H > 0 && W > 0 && is_u8 >= 0 && is_u8 <= 2, "augment_images: bad arguments");
    MVN_UNSUPPORTED(rot_k == nullptr || H == W, "augment_images: 90-degree rotations need square images, got %dx%d", H, W);
    ProfScope prof(PROF_ROW, (cudaStream_t)stream);
    augment_images_kernel<<<grid_for((size_t)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(img, is_u8, noise_u, rot_k, range_dev, seed, B, C, H, W, out);
    MVN_LAUNCH_CHECK();
    return 0;
}
