// Microbenchmark: issue rate of the legacy warp-level tensor path on sm_100a (mma.sync m16n8k8 tf32, m16n8k16 bf16, m16n8k4 tf32)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int ACC>
__global__ void k(float* out, int iters) {
    float c[ACC][4];
    for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    float s = 0;
    for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int ACC>
void run(const char* name, int warps_per_sm, double macs) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4 * 8);
    const int iters = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<KIND, ACC><<<148, warps_per_sm * 32>>>(out, 100);
    cudaEventRecord(a);
    k<KIND, ACC><<<148, warps_per_sm * 32>>>(out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double n = 148.0 * warps_per_sm * iters * ACC;
    printf("%-22s acc=%d warps/SM=%2d : %.3f ms  %.2f cycles/MMA/SMSP (at 1.965 GHz)  %.1f TFLOP/s dense\n", name, ACC, warps_per_sm, ms,
           ms * 1e-3 * 1.965e9 / (n / (148.0 * 4)), n * macs * 2 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    run<0, 1>("m16n8k8 tf32", 4, 2048); run<0, 4>("m16n8k8 tf32", 4, 2048); run<0, 4>("m16n8k8 tf32", 16, 2048); run<0, 8>("m16n8k8 tf32", 32, 2048);
    run<1, 1>("m16n8k16 bf16", 4, 2048); run<1, 4>("m16n8k16 bf16", 4, 2048); run<1, 4>("m16n8k16 bf16", 16, 2048); run<1, 8>("m16n8k16 bf16", 32, 2048);
    run<2, 1>("m16n8k4 tf32", 4, 512); run<2, 4>("m16n8k4 tf32", 16, 512);
    return 0;
}
