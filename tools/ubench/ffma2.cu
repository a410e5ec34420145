// Micro-benchmark: FP32 FMA issue throughput on sm_100a, scalar FFMA vs packed fma.rn.f32x2 (FFMA2), and the mixed case
// (FMA chain + shared-memory loads) that resembles the particle kernels.  Build: nvcc -arch=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int ILP>
__global__ void k_scalar(float* out, int iters, float b, float c) {
    float a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) a[j] = threadIdx.x * 1e-3f + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) a[j] = fmaf(a[j], b, c);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_packed(float* out, int iters, float b, float c) {
    unsigned long long a[ILP];
    unsigned long long bb = pack2(b, b), cc = pack2(c, c);
#pragma unroll
    for (int j = 0; j < ILP; j++) a[j] = pack2(threadIdx.x * 1e-3f + j, threadIdx.x * 2e-3f + j);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) a[j] = fma2(a[j], bb, cc);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) { float x, y; unpack2(a[j], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
float time_kernel(K k, int blocks, int threads, float* out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float* out;
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    for (int warps_per_sm_scale = 0; warps_per_sm_scale < 2; warps_per_sm_scale++) {
        int nb = warps_per_sm_scale == 0 ? blocks : 148;          // 64 warps / SM vs 8 warps / SM (2 per scheduler)
        double lanes = (double)nb * threads;
        float m1 = time_kernel(k_scalar<8>, nb, threads, out, iters);
        float m2 = time_kernel(k_packed<8>, nb, threads, out, iters);
        float m3 = time_kernel(k_scalar<2>, nb, threads, out, iters);
        float m4 = time_kernel(k_packed<2>, nb, threads, out, iters);
        printf("blocks %4d: scalar ILP8 %.3f ms %.1f TFLOP/s | packed ILP8 %.3f ms %.1f TFLOP/s | scalar ILP2 %.3f ms %.1f | packed ILP2 %.3f ms %.1f\n", nb,
               m1, 2.0 * lanes * iters * 8 / m1 * 1e-9, m2, 4.0 * lanes * iters * 8 / m2 * 1e-9,
               m3, 2.0 * lanes * iters * 2 / m3 * 1e-9, m4, 4.0 * lanes * iters * 2 / m4 * 1e-9);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
