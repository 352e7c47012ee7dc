// Latency / throughput of the legacy warp-level tensor path on sm_100a (mma.sync m16n8k8 tf32, ldmatrix), measured
// with clock64: decides how the batched decode kernel blocks its skinny GEMMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void probe(float* out, long long* cyc, int iters) {
    float c[CHAINS][4];
    for (int k = 0; k < CHAINS; ++k) c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) mma(c[k], a0, a1, a2, a3, b0, b1);
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int k = 0; k < CHAINS; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void ffma_probe(float* out, long long* cyc, int iters) {
    float c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float a = threadIdx.x * 0.001f, b = 1.0001f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = fmaf(a, b, c[k]);
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    long long h;
#define RUN(CH, WARPS)                                                                                   \
    probe<CH><<<148, 32 * WARPS>>>(out, cyc, iters);                                                     \
    cudaDeviceSynchronize();                                                                             \
    probe<CH><<<148, 32 * WARPS>>>(out, cyc, iters);                                                     \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                                      \
    printf("mma.m16n8k8.tf32  chains/warp=%d warps/SM=%d : %.2f cycles per mma per warp, %.2f cycles per mma per SM\n", CH, WARPS, \
           (double)h / (iters * CH), (double)h / (iters * CH * WARPS));
    RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(4, 4) RUN(1, 8) RUN(2, 8) RUN(4, 8) RUN(8, 8) RUN(4, 16)
    ffma_probe<<<148, 256>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ffma 8 chains 8 warps: %.2f cycles per warp-FFMA per warp\n", (double)h / (iters * 8));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
