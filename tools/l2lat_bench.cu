// L2 round-trip latency of the load flavours usable for cross-SM polling (sm_100a), with and without a
// background bulk-copy stream on every SM.  One measuring thread per CTA chases pointers through a 4 MB
// L2-resident buffer that ANOTHER kernel wrote (so lines are in L2, not in this SM's L1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2lat_bench tools/l2lat_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__device__ __forceinline__ uint32_t ld(const uint32_t* p) {
    uint32_t v;
    if (MODE == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else if (MODE == 1) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else if (MODE == 2) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else if (MODE == 3) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE, int ILP>
__global__ void chase(const uint32_t* buf, int iters, const unsigned char* stream, size_t stream_bytes, int load, unsigned long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* ring = smem + 256;
    __shared__ int stop_s;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        stop_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 32 && load > 0) {  // background streamer: `load` x 16 KB bulk copies in flight
        const size_t per = stream_bytes / gridDim.x / 16384 * 16384;
        const unsigned char* base = stream + (size_t)blockIdx.x * per;
        size_t off = 0;
        uint32_t t = 0, w = 0;
        for (; t < (uint32_t)load; ++t) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[t])), "r"(16384));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + t * 16384)), "l"(base + off), "r"(16384), "r"(smem_u32(&bars[t])) : "memory");
            off = (off + 16384) % per;
        }
        while (!*(volatile int*)&stop_s) {
            const uint32_t slot = w % load, par = (w / load) & 1u;
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[slot])), "r"(par) : "memory");
            if (ok) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[slot])), "r"(16384));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + slot * 16384)), "l"(base + off), "r"(16384), "r"(smem_u32(&bars[slot])) : "memory");
                off = (off + 16384) % per;
                ++w; ++t;
            }
        }
        for (; w < t; ++w) {
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[w % load])), "r"((w / load) & 1u) : "memory");
        }
    }
    if (tid == 0) {
        uint32_t idx = (blockIdx.x * 7919u) & ((1u << 20) - 1);
        for (int i = 0; i < 64; ++i) idx = ld<MODE>(buf + idx);  // warm-up
        const long long t0 = clock64();
        if (ILP == 1) {
            for (int i = 0; i < iters; ++i) idx = ld<MODE>(buf + idx);
        } else {  // four independent loads per step (different lines), the next step depends on all of them
            for (int i = 0; i < iters; ++i) {
                const uint32_t a = ld<MODE>(buf + idx), b = ld<MODE>(buf + ((idx + 4096u) & ((1u << 20) - 1)));
                const uint32_t c = ld<MODE>(buf + ((idx + 8192u) & ((1u << 20) - 1))), d = ld<MODE>(buf + ((idx + 12288u) & ((1u << 20) - 1)));
                idx = (a ^ (b & 32u) ^ (c & 64u) ^ (d & 96u)) & ((1u << 20) - 1) & ~31u;
            }
        }
        const long long t1 = clock64();
        out[blockIdx.x] = (unsigned long long)(t1 - t0) + (idx == 0xffffffffu);
        stop_s = 1;
    }
}

__global__ void fill(uint32_t* buf, uint32_t n) {  // random-ish permutation step: next = (i * 40503 + 12345) mod n, 32-word aligned lines
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = ((i * 40503u + 12345u) * 32u) & (n - 1);
}

template <int MODE, int ILP>
static void run(const char* name, const uint32_t* buf, const unsigned char* stream, size_t sb, int load, unsigned long long* out, int G) {
    const int iters = 2000;
    const size_t smem = 256 + 8 * 16384;
    CK(cudaFuncSetAttribute(chase<MODE, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chase<MODE, ILP><<<G, 64, smem>>>(buf, iters, stream, sb, load, out);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(G);
    CK(cudaMemcpy(h.data(), out, G * 8, cudaMemcpyDeviceToHost));
    double s = 0, mx = 0;
    for (int i = 0; i < G; ++i) { s += h[i]; if (h[i] > mx) mx = h[i]; }
    printf("%-22s ilp=%d load=%d  mean %.0f cycles / step, max-CTA %.0f\n", name, ILP, load, s / G / iters, mx / iters);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    uint32_t* buf;
    const uint32_t n = 1u << 20;  // 4 MB
    CK(cudaMalloc(&buf, n * 4));
    fill<<<256, 256>>>(buf, n);
    unsigned char* stream;
    const size_t sb = (size_t)2 << 30;
    CK(cudaMalloc(&stream, sb));
    CK(cudaMemset(stream, 1, sb));
    unsigned long long* out;
    CK(cudaMalloc(&out, G * 8));
    for (int load : {0, 4}) {
        run<0, 1>("ld.relaxed.gpu", buf, stream, sb, load, out, G);
        run<0, 4>("ld.relaxed.gpu", buf, stream, sb, load, out, G);
        run<1, 1>("ld.global.cg", buf, stream, sb, load, out, G);
        run<1, 4>("ld.global.cg", buf, stream, sb, load, out, G);
        run<2, 4>("ld.volatile", buf, stream, sb, load, out, G);
        run<3, 4>("ld.acquire.gpu", buf, stream, sb, load, out, G);
    }
    return 0;
}
