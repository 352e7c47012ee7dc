// Probe: raw contents of an mbarrier word across arrivals / phase completions (is the phase parity readable by a plain load?)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(unsigned long long* out) {
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) float buf[64];
    int k = 0;
    auto rd = [&]() { unsigned long long w; asm volatile("ld.volatile.shared::cta.b64 %0, [%1];" : "=l"(w) : "r"(s32(&bar)) : "memory"); out[k++] = w; };
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(2));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        rd();                                                                                    // 0 fresh
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
        rd();                                                                                    // 1 one of two arrivals
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
        rd();                                                                                    // 2 phase 0 complete
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
        rd();                                                                                    // 3 phase 1 complete
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(256) : "memory");
        rd();                                                                                    // 4 one arrival + 256 tx pending
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
        rd();                                                                                    // 5 arrivals done, tx pending
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf)), "l"(out + 64), "r"(256), "r"(s32(&bar)) : "memory");
        for (int i = 0; i < 2000; ++i) { unsigned long long w; asm volatile("ld.volatile.shared::cta.b64 %0, [%1];" : "=l"(w) : "r"(s32(&bar)) : "memory"); if (w != out[k - 1]) break; }
        rd();                                                                                    // 6 after the copy completed
    }
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 4096); cudaMemset(d, 0, 4096);
    probe<<<1, 32>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[8]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(e));
    const char* names[] = {"fresh(count 2)", "1 arrival", "phase0 complete", "phase1 complete", "arrive.expect_tx 256", "+arrive (tx pending)", "tx complete"};
    for (int i = 0; i < 7; ++i) printf("%-22s %016llx\n", names[i], h[i]);
    return 0;
}
