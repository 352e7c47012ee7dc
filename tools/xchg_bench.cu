// Exchange-protocol microbenchmark for the fused decode kernel (sm_100a).
//
// 148 persistent CTAs repeat: write my slice of an N-float vector -> obtain the whole vector.
// That is the all-to-all step between two GEMV phases of decode_mega.cu; its latency (x 151 per
// token) bounds the decode step.  Protocols:
//   0  LL: {value, tag} words, every consumer thread spins on the words it needs
//   1  LL + __nanosleep(backoff) between polls
//   2  LL data, arrival counter: writers st tagged data + red.relaxed.add (no fence); ONE thread per
//      CTA spins on the counter, bar.sync, then every thread loads its tagged words (re-polls if stale)
//   3  fence barrier: plain stores, __threadfence, red.release.add; one thread spins with ld.acquire;
//      bar.sync; ld.cg data
//   4  LL, only warp 0 polls (whole vector) and fills shared memory; the others wait on bar.sync
// Option --load K: an extra producer thread per CTA keeps K x 32 KB cp.async.bulk copies in flight
// (streams a big buffer) to emulate the weight stream running underneath.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/xchg_bench tools/xchg_bench.cu
//   ./tools/xchg_bench [--n 1024] [--iters 2000] [--load 0|1|2|..]
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e = (x);                                                                    \
        if (e != cudaSuccess) {                                                                 \
            fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

#define NTHREADS 256
#define RING_BYTES (6 * 32768)
#define MAXSLOT 24

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void st_tagged(float* buf, int idx, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(buf + 2 * (size_t)idx), "r"(__float_as_uint(v)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_poll16(const float* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_add(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct Params {
    float* buf[2];        // tagged exchange buffers (2 x N x 2 floats), alternated per iteration
    float* plain[2];      // plain buffers (protocol 3)
    unsigned* counter;    // [iters] zero-initialised arrival counters (protocols 2, 3)
    const float* stream;  // big buffer for the background load
    size_t stream_bytes;
    int N, iters, proto, backoff, load, tile_bytes, reps;
    unsigned long long* out;  // [grid] elapsed ns
    float* sink;
    volatile int* stopflag;
};

__global__ void __launch_bounds__(NTHREADS + 32, 1) xchg_kernel(Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* xs = reinterpret_cast<float*>(smem + 256);                        // [N]
    unsigned char* ring = smem + 256 + 4096 * 4;                             // load ring
    __shared__ int stop_s;
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < MAXSLOT; ++i) mbar_init(&bars[i], 1);
        stop_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= NTHREADS) {
        // background streamer: keep `load` tiles in flight, reissue as each lands
        if (tid == NTHREADS && p.load > 0) {
            const uint32_t TILE_BYTES = (uint32_t)p.tile_bytes;
            const size_t per_cta = p.stream_bytes / G / TILE_BYTES * TILE_BYTES;
            const unsigned char* base = reinterpret_cast<const unsigned char*>(p.stream) + (size_t)cta * per_cta;
            size_t off = 0;
            uint32_t t = 0;
            for (; t < (uint32_t)p.load; ++t) {
                mbar_arrive_expect_tx(&bars[t], TILE_BYTES);
                bulk_g2s(ring + (size_t)t * TILE_BYTES, base + off, TILE_BYTES, &bars[t]);
                off = (off + TILE_BYTES) % per_cta;
            }
            uint32_t w = 0;  // oldest outstanding
            while (!*(volatile int*)&stop_s) {
                const uint32_t slot = w % p.load, par = (w / p.load) & 1u;
                if (mbar_try_wait(&bars[slot], par)) {
                    mbar_arrive_expect_tx(&bars[slot], TILE_BYTES);
                    bulk_g2s(ring + (size_t)slot * TILE_BYTES, base + off, TILE_BYTES, &bars[slot]);
                    off = (off + TILE_BYTES) % per_cta;
                    ++w;
                    ++t;
                }
            }
            for (; w < t; ++w) {
                while (!mbar_try_wait(&bars[w % p.load], (w / p.load) & 1u)) {}
            }
            if (cta == 0) p.out[G] = (unsigned long long)t;  // tiles streamed by CTA 0
        }
        return;
    }
    const int N = p.N;
    // my slice (never empty for N >= G: a CTA that writes nothing is not waited for and could fall
    // two iterations behind, after which its buffer has been overwritten with newer tags)
    const int c0 = (int)((long long)cta * N / G), c1 = (int)((long long)(cta + 1) * N / G);
    float acc = 0.f;
    unsigned long long t0 = 0;
    for (int it = 0; it < p.iters; ++it) {
        if (it == 16) {  // skip warm-up iterations
            asm volatile("bar.sync 1, 256;" ::: "memory");
            t0 = gtime();
        }
        const uint32_t tag = (uint32_t)it + 1u;
        float* buf = p.buf[it & 1];
        const float* rbuf = buf + (size_t)(cta % p.reps) * 2 * N;  // the replica this CTA reads
        // ---- write my slice ----
        if (p.proto == 3) {
            float* pl = p.plain[it & 1];
            if (c0 + tid < c1) pl[c0 + tid] = acc + (float)(c0 + tid + it);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                __threadfence();
                red_release_add(p.counter + it, 1u);
            }
        } else {
            if (c0 + tid < c1) {
                const float v = acc + (float)(c0 + tid + it);
                for (int r = 0; r < p.reps; ++r) st_tagged(buf + (size_t)r * 2 * N, c0 + tid, v, tag);
            }
            if (p.proto == 2) {
                // one arrival per CTA after its stores were issued (same thread order is not required: tags validate)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid == 0) red_relaxed_add(p.counter + it, 1u);
            }
        }
        // ---- obtain the whole vector: thread t needs elements 4t.. (+1024 v) ----
        float sum = 0.f;
        if (p.proto == 0 || p.proto == 1) {
            for (int k = 4 * tid; k < N; k += 4 * NTHREADS) {
                const float* q = rbuf + 2 * (size_t)k;
                uint4 a, b;
                while (true) {
                    a = ld_poll16(q);
                    b = ld_poll16(q + 4);
                    if (a.y == tag && a.w == tag && b.y == tag && b.w == tag) break;
                    if (p.proto == 1) __nanosleep(p.backoff);
                }
                sum += __uint_as_float(a.x) + __uint_as_float(a.z) + __uint_as_float(b.x) + __uint_as_float(b.z);
            }
        } else if (p.proto == 2) {
            if (tid == 0) {
                while (ld_relaxed(p.counter + it) < (unsigned)G) {}
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int k = 4 * tid; k < N; k += 4 * NTHREADS) {
                const float* q = rbuf + 2 * (size_t)k;
                uint4 a, b;
                while (true) {
                    a = ld_poll16(q);
                    b = ld_poll16(q + 4);
                    if (a.y == tag && a.w == tag && b.y == tag && b.w == tag) break;
                }
                sum += __uint_as_float(a.x) + __uint_as_float(a.z) + __uint_as_float(b.x) + __uint_as_float(b.z);
            }
        } else if (p.proto == 3) {
            if (tid == 0) {
                while (ld_acquire(p.counter + it) < (unsigned)G) {}
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float* pl = p.plain[it & 1];
            for (int k = 4 * tid; k < N; k += 4 * NTHREADS) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(pl + k));
                sum += v.x + v.y + v.z + v.w;
            }
        } else {  // 4: warp 0 polls everything into smem
            if (tid < 32) {
                for (int k = 4 * tid; k < N; k += 128) {
                    const float* q = rbuf + 2 * (size_t)k;
                    uint4 a, b;
                    while (true) {
                        a = ld_poll16(q);
                        b = ld_poll16(q + 4);
                        if (a.y == tag && a.w == tag && b.y == tag && b.w == tag) break;
                    }
                    *reinterpret_cast<float4*>(xs + k) =
                        make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int k = 4 * tid; k < N; k += 4 * NTHREADS) {
                const float4 v = *reinterpret_cast<const float4*>(xs + k);
                sum += v.x + v.y + v.z + v.w;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // the real kernel has a block barrier between obtaining the vector and storing the next slice
        // (without it a writer thread could run ahead of its own CTA and overwrite a buffer others still poll)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        acc = sum * 1e-9f;  // dependency: next iteration's values depend on this one's reads
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) {
        p.out[cta] = gtime() - t0;
        stop_s = 1;
    }
    if (acc == 123.456f) p.sink[0] = acc;
}

int main(int argc, char** argv) {
    int N = 1024, iters = 2000, load = 0, only = -1, tile = 32768;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--n")) N = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--iters")) iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--load")) load = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--proto")) only = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--tile")) tile = atoi(argv[++i]);
    }
    if (load > MAXSLOT) load = MAXSLOT;
    if ((size_t)load * tile > RING_BYTES) load = RING_BYTES / tile;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    Params p;
    memset(&p, 0, sizeof p);
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&p.buf[i], (size_t)N * 8 * 8));
        CK(cudaMalloc(&p.plain[i], (size_t)N * 4));
    }
    CK(cudaMalloc(&p.counter, (size_t)iters * 4));
    CK(cudaMalloc(&p.out, (G + 1) * 8));
    CK(cudaMalloc(&p.sink, 4));
    p.stream_bytes = (size_t)2 << 30;
    CK(cudaMalloc((void**)&p.stream, p.stream_bytes));
    CK(cudaMemset((void*)p.stream, 0, p.stream_bytes));
    p.N = N;
    p.iters = iters;
    p.load = load;
    p.tile_bytes = tile;
    const size_t smem = 256 + 4096 * 4 + (size_t)RING_BYTES;
    CK(cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* names[] = {"LL all threads poll", "LL + nanosleep(100)", "LL + arrival counter (1 poller)", "fence barrier (v1)",
                           "LL, warp 0 polls -> smem", "LL + nanosleep(400)", "LL, 2 replicas", "LL, 4 replicas", "LL, 8 replicas",
                           "LL + counter, 4 replicas"};
    printf("grid %d, N = %d floats, %d iterations, background load = %d x %d B tiles in flight per SM\n", G, N, iters, load, tile);
    for (int proto = 0; proto < 10; ++proto) {
        if (only >= 0 && proto != only) continue;
        if (proto == 4) continue;
        p.proto = proto == 5 ? 1 : (proto >= 6 && proto <= 8 ? 0 : (proto == 9 ? 2 : proto));
        p.backoff = proto == 5 ? 400 : 100;
        p.reps = proto == 6 ? 2 : (proto == 7 || proto == 9 ? 4 : (proto == 8 ? 8 : 1));
        for (int i = 0; i < 2; ++i) CK(cudaMemset(p.buf[i], 0, (size_t)N * 8 * 8));
        CK(cudaMemset(p.counter, 0, (size_t)iters * 4));
        CK(cudaMemset(p.out, 0, (G + 1) * 8));
        void* args[] = {&p};
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void*)xchg_kernel, dim3(G), dim3(NTHREADS + 32), args, smem, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<unsigned long long> out(G + 1);
        CK(cudaMemcpy(out.data(), p.out, (G + 1) * 8, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (int i = 0; i < G; ++i) mx = out[i] > mx ? out[i] : mx;
        const double per_iter = (double)mx / (iters - 16);
        const double gbs = load ? (double)out[G] * tile * G / (ms * 1e-3) / 1e9 : 0.0;
        fflush(stdout);
        printf("proto %d  %-34s  %8.1f ns / exchange   (kernel %.3f ms, background stream %.0f GB/s)\n", proto, names[proto], per_iter,
               ms, gbs);
    }
    fflush(stdout);
    return 0;
}
