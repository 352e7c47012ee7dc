// Device-side building blocks shared by the fused persistent decode kernels (decode_mega.cu: one row;
// decode_batch.cu: up to 8 rows): the shared-memory weight ring and its TMA producer, the tagged exchange
// through L2, hop counters, LayerNorm statistics, the single-query attention item.
// Each including translation unit defines GV_RING_NSLOT (ring depth) and GV_MEGA_NS (a namespace of its own: the
// helpers depend on the ring depth, so the two kernels must not share inline definitions).
#pragma once
#include "common.cuh"
#include "mega.cuh"
#include "sampling.cuh"
#include "stream_layout.h"

#ifndef GV_RING_NSLOT
#error "define GV_RING_NSLOT before including mega_dev.cuh"
#endif
#ifndef GV_MEGA_NS
#error "define GV_MEGA_NS before including mega_dev.cuh"
#endif

namespace gv {
namespace GV_MEGA_NS {

#define MEGA_WARPS 8
#define MEGA_CONSUMERS (MEGA_WARPS * 32)
#define MEGA_THREADS (MEGA_CONSUMERS + 128)  // + one producer warpgroup (one working thread): register reallocation is per warpgroup
#define MEGA_SPIN_LIMIT (1u << 26)
#define NSLOT GV_RING_NSLOT
// Bounded waits.  Production: a wait that exceeds the spin limit faults the kernel (a pipeline bug must not hang the box).
// Debug builds (tools/build_diag.sh, tools/wait_diag.py) record {line, cta, thread, a, b, late} of the waits that time out
// in a pinned HOST buffer bound with genvc_debug_wait_bind() (host memory survives a faulted context):
//   -DGV_WAIT_DIAG=1  short spin limit; the first timeout raises a flag that every other wait polls: each spinning thread records
//                     where it stands exactly once, lingers, and the kernel faults: a deadlock becomes a list of who waited
//                     for what;
//   -DGV_WAIT_DIAG=2  production code in the spin loops (no flag polling): the waits that time out (2^22 spins) record, the first
//                     one lingers 5 s so that every other stuck wait reaches its own limit and records too.
#ifndef GV_WAIT_DIAG
#define GV_WAIT_DIAG 0
#endif
#if GV_WAIT_DIAG
#define GV_DIAG_RECORDS 65536
__device__ unsigned* g_wait_buf;  // [8 + 8 * GV_DIAG_RECORDS] words of pinned host memory: [0] abort flag, [1] record count
#if GV_WAIT_DIAG == 2
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void wait_diag_record(unsigned line, unsigned a, unsigned b, unsigned late) {
    unsigned* buf = g_wait_buf;
    if (buf == nullptr) return;
    const unsigned k = atomicAdd_system(&buf[1], 1u);
    if (k < GV_DIAG_RECORDS) {
        volatile unsigned* r = buf + 8 + 8 * k;
        r[0] = line; r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = a; r[4] = b; r[5] = late;
    }
    __threadfence_system();
    atomicExch_system(&buf[0], 1u);
}
// record where this thread stands, linger so that every other stuck (or flag-polling) thread can record too, then fault
#if GV_WAIT_DIAG == 2
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void wait_fault(unsigned line, unsigned a, unsigned b, unsigned late) {
    wait_diag_record(line, a, b, late);
    for (int k = 0; k < (GV_WAIT_DIAG == 2 ? 5000000 : 200000); ++k) __nanosleep(1000);
    __trap();
}
#if GV_WAIT_DIAG == 1
__device__ __forceinline__ void wait_check(uint32_t& spins, unsigned line, unsigned a, unsigned b) {
    if (((++spins) & 0x3ffu) == 0u) {
        const bool flag = g_wait_buf != nullptr && *(volatile unsigned*)g_wait_buf != 0u;
        if (flag || spins > (1u << 21)) wait_fault(line, a, b, flag ? 1u : 0u);
    }
}
#define GV_SPIN(spins, line, a, b) wait_check(spins, line, a, b)
#else
#define GV_SPIN(spins, line, a, b) if (++spins > (1u << 22)) wait_fault(line, a, b, 0u)
#endif
#else
#define GV_SPIN(spins, line, a, b) if (++spins > MEGA_SPIN_LIMIT) __trap()
#endif
#define UPT GV_MEGA_UPT
// scratch region (never live together): attention per-warp (max, sum) + PV partials (1 KB + 8 * hd floats <= 9 KB) |
// mlp.c_proj group partials [2][D] floats | partial-sum gather [G][8] floats | sampling sort keys [GV_SORT_N] u64 = 16 KB.
// The sort keys run over into the two residual vectors that follow the scratch region (dead while a token is sampled),
// which is what makes room for a twelfth ring slot.
__host__ __device__ inline size_t mega_scratch_bytes(int D) {
    const size_t spill = 2 * (size_t)D * sizeof(float);                    // xres0 + xres1
    const size_t need = (size_t)GV_SORT_N * 8 > spill ? (size_t)GV_SORT_N * 8 - spill : 0;
    return need > 9728 ? need : 9728;
}

enum { TG_XQ = 0, TG_AO = 1, TG_X1 = 2, TG_PP = 3, TG_X2 = 4 };

struct Ring {
    float* slots;
    uint64_t* full;
    uint64_t* empty;
    uint32_t* landed;  // number of leading tiles the producer has SEEN complete (monotonic hint for the consumers)
    int slot_floats;
};
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Tile `idx` is ready: either the producer already saw its full barrier complete (one shared-memory load,
// the common case: tiles are requested microseconds ahead) or this thread waits on the barrier itself
// (mbarrier.try_wait costs several hundred cycles even when the phase is long complete).
// The parity test is only meaningful once the slot's PREVIOUS occupant (tile idx - NSLOT) has landed: a barrier that is still
// two phases behind answers "complete" for the parity of tile idx (phase aliasing).  A warp can be that far ahead of the
// producer since the weight loads of a phase are hoisted above the hop before it (gemv_preload): three phases' worth of
// tiles can be waited on before the other warp group's tiles of the first one have landed.  The producer's `landed`
// counter (monotonic) closes the gap.
__device__ __forceinline__ void tile_ready_wait(const Ring& r, uint32_t idx) {
    uint32_t l = ld_acquire_cta_shared(r.landed);
    if (l > idx) return;
    uint32_t spins = 0;
    while (l + NSLOT <= idx) {
        GV_SPIN(spins, __LINE__, idx, l);
        l = ld_acquire_cta_shared(r.landed);
    }
    spins = 0;
    while (!mbar_try_wait(&r.full[idx % NSLOT], (idx / NSLOT) & 1u)) GV_SPIN(spins, __LINE__, idx, ld_acquire_cta_shared(r.landed));
}

// Warp-uniform wait (GV_UNIFORM_TILE_WAIT, single-row kernel): EVERY lane polls the producer's `landed` counter (one
// broadcast shared-memory load per round) until tile `last_idx` -- and with it every earlier tile -- is in; the loop exit
// is the same for all lanes, so the warp never splits on a weight wait.  (Lane-divergent mbarrier polling ahead of a block
// barrier was the other ingredient of the barrier-slip deadlock described at ld_tagged_vec_u.)
__device__ __forceinline__ void tile_ready_wait_u(const Ring& r, uint32_t last_idx) {
    uint32_t spins = 0;
    uint32_t l = ld_acquire_cta_shared(r.landed);
    while (l <= last_idx) {
        GV_SPIN(spins, __LINE__, last_idx, l);
        l = ld_acquire_cta_shared(r.landed);
    }
}
#ifndef GV_UNIFORM_TILE_WAIT
#define GV_UNIFORM_TILE_WAIT 0
#endif

// ---------------------------------------------------------------------------------------------
// tagged exchange through L2: element i of a buffer lives at floats [2i, 2i+1] = {value, tag}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_tagged(float* buf, int idx, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(buf + 2 * (size_t)idx), "r"(__float_as_uint(v)),
                 "r"(tag)
                 : "memory");
}
// elements idx, idx+1 (idx even) in one 16-byte store
__device__ __forceinline__ void st_tagged2(float* buf, int idx, float v0, float v1, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(buf + 2 * (size_t)idx),
                 "r"(__float_as_uint(v0)), "r"(tag), "r"(__float_as_uint(v1)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_x16(const float* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ uint2 ld_x8(const float* p) {
    uint2 r;
    asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ bool tags_ok(const uint4& a, uint32_t tag, uint32_t tmask) {
    return (((a.y ^ tag) | (a.w ^ tag)) & tmask) == 0u;
}
// elements idx, idx+1 (idx even): spin until both carry `tag`
__device__ __forceinline__ float2 ld_tagged2(const float* buf, int idx, uint32_t tag, uint32_t tmask) {
    const float* p = buf + 2 * (size_t)idx;
    uint4 a = ld_x16(p);
    uint32_t spins = 0;
    while (!tags_ok(a, tag, tmask)) {
        GV_SPIN(spins, __LINE__, (unsigned)idx, tag - a.y);
        a = ld_x16(p);
    }
    return make_float2(__uint_as_float(a.x), __uint_as_float(a.z));
}
__device__ __forceinline__ float ld_tagged1(const float* buf, int idx, uint32_t tag, uint32_t tmask) {
    const float* p = buf + 2 * (size_t)idx;
    uint2 a = ld_x8(p);
    uint32_t spins = 0;
    while (((a.y ^ tag) & tmask) != 0u) {
        GV_SPIN(spins, __LINE__, (unsigned)idx, tag - a.y);
        a = ld_x8(p);
    }
    return __uint_as_float(a.x);
}
// NE (1, 2 or 4) consecutive elements starting at idx (idx % NE == 0)
template <int NE>
__device__ __forceinline__ void ld_tagged_vec(const float* buf, int idx, uint32_t tag, uint32_t tmask, float* out) {
    if constexpr (NE == 1) {
        out[0] = ld_tagged1(buf, idx, tag, tmask);
    } else if constexpr (NE == 2) {
        const float2 v = ld_tagged2(buf, idx, tag, tmask);
        out[0] = v.x;
        out[1] = v.y;
    } else {
        static_assert(NE == 4, "NE must be 1, 2 or 4");
        const float* p = buf + 2 * (size_t)idx;
        uint4 a = ld_x16(p), b = ld_x16(p + 4);
        uint32_t spins = 0;
        while (!(tags_ok(a, tag, tmask) && tags_ok(b, tag, tmask))) {
            GV_SPIN(spins, __LINE__, (unsigned)idx, tag - a.y);
            a = ld_x16(p);
            b = ld_x16(p + 4);
        }
        out[0] = __uint_as_float(a.x);
        out[1] = __uint_as_float(a.z);
        out[2] = __uint_as_float(b.x);
        out[3] = __uint_as_float(b.z);
    }
}

// Warp-uniform variants (single-row kernel): every lane of the warp calls them together (`valid` = this lane has an element)
// and the warp leaves the poll loop as one -- lanes whose words are already there keep re-reading them until the slowest
// lane's words arrive.  A per-lane exit leaves the warp split until the compiler's reconvergence point; on this toolchain a
// block barrier further down was then seen releasing seven warps without the eighth (a warp counted twice), which deadlocks
// the CTA.  The vote costs a few cycles per poll round; the extra polls hit lines that are complete.
template <int NE>
__device__ __forceinline__ void ld_tagged_vec_u(const float* buf, int idx, bool valid, uint32_t tag, uint32_t tmask, float* out) {
    static_assert(NE == 2 || NE == 4, "NE must be 2 or 4");
    const float* p = buf + 2 * (size_t)(valid ? idx : 0);
    uint4 a = make_uint4(0u, tag, 0u, tag), b = a;
    uint32_t spins = 0;
    for (;;) {
        bool ok = true;
        if (valid) {
            a = ld_x16(p);
            if constexpr (NE == 4) b = ld_x16(p + 4);
            ok = tags_ok(a, tag, tmask) && tags_ok(b, tag, tmask);
        }
        if (__all_sync(0xffffffffu, ok)) break;
        GV_SPIN(spins, __LINE__, (unsigned)idx, tag - a.y);
    }
    out[0] = __uint_as_float(a.x);
    out[1] = __uint_as_float(a.z);
    if constexpr (NE == 4) {
        out[2] = __uint_as_float(b.x);
        out[3] = __uint_as_float(b.z);
    }
}

// ---------------------------------------------------------------------------------------------
// self-counting fixed-point all-reduce (mlp.c_proj partial sums).  Every CTA adds, per output, the 64-bit word
//   (round(partial * 2^34) << 8) + 1
// into one accumulator with a relaxed L2 reduction.  Integer addition is associative, so the sum does not depend
// on arrival order (bit-reproducible, unlike float atomics) and is exact to 2^-35 per term; the low byte counts
// contributors, so a reader knows from the word itself when all G (< 256) CTAs have arrived — no fence, no
// separate reducer CTAs, one hop instead of three.
// ---------------------------------------------------------------------------------------------
#ifndef GV_PP_COUNTER
#define GV_PP_COUNTER 1  /* 0 = reducer CTAs poll the partial-sum tags directly: measured slower (0.643 vs 0.631 ms/token) */
#endif
#ifndef GV_ATOMIC_RED
#define GV_ATOMIC_RED 0  /* measured on B200: 148 x 1024 u64 reductions onto 8 KB cost ~6 us per layer (L2 serialises per line); the reducer-CTA path below is faster */
#endif
#define GV_FIX_SCALE 17179869184.0f              /* 2^34 */
#define GV_FIX_INV (1.0 / 17179869184.0)
__device__ __forceinline__ void red_fix_add(unsigned long long* acc, float v) {
    const long long q = __float2ll_rn(v * GV_FIX_SCALE);
    const unsigned long long w = ((unsigned long long)q << 8) + 1ull;
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc), "l"(w) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_x2u64(const unsigned long long* p) {
    ulonglong2 r;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float fix_value(unsigned long long w) {
    return (float)((double)((long long)w >> 8) * GV_FIX_INV);
}

// Block barrier of the consumer warps.  -DGV_PROG (debug): every warp counts its arrivals in shared memory and checks after
// the barrier that all eight warps have arrived at least as often -- a barrier that releases without one of them (a warp
// counted twice, e.g. arriving in two halves) is recorded with its source line in g_slip.
#ifdef GV_PROG
__device__ unsigned g_slip[8 + 8 * 64];
__device__ __forceinline__ volatile unsigned* bar_cnt() {
    __shared__ unsigned s_cnt[8];
    return s_cnt;
}
__device__ __forceinline__ void bar_sync_chk(unsigned site) {
    volatile unsigned* c = bar_cnt();
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    unsigned my = 0;
    if (l == 0) {
        my = c[w] + 1u;
        c[w] = my;
    }
    bar_sync(1, MEGA_CONSUMERS);
    if (l == 0) {
        for (int q = 0; q < 8; ++q) {
            const unsigned o = c[q];
            if ((int)(o - my) < 0) {
                const unsigned k = atomicAdd(&g_slip[0], 1u);
                if (k < 64) {
                    unsigned* r = g_slip + 8 + 8 * k;
                    r[0] = site; r[1] = blockIdx.x; r[2] = (unsigned)w; r[3] = (unsigned)q; r[4] = my; r[5] = o;
                }
            }
        }
    }
}
#define BAR1() bar_sync_chk(__LINE__)
#else
#define BAR1() bar_sync(1, MEGA_CONSUMERS)
#endif

// ---------------------------------------------------------------------------------------------
// hops: arrival counter (a hint: one poller per CTA) + block barrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_add(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// One arrival per CTA, by thread 0 as soon as ITS warp's exchange stores are issued: the counter is only a hint (the
// tags validate the data), so it may run ahead of the other warps' stores by their skew -- no block barrier here.
// Every hop_arrive is followed by a hop_wait (with its barrier) before shared memory is reused.
__device__ __forceinline__ void hop_arrive(unsigned* cnt, int tid) {
    if (tid == 0) red_relaxed_add(cnt, 1u);
}
// `near`: the CTA is released when all but `near` arrivals are in; its threads then spin on the tagged words themselves
// for the tail, so the data loads overlap the last arrivals instead of following the counter by an L2 round trip
// (measured: 4 of 148 is the sweet spot, 0.442 -> 0.427 ms/token; 8 and more lose again -- many pollers on lines that
// are still being written slow the writers).  The tags, not the counter, validate the data: any `near` is correct.
// `settle_ns`: the arrival counter is bumped without a fence, so it can overtake the last data stores on their way to
// L2; a first data load that finds a stale tag costs a whole extra round trip, a short pause after the counter
// reaches its target is cheaper.
__device__ __forceinline__ void hop_wait(const unsigned* cnt, unsigned target, int tid, uint32_t tmask, unsigned settle_ns,
                                         volatile int* hold = nullptr, unsigned near = 0u) {
    if (tid == 0 && tmask != 0u) {
        if (hold) *hold = 1;  // the weight producer stops issuing bulk copies: they slow this SM's ordinary loads down
        uint32_t spins = 0;
        while (ld_relaxed_u32(cnt) + near < target) {
            GV_SPIN(spins, __LINE__, target, ld_relaxed_u32(cnt));
        }
        if (settle_ns) __nanosleep(settle_ns);
    }
    BAR1();
}

// ---------------------------------------------------------------------------------------------
// weight ring (consumer side).  Tile t of a phase holds units 4t .. 4t+3 and is read by exactly the
// four warps of group t % 2 (warp = 4 * group + unit % 4); each of them arrives once on the tile's
// empty barrier (count = 4).  Tiles are addressed by their global index (ring slot = index % NSLOT).
// ---------------------------------------------------------------------------------------------
struct Cons {
    uint32_t gt;               // global index of the first tile of the current phase (same in every thread)
    unsigned long long* wacc;  // debug: accumulates ns spent waiting for weight tiles (null = off)
};
__device__ __forceinline__ const float* slot_ptr(const Ring& r, uint32_t idx) {
    return r.slots + (size_t)(idx % NSLOT) * r.slot_floats;
}
// Wait for all tiles t = g, g + 2, g + 4, g + 6 (< ntiles) of this warp's group at once: lane k waits
// for tile g + 2k (32 lanes polling one mbarrier would serialise), __syncwarp orders the rest.
__device__ __forceinline__ void group_wait(const Ring& r, const Cons& cs, int g, int ntiles, int lane) {
    long long t0 = 0;
    if (cs.wacc != nullptr && lane == 0) t0 = clock64();  // debug timeline
#if GV_UNIFORM_TILE_WAIT
    if (g < ntiles) tile_ready_wait_u(r, cs.gt + (uint32_t)(g + 2 * ((ntiles - 1 - g) / 2)));  // the group's last tile
#else
    if (lane < 4) {
        const int t = g + 2 * lane;
        if (t < ntiles) tile_ready_wait(r, cs.gt + (uint32_t)t);
    }
#endif
    __syncwarp();
    if (cs.wacc != nullptr && lane == 0) *cs.wacc += (unsigned long long)(clock64() - t0);
}
// Release them (this warp's arrival on each tile's empty barrier) once the warp has read its units.
__device__ __forceinline__ void group_release(const Ring& r, const Cons& cs, int g, int ntiles, int lane) {
    __syncwarp();
    if (lane < 4) {
        const int t = g + 2 * lane;
        if (t < ntiles) mbar_arrive(&r.empty[(cs.gt + (uint32_t)t) % NSLOT]);
    }
}
// single tile read by all warps (logits-head LayerNorm parameters)
__device__ __forceinline__ const float* tile_wait(const Ring& r, const Cons& cs, uint32_t idx, int lane) {
#if GV_UNIFORM_TILE_WAIT
    if (cs.wacc != nullptr) {  // debug timeline
        const long long t0 = clock64();
        tile_ready_wait_u(r, idx);
        if (lane == 0) *cs.wacc += (unsigned long long)(clock64() - t0);
    } else {
        tile_ready_wait_u(r, idx);
    }
    __syncwarp();
    return slot_ptr(r, idx);
#endif
    if (lane == 0) {
        if (cs.wacc != nullptr) {  // debug timeline
            const long long t0 = clock64();
            tile_ready_wait(r, idx);
            *cs.wacc += (unsigned long long)(clock64() - t0);
        } else {
            tile_ready_wait(r, idx);
        }
    }
    __syncwarp();
    return slot_ptr(r, idx);
}
__device__ __forceinline__ void tile_release(const Ring& r, uint32_t idx, int lane, uint32_t count = 1u) {
    __syncwarp();
    if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&r.empty[idx % NSLOT])), "r"(count) : "memory");
}

// sum of the per-warp partials red[0..7] (written before a block barrier)
__device__ __forceinline__ float sum8(const float* red) {
    const float4 a = *reinterpret_cast<const float4*>(red), b = *reinterpret_cast<const float4*>(red + 4);
    return ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
}

// ---------------------------------------------------------------------------------------------
// LayerNorm statistics of the vector held four elements per thread (thread t owns x[4t .. 4t+3]).
// One pass over data shifted by `shift` (any value near the mean keeps E[d^2] - E[d]^2 free of
// cancellation; the callers pass the mean this LayerNorm saw one layer earlier):
//   stats_partial : per-warp sums of d and d^2 -> red[0..7], red[8..15]   (then ONE block barrier)
//   stats_finish  : mean, rstd from the 16 partials (every thread, after the barrier)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stats_partial(const float4& x, bool valid, float shift, float* red, int lane, int warp) {
    float s1 = 0.0f, s2 = 0.0f;
    if (valid) {
        const float d0 = x.x - shift, d1 = x.y - shift, d2 = x.z - shift, d3 = x.w - shift;
        s1 = (d0 + d1) + (d2 + d3);
        s2 = fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
        red[warp] = s1;
        red[8 + warp] = s2;
    }
}
__device__ __forceinline__ void stats_finish(const float* red, float inv_d, float shift, float& mean, float& rstd) {
    const float m = sum8(red) * inv_d;
    float var = fmaf(-m, m, sum8(red + 8) * inv_d);
    var = fmaxf(var, 0.0f);
    mean = shift + m;
    rstd = rsqrtf(var + 1e-5f);
    rstd = rstd * fmaf(-0.5f * (var + 1e-5f) * rstd, rstd, 1.5f);  // one Newton step: full fp32 accuracy
}
// explicit LayerNorm (logits head only: its output is the latent handed to the vocoder); two block barriers
__device__ __forceinline__ void ln_quad(float4& x, bool valid, int D, const float* w, const float* b, float* red, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    float s = valid ? ((x.x + x.y) + (x.z + x.w)) : 0.0f;
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    BAR1();
    const float mean = sum8(red) / (float)D;
    const float d0 = x.x - mean, d1 = x.y - mean, d2 = x.z - mean, d3 = x.w - mean;
    float q = valid ? (fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3)) : 0.0f;
    q = warp_sum(q);
    if (lane == 0) red[8 + warp] = q;
    BAR1();
    const float var = sum8(red + 8) / (float)D;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    if (valid) {
        const float4 ww = *reinterpret_cast<const float4*>(w + 4 * tid);
        const float4 bb = *reinterpret_cast<const float4*>(b + 4 * tid);
        x.x = d0 * rstd * ww.x + bb.x;
        x.y = d1 * rstd * ww.y + bb.y;
        x.z = d2 * rstd * ww.z + bb.z;
        x.w = d3 * rstd * ww.w + bb.w;
    }
}

// ---------------------------------------------------------------------------------------------
// GEMV, K = D, one warp per unit: unit u of the phase is handled by warp u % 8 (u < nunits <= 32,
// up to four units per warp).  `xs` is the activation vector in shared memory (complete: the
// caller synchronised).  For every unit, lane 8q (q = u / 8) of the owning warp calls
// epi(u, dot, c2, c1) with  dot = sum_k x_k W'[k][unit]  and the unit's two epilogue constants
// (pack_stream_kernel).
// ---------------------------------------------------------------------------------------------
// EARLY: release every tile right after its unit (the producer refills while this phase goes on: needed when another
// GEMV follows without a hop, i.e. FC -> P2); otherwise one release for all of the warp's tiles at the end (each
// __syncwarp + elected mbarrier.arrive costs ~250 cycles of this warp's chain).
// Registers: at 168 per thread (the launch allocation of a 12-warp CTA) ptxas sinks every LDS.128 next to its FFMAs (one
// or two shared-memory loads in flight per warp).  The CTA therefore launches with a 4-warp producer warpgroup that
// gives its registers away (setmaxnreg.dec 40) and the two consumer warpgroups take 232 each (setmaxnreg.inc): ptxas
// then batches 8 LDS.128 per warp and every phase of the layer gets ~40 % shorter (0.60 -> 0.45 ms/token).
template <int NXV, bool EARLY, class Epi>
__device__ __forceinline__ void gemv_dot(const Ring& ring, const Cons& cs, int nunits, const float* xs, int warp, int lane,
                                         Epi epi) {
    constexpr int D = NXV * 128, UF = D + 4;
    float4 xv[NXV];
#pragma unroll
    for (int i = 0; i < NXV; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (i * 32 + lane) * 4);
    float tot[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float2 cc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cc[k] = make_float2(0.f, 0.f);
    const int ntiles = (nunits + UPT - 1) / UPT;
    const int g = warp >> 2, r = warp & 3;
    group_wait(ring, cs, g, ntiles, lane);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int t = g + 2 * k;
        if (t * UPT + r < nunits) {
            const float* col = slot_ptr(ring, cs.gt + (uint32_t)t) + r * UF;
            cc[k] = *reinterpret_cast<const float2*>(col + D);
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
            for (int i = 0; i < NXV; ++i) {
                const float4 wv = *reinterpret_cast<const float4*>(col + (i * 32 + lane) * 4);
                a0 = fmaf(wv.x, xv[i].x, a0);
                a1 = fmaf(wv.y, xv[i].y, a1);
                a2 = fmaf(wv.z, xv[i].z, a2);
                a3 = fmaf(wv.w, xv[i].w, a3);
            }
            tot[k] = (a0 + a1) + (a2 + a3);
        }
        if constexpr (EARLY) {
            if (t < ntiles) tile_release(ring, cs.gt + (uint32_t)t, lane);
        }
    }
    if constexpr (!EARLY) group_release(ring, cs, g, ntiles, lane);
    // transposing butterfly: lanes [8q, 8q+8) end up reducing tot[q]
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    const float k0 = up16 ? tot[2] : tot[0], s0 = up16 ? tot[0] : tot[2];
    const float k1 = up16 ? tot[3] : tot[1], s1 = up16 ? tot[1] : tot[3];
    const float h0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);  // tot[0] (lanes < 16) / tot[2]
    const float h1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);  // tot[1] (lanes < 16) / tot[3]
    const float keep = up8 ? h1 : h0, send = up8 ? h0 : h1;
    float v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    const int q = lane >> 3;
    const int u = warp + 8 * q;
    const float2 c = up16 ? (up8 ? cc[3] : cc[2]) : (up8 ? cc[1] : cc[0]);
    if ((lane & 7) == 0 && u < nunits) epi(u, v, c.x, c.y);
}

// ---------------------------------------------------------------------------------------------
// The same GEMV in two halves, so that the half that does not need the activations runs BEFORE the hop that delivers them:
//   gemv_preload : wait for the phase's tiles, copy this warp's (up to four) units from the ring into registers
//                  (4 x D / 32 weights per lane), release the tiles -- the producer refills them while the CTA waits
//   gemv_finish  : activation vector -> registers, 32 FFMA per lane and unit, the transposing shuffle tree, epilogue
// A phase then costs the FFMAs and the shuffle tree after its hop instead of the shared-memory round trips as well.
// ---------------------------------------------------------------------------------------------
// KU = units a warp can own in the phase (compile-time bound: units per CTA <= 8 KU), so that only that many register
// sets exist: a phase with one unit per warp must not carry 128 registers across a function call.
template <int NXV, int KU>
struct GemvRegs {
    float4 w[KU][NXV];
    float2 cc[KU];
};
template <int NXV, int KU>
__device__ __forceinline__ void gemv_preload(const Ring& ring, const Cons& cs, int nunits, int warp, int lane, GemvRegs<NXV, KU>& W) {
    constexpr int D = NXV * 128, UF = D + 4;
    const int ntiles = (nunits + UPT - 1) / UPT;
    const int g = warp >> 2, r = warp & 3;
    group_wait(ring, cs, g, ntiles, lane);
#pragma unroll
    for (int k = 0; k < KU; ++k) {
        const int t = g + 2 * k;
        W.cc[k] = make_float2(0.f, 0.f);
        if (t * UPT + r < nunits) {
            const float* col = slot_ptr(ring, cs.gt + (uint32_t)t) + r * UF;
            W.cc[k] = *reinterpret_cast<const float2*>(col + D);
#pragma unroll
            for (int i = 0; i < NXV; ++i) W.w[k][i] = *reinterpret_cast<const float4*>(col + (i * 32 + lane) * 4);
        } else {
#pragma unroll
            for (int i = 0; i < NXV; ++i) W.w[k][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    group_release(ring, cs, g, ntiles, lane);
}
template <int NXV, int KU, class Epi>
__device__ __forceinline__ void gemv_finish(int nunits, const float* xs, int warp, int lane, const GemvRegs<NXV, KU>& W, Epi epi) {
    float4 xv[NXV];
#pragma unroll
    for (int i = 0; i < NXV; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (i * 32 + lane) * 4);
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    float2 cc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cc[k] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < KU; ++k) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for (int i = 0; i < NXV; ++i) {
            a0 = fmaf(W.w[k][i].x, xv[i].x, a0);
            a1 = fmaf(W.w[k][i].y, xv[i].y, a1);
            a2 = fmaf(W.w[k][i].z, xv[i].z, a2);
            a3 = fmaf(W.w[k][i].w, xv[i].w, a3);
        }
        tot[k] = (a0 + a1) + (a2 + a3);
        cc[k] = W.cc[k];
    }
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    const float k0 = up16 ? tot[2] : tot[0], s0 = up16 ? tot[0] : tot[2];
    const float k1 = up16 ? tot[3] : tot[1], s1 = up16 ? tot[1] : tot[3];
    const float h0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
    const float h1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
    const float keep = up8 ? h1 : h0, send = up8 ? h0 : h1;
    float v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    const int q = lane >> 3;
    const int u = warp + 8 * q;
    const float2 c = up16 ? (up8 ? cc[3] : cc[2]) : (up8 ? cc[1] : cc[0]);
    if ((lane & 7) == 0 && u < nunits) epi(u, v, c.x, c.y);
}
// register sets per warp and phase when the grid has at least 128 CTAs (launch_decode_mega checks): ceil(N / 128 / 8)
template <int NXV> struct GemvKU {
    static constexpr int QKV = (3 * NXV + 7) / 8, PROJ = (NXV + 7) / 8, FC = (4 * NXV + 7) / 8, HEAD = 2;
};

// mlp.c_proj split along K: unit k = row of W_proj2 owned by this CTA.  Warp group g takes tiles
// t = g, g + 2, ...; thread j of the group (0..127) owns outputs [4j, 4j+4) and [D/2 + 4j, D/2 + 4j + 4) and
// accumulates u_k * W[k][.] over the group's rows; the two group partials land in part[g][D].
template <int NXV>
__device__ __forceinline__ void gemv_outer(const Ring& ring, const Cons& cs, int nunits, const float* us, int tid, int lane,
                                           int warp, float* part) {
    constexpr int D = NXV * 128, UF = D + 4, HALF = D / 2;
    const int g = warp >> 2, j = tid & 127;
    const bool valid = 4 * j < HALF;
    const int ntiles = (nunits + UPT - 1) / UPT;
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int t = g + 2 * k;
        if (t < ntiles) {
            const float* base = tile_wait(ring, cs, cs.gt + (uint32_t)t, lane) + 4 * j;
#pragma unroll
            for (int r = 0; r < UPT; ++r) {
                const int kk = t * UPT + r;
                if (kk < nunits && valid) {
                    const float uk = us[kk];
                    const float4 w0 = *reinterpret_cast<const float4*>(base + r * UF);
                    const float4 w1 = *reinterpret_cast<const float4*>(base + r * UF + HALF);
                    acc0.x = fmaf(uk, w0.x, acc0.x); acc0.y = fmaf(uk, w0.y, acc0.y);
                    acc0.z = fmaf(uk, w0.z, acc0.z); acc0.w = fmaf(uk, w0.w, acc0.w);
                    acc1.x = fmaf(uk, w1.x, acc1.x); acc1.y = fmaf(uk, w1.y, acc1.y);
                    acc1.z = fmaf(uk, w1.z, acc1.z); acc1.w = fmaf(uk, w1.w, acc1.w);
                }
            }
        }
        if (t < ntiles) tile_release(ring, cs.gt + (uint32_t)t, lane);
    }
    if (valid) {
        *reinterpret_cast<float4*>(part + g * D + 4 * j) = acc0;
        *reinterpret_cast<float4*>(part + g * D + HALF + 4 * j) = acc1;
    }
}

// ---------------------------------------------------------------------------------------------
// single-query attention over one (head, key range) item, 8 warps, built for a short dependency chain after q
// arrives.  Arithmetic of HF GPT2Attention._attn for q_len == 1:
//   s_j = (q . k_j) / sqrt(hd);  p = softmax_j(s);  o = sum_j p_j v_j       (un-normalised here)
//   * keys are cut into batches of 32: warp w owns rows 4w .. 4w+3 of every batch and keeps its own running
//     (max, sum, o[hd]) -- flash-decoding inside the CTA: no score buffer, no barrier before the softmax;
//   * before q exists: K and V rows of batch 0 are requested from the cache into registers;
//   * then the CTA polls the tagged q words (only the <= H * 8 attention CTAs read xq, so they poll the data directly:
//     one L2 round trip less than counter-then-data; warp w polls an eighth of q and the CTA assembles it in shared
//     memory: -0.6 % ms/token against every warp polling all of q); the warp that owns the position being decoded also
//     polls k / v of this very step from the exchange buffer and appends them to the cache; that position rides along
//     as a fifth row of its batch (its cache row is loaded as zeros and masked);
//   * per batch: 4 (+1) dots and one interleaved shuffle tree, online-softmax update, o += p v; K / V rows of the next
//     batch are requested as soon as the registers of the current one are free;
//   * the 8 per-warp partials meet in shared memory behind ONE barrier and are merged by thread d < hd.
// ---------------------------------------------------------------------------------------------
#define ATT_ROWS 4        // K rows per warp per batch (32 keys / 8 warps)
template <int HD>
struct AttLane {
    static constexpr int VEC = (HD >= 128) ? 4 : (HD / 32);  // floats per lane per chunk
    static constexpr int NCH = HD / (32 * VEC);              // chunks per lane
    static constexpr int DPL = VEC * NCH;                    // dims per lane
};
template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float* r) {
    if constexpr (VEC == 4) {
        const float4 v = ldcg4(p);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else if constexpr (VEC == 2) {
        const float2 v = ldcg2(p);
        r[0] = v.x; r[1] = v.y;
    } else {
        r[0] = ldcg(p);
    }
}
template <int VEC>
__device__ __forceinline__ void st_vec(float* p, const float* r) {
    if constexpr (VEC == 4) __stcg(reinterpret_cast<float4*>(p), make_float4(r[0], r[1], r[2], r[3]));
    else if constexpr (VEC == 2) __stcg(reinterpret_cast<float2*>(p), make_float2(r[0], r[1]));
    else __stcg(p, r[0]);
}
// DPL tagged elements of one lane (element e at base[2e] = {value, tag}); false while any tag is stale
template <int VEC, int NCH>
__device__ __forceinline__ bool ld_tagged_lane(const float* base, int lane, uint32_t tag, uint32_t tmask, float* r) {
    bool ok = true;
    if constexpr (VEC == 1) {
        const uint2 a = ld_x8(base + 2 * lane);
        ok = ((a.y ^ tag) & tmask) == 0u;
        r[0] = __uint_as_float(a.x);
    } else {
        uint4 a[NCH * (VEC / 2)];
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int w2 = 0; w2 < VEC / 2; ++w2) a[c * (VEC / 2) + w2] = ld_x16(base + 2 * ((c * 32 + lane) * VEC + 2 * w2));
#pragma unroll
        for (int w2 = 0; w2 < NCH * (VEC / 2); ++w2) {
            ok = ok && tags_ok(a[w2], tag, tmask);
            r[2 * w2] = __uint_as_float(a[w2].x);
            r[2 * w2 + 1] = __uint_as_float(a[w2].z);
        }
    }
    return ok;
}

// NORM: the item covers all keys of its head (no split): it writes the normalised output, no (max, sum) pair.
template <int HD, bool DBG, bool NORM = false>
__device__ __noinline__ void att_item(float* __restrict__ Kc, float* __restrict__ Vc, const float* xq, int D, int h, int j0, int j1, int S,
                         uint32_t tag_in, float* wml, float* wpart, int tid,
                         float* o_out, float* ml_out, int item, uint32_t tag_out, uint32_t tmask, unsigned long long* dbg) {
    using L = AttLane<HD>;
    long long ck[8];
    if constexpr (DBG) ck[0] = clock64();
    constexpr int VEC = L::VEC, NCH = L::NCH, DPL = L::DPL;
    const int warp = tid >> 5, lane = tid & 31;
    const int nk = j1 - j0;
    const int nb = (nk + 31) >> 5;
    const int jn = S - 1 - j0;  // relative index of the position being decoded (inside this item iff 0 <= jn < nk)
    const bool mine = jn >= 0 && jn < nk && ((jn & 31) >> 2) == warp;  // this warp owns the new position
    const int bn = jn >> 5;
    const float sqrt_hd = sqrtf((float)HD);

    auto load_row = [&](const float* base, int jr, float* dst) {  // cache row of relative key jr (zeros outside / new key)
        if (jr < nk && jr != jn) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) ld_vec<VEC>(base + (size_t)(j0 + jr) * HD + (c * 32 + lane) * VEC, dst + c * VEC);
        } else {
#pragma unroll
            for (int i = 0; i < DPL; ++i) dst[i] = 0.0f;
        }
    };
    // ---- batch 0 from the cache (independent of this step's q) ----
    float kr[ATT_ROWS][DPL], vr[ATT_ROWS][DPL];
#pragma unroll
    for (int u = 0; u < ATT_ROWS; ++u) load_row(Kc, warp * ATT_ROWS + u, kr[u]);
#pragma unroll
    for (int u = 0; u < ATT_ROWS; ++u) load_row(Vc, warp * ATT_ROWS + u, vr[u]);
    if constexpr (DBG) ck[1] = ck[2] = ck[7] = clock64();
    // ---- this step's q (and k / v of the position being decoded) ----
    float qr[DPL], knew[DPL], vnew[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) knew[i] = vnew[i] = 0.0f;
    {
        const float* qp = xq + 2 * (size_t)(h * HD);
        const float* kp = xq + 2 * (size_t)(D + h * HD);
        const float* vp = xq + 2 * (size_t)(2 * D + h * HD);
        // q is the same for every warp: warp w polls only elements [w * hd/8, (w+1) * hd/8) and the CTA assembles q in
        // shared memory (an eighth of the polling traffic on lines the QKV epilogues are still writing)
        constexpr int EPW = HD / MEGA_WARPS;
        float qmine = 0.0f;
        uint32_t spins = 0;
        bool ok = false;
        while (!ok) {
            ok = true;
            [[maybe_unused]] unsigned seen_tag = tag_in;
            if (lane < EPW) {
                const uint2 a = ld_x8(qp + 2 * (warp * EPW + lane));
                ok = ((a.y ^ tag_in) & tmask) == 0u;
                qmine = __uint_as_float(a.x);
                seen_tag = a.y;
            }
            [[maybe_unused]] unsigned bad = ok ? 0u : 1u;
            if (mine) {
                const bool ok_k = ld_tagged_lane<VEC, NCH>(kp, lane, tag_in, tmask, knew);
                const bool ok_v = ld_tagged_lane<VEC, NCH>(vp, lane, tag_in, tmask, vnew);
                ok = ok && ok_k && ok_v;
                bad |= (ok_k ? 0u : 2u) | (ok_v ? 0u : 4u);
            }
            ok = __all_sync(0xffffffffu, ok);
            GV_SPIN(spins, __LINE__, (unsigned)(h << 8) | bad, tag_in - seen_tag);
        }
        float* qs = wml + 16;
        if (lane < EPW) qs[warp * EPW + lane] = qmine;
        BAR1();
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < VEC; ++i) qr[c * VEC + i] = qs[(c * 32 + lane) * VEC + i];
    }
    if constexpr (DBG) ck[3] = clock64() + (long long)(qr[0] == 123.f);
    if (mine) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            st_vec<VEC>(Kc + (size_t)(S - 1) * HD + (c * 32 + lane) * VEC, knew + c * VEC);
            st_vec<VEC>(Vc + (size_t)(S - 1) * HD + (c * 32 + lane) * VEC, vnew + c * VEC);
        }
    }
    // ---- batches: scores, online softmax, PV (per warp) ----
    float m = -INFINITY, lsum = 0.0f, o[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[i] = 0.0f;
    for (int b = 0; b < nb; ++b) {
        const int r0 = b * 32 + warp * ATT_ROWS;
        const bool with_new = mine && b == bn;  // warp-uniform
        float sc[ATT_ROWS + 1];
#pragma unroll
        for (int u = 0; u < ATT_ROWS; ++u) {
            float a = 0.0f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) a = fmaf(qr[i], kr[u][i], a);
            sc[u] = a;
        }
        {
            float a = 0.0f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) a = fmaf(qr[i], knew[i], a);
            sc[ATT_ROWS] = a;
        }
        if (b + 1 < nb) {  // K registers are free: request the next batch
#pragma unroll
            for (int u = 0; u < ATT_ROWS; ++u) load_row(Kc, r0 + 32 + u, kr[u]);
        }
#pragma unroll
        for (int x = 16; x > 0; x >>= 1) {
#pragma unroll
            for (int u = 0; u <= ATT_ROWS; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], x);
        }
        if constexpr (DBG) {
            if (b == 0) ck[7] = clock64() + (long long)(sc[0] == 123.f) + (long long)(sc[3] == 123.f);
        }
        float mb = -INFINITY;
#pragma unroll
        for (int u = 0; u <= ATT_ROWS; ++u) {
            const bool valid = u < ATT_ROWS ? (r0 + u < nk && r0 + u != jn) : with_new;
            // s / sqrt(hd): for hd = 64, 256 the divisor is a power of two and the product with its reciprocal is the same
            const float sv = (HD == 64 || HD == 256) ? sc[u] * (1.0f / sqrt_hd) : sc[u] / sqrt_hd;
            sc[u] = valid ? sv : -INFINITY;
            mb = fmaxf(mb, sc[u]);
        }
        if (mb > -INFINITY) {  // warp-uniform
            const float mn = fmaxf(m, mb);
            const float c = expf(m - mn);  // first batch: exp(-inf) = 0
            float pj[ATT_ROWS + 1];
            lsum *= c;
#pragma unroll
            for (int u = 0; u <= ATT_ROWS; ++u) {
                pj[u] = expf(sc[u] - mn);  // rows outside the range: exp(-inf) = 0
                lsum += pj[u];
            }
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                float a = o[i] * c;
#pragma unroll
                for (int u = 0; u < ATT_ROWS; ++u) a = fmaf(pj[u], vr[u][i], a);
                o[i] = fmaf(pj[ATT_ROWS], vnew[i], a);  // p = 0 unless this warp holds the new position in this batch
            }
            m = mn;
        }
        if constexpr (DBG) {
            if (b == 0) ck[2] = clock64() + (long long)(lsum == 123.f);
        }
        if (b + 1 < nb) {
#pragma unroll
            for (int u = 0; u < ATT_ROWS; ++u) load_row(Vc, r0 + 32 + u, vr[u]);
        }
    }
    if constexpr (DBG) ck[4] = clock64() + (long long)(lsum == 123.f);
    // ---- merge the 8 per-warp partials ----
    if (lane == 0) {
        wml[2 * warp] = m;
        wml[2 * warp + 1] = lsum;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float* dst = wpart + warp * HD + (c * 32 + lane) * VEC;
        if constexpr (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[c * 4], o[c * 4 + 1], o[c * 4 + 2], o[c * 4 + 3]);
        else if constexpr (VEC == 2) *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
        else *dst = o[0];
    }
    BAR1();
    if constexpr (DBG) ck[5] = clock64();
    if (tid < HD) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < MEGA_WARPS; ++w) M = fmaxf(M, wml[2 * w]);
        float Lsum = 0.0f, os = 0.0f;
#pragma unroll
        for (int w = 0; w < MEGA_WARPS; ++w) {
            const float cw = expf(wml[2 * w] - M);  // warps without keys: exp(-inf) = 0
            Lsum = fmaf(wml[2 * w + 1], cw, Lsum);
            os = fmaf(wpart[w * HD + tid], cw, os);
        }
        if constexpr (NORM) {
            st_tagged(o_out, item * HD + tid, os / Lsum, tag_out);
        } else {
            st_tagged(o_out, item * HD + tid, os, tag_out);
            if (tid == 0) st_tagged2(ml_out, item * 2, M, Lsum, tag_out);
        }
    }
    if constexpr (DBG) {
      ck[6] = clock64();
      if (dbg != nullptr && tid == 0) {
        // [prefetch issue, poll until q seen, dots + shuffles of batch 0, softmax weights + PV, later batches, barrier + merge + store]
        dbg[0] = (unsigned long long)(ck[1] - ck[0]);
        dbg[1] = (unsigned long long)(ck[3] - ck[1]);
        dbg[2] = (unsigned long long)(ck[7] - ck[3]);
        dbg[3] = (unsigned long long)(ck[2] - ck[7]);
        dbg[4] = (unsigned long long)(ck[4] - ck[2]);
        dbg[5] = (unsigned long long)(ck[6] - ck[4]);
      }
    }
}

// ---------------------------------------------------------------------------------------------
// Scores-only attention item (projected-value cache variant of the single-row kernel, decode_mega.cu): the item
// computes s_j = (q . k_j) / sqrt(hd) for its key range and publishes the raw scaled scores; softmax and the weighted
// sum run in every CTA against the cache of PROJECTED values (v_j . W_proj, per head), so there is no P.V product, no
// partial-output merge and no barrier after q here: prefetch K rows, poll q, dots, one shuffle tree, tagged stores.
// The warp that owns the position being decoded also takes k / v of this step from the exchange buffer, appends them
// to the K / V caches (the V cache stays complete for the per-op path) and scores the new key.
// ---------------------------------------------------------------------------------------------
template <int HD>
__device__ __noinline__ void score_item(float* __restrict__ Kc, float* __restrict__ Vc, const float* xq, int D, int h, int j0, int j1,
                                        int S, uint32_t tag_in, float* qs, int tid, float* s_out, uint32_t tag_out, uint32_t tmask) {
    using L = AttLane<HD>;
    constexpr int VEC = L::VEC, NCH = L::NCH, DPL = L::DPL;
    const int warp = tid >> 5, lane = tid & 31;
    const int nk = j1 - j0;
    const int nb = (nk + 31) >> 5;
    const int jn = S - 1 - j0;  // relative index of the position being decoded (inside this item iff 0 <= jn < nk)
    const bool mine = jn >= 0 && jn < nk && ((jn & 31) >> 2) == warp;  // this warp owns the new position
    const int bn = jn >> 5;
    const float sqrt_hd = sqrtf((float)HD);
    auto load_row = [&](const float* base, int jr, float* dst) {  // cache row of relative key jr (zeros outside / new key)
        if (jr < nk && jr != jn) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) ld_vec<VEC>(base + (size_t)(j0 + jr) * HD + (c * 32 + lane) * VEC, dst + c * VEC);
        } else {
#pragma unroll
            for (int i = 0; i < DPL; ++i) dst[i] = 0.0f;
        }
    };
    float kr[ATT_ROWS][DPL];
#pragma unroll
    for (int u = 0; u < ATT_ROWS; ++u) load_row(Kc, warp * ATT_ROWS + u, kr[u]);
    float qr[DPL], knew[DPL], vnew[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) knew[i] = vnew[i] = 0.0f;
    {
        const float* qp = xq + 2 * (size_t)(h * HD);
        const float* kp = xq + 2 * (size_t)(D + h * HD);
        const float* vp = xq + 2 * (size_t)(2 * D + h * HD);
        constexpr int EPW = HD / MEGA_WARPS;
        float qmine = 0.0f;
        uint32_t spins = 0;
        bool ok = false;
        while (!ok) {
            ok = true;
            if (lane < EPW) {
                const uint2 a = ld_x8(qp + 2 * (warp * EPW + lane));
                ok = ((a.y ^ tag_in) & tmask) == 0u;
                qmine = __uint_as_float(a.x);
            }
            if (mine) {
                const bool ok_k = ld_tagged_lane<VEC, NCH>(kp, lane, tag_in, tmask, knew);
                const bool ok_v = ld_tagged_lane<VEC, NCH>(vp, lane, tag_in, tmask, vnew);
                ok = ok && ok_k && ok_v;
            }
            ok = __all_sync(0xffffffffu, ok);
            GV_SPIN(spins, __LINE__, 0u, 0u);
        }
        if (lane < EPW) qs[warp * EPW + lane] = qmine;
        BAR1();
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < VEC; ++i) qr[c * VEC + i] = qs[(c * 32 + lane) * VEC + i];
    }
    if (mine) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            st_vec<VEC>(Kc + (size_t)(S - 1) * HD + (c * 32 + lane) * VEC, knew + c * VEC);
            st_vec<VEC>(Vc + (size_t)(S - 1) * HD + (c * 32 + lane) * VEC, vnew + c * VEC);
        }
    }
    for (int b = 0; b < nb; ++b) {
        const int r0 = b * 32 + warp * ATT_ROWS;
        const bool with_new = mine && b == bn;  // warp-uniform
        float sc[ATT_ROWS + 1];
#pragma unroll
        for (int u = 0; u < ATT_ROWS; ++u) {
            float a = 0.0f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) a = fmaf(qr[i], kr[u][i], a);
            sc[u] = a;
        }
        {
            float a = 0.0f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) a = fmaf(qr[i], knew[i], a);
            sc[ATT_ROWS] = a;
        }
        if (b + 1 < nb) {
#pragma unroll
            for (int u = 0; u < ATT_ROWS; ++u) load_row(Kc, r0 + 32 + u, kr[u]);
        }
#pragma unroll
        for (int x = 16; x > 0; x >>= 1) {
#pragma unroll
            for (int u = 0; u <= ATT_ROWS; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], x);
        }
        // lane u publishes row u of the batch (lane ATT_ROWS: the new position)
        float mys = sc[0];
#pragma unroll
        for (int u = 1; u <= ATT_ROWS; ++u) mys = (lane == u) ? sc[u] : mys;
        // s / sqrt(hd): for hd = 64, 256 the divisor is a power of two and the product with its reciprocal is the same
        mys = (HD == 64 || HD == 256) ? mys * (1.0f / sqrt_hd) : mys / sqrt_hd;
        if (lane < ATT_ROWS) {
            if (r0 + lane < nk && r0 + lane != jn) st_tagged(s_out, j0 + r0 + lane, mys, tag_out);
        } else if (lane == ATT_ROWS && with_new) {
            st_tagged(s_out, S - 1, mys, tag_out);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// producer: one thread walks the CTA's weight stream through the ring
// ---------------------------------------------------------------------------------------------
struct Producer {
    const Ring& ring;
    const float* region = nullptr;   // this CTA's contiguous weight stream [region, region + region_floats)
    long long region_floats = 0;
    long long ahead_floats = 0;      // L2 prefetch distance (0 = off)
    uint32_t t = 0, slot = 0, phase = 0;
    uint32_t land = 0;  // tiles [0, land) have been seen complete (published to the consumers through ring.landed)
    uint32_t window;    // at most this many tiles requested but not landed
    volatile int* stop;
    volatile int* hold = nullptr;  // non-null: do not issue new copies while *hold != 0 (consumers are polling / loading a hop)
    uint64_t policy;
    __device__ Producer(const Ring& r, volatile int* s, uint32_t w) : ring(r), window(w), stop(s) {
        policy = l2_policy_evict_first();
    }
    // non-blocking: move `land` over every tile whose full barrier has completed and publish it
    __device__ void advance() {
        const uint32_t l0 = land;
        while (land < t && mbar_test_wait(&ring.full[land % NSLOT], (land / NSLOT) & 1u)) ++land;
        if (land != l0) st_release_cta_shared(ring.landed, land);
    }
    // returns false when the consumers asked to stop
    __device__ bool issue(const float* src, uint32_t floats, bool stream_once) {
        uint32_t spins = 0;
        while (!mbar_test_wait(&ring.empty[slot], phase ^ 1u)) {
            advance();
            if (*stop) return false;
            GV_SPIN(spins, __LINE__, t, land);
        }
        if (*stop) return false;
        if (hold != nullptr) {
            spins = 0;
            while (*hold) {
                advance();
                if (*stop) return false;
                GV_SPIN(spins, __LINE__, 0u, 0u);
            }
        }
        spins = 0;
        while (t - land >= window) {  // at most `window` tiles requested but not landed
            advance();
            GV_SPIN(spins, __LINE__, 0u, 0u);
        }
        mbar_arrive_expect_tx(&ring.full[slot], floats * 4u);
        float* dst = ring.slots + (size_t)slot * ring.slot_floats;
        if (stream_once) {
            bulk_g2s_hint(dst, src, floats * 4u, &ring.full[slot], policy);
            if (ahead_floats > 0) {
                // HBM -> L2 for the bytes this CTA will want `ahead_floats` later (wraps to the next token's pass):
                // the smem ring then refills at L2 latency, and HBM streaming no longer depends on ring depth
                long long o = (src - region) + ahead_floats;
                if (o >= region_floats) o -= region_floats;
                const long long n = min((long long)floats, region_floats - o);
                bulk_prefetch_l2(region + o, (uint32_t)n * 4u);
                if (n < (long long)floats) bulk_prefetch_l2(region, (uint32_t)((long long)floats - n) * 4u);
            }
        } else {
            bulk_g2s(dst, src, floats * 4u, &ring.full[slot]);
        }
        ++t;
        if (++slot == NSLOT) {
            slot = 0;
            phase ^= 1u;
        }
        return true;
    }
    __device__ bool issue_units(const float*& src, int nunits, int uf) {
        for (int u0 = 0; u0 < nunits; u0 += UPT) {
            const int nu = min(UPT, nunits - u0);
            if (!issue(src, (uint32_t)(nu * uf), true)) return false;
            src += (long long)nu * uf;
        }
        return true;
    }
};

__device__ __forceinline__ bool produce_forward(Producer& pr, const float* stream, const float* blob, long long lnf_off, int L, int D,
                                                const StreamDims& sd, int cta) {
    const float* base = stream + cta_base(sd, cta);
    const long long lfl = cta_layer_floats(sd, cta);
    const int uf = unit_floats(D);
    int nun[5];
    for (int ph = 0; ph < 5; ++ph) nun[ph] = ph_units(sd, ph, cta);
    for (int l = 0; l < L; ++l) {
        const float* lw = base + (long long)l * lfl;
        for (int ph = PH_QKV; ph <= PH_P2; ++ph)
            if (!pr.issue_units(lw, nun[ph], uf)) return false;
    }
    if (!pr.issue(blob + lnf_off, 4 * D, false)) return false;  // ln_f and final_norm parameters
    const float* hw = base + (long long)L * lfl;
    return pr.issue_units(hw, nun[PH_HEAD], uf);
}

}  // namespace GV_MEGA_NS
}  // namespace gv
