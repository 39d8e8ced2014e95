// Common device/host helpers for liblira_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <string>

namespace lira {

// ------------------------------------------------------------------------------------------
// host error plumbing (C ABI returns int status + lira_last_error(), SURVEY.md 8b "Errors")
// ------------------------------------------------------------------------------------------
void set_error(const std::string& msg);

#define LIRA_CUDA_OK(expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::lira::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +      \
                              __FILE__ + ":" + std::to_string(__LINE__) + ")");                \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

#define LIRA_REQUIRE(cond, msg)                                                                \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::lira::set_error(std::string(msg));                                               \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

// ------------------------------------------------------------------------------------------
// device: ordered keys. A candidate is (score, position); smaller score is better
// (L2: squared distance; IP: -inner product). The 64-bit key orders by score, then position,
// which is exactly the Faiss heap result order the oracle restates (lower position wins ties).
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

static constexpr unsigned long long KEY_INF = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ uint32_t f32_to_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t pos) {
    return ((unsigned long long)f32_to_ordered(score) << 32) | (unsigned long long)pos;
}
__device__ __forceinline__ float key_score(unsigned long long k) { return ordered_to_f32((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_pos(unsigned long long k) { return (uint32_t)(k & 0xFFFFFFFFull); }

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long v, int delta) {
    uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, delta);
    uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), delta);
    return ((unsigned long long)hi << 32) | lo;
}

// Sorted (ascending) list of 32*S keys held by one warp: element p lives in key[p / 32] of
// lane p % 32. Inserts x (x must differ from every key present), dropping the largest.
template <int S>
__device__ __forceinline__ void warp_sorted_insert(unsigned long long (&key)[S], unsigned long long x, int lane) {
#pragma unroll
    for (int s = S - 1; s >= 0; --s) {
        unsigned long long up = shfl_up_u64(key[s], 1);
        if (s > 0) {
            unsigned long long carry = shfl_u64(key[s - 1], 31);
            if (lane == 0) up = carry;
        }
        const bool first = (s == 0) && (lane == 0);
        const bool keep = key[s] < x;
        const bool prev_less = first ? true : (up < x);
        key[s] = keep ? key[s] : (prev_less ? x : up);
    }
}

// element p of the distributed list, broadcast to the whole warp
template <int S>
__device__ __forceinline__ unsigned long long warp_sorted_get(const unsigned long long (&key)[S], int p) {
    unsigned long long v = key[0];
#pragma unroll
    for (int s = 1; s < S; ++s)
        if ((p >> 5) == s) v = key[s];
    return shfl_u64(v, p & 31);
}

// ------------------------------------------------------------------------------------------
// device: PTX wrappers -- mbarrier, TMA (cp.async.bulk.tensor), cp.async
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Bounded wait: a protocol bug must trap (-> CUDA error on the host), never hang the GPU box.
// try_wait is issued WITHOUT a suspend-time hint: with a hint ptxas emits NANOSLEEP between polls and every
// wait then costs up to the hinted time in wake-up latency (measured: ~2 us per wait, 5x on the whole scan).
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity);
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_addr(smem_u32(bar), parity); }
// same, on a precomputed shared-window address (hot loops: the generic -> shared conversion is not free)
__device__ __forceinline__ void mbar_arrive_addr(uint32_t addr) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
    uint32_t done = 0;
    uint64_t t0 = 0;
#pragma unroll 1
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if ((spin & 0xFFFu) == 0xFFFu) {  // every 4096 failed polls: wall-clock guard (5 s)
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 5000000000ull) __trap();
        }
    }
}

// 2-D tiled TMA load: box lands in smem, completion bytes are counted on `bar`.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// 16-byte cp.async; src_bytes = 0 zero-fills the destination (used for rows/columns past the end).
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
// arrive on `bar` once all of this thread's earlier cp.async have landed (count pre-charged at init)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One lane of a CONVERGED warp (true for exactly one lane). Single-thread instructions (TMA, tcgen05.mma / commit)
// are issued under it while the surrounding loop stays warp-uniform: operands computed in uniform code go straight to
// uniform registers, whereas code under `if (lane == 0)` is divergent to the compiler and every such instruction is
// then wrapped in a per-thread serialisation loop (measured: 4x on the MMA issue path).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start (its
// CTAs become resident and run their own prologue) while the previous kernel of the stream is still draining. pdl_wait()
// returns once that kernel has completed and its writes are visible: EVERY thread of such a kernel executes it before the
// first access to anything another kernel produces or still reads, and before it can exit (completion of this grid then
// implies completion of every earlier one). pdl_launch() lets the NEXT kernel of the stream be scheduled as soon as every
// CTA of this one has issued it. Both are no-ops for a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#endif  // __CUDACC__

}  // namespace lira
