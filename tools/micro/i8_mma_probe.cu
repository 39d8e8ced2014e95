// Probe: does tcgen05.mma kind::i8 (u8 x u8 -> s32, M = 128, N = 256, K = 32) run on this GPU, and what does it compute on
// constant operands (A = 1, B = 2 everywhere -> every accumulator element = 64)?   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void probe(int* out, uint32_t idesc, int kind) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    uint8_t* sA = smem;            // 128 x 128 B
    uint8_t* sB = smem + 16384;    // 256 x 128 B
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sA[i] = 1;
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) sB[i] = 2;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tslot;
    if (threadIdx.x == 0) {
        for (int j = 0; j < 4; ++j) {
            uint64_t a = desc128(smem_u32(sA) + j * 32), b = desc128(smem_u32(sB) + j * 32);
            uint32_t acc = j ? 1u : 0u;
            if (kind == 0)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x < 128) {
        uint32_t r0, r1, r2, r3;
        const uint32_t taddr = tb + ((uint32_t)(threadIdx.x & ~31u) << 16) + 252;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[threadIdx.x] = (int)r3;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
int main() {
    int* d; cudaMalloc(&d, 512);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
    const uint32_t idesc_i8 = (2u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_f8 = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    for (int kind = 0; kind < 2; ++kind) {
        cudaMemset(d, 0xff, 512);
        probe<<<1, 128, 49152, 0>>>(d, kind == 0 ? idesc_i8 : idesc_f8, kind);
        cudaError_t e = cudaDeviceSynchronize();
        int h[128]; cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost);
        printf("kind %s: %s; out[0]=%d out[127]=%d (as float %g)\n", kind == 0 ? "i8" : "f8f6f4", cudaGetErrorString(e), h[0], h[127], *(float*)&h[0]);
        if (e != cudaSuccess) { cudaDeviceReset(); cudaMalloc(&d, 512); }
    }
    return 0;
}
