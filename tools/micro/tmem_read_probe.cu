// Probe: TMEM -> register read bandwidth of one SM with the access pattern of the scan epilogues
// (tcgen05.ld.32x32b.x32: a warp reads its 32 lanes x 32 columns = 4 KiB per instruction), for 4, 8 and 16 reading warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_read_probe tmem_read_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(long long* out, int n_warps, int iters, int two_in_flight) {
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tslot;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < n_warps) {
        const uint32_t taddr = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 32;
        __syncwarp();
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t r[32], s[32];
            const uint32_t a = taddr + (uint32_t)(i & 1) * 128 + (uint32_t)((i >> 1) & 1) * 256;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                           "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                           "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(a) : "memory");
            if (two_in_flight) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(s[9]),
                               "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15]), "=r"(s[16]), "=r"(s[17]), "=r"(s[18]),
                               "=r"(s[19]), "=r"(s[20]), "=r"(s[21]), "=r"(s[22]), "=r"(s[23]), "=r"(s[24]), "=r"(s[25]), "=r"(s[26]), "=r"(s[27]),
                               "=r"(s[28]), "=r"(s[29]), "=r"(s[30]), "=r"(s[31]) : "r"(a ^ 64u) : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
            if (two_in_flight) {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= s[j];
            }
        }
        t1 = clock64();
    }
    __syncthreads();
    if (lane == 0 && warp < n_warps) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = acc; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 64 * 8);
    const int iters = 20000;
    for (int two = 0; two < 2; ++two)
        for (int nw : {1, 4, 8, 16}) {
            cudaMemset(d, 0, 64 * 8);
            probe<<<1, 512>>>(d, nw, iters, two);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[64]; cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w * 2] > mx ? h[w * 2] : mx;
            const double bytes = (double)nw * iters * 4096.0 * (two ? 2 : 1);
            printf("%s: %2d warps, %d load(s) in flight per warp: %.1f B/clk/SM (%lld cycles)\n", cudaGetErrorString(e), nw, two + 1, bytes / (double)mx, mx);
        }
    return 0;
}
