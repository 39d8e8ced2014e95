// Probe: 1-D tiled TMA load (int32, box 256) at an arbitrary element coordinate; REDUX.OR; 3-input integer max.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int* out) {
    __shared__ __align__(128) int buf[256];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bar)), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                     ::"r"(smem_u32(buf)), "l"(&tm), "r"(smem_u32(&bar)), "r"(c0) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    int v = buf[threadIdx.x];
    const uint32_t orv = __reduce_or_sync(0xffffffffu, (uint32_t)v);
    const int m3 = max(max(v, buf[(threadIdx.x + 1) & 255]), buf[(threadIdx.x + 2) & 255]);
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) { out[256] = (int)orv; out[257] = m3; }
}
int main() {
    int n = 1000;
    int* d; cudaMalloc(&d, n * 4 + 64);
    int* h = new int[n];
    for (int i = 0; i < n; ++i) h[i] = i;
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    int* out; cudaMalloc(&out, 260 * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)p;
    CUtensorMap tm;
    cuuint64_t dims[1] = {(cuuint64_t)n}; cuuint64_t strides[1] = {0}; cuuint32_t box[1] = {256}; cuuint32_t es[1] = {1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 1, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    for (int c0 : {0, 100, 101, 900}) {
        probe<<<1, 256>>>(tm, c0, out);
        cudaError_t e = cudaDeviceSynchronize();
        int ho[260]; cudaMemcpy(ho, out, 260 * 4, cudaMemcpyDeviceToHost);
        printf("c0=%d: %s; out[0]=%d out[1]=%d out[255]=%d or=%d m3=%d\n", c0, cudaGetErrorString(e), ho[0], ho[1], ho[255], ho[256], ho[257]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
