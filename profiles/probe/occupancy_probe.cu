#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_plain(uint32_t *o) { extern __shared__ unsigned char sm[]; o[threadIdx.x] = sm[threadIdx.x]; }
__global__ void __launch_bounds__(256) k_tmem(uint32_t *o) {
    extern __shared__ unsigned char sm[];
    uint32_t *slot = (uint32_t *)sm;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    uint32_t t = *slot;
    o[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(128u) : "memory");
}
__global__ void __launch_bounds__(256) k_mbar(uint32_t *o) {   // mbarrier + bulk copy, no tcgen05
    extern __shared__ unsigned char sm[];
    uint64_t *bar = (uint64_t *)sm;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(1) : "memory");
    __syncthreads();
    o[threadIdx.x] = sm[threadIdx.x + 64];
}
int main() {
    int n;
    for (size_t s : {16384ul, 65536ul, 110000ul}) {
        cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000); cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000); cudaFuncSetAttribute(k_mbar, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_plain, 256, s); printf("plain smem %zu -> %d\n", s, n);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tmem, 256, s); printf("tmem  smem %zu -> %d\n", s, n);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_mbar, 256, s); printf("mbar  smem %zu -> %d\n", s, n);
    }
    return 0;
}
