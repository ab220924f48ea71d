// How long do small tcgen05.mma instructions take?  One CTA per SM; `issuers` warps (one per warpgroup) each issue `nm`
// MMAs of shape M=128, N, K=16 (bf16, fp32 accumulate) into their own TMEM columns, commit to their own mbarrier and wait.
// Reports cycles from first issue to (a) all issued, (b) commit observed, for the slowest issuer of CTA 0.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu ; run: ./mma_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) {
    return (uint64_t)((a & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

__global__ void __launch_bounds__(512, 1) probe(int N, int nm, int issuers, long long *out, int stagger = 0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar[4];
    const int tid = threadIdx.x, wg = tid >> 7;
    for (int i = tid; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 4) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[tid])));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 3; rep++) {
        __syncthreads();
        if ((tid & 127) < 32 && wg < issuers) {
            { const long long until = clock64() + (long long)stagger * wg; while (clock64() < until) { } }
            t0 = clock64();
            if (elect_one()) {
                const uint64_t ad = desc(smem_u32(smem)), bd = desc(smem_u32(smem + 16384));
                const uint32_t d = tmem + (uint32_t)((wg * N) & 511);
                for (int i = 0; i < nm; i++) mma(d, ad + (uint64_t)(2 * (i & 3)), bd + (uint64_t)(2 * (i & 3)), idesc, i > 0);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[wg])) : "memory");
            }
            __syncwarp();
            t1 = clock64();
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar[wg])), "r"((uint32_t)(rep & 1)) : "memory");
            t2 = clock64();
        }
    }
    if (blockIdx.x == 0 && (tid & 127) == 0 && wg < issuers) { out[2 * wg] = t1 - t0; out[2 * wg + 1] = t2 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long *out; cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
    printf("%4s %4s %8s | cycles until issued / until commit observed (slowest issuer)\n", "N", "nm", "issuers");
    const int Ns[] = {64, 128, 256}, nms[] = {1, 4, 8, 16, 64}, iss[] = {1, 4};
    for (int N : Ns) for (int is : iss) for (int nm : nms) {
        if (is * N > 512) continue;
        for (int i = 0; i < 8; i++) out[i] = 0;
        probe<<<148, 512, 16384 + 32768 + 1024>>>(N, nm, is, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long a = 0, b = 0;
        for (int w = 0; w < is; w++) { if (out[2 * w] > a) a = out[2 * w]; if (out[2 * w + 1] > b) b = out[2 * w + 1]; }
        printf("%4d %4d %8d | %6lld %6lld   (%.1f cycles per MMA over all issuers)\n", N, nm, is, a, b, (double)b / (nm * is));
    }
    printf("4 issuers x 4 MMAs (N=64) started `stagger` cycles apart: cycles until issued / until commit observed, per issuer\n");
    for (int st : {0, 150, 300, 600}) {
        for (int i = 0; i < 8; i++) out[i] = 0;
        probe<<<148, 512, 16384 + 32768 + 1024>>>(64, 4, 4, out, st);
        cudaDeviceSynchronize();
        printf("stagger %4d |", st);
        for (int w = 0; w < 4; w++) printf(" %5lld/%5lld", out[2 * w], out[2 * w + 1]);
        printf("\n");
    }
    return 0;
}
