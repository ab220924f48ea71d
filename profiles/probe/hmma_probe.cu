// How fast is a register-resident chain of small Dense layers on the legacy warp-level tensor path (mma.sync m16n8k16 bf16, SASS HMMA)
// on B200?  One warp = 16 rows (trees); a 64 -> 64 layer = 8 n-tiles x 4 k-steps x TERMS MMAs (TERMS = 1: bf16, 3: split-precision
// hi*lo + lo*hi + hi*hi); the accumulator fragments of a layer ARE the A fragments of the next one (bias + relu + bf16 hi/lo split in
// registers), the weights come from shared memory in fragment order (one 16-byte load per lane per MMA group).
// Reports cycles per layer for `warps` warps per CTA, one CTA per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe hmma_probe.cu ; run: ./hmma_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#ifndef UNROLL
#define UNROLL 1
#endif
#define PRAGMA_(x) _Pragma(#x)
#define PRAGMA(x) PRAGMA_(x)

__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
// x0, x1 -> packed bf16 hi parts and packed bf16 lo parts (x - hi)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &h, uint32_t &l) {
    h = pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
    l = pack_bf16x2(x0 - h0, x1 - h1);
}

template <int TERMS>
__global__ void __launch_bounds__(256, 1) probe(int layers, long long *out, float *sink) {
    extern __shared__ __align__(16) uint4 wsm[];    // [layer % 2][j][s][lane] = {hi b0, hi b1, lo b0, lo b1}
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 2 * 32 * 32; i += blockDim.x) wsm[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x38003800u, 0x38003800u);
    __syncthreads();
    uint32_t ah[4][4], al[4][4];
    for (int s = 0; s < 4; s++) for (int i = 0; i < 4; i++) { ah[s][i] = 0x3c003c00u + lane; al[s][i] = 0x38003800u; }
    long long t0 = clock64();
PRAGMA(unroll UNROLL)
    for (int L = 0; L < layers; L++) {
        float d[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) { d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.0f; }
        const uint4 *w = wsm + (L & 1) * 1024;
#pragma unroll
        for (int s = 0; s < 4; s++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint4 b = w[(j * 4 + s) * 32 + lane];
                if (TERMS == 3) { hmma(d[j], ah[s], b.z, b.w); hmma(d[j], al[s], b.x, b.y); }
                hmma(d[j], ah[s], b.x, b.y);
            }
        }
        // epilogue: bias + relu + split -> next layer's A fragments
#pragma unroll
        for (int s = 0; s < 4; s++) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int j = 2 * s + half;
                const float b0 = 0.001f * (float)j, b1 = 0.002f;
                float x0 = fmaxf(d[j][0] * 1e-3f + b0, 0.0f), x1 = fmaxf(d[j][1] * 1e-3f + b1, 0.0f), x2 = fmaxf(d[j][2] * 1e-3f + b0, 0.0f), x3 = fmaxf(d[j][3] * 1e-3f + b1, 0.0f);
                split2(x0, x1, ah[s][2 * half], al[s][2 * half]);
                split2(x2, x3, ah[s][2 * half + 1], al[s][2 * half + 1]);
            }
        }
    }
    long long t1 = clock64();
    float acc = 0.0f;
    for (int s = 0; s < 4; s++) for (int i = 0; i < 4; i++) acc += __uint_as_float(ah[s][i] << 16) + __uint_as_float(al[s][i] << 16);
    sink[blockIdx.x * blockDim.x + tid] = acc;
    if (blockIdx.x == 0 && lane == 0) out[tid >> 5] = t1 - t0;
}

int main() {
    long long *out; float *sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 148 * 256 * 4);
    const int layers = 64;
    printf("terms warps | cycles per 64x64 layer (16 rows per warp), slowest warp of CTA 0; HMMA per layer per warp = 32 x terms\n");
    for (int terms = 1; terms <= 3; terms += 2)
        for (int warps = 1; warps <= 8; warps *= 2) {
            for (int rep = 0; rep < 2; rep++) {
                if (terms == 1) probe<1><<<148, warps * 32, 32768>>>(layers, out, sink); else probe<3><<<148, warps * 32, 32768>>>(layers, out, sink);
                cudaDeviceSynchronize();
            }
            long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < warps; i++) mx = h[i] > mx ? h[i] : mx;
            printf("%5d %5d | %8.1f   (%.2f cycles per HMMA per warp)\n", terms, warps, (double)mx / layers, (double)mx / layers / (32.0 * terms));
        }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
