// dependent fmaf chain latency on one warp (is the 64-long chain of mz_lat_apply 4 cycles per link?)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const float *w, const float *x, float *out, long long *cyc, int warps_active) {
    __shared__ float sx[64];
    if (threadIdx.x < 64) sx[threadIdx.x] = x[threadIdx.x];
    float wr[64];
#pragma unroll
    for (int i = 0; i < 64; i++) wr[i] = w[i * 64 + (threadIdx.x & 63)];
    __syncthreads();
    if ((int)(threadIdx.x >> 5) >= warps_active) return;
    long long best = 1 << 30; float acc = 0;
    for (int rep = 0; rep < 8; rep++) {
        float4 xv[16];
#pragma unroll
        for (int i = 0; i < 16; i++) xv[i] = reinterpret_cast<float4 *>(sx)[i];
        long long t0 = clock64();
        acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; i++) { acc = fmaf(wr[4 * i], xv[i].x, acc); acc = fmaf(wr[4 * i + 1], xv[i].y, acc); acc = fmaf(wr[4 * i + 2], xv[i].z, acc); acc = fmaf(wr[4 * i + 3], xv[i].w, acc); }
        // make the end of the chain observable before the clock read
        if (acc == 123.456f) sx[0] = acc;
        long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
        sx[threadIdx.x & 63] = acc * 0.5f; __syncwarp();
    }
    out[threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = best;
}
int main() {
    float *w, *x, *o; long long *c;
    cudaMalloc(&w, 64 * 64 * 4); cudaMalloc(&x, 256); cudaMalloc(&o, 512); cudaMalloc(&c, 64);
    cudaMemset(w, 0, 64 * 64 * 4); cudaMemset(x, 0, 256);
    for (int wa = 1; wa <= 4; wa++) {
        k<<<1, 128>>>(w, x, o, c, wa); cudaDeviceSynchronize();
        long long h[4]; cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
        printf("%d warps: 64 dependent fmaf + x in registers: %lld cycles (%.2f per link)\n", wa, h[0], h[0] / 64.0);
    }
    return 0;
}
