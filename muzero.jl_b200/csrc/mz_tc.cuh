// mz_tc.cuh -- tensor-core network path (MZ_NN_BF16_TC): the Dense layers of the three networks as tcgen05.mma
// (5th-gen tensor cores, sm_100a) with accumulators in TMEM.
//
// Orientation: D[out feature m][tree n] = sum_k W[m][k] * X[n][k], i.e. the WEIGHTS are the A operand (M = 64 output
// features = width_hidden), the 32 trees of the CTA are the N dimension, so a CTA keeps its 32 trees and all 148 SMs
// stay busy.  All weights of all three networks (156 KB as bf16) are staged ONCE per kernel by TMA bulk copies into
// shared memory in the UMMA K-major SWIZZLE_128B layout (pre-swizzled on the host) and stay resident; activations are
// bf16 [32 trees x 64] tiles in the same layout, produced by the epilogue of the previous layer:
//     tcgen05.mma (1 thread)  ->  tcgen05.commit -> mbarrier  ->  tcgen05.ld (TMEM -> registers)
//     -> bias + activation -> bf16 -> st.shared into the next layer's B tile -> fence.proxy.async -> barrier.
// The two 128-thread groups of the CTA run the prediction and the dynamics chain concurrently on disjoint TMEM columns.
// Final layers write fp32 to the same output buffers the exact path uses, so the tree code is shared.
#pragma once
#include <cuda_bf16.h>
#include "mz_device.cuh"

#define MZ_TC_TILE_BYTES 4096          // [32 rows x 64 bf16]
#define MZ_TC_TMEM_COLS 128            // two groups x two accumulators of 32 fp32 columns
// instruction descriptor, kind::f16: D=F32 (bits 4-5 = 1), A=BF16 (bits 7-9 = 1), B=BF16 (bits 10-12 = 1), K-major A and B,
// N>>3 at bits 17-22, M>>4 at bits 24-28  (cute::UMMA::InstrDescriptor)
#define MZ_TC_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((64u >> 4) << 24))

__device__ __forceinline__ void mz_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mz_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mz_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64).  K-major SW128: SBO = 1024 B between 8-row groups, LBO unused (1).
__device__ __forceinline__ uint64_t mz_tc_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mz_tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(MZ_TC_IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mz_tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mz_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mz_tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mz_tc_alloc(uint32_t *slot) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mz_smem_u32(slot)), "r"((uint32_t)MZ_TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mz_tc_dealloc(uint32_t taddr) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)MZ_TC_TMEM_COLS) : "memory");
}
// byte offset of element (row n, k) in a K-major SWIZZLE_128B tile (same function as mzh::tc_tile_offset)
__device__ __forceinline__ uint32_t mz_tc_tile_offset(int n, int k) {
    return (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2);
}
__device__ __forceinline__ void mz_tc_store_bf16(uint32_t tile, int n, int k, float v) {
    unsigned short b = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(tile + mz_tc_tile_offset(n, k)), "h"(b) : "memory");
}

__device__ __noinline__ float mz_tanhf_noinline(float x) { return mz_tanhf(x); }

struct mz_tc_pipe {             // one per group
    uint32_t w_base;            // shared address of the weight image
    const float *bias;          // shared fp32 bias block
    uint64_t *mbar;             // MMA-done barrier of this group
    uint32_t tmem_d;            // this group's 32 accumulator columns
    uint32_t q;                 // layers executed by this group (barrier parity)
    int grp, gtid;
};

// TMEM -> registers, shape 16x256b.x4: the warp reads 16 TMEM lanes x 32 columns; thread t receives, for column block
// q = 0..3, v[4q+0..1] = (lane t/4,   columns 8q + 2(t%4) + {0,1}) and v[4q+2..3] = (lane t/4 + 8, same columns)
// (cute::SM100_TMEM_LOAD_16dp256b4x).  With M = 64 accumulators only lanes 0..15 of each warp's TMEM partition hold
// rows, so this shape keeps all 32 threads of every warp busy in the epilogue.
__device__ __forceinline__ void mz_tc_ld16x256(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void mz_tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct mz_tc_job { int layer; uint32_t src, dst_tile; float *dst_f32; };   // layer < 0: no job

// bias + activation + store of one job's accumulator fragment (16 values per thread)
__device__ __forceinline__ void mz_tc_epilogue(const mz_tc_pipe &s, const mz_params &P, const mz_tc_job &j, const uint32_t (&v)[16]) {
    const mz_layer &L = P.layers[j.layer];
    const int w = s.gtid >> 5, t = s.gtid & 31, c = 2 * (t & 3);
    const float lo = L.act == MZ_ACT_RELU ? 0.0f : -INFINITY;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int m = 16 * w + (t >> 2) + 8 * half;
        if (m < L.out) {
            const float b = s.bias[P.tc_bias_off[j.layer] + m];
            float x[8];
#pragma unroll
            for (int q = 0; q < 4; q++) { x[2 * q] = fmaxf(__uint_as_float(v[4 * q + 2 * half]) + b, lo); x[2 * q + 1] = fmaxf(__uint_as_float(v[4 * q + 2 * half + 1]) + b, lo); }
            if (L.act == MZ_ACT_TANH) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = mz_tanhf_noinline(x[i]);
            }
            if (j.dst_tile) {
                const uint32_t colbase = j.dst_tile + (uint32_t)((m & 7) * 2), chunk = (uint32_t)(m >> 3);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int n = 8 * (i >> 1) + c + (i & 1);
                    unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(x[i]));
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(colbase + (uint32_t)((n >> 3) * 1024 + (n & 7) * 128) + ((chunk ^ (uint32_t)(n & 7)) << 4)), "h"(h) : "memory");
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) *reinterpret_cast<float2 *>(j.dst_f32 + m * MZ_ROWS + 8 * q + c) = make_float2(x[2 * q], x[2 * q + 1]);
            }
        }
    }
}

// One round for one group: up to two independent Dense layers (e.g. the first layers of the two heads of a network,
// which read the same trunk tile) are issued back to back on the tensor core into the group's two 32-column
// accumulators, committed once, and their epilogues share one barrier round trip.
__device__ __noinline__ void mz_tc_round(mz_tc_pipe &s, const mz_params &P, mz_tc_job j0, mz_tc_job j1) {
    if (s.gtid == 0) {
        mz_tc_fence_after();
        {
            const uint64_t adesc = mz_tc_desc(s.w_base + (uint32_t)P.tc_a_off[j0.layer]), bdesc = mz_tc_desc(j0.src);
            const int ks = P.tc_ksteps[j0.layer];
            for (int k = 0; k < ks; k++) mz_tc_mma(s.tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), k > 0 ? 1u : 0u);   // +32 B per K=16 step
        }
        if (j1.layer >= 0) {
            const uint64_t adesc = mz_tc_desc(s.w_base + (uint32_t)P.tc_a_off[j1.layer]), bdesc = mz_tc_desc(j1.src);
            const int ks = P.tc_ksteps[j1.layer];
            for (int k = 0; k < ks; k++) mz_tc_mma(s.tmem_d + 32u, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), k > 0 ? 1u : 0u);
        }
        mz_tc_commit(s.mbar);
    }
    mz_mbar_wait(s.mbar, s.q & 1u);
    mz_tc_fence_after();
    __syncwarp();
    const uint32_t lane_base = s.tmem_d + ((uint32_t)(32 * (s.gtid >> 5)) << 16);     // M = 64: rows 16w..16w+15 live in lanes 32w..32w+15
    uint32_t v0[16], v1[16];
    mz_tc_ld16x256(lane_base, v0);
    if (j1.layer >= 0) mz_tc_ld16x256(lane_base + 32u, v1);
    mz_tc_wait_ld();
    mz_tc_epilogue(s, P, j0, v0);
    if (j1.layer >= 0) mz_tc_epilogue(s, P, j1, v1);
    mz_fence_proxy_async();
    mz_tc_fence_before();
    mz_group_sync(s.grp);
    s.q++;
}
__device__ __forceinline__ mz_tc_job mz_tc_mkjob(int layer, uint32_t src, uint32_t dst_tile, float *dst_f32) { mz_tc_job j; j.layer = layer; j.src = src; j.dst_tile = dst_tile; j.dst_f32 = dst_f32; return j; }
__device__ __forceinline__ void mz_tc_chain(mz_tc_pipe &s, const mz_params &P, int first, int n, uint32_t src, float *dst_f32, uint32_t t0, uint32_t t1) {
    uint32_t cur = src;
    for (int i = 0; i < n; i++) {
        const bool last = i == n - 1;
        uint32_t d = (i & 1) ? t1 : t0;
        mz_tc_round(s, P, mz_tc_mkjob(first + i, cur, last ? 0u : d, last ? dst_f32 : nullptr), mz_tc_mkjob(-1, 0u, 0u, nullptr));
        cur = d;
    }
}
// trunk -> bufT tile, then the two heads advance in lockstep, one round per depth level (head 2 may be at most two layers
// deep: it has the single scratch tile tx); representation: the trunk's last layer -> h1dst
__device__ __forceinline__ void mz_tc_net(mz_tc_pipe &s, const mz_params &P, int net, uint32_t src, uint32_t bufT, float *h1dst, float *h2dst,
                                          uint32_t t0, uint32_t t1, uint32_t tx) {
    const mz_net &N = P.nets[net];
    const int f = N.first;
    if (N.n_h1 == 0) { mz_tc_chain(s, P, f, N.n_trunk, src, h1dst, t0, t1); return; }
    uint32_t cur = src;
    for (int i = 0; i < N.n_trunk; i++) {
        uint32_t d = (i == N.n_trunk - 1) ? bufT : ((i & 1) ? t1 : t0);
        mz_tc_round(s, P, mz_tc_mkjob(f + i, cur, d, nullptr), mz_tc_mkjob(-1, 0u, 0u, nullptr));
        cur = d;
    }
    const int f1 = f + N.n_trunk, f2 = f1 + N.n_h1;
    if (N.n_h2 > 2 || N.n_h2 > N.n_h1) {   // general fallback: heads one after the other
        mz_tc_chain(s, P, f1, N.n_h1, bufT, h1dst, t0, t1);
        mz_tc_chain(s, P, f2, N.n_h2, bufT, h2dst, t0, t1);
        return;
    }
    uint32_t cur1 = bufT, cur2 = bufT;
    for (int i = 0; i < N.n_h1; i++) {
        const bool last1 = i == N.n_h1 - 1, has2 = i < N.n_h2, last2 = i == N.n_h2 - 1;
        // when head 1 ends here its output is fp32, so its ping-pong tile of this round is free for head 2
        const uint32_t d1 = (i & 1) ? t1 : t0;
        const uint32_t d2 = last1 ? d1 : tx;
        mz_tc_round(s, P, mz_tc_mkjob(f1 + i, cur1, last1 ? 0u : d1, last1 ? h1dst : nullptr),
                    has2 ? mz_tc_mkjob(f2 + i, cur2, last2 ? 0u : d2, last2 ? h2dst : nullptr) : mz_tc_mkjob(-1, 0u, 0u, nullptr));
        cur1 = d1; cur2 = d2;
    }
}

struct mz_tc_plan {
    uint32_t w_base; float *bias; uint64_t *mbar_w; uint64_t *mbar_mma[2]; uint32_t *tmem_slot;
    uint32_t inS, in0, in1, bufT[2], t0[2], t1[2];     // bf16 tiles (shared addresses)
    unsigned char *tiles_ptr;                          // generic pointer to inS (all 9 tiles are contiguous)
    float *outV, *outL, *outR, *outH; double *pbc0, *sqrtN; uint16_t *path;
};
__host__ __device__ inline size_t mz_tc_smem_bytes(int image_bytes, int bias_floats, int hidden_pad, int S) {
    size_t w = ((size_t)image_bytes + 1023) & ~(size_t)1023;
    size_t small = (size_t)(4 + 16 + 4) * MZ_ROWS * 4 + (size_t)hidden_pad * MZ_ROWS * 4;
    size_t tab = (((size_t)S + 2) * 8 * 2 + 127) & ~(size_t)127;
    size_t path = (((size_t)S + 2) * 2 * MZ_ROWS + 127) & ~(size_t)127;
    size_t bias = ((size_t)bias_floats * 4 + 127) & ~(size_t)127;
    return 1024 + w + 9 * MZ_TC_TILE_BYTES + bias + 128 + small + tab + path + 128;
}
__device__ __forceinline__ mz_tc_plan mz_tc_carve(unsigned char *raw, int image_bytes, int bias_floats, int hidden_pad, int S) {
    mz_tc_plan p;
    uint32_t a = mz_smem_u32(raw);
    unsigned char *base = raw + (((a + 1023u) & ~1023u) - a);           // SWIZZLE_128B tiles need 1024-byte alignment
    size_t w = ((size_t)image_bytes + 1023) & ~(size_t)1023;
    unsigned char *c = base;
    p.w_base = mz_smem_u32(c); c += w;
    p.tiles_ptr = c;
    p.inS = mz_smem_u32(c); c += MZ_TC_TILE_BYTES; p.in0 = mz_smem_u32(c); c += MZ_TC_TILE_BYTES; p.in1 = mz_smem_u32(c); c += MZ_TC_TILE_BYTES;
    for (int g = 0; g < 2; g++) { p.bufT[g] = mz_smem_u32(c); c += MZ_TC_TILE_BYTES; p.t0[g] = mz_smem_u32(c); c += MZ_TC_TILE_BYTES; p.t1[g] = mz_smem_u32(c); c += MZ_TC_TILE_BYTES; }
    p.bias = (float *)c; c += ((size_t)bias_floats * 4 + 127) & ~(size_t)127;
    p.mbar_w = (uint64_t *)c; p.mbar_mma[0] = (uint64_t *)(c + 16); p.mbar_mma[1] = (uint64_t *)(c + 32); p.tmem_slot = (uint32_t *)(c + 64); c += 128;
    p.outV = (float *)c; c += 4 * MZ_ROWS * 4;
    p.outL = (float *)c; c += 16 * MZ_ROWS * 4;
    p.outR = (float *)c; c += 4 * MZ_ROWS * 4;
    p.outH = (float *)c; c += (size_t)hidden_pad * MZ_ROWS * 4;
    p.pbc0 = (double *)c; p.sqrtN = p.pbc0 + (S + 2); c += (((size_t)S + 2) * 8 * 2 + 127) & ~(size_t)127;
    p.path = (uint16_t *)c;
    return p;
}
