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
__device__ __forceinline__ bool mz_elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
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

// ---- table-driven execution ------------------------------------------------------------------------------------
// The three networks are compiled ONCE per kernel (by one thread) into a table of "rounds" in shared memory.  A round is
// up to two independent Dense layers (e.g. the first layers of the two heads of a network, which read the same trunk
// tile): they are issued back to back into the group's two 32-column accumulators, committed once, and their epilogues
// share one barrier round trip.  Everything a round needs (UMMA descriptors, K steps, bias address, destination) is in
// its 80-byte descriptor, so the per-round code is one compact, register-only loop body.
struct __align__(16) mz_tc_rdesc {
    unsigned long long adesc[2], bdesc[2];
    uint32_t dst_tile[2];      // bf16 B tile of the next layer (shared address) or 0
    uint32_t dst_f32[2];       // fp32 [m*32 + n] output (shared address) or 0
    uint32_t bias[2];          // shared address of the layer's fp32 bias
    int16_t ks[2], out[2], act[2], njobs, pad_;
};
#define MZ_TC_MAX_ROUNDS 40

#ifdef MZ_PHASE_TIMERS
#define MZ_RT(i) do { if (tk) { long long c_ = clock64(); if ((i) > 0) tk[(i) - 1] += c_ - tprev; tprev = c_; } } while (0)
#else
#define MZ_RT(i)
#endif


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
__device__ __forceinline__ float mz_lds32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }

// bias + activation + store of one job's accumulator fragment (16 values per thread, rows m0 = 16w + t/4 and m0 + 8)
__device__ __forceinline__ void mz_tc_epilogue(const uint32_t (&v)[16], int out, int act, uint32_t bias, uint32_t dst_tile, uint32_t dst_f32, int w, int t) {
    const int c = 2 * (t & 3);
    const float lo = act == MZ_ACT_RELU ? 0.0f : -INFINITY;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int m = 16 * w + (t >> 2) + 8 * half;
        if (m < out) {
            const float b = mz_lds32(bias + (uint32_t)m * 4u);
            float x[8];
#pragma unroll
            for (int q = 0; q < 4; q++) { x[2 * q] = fmaxf(__uint_as_float(v[4 * q + 2 * half]) + b, lo); x[2 * q + 1] = fmaxf(__uint_as_float(v[4 * q + 2 * half + 1]) + b, lo); }
            if (act == MZ_ACT_TANH) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = mz_tanhf_noinline(x[i]);
            }
            if (dst_tile) {
                const uint32_t colbase = dst_tile + (uint32_t)((m & 7) * 2), chunk = (uint32_t)(m >> 3);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int n = 8 * (i >> 1) + c + (i & 1);
                    unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(x[i]));
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(colbase + (uint32_t)((n >> 3) * 1024 + (n & 7) * 128) + ((chunk ^ (uint32_t)(n & 7)) << 4)), "h"(h) : "memory");
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(dst_f32 + (uint32_t)((m * MZ_ROWS + 8 * q + c) * 4)), "f"(x[2 * q]), "f"(x[2 * q + 1]) : "memory");
            }
        }
    }
}

// Runs rounds [first, first+count) of the table for one 128-thread group.  All state is in registers.
__device__ __noinline__ uint32_t mz_tc_run(const mz_tc_rdesc *prog, int first, int count, uint32_t tmem_d, uint32_t mbar, uint32_t q, int grp, int gtid,
                                           long long *tk) {
    long long tprev = 0; (void)tprev; (void)tk;
    const int w = gtid >> 5, t = gtid & 31;
    const uint32_t lane_base = tmem_d + ((uint32_t)(32 * w) << 16);     // M = 64: rows 16w..16w+15 live in lanes 32w..32w+15
    for (int r = first; r < first + count; r++) {
        const mz_tc_rdesc *R = prog + r;
        const int njobs = R->njobs;
        MZ_RT(0);
        if (w == 0) {   // warp-uniform branch + elect.sync: the MMA operands stay in uniform registers (no per-lane waterfall loop)
            mz_tc_fence_after();
            const uint64_t ad0 = R->adesc[0], bd0 = R->bdesc[0], ad1 = R->adesc[1], bd1 = R->bdesc[1];
            if (mz_elect_one()) {
                // always 4 K-steps: operand tiles are zero-filled beyond a layer's real K, so the extra steps add exact zeros
#pragma unroll
                for (int k = 0; k < 4; k++) mz_tc_mma(tmem_d, ad0 + (uint64_t)(2 * k), bd0 + (uint64_t)(2 * k), k > 0 ? 1u : 0u);   // +32 B per K=16 step
                if (njobs > 1) {
#pragma unroll
                    for (int k = 0; k < 4; k++) mz_tc_mma(tmem_d + 32u, ad1 + (uint64_t)(2 * k), bd1 + (uint64_t)(2 * k), k > 0 ? 1u : 0u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
            }
            __syncwarp();
        }
        MZ_RT(1);
        {
            uint32_t ok = 0, spin = 0;
            while (!ok) {
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(mbar), "r"(q & 1u) : "memory");
                if (++spin > (1u << 24)) __trap();
            }
        }
        MZ_RT(2);
        mz_tc_fence_after();
        __syncwarp();
        uint32_t v0[16], v1[16];
        mz_tc_ld16x256(lane_base, v0);
        if (njobs > 1) mz_tc_ld16x256(lane_base + 32u, v1);
        mz_tc_wait_ld();
        MZ_RT(3);
        mz_tc_epilogue(v0, R->out[0], R->act[0], R->bias[0], R->dst_tile[0], R->dst_f32[0], w, t);
        if (njobs > 1) mz_tc_epilogue(v1, R->out[1], R->act[1], R->bias[1], R->dst_tile[1], R->dst_f32[1], w, t);
        MZ_RT(4);
        mz_fence_proxy_async();
        MZ_RT(5);
        mz_tc_fence_before();
        mz_group_sync(grp);
        MZ_RT(6);
        q++;
    }
    return q;
}

// ---- table construction (one thread, once per kernel) ------------------------------------------------------------
struct mz_tc_builder { mz_tc_rdesc *prog; int n; uint32_t w_base, bias_base; };
__device__ __forceinline__ void mz_tc_emit(mz_tc_builder &B, const mz_params &P, int l0, uint32_t src0, uint32_t dt0, uint32_t df0,
                                           int l1, uint32_t src1, uint32_t dt1, uint32_t df1) {
    mz_tc_rdesc &R = B.prog[B.n++];
    const int ls[2] = {l0, l1}; const uint32_t srcs[2] = {src0, src1}, dts[2] = {dt0, dt1}, dfs[2] = {df0, df1};
    R.njobs = l1 >= 0 ? 2 : 1; R.pad_ = 0;
    for (int j = 0; j < 2; j++) {
        const int l = ls[j] >= 0 ? ls[j] : l0;
        R.adesc[j] = mz_tc_desc(B.w_base + (uint32_t)P.tc_a_off[l]); R.bdesc[j] = mz_tc_desc(srcs[j] ? srcs[j] : src0);
        R.ks[j] = (int16_t)P.tc_ksteps[l]; R.out[j] = (int16_t)P.layers[l].out; R.act[j] = (int16_t)P.layers[l].act;
        R.bias[j] = B.bias_base + (uint32_t)P.tc_bias_off[l] * 4u; R.dst_tile[j] = dts[j]; R.dst_f32[j] = dfs[j];
    }
}
// rounds of one network: trunk -> bufT tile, then the two heads in lockstep (head 2 at most two layers deep: it has the
// single scratch tile tx); representation: the trunk's last layer -> h1 (fp32).  Returns the number of rounds emitted.
__device__ __forceinline__ int mz_tc_build_net(mz_tc_builder &B, const mz_params &P, int net, uint32_t src, uint32_t bufT, uint32_t h1dst, uint32_t h2dst,
                                               uint32_t t0, uint32_t t1, uint32_t tx) {
    const mz_net &N = P.nets[net];
    const int f = N.first, n0 = B.n;
    uint32_t cur = src;
    for (int i = 0; i < N.n_trunk; i++) {
        const bool last = i == N.n_trunk - 1;
        const uint32_t d = last ? bufT : ((i & 1) ? t1 : t0);
        if (last && N.n_h1 == 0) mz_tc_emit(B, P, f + i, cur, 0u, h1dst, -1, 0u, 0u, 0u);
        else mz_tc_emit(B, P, f + i, cur, d, 0u, -1, 0u, 0u, 0u);
        cur = d;
    }
    if (N.n_h1 == 0) return B.n - n0;
    const int f1 = f + N.n_trunk, f2 = f1 + N.n_h1;
    if (N.n_h2 > 2 || N.n_h2 > N.n_h1) {   // general fallback: heads one after the other
        uint32_t c1 = bufT;
        for (int i = 0; i < N.n_h1; i++) { const bool last = i == N.n_h1 - 1; const uint32_t d = (i & 1) ? t1 : t0; mz_tc_emit(B, P, f1 + i, c1, last ? 0u : d, last ? h1dst : 0u, -1, 0u, 0u, 0u); c1 = d; }
        uint32_t c2 = bufT;
        for (int i = 0; i < N.n_h2; i++) { const bool last = i == N.n_h2 - 1; const uint32_t d = (i & 1) ? t1 : t0; mz_tc_emit(B, P, f2 + i, c2, last ? 0u : d, last ? h2dst : 0u, -1, 0u, 0u, 0u); c2 = d; }
        return B.n - n0;
    }
    uint32_t cur1 = bufT, cur2 = bufT;
    for (int i = 0; i < N.n_h1; i++) {
        const bool last1 = i == N.n_h1 - 1, has2 = i < N.n_h2, last2 = i == N.n_h2 - 1;
        const uint32_t d1 = (i & 1) ? t1 : t0;
        const uint32_t d2 = last1 ? d1 : tx;     // when head 1 ends here its output is fp32, so its ping-pong tile is free for head 2
        mz_tc_emit(B, P, f1 + i, cur1, last1 ? 0u : d1, last1 ? h1dst : 0u, has2 ? f2 + i : -1, cur2, (has2 && !last2) ? d2 : 0u, (has2 && last2) ? h2dst : 0u);
        cur1 = d1; cur2 = d2;
    }
    return B.n - n0;
}

struct mz_tc_plan {
    uint32_t w_base; float *bias; uint64_t *mbar_w; uint64_t *mbar_mma[2]; uint32_t *tmem_slot;
    uint32_t inS, in0, in1, bufT[2], t0[2], t1[2];     // bf16 tiles (shared addresses)
    unsigned char *tiles_ptr;                          // generic pointer to inS (all 9 tiles are contiguous)
    float *outV, *outL, *outR, *outH; double *pbc0, *sqrtN; uint16_t *path; mz_tc_rdesc *prog;
};
__host__ __device__ inline size_t mz_tc_smem_bytes(int image_bytes, int bias_floats, int hidden_pad, int S) {
    size_t w = ((size_t)image_bytes + 1023) & ~(size_t)1023;
    size_t small = (size_t)(4 + 16 + 4) * MZ_ROWS * 4 + (size_t)hidden_pad * MZ_ROWS * 4;
    size_t tab = (((size_t)S + 2) * 8 * 2 + 127) & ~(size_t)127;
    size_t path = (((size_t)S + 2) * 2 * MZ_ROWS + 127) & ~(size_t)127;
    size_t bias = ((size_t)bias_floats * 4 + 127) & ~(size_t)127;
    return 1024 + w + 9 * MZ_TC_TILE_BYTES + bias + 128 + small + tab + path + MZ_TC_MAX_ROUNDS * sizeof(mz_tc_rdesc) + 128;
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
    p.path = (uint16_t *)c; c += (((size_t)S + 2) * 2 * MZ_ROWS + 127) & ~(size_t)127;
    p.prog = (mz_tc_rdesc *)c;
    return p;
}
