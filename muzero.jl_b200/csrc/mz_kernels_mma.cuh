// mz_kernels_mma.cuh -- mz_k_search_mma: the fused search kernel of mz_kernels.cuh with the networks on the tensor cores at
// near-Float32 accuracy (nn_mode = MZ_NN_SPLIT_MMA).
//
// Why this shape.  A simulation is a dependent chain: select -> prediction(parent) || dynamics(parent, a) -> expand -> backup, and the
// dynamics network alone is 8 dependent 64-wide layers.  With 4096 games an SM holds 28 trees, so throughput = trees per SM / chain
// latency and nothing else.  The tcgen05 kernel (mz_kernels_tc.cuh) pays ~2.5 k cycles per layer round for the asynchronous round trip
// MMA issue -> commit -> mbarrier -> tcgen05.ld -> epilogue -> st.shared -> proxy fence -> barrier; this kernel keeps a whole chain in
// the registers of ONE warp instead:
//   * a warp owns a tile of 16 trees; D[tree][feature] = sum_k X[tree][k] W[feature][k] as mma.sync.m16n8k16 (bf16, fp32 accumulate):
//     the accumulator fragment of a layer has exactly the register layout of the A fragment of the next layer, so bias + activation +
//     bf16 conversion happen in registers and a layer costs no shared-memory round trip and no barrier at all;
//   * both operands are split x = hi + lo (two bf16) and hi*lo + lo*hi + hi*hi are accumulated: 16 mantissa bits per operand.  Visit
//     counts agree with the Float32 oracle on > 99 % of roots (bf16 alone: 72 %; tests/test_gpu_mma.py);
//   * the 8 warps of the CTA are specialised: per 16-tree tile one warp each for {dynamics trunk + state head, dynamics trunk + reward
//     head, prediction trunk + value head, prediction trunk + policy head} -- the trunks are computed twice so that the two heads of a
//     network run side by side: the chain is 8 layers long instead of 10;
//   * weights (bf16 hi + lo = 4 bytes per weight, 228 KB for prediction + dynamics: more than an SM has) stream from L2 through a
//     three-slot shared-memory ring per network by TMA bulk copies in "fragment order" (mzh::pack_weights_mma): a lane fetches its B
//     fragments of one (n-tile, k-step) -- hi and lo -- with one conflict-free 16-byte load.  The warp that frees a slot last refills it.
// Tree phases, the move epilogue and the data layout are shared with mz_k_search.
#pragma once
#include "mz_kernels.cuh"

#ifdef MZ_PHASE_TIMERS
#define MZ_RT(i) do { if (tk) { long long c_ = clock64(); if ((i) > 0) tk[(i) - 1] += c_ - tprev; tprev = c_; } } while (0)
#else
#define MZ_RT(i)
#endif
#define MZ_MMA_THREADS (MZ_THREADS + 32)   // 8 worker warps + the weight producer warp
#define MZ_MMA_STRIDE 72     // floats per row of a staged network input ([tree][k], k < 64): 72 = 8 mod 32 keeps the 8-byte fragment loads conflict-free

struct mz_search_mma_args {
    mz_search_args base;
    const unsigned char *image;   // fragment-ordered bf16 hi/lo weights (global)
    const float *bias;            // fp32 biases, 64 per layer (global)
    int32_t rows;                 // trees per CTA (<= 32)
    int32_t pbc_in_smem;          // the (S+2)^2 PUCT table fits into shared memory
};

__device__ __forceinline__ void mz_hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    // volatile: the issue order below (eight independent accumulators between two MMAs on the same one) must survive the compiler
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x0, x1) -> packed bf16 hi parts (x0 in the low half) and packed bf16 lo parts (x - hi), both round to nearest even
__device__ __forceinline__ void mz_split2(float x0, float x1, uint32_t &h, uint32_t &l) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
}

// mbarrier wait that backs off: a warp that waits here (for weights it does not even use, or for a slot to drain) must not take issue
// slots from the warps of its SM sub-partition that are computing
__device__ __forceinline__ void mz_mbar_wait_sleep(uint64_t *bar, uint32_t parity, unsigned ns) {
    for (uint32_t spin = 0; !mz_mbar_try_wait(bar, parity); spin++) {
        __nanosleep(ns);
        if (spin > (1u << 22)) __trap();
    }
}
struct mz_mma_ring {
    uint32_t slot0;               // shared address of slot 0
    uint32_t slot_bytes;
    uint64_t *full, *empty;       // [MZ_MMA_SLOTS] each: weights have landed / all consumer warps are done with them
    const unsigned char *image;
    int lead_st, lead_n, loop_st, loop_n, total;
};
#define MZ_MMA_CONSUMERS 4        // warps per ring: 2 tiles x 2 heads
__device__ __forceinline__ int mz_mma_layer_at(const mz_mma_plan &M, const mz_mma_ring &R, int q) {
    return q < R.lead_n ? M.layer[R.lead_st][q] : M.layer[R.loop_st][(q - R.lead_n) % R.loop_n];
}
__device__ __forceinline__ void mz_mma_issue(const mz_mma_plan &M, const mz_mma_ring &R, int q) {
    const int slot = q % MZ_MMA_SLOTS, L = mz_mma_layer_at(M, R, q);
    mz_mbar_expect_tx(&R.full[slot], (uint32_t)M.w_bytes[L]);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(R.slot0 + (uint32_t)slot * R.slot_bytes), "l"(R.image + M.w_off[L]), "r"((uint32_t)M.w_bytes[L]), "r"(mz_smem_u32(&R.full[slot])) : "memory");
}
// The weight producer: one lane per ring, for the whole kernel.  Position q of the ring's layer sequence goes to slot q % SLOTS as soon
// as every consumer warp has released the slot's previous contents.
__device__ __forceinline__ void mz_mma_producer(const mz_mma_plan &M, const mz_mma_ring &R) {
    for (int q = 0; q < R.total; q++) {
        const int slot = q % MZ_MMA_SLOTS;
        if (q >= MZ_MMA_SLOTS) { mz_mbar_wait_sleep(&R.empty[slot], (uint32_t)(q / MZ_MMA_SLOTS - 1) & 1u, 100); mz_fence_proxy_async(); }
        mz_mma_issue(M, R, q);
    }
}

__device__ __forceinline__ float2 mz_lds64f(uint32_t addr) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr)); return v; }
__device__ __forceinline__ void mz_sts32f(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// first layer of a chain: A fragments (hi / lo) of k-steps 0..ks-1 from the staged fp32 input in[row][k] (shared address)
__device__ __forceinline__ void mz_mma_load_input(uint32_t in, int ks, int tile, int lane, uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t r0 = in + (uint32_t)(((tile * 16 + g) * MZ_MMA_STRIDE + 2 * t) * 4), r1 = r0 + 8 * MZ_MMA_STRIDE * 4;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        if (s < ks) {
            const float2 v0 = mz_lds64f(r0 + 64 * s), v1 = mz_lds64f(r1 + 64 * s), v2 = mz_lds64f(r0 + 64 * s + 32), v3 = mz_lds64f(r1 + 64 * s + 32);
            mz_split2(v0.x, v0.y, ah[s][0], al[s][0]); mz_split2(v1.x, v1.y, ah[s][1], al[s][1]);
            mz_split2(v2.x, v2.y, ah[s][2], al[s][2]); mz_split2(v3.x, v3.y, ah[s][3], al[s][3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) { ah[s][i] = 0u; al[s][i] = 0u; }
        }
    }
}
// one k-step: D[j] += X_s W_{j,s}^T for the n-tiles j < nt, with split operands; wl = this lane's B fragments of (n-tile 0, k-step s)
template <bool FULL>
__device__ __forceinline__ void mz_mma_kstep(uint32_t wl, uint32_t jstride, int nt, const uint32_t (&ah)[4], const uint32_t (&al)[4], float (&d)[8][4]) {
    uint4 b[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (FULL || j < nt) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b[j].x), "=r"(b[j].y), "=r"(b[j].z), "=r"(b[j].w) : "r"(wl + (uint32_t)j * jstride));
    // All eight loads are in flight before the first MMA: without this fence ptxas sinks every load next to its three MMAs and reuses one
    // register quad for all n-tiles, which serialises load latency -> MMA -> MMA -> MMA eight times per k-step (3x slower, measured).
    __syncwarp();
    // the small cross terms first, hi * hi on top; eight independent accumulators per pass keep the tensor pipe fed
#pragma unroll
    for (int j = 0; j < 8; j++) if (FULL || j < nt) mz_hmma(d[j], ah, b[j].z, b[j].w);     // x_hi * w_lo
#pragma unroll
    for (int j = 0; j < 8; j++) if (FULL || j < nt) mz_hmma(d[j], al, b[j].x, b[j].y);     // x_lo * w_hi
#pragma unroll
    for (int j = 0; j < 8; j++) if (FULL || j < nt) mz_hmma(d[j], ah, b[j].x, b[j].y);     // x_hi * w_hi
}
// D[16 trees][8 nt features] = X W^T; w = the layer's fragment block in shared memory.  The k-step loop is a real loop (compact code: the
// instruction cache matters more than the 32 moves): each pass uses the fragments in position 0 and rotates the arrays.
__device__ __forceinline__ void mz_mma_layer(uint32_t w, int ks, int nt, int lane, uint32_t (&ah)[4][4], uint32_t (&al)[4][4], float (&d)[8][4]) {
#pragma unroll
    for (int j = 0; j < 8; j++) { d[j][0] = 0.0f; d[j][1] = 0.0f; d[j][2] = 0.0f; d[j][3] = 0.0f; }
    uint32_t wl = w + (uint32_t)lane * 16u;
    const uint32_t jstride = (uint32_t)ks * 512u;
#pragma unroll 1
    for (int s = 0; s < ks; s++, wl += 512u) {
        if (nt == 8) mz_mma_kstep<true>(wl, jstride, 8, ah[0], al[0], d);
        else mz_mma_kstep<false>(wl, jstride, nt, ah[0], al[0], d);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t th = ah[0][i], tl = al[0][i];
            ah[0][i] = ah[1][i]; ah[1][i] = ah[2][i]; ah[2][i] = ah[3][i]; ah[3][i] = th;
            al[0][i] = al[1][i]; al[1][i] = al[2][i]; al[2][i] = al[3][i]; al[3][i] = tl;
        }
    }
}
// hidden layer: bias + activation, then the accumulator fragments become the next layer's A fragments (hi / lo) in registers
__device__ __forceinline__ void mz_mma_to_frags(const float (&d)[8][4], uint32_t bias, int nt, int act, int lane, uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int s = j >> 1, h = (j & 1) * 2;
        if (j < nt) {
            const float2 b = mz_lds64f(bias + (uint32_t)((8 * j + 2 * t) * 4));
            float x0 = d[j][0] + b.x, x1 = d[j][1] + b.y, x2 = d[j][2] + b.x, x3 = d[j][3] + b.y;
            if (act == MZ_ACT_RELU) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); x2 = fmaxf(x2, 0.0f); x3 = fmaxf(x3, 0.0f); }
            else if (act == MZ_ACT_TANH) { x0 = mz_tanhf_ni(x0); x1 = mz_tanhf_ni(x1); x2 = mz_tanhf_ni(x2); x3 = mz_tanhf_ni(x3); }
            mz_split2(x0, x1, ah[s][h], al[s][h]);            // row g
            mz_split2(x2, x3, ah[s][h + 1], al[s][h + 1]);    // row g + 8
        } else { ah[s][h] = 0u; al[s][h] = 0u; ah[s][h + 1] = 0u; al[s][h + 1] = 0u; }
    }
}
// last layer of a chain: fp32 outputs out[feature * MZ_ROWS + tree] (shared address; the layout the tree phases read)
__device__ __noinline__ void mz_mma_to_output(const float (&d)[8][4], uint32_t bias, int nt, int out_n, int act, int tile, int lane, uint32_t out) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < nt) {
            const float2 b = mz_lds64f(bias + (uint32_t)((8 * j + 2 * t) * 4));
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int f = 8 * j + 2 * t + e;
                if (f < out_n) {
                    float v0 = d[j][e] + (e ? b.y : b.x), v1 = d[j][2 + e] + (e ? b.y : b.x);
                    if (act == MZ_ACT_RELU) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
                    else if (act == MZ_ACT_TANH) { v0 = mz_tanhf_ni(v0); v1 = mz_tanhf_ni(v1); }
                    mz_sts32f(out + (uint32_t)((f * MZ_ROWS + tile * 16 + g) * 4), v0); mz_sts32f(out + (uint32_t)((f * MZ_ROWS + tile * 16 + g + 8) * 4), v1);
                }
            }
        }
    }
}

// One pass over stream `st` (a whole network) for one warp: head = 0 / 1 selects the chain (trunk + that head) this warp computes.
// Every consumer warp of the ring visits every stream position: waits for the weights, computes if the entry is its own, releases the slot.
__device__ __noinline__ int mz_mma_run(const mz_params &P, const mz_mma_plan &M, const mz_mma_ring R, int st, int pos, int head, uint32_t in,
                                       uint32_t out, uint32_t bias, int tile, int lane, long long *tk = nullptr) {
    long long tprev = 0; (void)tprev; (void)tk;
    uint32_t ah[4][4], al[4][4];
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
        for (int i = 0; i < 4; i++) { ah[s][i] = 0u; al[s][i] = 0u; }
    const int n = M.n[st];
    int slot = pos % MZ_MMA_SLOTS;
    for (int e = 0; e < n; e++, pos++) {
        MZ_RT(0);
        mz_mbar_wait_sleep(&R.full[slot], (uint32_t)(pos / MZ_MMA_SLOTS) & 1u, ((M.use[st][e] >> head) & 1) ? 20 : 200);
        MZ_RT(1);
        if ((M.use[st][e] >> head) & 1) {
            const int L = M.layer[st][e], ks = M.ks[L], nt = M.nt[L];
            if ((M.first[st][e] >> head) & 1) mz_mma_load_input(in, ks, tile, lane, ah, al);
            MZ_RT(2);
            float d[8][4];
            mz_mma_layer(R.slot0 + (uint32_t)slot * R.slot_bytes, ks, nt, lane, ah, al, d);
            MZ_RT(3);
            if ((M.last[st][e] >> head) & 1) mz_mma_to_output(d, bias + (uint32_t)L * 256u, nt, P.layers[L].out, P.layers[L].act, tile, lane, out);
            else mz_mma_to_frags(d, bias + (uint32_t)L * 256u, nt, P.layers[L].act, lane, ah, al);
            MZ_RT(4);
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mz_smem_u32(&R.empty[slot])) : "memory");
        slot = slot + 1 == MZ_MMA_SLOTS ? 0 : slot + 1;
        MZ_RT(5);
    }
    return pos;
}

struct mz_mma_smem {
    uint32_t slot0[2]; uint64_t *full[2], *empty[2];
    float *bias, *in0, *in1, *outV, *outL, *outR, *outH;
    double *pbc; uint16_t *path;
};
__host__ __device__ inline size_t mz_mma_smem_bytes(int slot_bytes, int bias_floats, int hidden_pad, int S, int pbc_in_smem) {
    size_t ring = 2 * (size_t)MZ_MMA_SLOTS * (size_t)slot_bytes;
    size_t bias = ((size_t)bias_floats * 4 + 127) & ~(size_t)127;
    size_t in = 2 * (size_t)MZ_ROWS * MZ_MMA_STRIDE * 4;
    size_t small = (size_t)(4 + 16 + 4) * MZ_ROWS * 4 + (size_t)hidden_pad * MZ_ROWS * 4;
    size_t path = (((size_t)S + 2) * 2 * MZ_ROWS + 127) & ~(size_t)127;
    size_t pbc = pbc_in_smem ? ((((size_t)S + 2) * ((size_t)S + 2) * 8 + 127) & ~(size_t)127) : 0;
    return 128 + ring + 256 + bias + in + small + path + pbc;
}
__device__ __forceinline__ mz_mma_smem mz_mma_carve(unsigned char *raw, int slot_bytes, int bias_floats, int hidden_pad, int S, int pbc_in_smem) {
    mz_mma_smem p;
    const uint32_t a = mz_smem_u32(raw);
    unsigned char *c = raw + (((a + 127u) & ~127u) - a);
    for (int r = 0; r < 2; r++) { p.slot0[r] = mz_smem_u32(c); c += (size_t)MZ_MMA_SLOTS * slot_bytes; }
    for (int r = 0; r < 2; r++) { p.full[r] = (uint64_t *)(c + 32 * r); p.empty[r] = (uint64_t *)(c + 128 + 32 * r); }
    c += 256;
    p.bias = (float *)c; c += ((size_t)bias_floats * 4 + 127) & ~(size_t)127;
    p.in0 = (float *)c; c += (size_t)MZ_ROWS * MZ_MMA_STRIDE * 4;
    p.in1 = (float *)c; c += (size_t)MZ_ROWS * MZ_MMA_STRIDE * 4;
    p.outV = (float *)c; c += 4 * MZ_ROWS * 4;
    p.outL = (float *)c; c += 16 * MZ_ROWS * 4;
    p.outR = (float *)c; c += 4 * MZ_ROWS * 4;
    p.outH = (float *)c; c += (size_t)hidden_pad * MZ_ROWS * 4;
    p.path = (uint16_t *)c; c += (((size_t)S + 2) * 2 * MZ_ROWS + 127) & ~(size_t)127;
    p.pbc = (double *)c;
    return p;
}

// common set-up of a kernel that runs networks on this path: barriers, counters, biases, zeroed inputs; returns after a CTA barrier
__device__ __forceinline__ void mz_mma_setup(const mz_mma_smem &sp, const mz_mma_plan &M, const float *bias_glob, int nthreads) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int r = 0; r < 2; r++) for (int i = 0; i < MZ_MMA_SLOTS; i++) { mz_mbar_init(&sp.full[r][i], 1); mz_mbar_init(&sp.empty[r][i], MZ_MMA_CONSUMERS); }
        mz_fence_mbar_init();
    }
    for (int i = tid; i < M.bias_floats; i += nthreads) sp.bias[i] = bias_glob[i];
    for (int i = tid; i < 2 * MZ_ROWS * MZ_MMA_STRIDE; i += nthreads) sp.in0[i] = 0.0f;     // in0 and in1 are contiguous
    __syncthreads();
}

// barrier over the 256 worker threads (the producer warp never joins it)
__device__ __forceinline__ void mz_mma_workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(MZ_MMA_THREADS) mz_k_search_mma(const __grid_constant__ mz_params P, const __grid_constant__ mz_mma_plan M, const mz_search_mma_args ma) {
    extern __shared__ __align__(128) unsigned char mz_smem_mma[];
    const mz_search_args &a = ma.base;
    const int R = ma.rows;
    if (mz_cta_idle<MODE>(P, a, R)) return;
    const mz_mma_smem sp = mz_mma_carve(mz_smem_mma, M.slot_bytes, M.bias_floats, P.hidden_pad, P.S, ma.pbc_in_smem);
    const int tid = threadIdx.x;
    const int r = tid >> 3, ln = tid & (MZ_LANES - 1);
    const uint32_t segmask = 0xffu << ((tid & 31) & ~7);
    const int64_t g = (int64_t)blockIdx.x * R + r;
    mz_mma_setup(sp, M, ma.bias, MZ_MMA_THREADS);
    // warp roles in the network phases: tile = 16-tree tile; role 0 / 1: dynamics (and, at the root, representation) chain of head 1 / 2,
    // role 2 / 3: prediction chain of head 1 / 2.  Roles 0 and 2 (the long chains) sit on different SM sub-partitions than 1 and 3.
    const int warp = tid >> 5, lane = tid & 31, tile = warp & 1, role = warp >> 1, head = role & 1;
    int ring_id = role >> 1;
    mz_mma_ring ring;
    if (warp == 8) { if (lane >= 2) return; ring_id = lane; }            // the producer warp: lane r feeds ring r for the whole kernel
    ring.slot0 = sp.slot0[ring_id]; ring.slot_bytes = (uint32_t)M.slot_bytes; ring.full = sp.full[ring_id]; ring.empty = sp.empty[ring_id]; ring.image = ma.image;
    if (ring_id == 0) { ring.lead_st = 0; ring.lead_n = M.n[0]; ring.loop_st = 2; ring.loop_n = M.n[2]; ring.total = M.n[0] + P.S * M.n[2]; }
    else { ring.lead_st = 1; ring.lead_n = M.n[1]; ring.loop_st = 1; ring.loop_n = M.n[1]; ring.total = (P.S + 1) * M.n[1]; }
    if (warp == 8) { mz_mma_producer(M, ring); return; }
    int pos = 0;
    long long *tk = nullptr;
#ifdef MZ_PHASE_TIMERS
    long long rt[6] = {0, 0, 0, 0, 0, 0};
    if (tid == 0 || tid == 64) tk = rt;                                  // observers: lane 0 of the dynamics warps of tile 0 (state head / reward head)
#endif
    const double *pbc = a.pbc0;
    if (ma.pbc_in_smem) {
        const int n = (P.S + 2) * (P.S + 2);
        for (int i = tid; i < n; i += MZ_THREADS) sp.pbc[i] = a.pbc0[i];
        pbc = sp.pbc;
    }
    uint16_t *path = sp.path + (size_t)r * (P.S + 2);

    // ---- per-tree state, replicated in the 8 lanes of the tree ----
    bool active = false; uint32_t legal = 0, game = 0, move = 0; int to_play = 1;
    mz_tree tree; tree.A = nullptr; tree.hidden = nullptr;
    if (r < R && g < a.n) {
        tree = mz_tree_at(P, a.tree_pool, g);
        if (MODE == MZ_MODE_API) {
            active = true; legal = a.legal[g]; to_play = a.to_play[g]; game = (uint32_t)a.game_id[g]; move = (uint32_t)a.move_idx[g];
        } else {
            active = a.slots.status[g] == MZ_SLOT_ACTIVE && (P.arena_player == 0 || a.slots.player[g] == P.arena_player);
            if (active) {
                mz_board b; b.p1 = a.slots.p1[g]; b.p2 = a.slots.p2[g]; b.player = a.slots.player[g];
                legal = mz_env_legal_b(P, b); to_play = b.player;
                game = (uint32_t)a.slots.game_id[g]; move = (uint32_t)a.slots.T[g] + 1u;
            }
        }
        if (legal == 0) active = false;
    }
    (void)to_play;
    uint32_t posmask = 0;
    for (int j = 0; j < P.A; j++) if ((legal >> (P.order[j] - 1)) & 1u) posmask |= 1u << j;

    // ---- stage the stacked observations, row-major [tree][k] ----
    for (int i = tid; i < R * P.stack_size; i += MZ_THREADS) {
        float v = 0.0f; int rr, k;
        if (MODE == MZ_MODE_API) {
            rr = i / P.stack_size; k = i % P.stack_size;
            const int64_t gg = (int64_t)blockIdx.x * R + rr;
            if (gg < a.n) v = a.stacked[gg * P.stack_size + k];
        } else {
            k = i / R; rr = i % R;
            const int64_t gg = (int64_t)blockIdx.x * R + rr;
            if (gg < a.n && a.slots.status[gg] == MZ_SLOT_ACTIVE)
                v = mz_stacked_value(P, a.slots.h_p1 + gg * P.Tmax, a.slots.h_p2 + gg * P.Tmax, a.slots.h_action + gg * P.Tmax, a.slots.T[gg] + 1, k);
        }
        sp.in0[rr * MZ_MMA_STRIDE + k] = v;
    }
    mz_mma_workers_sync();

    // ---- root: representation -> h0; prediction(h0) -> (v0, p0) ----
    if (ring_id == 0) pos = mz_mma_run(P, M, ring, 0, pos, head, mz_smem_u32(sp.in0), mz_smem_u32(sp.outH), mz_smem_u32(sp.bias), tile, lane);
    mz_mma_workers_sync();
    for (int i = tid; i < R * P.hidden; i += MZ_THREADS) { const int k = i / R, rr = i % R; sp.in1[rr * MZ_MMA_STRIDE + k] = sp.outH[k * MZ_ROWS + rr]; }
    mz_mma_workers_sync();
    if (ring_id == 1) pos = mz_mma_run(P, M, ring, 1, pos, head, mz_smem_u32(sp.in1), mz_smem_u32(head == 0 ? sp.outV : sp.outL), mz_smem_u32(sp.bias), tile, lane);
    mz_mma_workers_sync();

    mz_minmax mm; mm.mn = INFINITY; mm.mx = -INFINITY;
    unsigned long long depth_sum = 0;
    MZ_TIMER_DECL;
    if (active) {
        for (int k = ln; k < P.hidden; k += MZ_LANES) tree.hidden[k] = sp.outH[k * MZ_ROWS + r];
        if (ln == 0) {
            mz_f4 root; root.x = mz_bits2f(mz_nx_pack(0, -1, 0)); root.y = 0.0f; root.z = 0.0f; root.w = 0.0f;
            tree.A[0] = root;
        }
        __syncwarp(segmask);
        mz_tree_expand_lanes(P, tree, 0, 0, legal, sp.outL + r, 0.0f, 0.0f, ln, segmask);
        if (ln == 0 && a.exploration && P.exploration_eps != 0.0f) mz_tree_add_noise(P, tree, legal, game, move);
        __syncwarp(segmask);
    }

    // ---- simulations ----
    for (int sim = 1; sim <= P.S; sim++) {
        mz_leaf leaf; leaf.node = 0; leaf.parent = 0; leaf.action = 1; leaf.depth = 0; leaf.prior = 0.0f; leaf.parent_x = 0;
        MZ_TIMER(0);
        if (active) {
            leaf = mz_tree_select_lanes(P, tree, pbc, a.sqrtN, legal, posmask, mm, game, move, (uint32_t)sim, ln, segmask, path);
            depth_sum += (unsigned long long)leaf.depth;
            MZ_TIMER(1);
            const int pe = mz_nx_exp(leaf.parent_x), dbl = mz_nx_dbl(leaf.parent_x);
            const float *h = tree.hidden + (size_t)pe * P.hidden_pad;
            const float sc = mz_bits2f((uint32_t)(127 + dbl) << 23);
            for (int k = ln; k < P.hidden; k += MZ_LANES) {
                const float v = h[k] * sc;
                sp.in1[r * MZ_MMA_STRIDE + k] = v;                 // prediction(parent.hidden_state) (Q5)
                sp.in0[r * MZ_MMA_STRIDE + k] = v * 2.0f;          // make_state_action: state .*= 2 (Q6)
            }
            const float plane = P.act_plane_play[leaf.action];
            for (int k = P.obs_size + ln; k < P.sa_size; k += MZ_LANES) sp.in0[r * MZ_MMA_STRIDE + k] = plane;
            __syncwarp(segmask);
            if (ln == 0) reinterpret_cast<uint32_t *>(&tree.A[leaf.parent])[0] = leaf.parent_x + (1u << 24);   // one more doubling (Q6)
        }
        MZ_TIMER(2);
        mz_mma_workers_sync();
        MZ_TIMER(3);
        pos = mz_mma_run(P, M, ring, ring_id == 0 ? 2 : 1, pos, head, mz_smem_u32(ring_id == 0 ? sp.in0 : sp.in1),
                         mz_smem_u32(ring_id == 0 ? (head == 0 ? sp.outH : sp.outR) : (head == 0 ? sp.outV : sp.outL)), mz_smem_u32(sp.bias), tile, lane, tk);
        MZ_TIMER(4);
        mz_mma_workers_sync();
        MZ_TIMER(5);
        if (active) {
            float *nh = tree.hidden + (size_t)sim * P.hidden_pad;
            for (int k = ln; k < P.hidden; k += MZ_LANES) nh[k] = sp.outH[k * MZ_ROWS + r];
            MZ_TIMER(6);
            mz_tree_expand_lanes(P, tree, leaf.node, sim, legal, sp.outL + r, sp.outR[r], leaf.prior, ln, segmask);
            MZ_TIMER(7);
            mz_tree_backup_lanes(P, tree, path, leaf.depth, sp.outV[r], mm, ln, segmask);
        }
        MZ_TIMER(8);
    }

    MZ_TIMER_FLUSH(a.stats);
#ifdef MZ_PHASE_TIMERS
    if (a.stats && tk) { int o_ = tid == 0 ? 16 : 24; for (int i_ = 0; i_ < 5; i_++) atomicAdd(&a.stats[o_ + i_], (unsigned long long)rt[i_]); atomicAdd(&a.stats[o_ + 6], (unsigned long long)(P.S)); }
#endif
    // ---- results (lane 0 of each tree), identical to mz_k_search ----
    if (active && ln == 0) {
        int32_t vc[MZ_MAX_A]; int sum_visits = 0, nlegal = 0;
        for (int i = 0; i < P.A; i++) {
            vc[i] = ((legal >> i) & 1u) ? (int32_t)mz_nx_visit(mz_f2bits(tree.A[1 + i].x)) : 0;
            sum_visits += vc[i]; nlegal += (int)((legal >> i) & 1u);
        }
        mz_f4 root = tree.A[0];
        const int rvc = mz_nx_visit(mz_f2bits(root.x));
        const float rv = rvc == 0 ? 0.0f : root.y / (float)rvc;
        if (a.stats) {
            atomicAdd(&a.stats[0], depth_sum); atomicAdd(&a.stats[1], (unsigned long long)P.S);
            atomicAdd(&a.stats[2], (unsigned long long)nlegal); atomicAdd(&a.stats[3], 1ull);
        }
        if (MODE == MZ_MODE_API) {
            for (int i = 0; i < P.A; i++) {
                a.visit_counts[g * P.A + i] = vc[i];
                if (a.root_priors) a.root_priors[g * P.A + i] = ((legal >> i) & 1u) ? tree.A[1 + i].z : 0.0f;
            }
            a.root_value[g] = rv;
        } else mz_slot_epilogue(P, a.slots, g, vc, sum_visits, legal, rv, a.temperature, game, move);
    }
}

// batched network callable on this path (network outputs against the oracle's split-precision emulation and the Float32 oracle)
struct mz_nn_mma_args { const unsigned char *image; const float *bias; int32_t B, net; const float *in; float *out1; float *out2; int32_t repeat; };
__global__ void __launch_bounds__(160) mz_k_nn_forward_mma(const __grid_constant__ mz_params P, const __grid_constant__ mz_mma_plan M, const mz_nn_mma_args a) {
    extern __shared__ __align__(128) unsigned char mz_smem_mma[];
    const mz_mma_smem sp = mz_mma_carve(mz_smem_mma, M.slot_bytes, M.bias_floats, P.hidden_pad, P.S, 0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, tile = warp & 1, head = warp >> 1;
    mz_mma_setup(sp, M, a.bias, 160);
    mz_mma_ring ring;
    ring.slot0 = sp.slot0[0]; ring.slot_bytes = (uint32_t)M.slot_bytes; ring.full = sp.full[0]; ring.empty = sp.empty[0]; ring.image = a.image;
    ring.lead_st = a.net; ring.lead_n = M.n[a.net]; ring.loop_st = a.net; ring.loop_n = M.n[a.net]; ring.total = M.n[a.net] * (a.repeat > 1 ? a.repeat : 1);
    if (warp == 4) { if (lane == 0) mz_mma_producer(M, ring); return; }
    const int in = P.layers[P.nets[a.net].first].in;
    for (int i = tid; i < MZ_ROWS * in; i += 128) {
        const int rr = i / in, k = i % in;
        const int64_t gg = (int64_t)blockIdx.x * MZ_ROWS + rr;
        sp.in0[rr * MZ_MMA_STRIDE + k] = gg < a.B ? a.in[gg * in + k] : 0.0f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float *h1 = a.net == 1 ? sp.outV : sp.outH, *h2 = a.net == 1 ? sp.outL : sp.outR;
    int pos = 0;
    for (int it = 0; it < (a.repeat > 1 ? a.repeat : 1); it++)        // repeat > 1: timing probe (profiles/nn_rate.py), same result every pass
        pos = mz_mma_run(P, M, ring, a.net, pos, head, mz_smem_u32(sp.in0), mz_smem_u32(head == 0 ? h1 : h2), mz_smem_u32(sp.bias), tile, lane);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int64_t g = (int64_t)blockIdx.x * MZ_ROWS + tid;
    if (tid < MZ_ROWS && g < a.B) {
        if (a.net == 1) {
            float logits[MZ_MAX_A], policy[MZ_MAX_A];
            for (int i = 0; i < P.A; i++) logits[i] = sp.outL[i * MZ_ROWS + tid];
            mz_softmax(logits, P.A, policy);
            a.out1[g] = sp.outV[tid];
            for (int i = 0; i < P.A; i++) a.out2[g * P.A + i] = policy[i];
        } else {
            for (int k = 0; k < P.hidden; k++) a.out1[g * P.hidden + k] = sp.outH[k * MZ_ROWS + tid];
            if (a.net == 2) a.out2[g] = sp.outR[tid];
        }
    }
}
