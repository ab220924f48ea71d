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

#define MZ_MMA_STRIDE 72     // floats per row of a staged network input ([tree][k], k < 64): 72 = 8 mod 32 keeps the 8-byte fragment loads conflict-free

struct mz_search_mma_args {
    mz_search_args base;
    const unsigned char *image;   // fragment-ordered bf16 hi/lo weights (global)
    const float *bias;            // fp32 biases, 64 per layer (global)
    int32_t rows;                 // trees per CTA (<= 32)
    int32_t pbc_in_smem;          // the (S+2)^2 PUCT table fits into shared memory
};

__device__ __forceinline__ void mz_hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x0, x1) -> packed bf16 hi parts (x0 in the low half) and packed bf16 lo parts (x - hi), both round to nearest even
__device__ __forceinline__ void mz_split2(float x0, float x1, uint32_t &h, uint32_t &l) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
}

struct mz_mma_ring {
    uint32_t slot0;               // shared address of slot 0
    uint32_t slot_bytes;
    uint64_t *full;               // [MZ_MMA_SLOTS]
    int *cnt;                     // [MZ_MMA_SLOTS] warps that are done with the slot's current contents
    const unsigned char *image;
    int lead_st, lead_n, loop_st, loop_n, total, nwarps;
};
__device__ __forceinline__ int mz_mma_layer_at(const mz_mma_plan &M, const mz_mma_ring &R, int q) {
    return q < R.lead_n ? M.layer[R.lead_st][q] : M.layer[R.loop_st][(q - R.lead_n) % R.loop_n];
}
__device__ __forceinline__ void mz_mma_issue(const mz_mma_plan &M, const mz_mma_ring &R, int q) {
    const int slot = q % MZ_MMA_SLOTS, L = mz_mma_layer_at(M, R, q);
    mz_mbar_expect_tx(&R.full[slot], (uint32_t)M.w_bytes[L]);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(R.slot0 + (uint32_t)slot * R.slot_bytes), "l"(R.image + M.w_off[L]), "r"((uint32_t)M.w_bytes[L]), "r"(mz_smem_u32(&R.full[slot])) : "memory");
}

// first layer of a chain: A fragments (hi / lo) of k-steps 0..ks-1 from the staged fp32 input in[row][k]
__device__ __forceinline__ void mz_mma_load_input(const float *in, int ks, int tile, int lane, uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
    const int g = lane >> 2, t = lane & 3;
    const float *r0 = in + (size_t)(tile * 16 + g) * MZ_MMA_STRIDE + 2 * t, *r1 = r0 + 8 * MZ_MMA_STRIDE;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        if (s < ks) {
            const float2 v0 = *reinterpret_cast<const float2 *>(r0 + 16 * s), v1 = *reinterpret_cast<const float2 *>(r1 + 16 * s);
            const float2 v2 = *reinterpret_cast<const float2 *>(r0 + 16 * s + 8), v3 = *reinterpret_cast<const float2 *>(r1 + 16 * s + 8);
            mz_split2(v0.x, v0.y, ah[s][0], al[s][0]); mz_split2(v1.x, v1.y, ah[s][1], al[s][1]);
            mz_split2(v2.x, v2.y, ah[s][2], al[s][2]); mz_split2(v3.x, v3.y, ah[s][3], al[s][3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) { ah[s][i] = 0u; al[s][i] = 0u; }
        }
    }
}
// D[16 trees][8 nt features] = X W^T with split operands; w = the layer's fragment block in shared memory
__device__ __forceinline__ void mz_mma_layer(uint32_t w, int ks, int nt, int lane, const uint32_t (&ah)[4][4], const uint32_t (&al)[4][4], float (&d)[8][4]) {
#pragma unroll
    for (int j = 0; j < 8; j++) { d[j][0] = 0.0f; d[j][1] = 0.0f; d[j][2] = 0.0f; d[j][3] = 0.0f; }
    const uint32_t wl = w + (uint32_t)lane * 16u;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        if (s < ks) {
            uint4 b[8];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < nt) asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b[j].x), "=r"(b[j].y), "=r"(b[j].z), "=r"(b[j].w) : "r"(wl + (uint32_t)((j * ks + s) * 512)));
            // the small cross terms first, hi * hi on top; eight independent accumulators per pass keep the tensor pipe fed
#pragma unroll
            for (int j = 0; j < 8; j++) if (j < nt) mz_hmma(d[j], ah[s], b[j].z, b[j].w);     // x_hi * w_lo
#pragma unroll
            for (int j = 0; j < 8; j++) if (j < nt) mz_hmma(d[j], al[s], b[j].x, b[j].y);     // x_lo * w_hi
#pragma unroll
            for (int j = 0; j < 8; j++) if (j < nt) mz_hmma(d[j], ah[s], b[j].x, b[j].y);     // x_hi * w_hi
        }
    }
}
// hidden layer: bias + activation, then the accumulator fragments become the next layer's A fragments (hi / lo) in registers
__device__ __forceinline__ void mz_mma_to_frags(const float (&d)[8][4], const float *bias, int nt, int act, int lane, uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int s = j >> 1, h = (j & 1) * 2;
        if (j < nt) {
            const float2 b = *reinterpret_cast<const float2 *>(bias + 8 * j + 2 * t);
            float x0 = d[j][0] + b.x, x1 = d[j][1] + b.y, x2 = d[j][2] + b.x, x3 = d[j][3] + b.y;
            if (act == MZ_ACT_RELU) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); x2 = fmaxf(x2, 0.0f); x3 = fmaxf(x3, 0.0f); }
            else if (act == MZ_ACT_TANH) { x0 = mz_tanhf_ni(x0); x1 = mz_tanhf_ni(x1); x2 = mz_tanhf_ni(x2); x3 = mz_tanhf_ni(x3); }
            mz_split2(x0, x1, ah[s][h], al[s][h]);            // row g
            mz_split2(x2, x3, ah[s][h + 1], al[s][h + 1]);    // row g + 8
        } else { ah[s][h] = 0u; al[s][h] = 0u; ah[s][h + 1] = 0u; al[s][h + 1] = 0u; }
    }
}
// last layer of a chain: fp32 outputs out[feature * MZ_ROWS + tree] (the layout the tree phases read)
__device__ __forceinline__ void mz_mma_to_output(const float (&d)[8][4], const float *bias, int nt, int out_n, int act, int tile, int lane, float *out) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < nt) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int f = 8 * j + 2 * t + e;
                if (f < out_n) {
                    float v0 = d[j][e] + bias[f], v1 = d[j][2 + e] + bias[f];
                    if (act == MZ_ACT_RELU) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
                    else if (act == MZ_ACT_TANH) { v0 = mz_tanhf_ni(v0); v1 = mz_tanhf_ni(v1); }
                    out[f * MZ_ROWS + tile * 16 + g] = v0; out[f * MZ_ROWS + tile * 16 + g + 8] = v1;
                }
            }
        }
    }
}

// One pass over stream `st` (a whole network) for one warp: head = 0 / 1 selects the chain (trunk + that head) this warp computes.
// Every warp of the ring visits every stream position -- waits for its weights, computes if the entry is its own, and reports the slot
// free; the warp that reports last refills the slot with the layer MZ_MMA_SLOTS positions ahead.
__device__ __noinline__ void mz_mma_run(const mz_params &P, const mz_mma_plan &M, const mz_mma_ring &R, int st, int &pos, int head, const float *in,
                                        float *out, const float *bias, int tile, int lane) {
    uint32_t ah[4][4], al[4][4];
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
        for (int i = 0; i < 4; i++) { ah[s][i] = 0u; al[s][i] = 0u; }
    const int n = M.n[st];
    int slot = pos % MZ_MMA_SLOTS;
    for (int e = 0; e < n; e++, pos++) {
        mz_mbar_wait(&R.full[slot], (uint32_t)(pos / MZ_MMA_SLOTS) & 1u);
        if ((M.use[st][e] >> head) & 1) {
            const int L = M.layer[st][e], ks = M.ks[L], nt = M.nt[L];
            if ((M.first[st][e] >> head) & 1) mz_mma_load_input(in, ks, tile, lane, ah, al);
            float d[8][4];
            mz_mma_layer(R.slot0 + (uint32_t)slot * R.slot_bytes, ks, nt, lane, ah, al, d);
            if ((M.last[st][e] >> head) & 1) mz_mma_to_output(d, bias + L * 64, nt, P.layers[L].out, P.layers[L].act, tile, lane, out);
            else mz_mma_to_frags(d, bias + L * 64, nt, P.layers[L].act, lane, ah, al);
        }
        __syncwarp();
        if (lane == 0) {
            const int old = atomicAdd(&R.cnt[slot], 1);
            if (old == R.nwarps - 1) {                       // everybody is done with this slot
                R.cnt[slot] = 0;
                __threadfence_block();
                if (pos + MZ_MMA_SLOTS < R.total) { mz_fence_proxy_async(); mz_mma_issue(M, R, pos + MZ_MMA_SLOTS); }
            }
        }
        slot = slot + 1 == MZ_MMA_SLOTS ? 0 : slot + 1;
    }
}

struct mz_mma_smem {
    uint32_t slot0[2]; uint64_t *full[2]; int *cnt[2];
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
    for (int r = 0; r < 2; r++) { p.full[r] = (uint64_t *)(c + 32 * r); p.cnt[r] = (int *)(c + 128 + 16 * r); }
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
        for (int r = 0; r < 2; r++) for (int i = 0; i < MZ_MMA_SLOTS; i++) { mz_mbar_init(&sp.full[r][i], 1); sp.cnt[r][i] = 0; }
        mz_fence_mbar_init();
    }
    for (int i = tid; i < M.bias_floats; i += nthreads) sp.bias[i] = bias_glob[i];
    for (int i = tid; i < 2 * MZ_ROWS * MZ_MMA_STRIDE; i += nthreads) sp.in0[i] = 0.0f;     // in0 and in1 are contiguous
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(MZ_THREADS) mz_k_search_mma(const __grid_constant__ mz_params P, const __grid_constant__ mz_mma_plan M, const mz_search_mma_args ma) {
    extern __shared__ __align__(128) unsigned char mz_smem_mma[];
    const mz_search_args &a = ma.base;
    const int R = ma.rows;
    if (mz_cta_idle<MODE>(P, a, R)) return;
    const mz_mma_smem sp = mz_mma_carve(mz_smem_mma, M.slot_bytes, M.bias_floats, P.hidden_pad, P.S, ma.pbc_in_smem);
    const int tid = threadIdx.x;
    const int r = tid >> 3, ln = tid & (MZ_LANES - 1);
    const uint32_t segmask = 0xffu << ((tid & 31) & ~7);
    const int64_t g = (int64_t)blockIdx.x * R + r;
    mz_mma_setup(sp, M, ma.bias, MZ_THREADS);
    // warp roles in the network phases: tile = 16-tree tile; role 0 / 1: dynamics (and, at the root, representation) chain of head 1 / 2,
    // role 2 / 3: prediction chain of head 1 / 2.  Roles 0 and 2 (the long chains) sit on different SM sub-partitions than 1 and 3.
    const int warp = tid >> 5, lane = tid & 31, tile = warp & 1, role = warp >> 1, ring_id = role >> 1, head = role & 1;
    mz_mma_ring ring;
    ring.slot0 = sp.slot0[ring_id]; ring.slot_bytes = (uint32_t)M.slot_bytes; ring.full = sp.full[ring_id]; ring.cnt = sp.cnt[ring_id]; ring.image = ma.image;
    ring.nwarps = 4;
    if (ring_id == 0) { ring.lead_st = 0; ring.lead_n = M.n[0]; ring.loop_st = 2; ring.loop_n = M.n[2]; ring.total = M.n[0] + P.S * M.n[2]; }
    else { ring.lead_st = 1; ring.lead_n = M.n[1]; ring.loop_st = 1; ring.loop_n = M.n[1]; ring.total = (P.S + 1) * M.n[1]; }
    if (lane == 0 && tile == 0 && head == 0)
        for (int q = 0; q < MZ_MMA_SLOTS && q < ring.total; q++) mz_mma_issue(M, ring, q);
    int pos = 0;
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
    __syncthreads();

    // ---- root: representation -> h0; prediction(h0) -> (v0, p0) ----
    if (ring_id == 0) mz_mma_run(P, M, ring, 0, pos, head, sp.in0, sp.outH, sp.bias, tile, lane);
    __syncthreads();
    for (int i = tid; i < R * P.hidden; i += MZ_THREADS) { const int k = i / R, rr = i % R; sp.in1[rr * MZ_MMA_STRIDE + k] = sp.outH[k * MZ_ROWS + rr]; }
    __syncthreads();
    if (ring_id == 1) mz_mma_run(P, M, ring, 1, pos, head, sp.in1, head == 0 ? sp.outV : sp.outL, sp.bias, tile, lane);
    __syncthreads();

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
        __syncthreads();
        MZ_TIMER(3);
        if (ring_id == 0) mz_mma_run(P, M, ring, 2, pos, head, sp.in0, head == 0 ? sp.outH : sp.outR, sp.bias, tile, lane);
        else              mz_mma_run(P, M, ring, 1, pos, head, sp.in1, head == 0 ? sp.outV : sp.outL, sp.bias, tile, lane);
        MZ_TIMER(4);
        __syncthreads();
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
struct mz_nn_mma_args { const unsigned char *image; const float *bias; int32_t B, net; const float *in; float *out1; float *out2; };
__global__ void __launch_bounds__(128) mz_k_nn_forward_mma(const __grid_constant__ mz_params P, const __grid_constant__ mz_mma_plan M, const mz_nn_mma_args a) {
    extern __shared__ __align__(128) unsigned char mz_smem_mma[];
    const mz_mma_smem sp = mz_mma_carve(mz_smem_mma, M.slot_bytes, M.bias_floats, P.hidden_pad, P.S, 0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, tile = warp & 1, head = warp >> 1;
    mz_mma_setup(sp, M, a.bias, 128);
    mz_mma_ring ring;
    ring.slot0 = sp.slot0[0]; ring.slot_bytes = (uint32_t)M.slot_bytes; ring.full = sp.full[0]; ring.cnt = sp.cnt[0]; ring.image = a.image; ring.nwarps = 4;
    ring.lead_st = a.net; ring.lead_n = M.n[a.net]; ring.loop_st = a.net; ring.loop_n = M.n[a.net]; ring.total = M.n[a.net];
    if (tid == 0) for (int q = 0; q < MZ_MMA_SLOTS && q < ring.total; q++) mz_mma_issue(M, ring, q);
    const int in = P.layers[P.nets[a.net].first].in;
    for (int i = tid; i < MZ_ROWS * in; i += 128) {
        const int rr = i / in, k = i % in;
        const int64_t gg = (int64_t)blockIdx.x * MZ_ROWS + rr;
        sp.in0[rr * MZ_MMA_STRIDE + k] = gg < a.B ? a.in[gg * in + k] : 0.0f;
    }
    __syncthreads();
    float *h1 = a.net == 1 ? sp.outV : sp.outH, *h2 = a.net == 1 ? sp.outL : sp.outR;
    int pos = 0;
    mz_mma_run(P, M, ring, a.net, pos, head, sp.in0, head == 0 ? h1 : h2, sp.bias, tile, lane);
    __syncthreads();
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
