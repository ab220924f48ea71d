// mz_learner_tc.cuh -- grad_mode = MZ_GRAD_BPTT on the tensor cores (nn_mode = MZ_NN_SPLIT_MMA): the K-step unroll forward
// (src/Learning.jl:347-370) and the backward pass through it as tcgen05.mma rounds.  Plan: mz_common.h (mz_lr_plan), mz_host.h.
//
//   mz_k_learn_bptt_tc   32 samples per CTA.  Forward = the rounds of mz_k_learn_forward_sp (split precision: the predictions and the
//                        losses are the ones of the forward-only kernel); every round also stores the hi tile of each layer's input to
//                        global memory by TMA (X[slot]).  Backward, evaluation K-1 .. 0, prediction on group 0 || dynamics on group 1:
//                        dX = W^T dZ with the layer's hi weight block read through an M-MAJOR A descriptor (the K-major SWIZZLE_128B
//                        block [o][k] is the M-major layout of W^T), bf16 operands, fp32 accumulation; the epilogue applies the relu
//                        mask taken from X[slot] and writes the result as the dZ tile of the layer below (same 16-byte chunk stores and
//                        column-order flips as the forward, so dZ[slot] and X[slot] always agree on the order of the samples); the
//                        first layers of the two heads accumulate into ONE accumulator (gradient of the trunk output); every round
//                        stores its dZ input tiles (dZ[slot]).  d loss / d h_e = prediction part + 2 * dynamics part (Learning.jl:299).
//   mz_k_learn_dw        dW[layer] = sum over (CTA, evaluation) dZ[slot] X[slot]^T: M = 64 outputs, N = 64 inputs, K = the 32 samples of a
//                        tile -- the saved N-major SWIZZLE_64B tiles [feature][32 samples] ARE K-major SWIZZLE_64B operands with K = samples.
//                        One CTA per (layer, chunk of sample tiles), accumulation over all its tiles and evaluations in one TMEM
//                        accumulator, so a partial is written once (no read-modify-write); db = row sums of the dZ tiles.
//   mz_k_grad_reduce (mz_learner_bptt.cuh) then sums the chunks' partials in order and adds 2 * theta.
// Backward operands are plain bf16 (hi parts): gradients agree with the oracle's Float64 backward to ~1e-2 of the largest entry per
// network (tests/test_gpu_mma.py states the tolerance); the exact fp32 kernel (mz_k_learn_bptt) remains for tight parity.
#pragma once
#include "mz_kernels_sp.cuh"
#include "mz_learner_bptt.cuh"

#define MZ_LR_IDESC_BACK (MZ_TC_IDESC | (1u << 15) | (1u << 16))                            // M = 64, N = 32, A M-major, B N-major
#define MZ_LR_IDESC_DW ((MZ_TC_IDESC & ~(0x3fu << 17)) | ((64u >> 3) << 17))               // M = 64, N = 64, both K-major
#define MZ_LR_BDESC_BYTES 96
#define MZ_LR_LG_ROWS 24                            // fp32 loss-gradient rows per evaluation: value [4], policy [16], reward [4], each [row][32 samples]

struct __align__(16) mz_lr_bdesc {                  // device form of a backward job pair (one round)
    unsigned long long a[2][2], b[2][2];            // [job][t]: M-major weight descriptor, N-major dZ tile descriptor        (64 B)
    uint32_t dst[2], f32[2];                        // dZ tile of the layer below / fp32 input gradient                        (16 B)
    int16_t ks[2][2];                               //                                                                         ( 8 B)
    uint8_t rows[2], mask[2], perm[2], discard[2];  //                                                                         ( 8 B)
};
static_assert(sizeof(mz_lr_bdesc) == MZ_LR_BDESC_BYTES, "mz_lr_bdesc size");

// M-major SWIZZLE_128B A descriptor over a K-major weight block [o][64 k]: 128 contiguous bytes = 64 values of M (= k), 8 K rows (= o) per
// 1024-byte atom -> SBO = 1024; M = 64 is one group, LBO unused
__device__ __forceinline__ uint64_t mz_lr_adesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// K-major SWIZZLE_64B descriptor over a saved tile [row][32 samples]: 8 rows per 512-byte atom
__device__ __forceinline__ uint64_t mz_lr_kdesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void mz_lr_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// position of sample n in a tile of column order `perm` (the epilogue's involution: 8q + 2c + e <-> 8c + 2q + e)
__device__ __forceinline__ int mz_lr_pos(int n, int perm) { return perm ? (((n >> 1) & 3) << 3) | ((n >> 3) << 1) | (n & 1) : n; }

struct mz_lr_args {
    mz_sp_args sp;                                   // forward plan
    const mz_lr_bround *brounds;                     // backward rounds (global)
    int32_t bfirst[3], bn_rounds[3], btotal_rounds;
    int32_t slot_base[3], layers_in_net[3], slots_per_cta, n_eval;
    int32_t start_tile[3][2], start_layer[3][2], start_perm[3][2];
    int32_t first_layer[3];
    int32_t dh_extra;                                // 1: the hidden-state gradient buffers have their own region (they do not fit over the forward's bias + output area)
    int32_t lg_off;                                  // byte offset in the weight area (behind the backward weights) of the loss gradients of all evaluations
    mz_bptt_args f;                                  // batch, predictions (mz_bwd_loss_grad reads a.f)
    unsigned char *xsave, *dzsave;                   // [tiles][slots_per_cta][4 KB]
    unsigned long long *dbg;                         // profiling build: cycle stamps of CTA 0 (thread 0 and thread 128)
    const double *inv_g_sum;                         // sum_i w_i / g_i over the whole batch (mz_k_inv_g_sum)
};
// sum over the batch of (importance weight) / gradient_scale in a fixed order: Q21's mean_i(1/g_i) couples every sample's policy term to it
__global__ void __launch_bounds__(1024) mz_k_inv_g_sum(int B, const float *gscale, const float *weights, double *out) {
    __shared__ double red[1024];
    const int tid = threadIdx.x;
    double s = 0.0;
    for (int i = tid; i < B; i += 1024) s += (weights ? (double)weights[i] : 1.0) / (double)gscale[i];
    red[tid] = s; __syncthreads();
    for (int k = 512; k > 0; k >>= 1) { if (tid < k) red[tid] += red[tid + k]; __syncthreads(); }
    if (tid == 0) out[0] = red[0];
}
#ifdef MZ_PHASE_TIMERS
#define MZ_LSTAMP(i) do { if (a.dbg && blockIdx.x == 0 && (tid == 0 || tid == 128)) { const long long c_ = clock64(); a.dbg[(tid ? 16 : 0) + (i)] += (unsigned long long)(c_ - lt_); lt_ = c_; } } while (0)
#else
#define MZ_LSTAMP(i)
#endif

// shared memory on top of the search kernel's carve-up: the PUCT-table + path region is free in the learner
struct mz_lr_smem { mz_lr_bdesc *bprog; int16_t *fslot, *bslot; float *dhp, *dhd, *dh, *scratch; uint64_t *bbar; double *red; };

// one thread: the backward weights of a network (hi blocks, side by side) on one barrier
__device__ __forceinline__ void mz_lr_load_weights(const mz_lr_args &a, const unsigned char *image, uint32_t w_base, uint32_t bar, int net) {
    uint32_t total = 0;
    for (int r = a.bfirst[net]; r < a.bfirst[net] + a.bn_rounds[net]; r++) for (int i = 0; i < a.brounds[r].ncopy; i++) total += (uint32_t)a.brounds[r].copy[i].bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
    for (int r = a.bfirst[net]; r < a.bfirst[net] + a.bn_rounds[net]; r++)
        for (int i = 0; i < a.brounds[r].ncopy; i++) {
            const mz_sp_copy c = a.brounds[r].copy[i];
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(w_base + (uint32_t)c.dst_off), "l"(image + c.src_off), "r"((uint32_t)c.bytes), "r"(bar) : "memory");
        }
}

// ---- backward rounds [first, first + count) of one network for one group: issuer warp / epilogue warps (cf. mz_sp_run_issuer / mz_sp_run) ----
__device__ __noinline__ uint32_t mz_lr_back_issuer(const mz_lr_bdesc *prog, const int16_t *bslot, int first, int count, uint32_t tmem_d, uint32_t mbar, uint32_t q,
                                                   uint32_t wbar, int grp, unsigned char *dzsave) {
    if (mz_elect_one()) { uint32_t ok = 0, spin = 0; while (!ok) { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(wbar), "r"(0u) : "memory"); if (++spin > (1u << 24)) __trap(); } }
    __syncwarp();
    for (int r = 0; r < count; r++, q++) {
        const mz_lr_bdesc *R = prog + first + r;
        if (r > 0) mz_sp_bar_sync(grp);
        mz_tc_fence_after();
        if (mz_elect_one()) {
            for (int j = 0; j < 2; j++) {
                if (R->discard[j]) continue;
                bool acc = false;
                for (int t = 0; t < 2; t++) {
                    const int ks = R->ks[j][t];
                    for (int k = 0; k < ks; k++) {       // A: 16 K rows (o) = two 1024-byte atoms; B: 16 k rows of the dZ tile = 1024 B
                        mz_lr_mma(tmem_d + 64u * (uint32_t)j, R->a[j][t] + (uint64_t)(128 * k), R->b[j][t] + (uint64_t)(64 * k), MZ_LR_IDESC_BACK, acc ? 1u : 0u);
                        acc = true;
                    }
                }
            }
            mz_sp_store_wait_read();                                    // before the commit releases an epilogue that may overwrite what the previous stores read
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
            for (int j = 0; j < 2; j++) for (int t = 0; t < 2; t++) {   // this round's dZ input tiles -> global
                const int sl = bslot[4 * (first + r) + 2 * j + t];
                if (sl >= 0) mz_sp_store_tile(dzsave + (size_t)sl * MZ_SP_TILE_BYTES, (uint32_t)(R->b[j][t] & 0x3fffu) << 4);
            }
            mz_sp_store_commit();
        }
        __syncwarp();
    }
    return q;
}
__device__ __noinline__ uint32_t mz_lr_back(const mz_lr_bdesc *prog, const int16_t *bslot, int first, int count, uint32_t tmem_d, uint32_t mbar, uint32_t q, int grp, int gtid,
                                            const unsigned char *xsave) {
    const int w = gtid >> 5, t = gtid & 31, c = t & 3;
    const uint32_t lane_base = tmem_d + ((uint32_t)(32 * w) << 16);
    // relu masks: this thread's eight samples of rows m0 / m0 + 8 of X[slot of layer[0]]; loaded one round ahead of their use
    uint32_t xm[2][2][4], xn[2][2][4];
    auto load_masks = [&](int rr, uint32_t (&dstm)[2][2][4]) {
        const mz_lr_bdesc *Rn = prog + first + rr;
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int m = 16 * w + (t >> 2) + 8 * half;
                const int sl = bslot[4 * (first + rr) + 2 * j];
                const bool on = Rn->mask[j] && !Rn->discard[j] && sl >= 0 && m < Rn->rows[j];
#pragma unroll
                for (int qq = 0; qq < 4; qq++)
                    dstm[j][half][qq] = on ? __ldcg(reinterpret_cast<const uint32_t *>(xsave + (size_t)sl * MZ_SP_TILE_BYTES + m * 64 + (((qq ^ (m >> 1)) & 3) << 4) + 4 * c)) : 0x3f803f80u;
            }
    };
    load_masks(0, xn);
    for (int r = 0; r < count; r++, q++) {
        const mz_lr_bdesc *R = prog + first + r;
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int qq = 0; qq < 4; qq++) xm[j][half][qq] = xn[j][half][qq];
        if (r + 1 < count) load_masks(r + 1, xn);
        mz_sp_wait_mma(mbar, q);
        mz_tc_fence_after();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (R->discard[j]) continue;                                // (also the absent second job)
            uint32_t v[16];
            mz_tc_ld16x256(lane_base + 64u * (uint32_t)j, v);
            mz_tc_wait_ld();
            const int rows = R->rows[j], rows16 = (rows + 15) & ~15;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int m = 16 * w + (t >> 2) + 8 * half;
                float x[8];
#pragma unroll
                for (int qq = 0; qq < 4; qq++) {
                    const uint32_t mk = xm[j][half][qq];
                    x[2 * qq] = (mk & 0x7fffu) != 0 && !(mk & 0x8000u) ? __uint_as_float(v[4 * qq + 2 * half]) : 0.0f;
                    x[2 * qq + 1] = ((mk >> 16) & 0x7fffu) != 0 && !(mk & 0x80000000u) ? __uint_as_float(v[4 * qq + 2 * half + 1]) : 0.0f;
                }
                if (R->dst[j]) {
                    if (m < rows16) {
                        uint32_t h[4];
#pragma unroll
                        for (int qq = 0; qq < 4; qq++) { if (m < rows) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[qq]) : "f"(x[2 * qq + 1]), "f"(x[2 * qq])); else h[qq] = 0u; }
                        const uint32_t ad = R->dst[j] + (uint32_t)(m * 64 + ((c ^ (m >> 1)) & 3) * 16);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                    }
                } else if (R->f32[j] && m < rows) {
                    if (R->perm[j]) {
                        const uint32_t ad = R->f32[j] + (uint32_t)((m * MZ_SP_OS + 8 * c) * 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ad + 16u), "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]) : "memory");
                    } else {
#pragma unroll
                        for (int qq = 0; qq < 4; qq++)
                            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(R->f32[j] + (uint32_t)((m * MZ_SP_OS + 8 * qq + 2 * c) * 4)), "f"(x[2 * qq]), "f"(x[2 * qq + 1]) : "memory");
                    }
                }
            }
        }
        mz_fence_proxy_async();
        mz_tc_fence_before();
        if (r + 1 < count) mz_sp_bar_arrive(grp);
    }
    return q;
}

// fp32 gradient rows [k][32 samples] (stride `stride` floats) -> a bf16 dZ tile in column order `perm`; rows [nrows, rows16) are zeroed
// src2: the value is src + 2 * src2 (d loss / d h = prediction part + 2 * dynamics part: make_dynamics_input doubles the state, Learning.jl:299)
__device__ __forceinline__ void mz_lr_stage_dz(uint32_t tile, int perm, int nrows, const float *src, int stride, int gtid, const float *src2 = nullptr) {
    const int rows16 = (nrows + 15) & ~15;
    for (int i = gtid; i < rows16 * MZ_ROWS; i += MZ_GROUP) {
        const int k = i / MZ_ROWS, n = i % MZ_ROWS;
        float v = (k < nrows && src) ? src[k * stride + n] : 0.0f;
        if (k < nrows && src && src2) v = v + 2.0f * src2[k * stride + n];
        const unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(v));
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(tile + mz_sp_tile_offset(k, mz_lr_pos(n, perm))), "h"(h) : "memory");
    }
}

__global__ void __launch_bounds__(MZ_SP_THREADS) mz_k_learn_bptt_tc(const __grid_constant__ mz_params P, const __grid_constant__ mz_lr_args a) {
    extern __shared__ __align__(1024) unsigned char mz_smem_sp[];
    const mz_sp_args &A = a.sp;
    const mz_sp_plan_s sp = mz_sp_carve(mz_smem_sp, A.warea_bytes, A.bias_floats, A.total_rounds, P.hidden_pad, 0, 0);   // no PUCT table, no path lists
    const int tid = threadIdx.x, K1 = P.K + 1;
    // learner-only regions behind the forward round table
    mz_lr_smem ls;
    {
        unsigned char *c = reinterpret_cast<unsigned char *>(sp.prog) + (((size_t)A.total_rounds * MZ_SP_RDESC_BYTES + 127) & ~(size_t)127);
        ls.bprog = (mz_lr_bdesc *)c; c += ((size_t)a.btotal_rounds * MZ_LR_BDESC_BYTES + 127) & ~(size_t)127;
        ls.fslot = (int16_t *)c; c += 256; ls.bslot = (int16_t *)c; c += 512;
        ls.bbar = (uint64_t *)c; c += 128;
        ls.scratch = (float *)c; ls.red = (double *)c;                  // 4 KB: loss-gradient rows of the two groups; first the batch reduction
        // the hidden-state gradients live where the forward kept its biases and fp32 outputs (free once the forward is over)
        ls.dhp = a.dh_extra ? reinterpret_cast<float *>(c + 4096 + 128) : sp.bias; ls.dhd = ls.dhp + (size_t)P.hidden_pad * MZ_SP_OS; ls.dh = ls.dhd + (size_t)P.hidden_pad * MZ_SP_OS;
    }
    const uint32_t tmem_base = mz_sp_setup(sp, A, MZ_SP_THREADS);
    mz_sp_ctx C; C.prog = mz_smem_u32(sp.prog); C.image = A.image; C.bars = mz_smem_u32(sp.bars);
    const bool worker = tid < MZ_THREADS;
    const int grp = worker ? tid >> 7 : (tid - MZ_THREADS) >> 5, gtid = tid & (MZ_GROUP - 1);
    const bool issuer0 = !worker && (tid & 31) == 0;
    const uint32_t tmem_d = tmem_base + (uint32_t)(128 * grp), mbar_mma = mz_smem_u32(sp.mbar_mma[grp]);
    const uint32_t tiles_g = sp.tiles + (uint32_t)(grp * MZ_SP_TILES_PER_GROUP * 2 * MZ_SP_TILE_BYTES);      // six 4 KB tiles of this group
    const uint32_t in_pred = mz_sp_group_tile(sp, 0, 0), in_dyn = mz_sp_group_tile(sp, 1, 0);
    const bool tanh_v = P.layers[P.nets[1].first + P.nets[1].n_trunk + P.nets[1].n_h1 - 1].act == MZ_ACT_TANH;
    const bool tanh_r = P.layers[P.nets[2].first + P.nets[2].n_trunk + P.nets[2].n_h1 + P.nets[2].n_h2 - 1].act == MZ_ACT_TANH;
    const int64_t g = (int64_t)blockIdx.x * MZ_ROWS + tid;
    const bool row_ok = tid < MZ_ROWS && g < a.f.f.B;
    unsigned char *xs = a.xsave + (size_t)blockIdx.x * a.slots_per_cta * MZ_SP_TILE_BYTES, *dzs = a.dzsave + (size_t)blockIdx.x * a.slots_per_cta * MZ_SP_TILE_BYTES;
    const bool dyn = P.K > 0;
    // ---- tables: save slot of every forward job, backward descriptors, slots of the backward jobs' layers ----
    if (tid < A.total_rounds) {
        const mz_sp_round G = A.rounds[tid];
        const int net = tid >= A.first[2] ? 2 : tid >= A.first[1] ? 1 : 0;
        for (int j = 0; j < 2; j++) ls.fslot[2 * tid + j] = j < G.njobs ? (int16_t)(G.job[j].layer - a.first_layer[net]) : (int16_t)0;
    }
    if (tid < a.btotal_rounds) {
        const mz_lr_bround G = a.brounds[tid];
        const int net = tid >= a.bfirst[2] ? 2 : tid >= a.bfirst[1] ? 1 : 0, bg = net == 1 ? 0 : 1;
        const uint32_t tg = sp.tiles + (uint32_t)(bg * MZ_SP_TILES_PER_GROUP * 2 * MZ_SP_TILE_BYTES);
        mz_lr_bdesc R;
        for (int j = 0; j < 2; j++) {
            const bool have = j < G.njobs;
            const mz_lr_bjob &J = G.job[have ? j : 0];
            for (int t = 0; t < 2; t++) {
                const bool on = have && J.layer[t] >= 0;
                R.a[j][t] = mz_lr_adesc(sp.w_base + (uint32_t)(on && !J.discard ? J.a_off[t] : 0));
                R.b[j][t] = mz_sp_bdesc(tg + (uint32_t)((on ? J.src_tile[t] : 0) * MZ_SP_TILE_BYTES));
                R.ks[j][t] = (int16_t)(on && !J.discard ? J.ks[t] : 0);
                ls.bslot[4 * tid + 2 * j + t] = on ? (int16_t)(J.layer[t] - a.first_layer[net]) : (int16_t)-1;
            }
            R.dst[j] = have && J.dst_tile >= 0 ? tg + (uint32_t)(J.dst_tile * MZ_SP_TILE_BYTES) : 0u;
            R.f32[j] = have && J.f32_off >= 0 ? mz_smem_u32(ls.dhp) + 4u * (uint32_t)J.f32_off : 0u;
            R.rows[j] = (uint8_t)(have ? J.rows : 0); R.mask[j] = (uint8_t)(have ? J.mask : 0); R.perm[j] = (uint8_t)(have ? J.perm : 0);
            R.discard[j] = (uint8_t)(have ? J.discard : 1);
        }
        ls.bprog[tid] = R;
    }
    if (tid == 0) { for (int i = 0; i < 3; i++) mz_mbar_init(&ls.bbar[i], 1); mz_fence_mbar_init(); }
    __syncthreads();
    // mean_i(1/g_i) over the whole batch (Q21): one sum for all CTAs (mz_k_inv_g_sum)
    double *s_red = ls.red; (void)s_red;
    const float invB = 1.0f / (float)a.f.f.B;
    const float up_pol = (float)(a.inv_g_sum[0] / (double)a.f.f.B / (double)a.f.f.B);
#ifdef MZ_PHASE_TIMERS
    long long lt_ = clock64();
#endif

    // ================= forward (mz_k_learn_forward_sp + tile saving) =================
    if (issuer0) { if (grp == 0) mz_sp_prime(C, A, 1); else mz_sp_fill_many(C, A.first[0], A.n_rounds[0]); }
    uint32_t q = 0, pass = 0;
    for (int i = tid; i < MZ_ROWS * P.stack_size; i += MZ_SP_THREADS) {
        const int rr = i / P.stack_size, k = i % P.stack_size;
        const int64_t gg = (int64_t)blockIdx.x * MZ_ROWS + rr;
        mz_sp_stage(in_dyn, k, rr, gg < a.f.f.B ? a.f.f.batch.obs[gg * P.stack_size + k] : 0.0f);
    }
    mz_fence_proxy_async();
    __syncthreads();
    if (grp == 1) {
        if (worker) q = mz_sp_run(C, A.first[0], A.n_rounds[0], tmem_d, mbar_mma, q, grp, gtid, nullptr);
        else { q = mz_sp_run_issuer(C, A.first[0], A.n_rounds[0], tmem_d, mbar_mma, q, 0u, grp, xs + (size_t)a.slot_base[0] * MZ_SP_TILE_BYTES, mz_smem_u32(ls.fslot)); if (issuer0 && dyn) mz_sp_prime(C, A, 2); }
    }
    __syncthreads();
    // Per evaluation: stage h_e (+ action plane) -> rounds -> then, side by side, warp 7 turns the heads' outputs into prediction rows while
    // everybody else stages h_{e+1} for the next evaluation.
    auto stage_eval = [&](int e, int t0, int nthr) {                    // threads t0 .. of nthr: operand tiles of evaluation e from outH (= h_e)
        for (int i = t0; i < MZ_ROWS * P.hidden; i += nthr) {
            const int k = i / MZ_ROWS, rr = i % MZ_ROWS;
            const float h = sp.outH[k * MZ_SP_OS + rr];
            mz_sp_stage(in_pred, k, rr, h);
            if (dyn) mz_sp_stage(in_dyn, k, rr, h * 2.0f);             // make_dynamics_input (:293-304): state * 2 (a copy), action plane = Float32(a) / A
        }
        if (dyn) for (int i = t0; i < MZ_ROWS * (P.sa_size - P.obs_size); i += nthr) {
            const int k = P.obs_size + i / MZ_ROWS, rr = i % MZ_ROWS;
            const int64_t gg = (int64_t)blockIdx.x * MZ_ROWS + rr;
            mz_sp_stage(in_dyn, k, rr, gg < a.f.f.B ? a.f.f.batch.actions[gg * K1 + e] / (float)P.A : 0.0f);
        }
    };
    stage_eval(0, tid, MZ_SP_THREADS);
    for (int e = 0; e < a.n_eval; e++) {
        mz_fence_proxy_async();
        __syncthreads();
        MZ_LSTAMP(1);                                                   // forward: staging / rows of the previous evaluation
        if (grp == 0 || dyn) {
            const int net = grp == 0 ? 1 : 2;
            if (worker) q = mz_sp_run(C, A.first[net], A.n_rounds[net], tmem_d, mbar_mma, q, grp, gtid, nullptr);
            else q = mz_sp_run_issuer(C, A.first[net], A.n_rounds[net], tmem_d, mbar_mma, q, pass, grp,
                                      xs + (size_t)(a.slot_base[net] + e * a.layers_in_net[net]) * MZ_SP_TILE_BYTES, mz_smem_u32(ls.fslot));
            pass++;
        }
        MZ_LSTAMP(2);                                                   // forward: rounds (own group)
        __syncthreads();
        MZ_LSTAMP(3);                                                   // forward: wait for the other group
        if (tid >= MZ_THREADS - 32 && tid < MZ_THREADS) {   // warp 7: evaluation e is row e + 1, and also row 0 when e == 0 (Q19); rewards: row 0 = 0 (:352)
            const int n = tid - (MZ_THREADS - 32);
            const int64_t gn = (int64_t)blockIdx.x * MZ_ROWS + n;
            if (gn < a.f.f.B) {
                float logits[MZ_MAX_A], policy[MZ_MAX_A];
                for (int k = 0; k < P.A; k++) logits[k] = sp.outL[k * MZ_SP_OS + n];
                mz_softmax(logits, P.A, policy);
                const float v = tanh_v ? mz_tanhf(sp.outV[n]) : sp.outV[n];
                const float rw = dyn ? (tanh_r ? mz_tanhf(sp.outR[n]) : sp.outR[n]) : 0.0f;
                for (int rr = (e == 0 ? 0 : e + 1); rr <= (dyn ? e + 1 : 0); rr++) {
                    a.f.f.pred_values[gn * K1 + rr] = v;
                    for (int k = 0; k < P.A; k++) a.f.f.pred_policies[(gn * K1 + rr) * P.A + k] = policy[k];
                    a.f.f.pred_rewards[gn * K1 + rr] = rr == 0 ? 0.0f : rw;
                }
            }
        } else if (e + 1 < a.n_eval) stage_eval(e + 1, tid < MZ_THREADS - 32 ? tid : tid - 32, MZ_SP_THREADS - 32);
    }
    __syncthreads();
    MZ_LSTAMP(4);
    // ---- the forward's weight refills and tile stores must have landed; then the backward weights take over the weight area ----
    if (issuer0) { mz_sp_drain(C, A, grp == 0 ? 1 : 2, pass); mz_sp_store_wait_all(); }
    __threadfence();
    __syncthreads();
    if (issuer0 && (grp == 0 || dyn)) mz_lr_load_weights(a, A.image, sp.w_base, mz_smem_u32(&ls.bbar[grp == 0 ? 0 : 1]), grp == 0 ? 1 : 2);
    for (int i = tid; i < P.hidden_pad * MZ_SP_OS; i += MZ_SP_THREADS) { ls.dh[i] = 0.0f; ls.dhd[i] = 0.0f; ls.dhp[i] = 0.0f; }
    __syncthreads();

    // ---- d loss / d (pre-activations of the heads' last layers) for every evaluation, all at once: (evaluation, head, sample) tasks over the
    //      worker threads.  Prediction heads: row e + 1 (+ row 0 for e = 0, which shares prediction(h_0)); reward: row e + 1 of dynamics step e ----
    float *lg = reinterpret_cast<float *>(mz_smem_sp + (sp.w_base - mz_smem_u32(mz_smem_sp)) + a.lg_off);
    if (worker) {
        const bool rew = dyn && P.intermediate_rewards != 0;
        for (int task = tid; task < a.n_eval * 3 * MZ_ROWS; task += MZ_THREADS) {
            const int e = task / (3 * MZ_ROWS), kind = (task / MZ_ROWS) % 3, n = task % MZ_ROWS;
            const int64_t gg = (int64_t)blockIdx.x * MZ_ROWS + n;
            float *dz = lg + (size_t)(e * MZ_LR_LG_ROWS + (kind == 0 ? 0 : kind == 1 ? 4 : 20)) * MZ_ROWS;
            if (kind == 2 && !rew) continue;
            mz_bwd_loss_grad(P, a.f, kind == 0 ? MZ_PRE_VALUE : kind == 1 ? MZ_PRE_POLICY : MZ_PRE_REWARD, P.K > 0 ? e + 1 : 0, kind != 2 && e == 0 && P.K > 0,
                             gg, gg < a.f.f.B, invB, up_pol, dz, n);
        }
    }
    __syncthreads();
    MZ_LSTAMP(5);                                                       // drain + backward weights issued + loss gradients
    // ================= backward =================
    for (int e = a.n_eval - 1; e >= 0; e--) {
        // ---- stage the top gradients: prediction heads from the loss, dynamics: d loss / d h_{e+1} and the reward loss ----
        if (worker) {
            const int net = grp == 0 ? 1 : 2;
            for (int h = 0; h < 2; h++) {
                const int L = a.start_layer[net][h];
                if (L < 0 || (net == 2 && !dyn)) continue;
                const uint32_t tile = tiles_g + (uint32_t)(a.start_tile[net][h] * MZ_SP_TILE_BYTES);
                const int out = P.layers[L].out;
                if (net == 2 && h == 0) {                                                // state head: dh_{e+1} = prediction part + 2 * dynamics part of evaluation e + 1
                    mz_lr_stage_dz(tile, a.start_perm[net][h], out, e + 1 < a.n_eval ? ls.dhp : nullptr, MZ_SP_OS, gtid, dyn ? ls.dhd : nullptr);   // (zero for the last evaluation: h_K feeds nothing)
                } else {
                    const bool live = net == 1 || P.intermediate_rewards != 0;
                    const float *src = lg + (size_t)(e * MZ_LR_LG_ROWS + (net == 2 ? 20 : h == 0 ? 0 : 4)) * MZ_ROWS;
                    mz_lr_stage_dz(tile, a.start_perm[net][h], out, live ? src : nullptr, MZ_ROWS, gtid);
                }
            }
        }
        mz_fence_proxy_async();
        __syncthreads();
        MZ_LSTAMP(6);                                                   // backward: loss gradients + staging
        if (grp == 0 || dyn) {
            const int net = grp == 0 ? 1 : 2;
            const size_t so = (size_t)(a.slot_base[net] + e * a.layers_in_net[net]) * MZ_SP_TILE_BYTES;
            if (worker) q = mz_lr_back(ls.bprog, ls.bslot, a.bfirst[net], a.bn_rounds[net], tmem_d, mbar_mma, q, grp, gtid, xs + so);
            else q = mz_lr_back_issuer(ls.bprog, ls.bslot, a.bfirst[net], a.bn_rounds[net], tmem_d, mbar_mma, q, mz_smem_u32(&ls.bbar[grp == 0 ? 0 : 1]), grp, dzs + so);
        }
        MZ_LSTAMP(7);                                                   // backward: rounds (own group)
        __syncthreads();
        MZ_LSTAMP(8);                                                   // backward: wait for the other group
    }
    // ---- representation: its backward weights replace the dynamics' ----
    if (issuer0 && grp == 1) mz_lr_load_weights(a, A.image, sp.w_base, mz_smem_u32(&ls.bbar[2]), 0);
    if (worker && grp == 1) mz_lr_stage_dz(tiles_g + (uint32_t)(a.start_tile[0][0] * MZ_SP_TILE_BYTES), a.start_perm[0][0], P.hidden, ls.dhp, MZ_SP_OS, gtid, dyn ? ls.dhd : nullptr);
    mz_fence_proxy_async();
    __syncthreads();
    if (grp == 1) {
        const size_t so = (size_t)a.slot_base[0] * MZ_SP_TILE_BYTES;
        if (worker) q = mz_lr_back(ls.bprog, ls.bslot, a.bfirst[0], a.bn_rounds[0], tmem_d, mbar_mma, q, grp, gtid, xs + so);
        else q = mz_lr_back_issuer(ls.bprog, ls.bslot, a.bfirst[0], a.bn_rounds[0], tmem_d, mbar_mma, q, mz_smem_u32(&ls.bbar[2]), grp, dzs + so);
    }
    if (issuer0) mz_sp_store_wait_all();
    mz_tc_fence_before();
    __syncthreads();
    MZ_LSTAMP(9);                                                       // representation backward + teardown
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)MZ_SP_TMEM_COLS) : "memory");
}

// dynamic shared memory of mz_k_learn_bptt_tc: the search kernel's carve-up without PUCT table and path lists (S = 0) + the backward tables
__host__ __device__ inline bool mz_lr_alias_fits(int bias_floats, int hidden_pad);
__host__ __device__ inline size_t mz_lr_smem_bytes(int warea_bytes, int bias_floats, int total_rounds, int btotal_rounds, int hidden_pad) {
    return mz_sp_smem_bytes(warea_bytes, bias_floats, total_rounds, hidden_pad, 0, 0) + (((size_t)btotal_rounds * MZ_LR_BDESC_BYTES + 127) & ~(size_t)127) + 256 + 512 + 128 + 4096 + 128 +
           (mz_lr_alias_fits(bias_floats, hidden_pad) ? 0 : 3 * (size_t)hidden_pad * MZ_SP_OS * 4 + 128);
}
// the three hidden-state gradient buffers alias the forward's bias block + fp32 output area
__host__ __device__ inline bool mz_lr_alias_fits(int bias_floats, int hidden_pad) {
    return (((size_t)bias_floats * 4 + 127) & ~(size_t)127) + (size_t)(24 + hidden_pad) * MZ_SP_OS * 4 >= 3 * (size_t)hidden_pad * MZ_SP_OS * 4;
}

// ---- dW / db from the saved tiles --------------------------------------------------------------------------------------------------
struct mz_dw_args {
    const unsigned char *xsave, *dzsave; float *gpart;   // gpart: [chunks][P.total_floats], device weight layout
    int32_t tiles, chunks, slots_per_cta, n_eval;
    int32_t slot_base[3], layers_in_net[3], first_layer[3];
};
#define MZ_DW_STAGES 8                             // 8 KB each (dZ tile + X tile); dynamic shared memory
#define MZ_DW_GROUP 4                              // instances per synchronisation
__global__ void __launch_bounds__(128) mz_k_learn_dw(const __grid_constant__ mz_params P, const __grid_constant__ mz_dw_args a) {
    extern __shared__ __align__(1024) unsigned char mz_smem_dw[];
    constexpr int NS = MZ_DW_STAGES, GI = MZ_DW_GROUP;
    __shared__ __align__(8) uint64_t full[NS], empty[NS], done;
    __shared__ uint32_t tmem_slot;
    unsigned char *base = mz_smem_dw + ((1024u - (mz_smem_u32(mz_smem_dw) & 1023u)) & 1023u);
    auto bufp = [&](int s, int which) { return base + (size_t)(2 * s + which) * MZ_SP_TILE_BYTES; };     // [stage][dZ | X]
    const int L = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x, w = tid >> 5, t = tid & 31;
    const mz_layer &l = P.layers[L];
    const int net = L >= a.first_layer[2] ? 2 : L >= a.first_layer[1] ? 1 : 0;
    const int evals = net == 0 ? 1 : a.n_eval;
    const int t0 = (int)((long long)a.tiles * chunk / a.chunks), t1 = (int)((long long)a.tiles * (chunk + 1) / a.chunks);
    const int n_inst = (t1 - t0) * evals;
    if (tid == 0) { for (int i = 0; i < NS; i++) { mz_mbar_init(&full[i], 1); mz_mbar_init(&empty[i], 1); } mz_mbar_init(&done, 1); mz_fence_mbar_init(); }
    __syncwarp();
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mz_smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    mz_tc_fence_before();
    __syncthreads();
    mz_tc_fence_after();
    const uint32_t tmem = tmem_slot;
    auto slot_of = [&](int i, size_t &off) {
        const int tile = t0 + i / evals, e = i % evals;
        off = ((size_t)tile * a.slots_per_cta + a.slot_base[net] + e * a.layers_in_net[net] + (L - a.first_layer[net])) * MZ_SP_TILE_BYTES;
    };
    auto load = [&](int i) {
        const int s = i % NS; size_t off; slot_of(i, off);
        mz_mbar_expect_tx(&full[s], 2 * MZ_SP_TILE_BYTES);
        mz_bulk_g2s(bufp(s, 0), a.dzsave + off, MZ_SP_TILE_BYTES, &full[s]);
        mz_bulk_g2s(bufp(s, 1), a.xsave + off, MZ_SP_TILE_BYTES, &full[s]);
    };
    if (tid == 0) for (int i = 0; i < NS && i < n_inst; i++) load(i);
    float db = 0.0f;                                                                  // thread (o = tid >> 1, half = tid & 1): sum of dZ[o][16 samples]
    for (int i0 = 0; i0 < n_inst; i0 += GI) {
        const int i1 = i0 + GI < n_inst ? i0 + GI : n_inst;
        for (int i = i0; i < i1; i++) {
            const int s = i % NS;
            mz_mbar_wait(&full[s], (uint32_t)(i / NS) & 1u);
            // bias gradient: row sums of the dZ tile (fixed order)
            const int o = tid >> 1, hf = tid & 1;
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                const int chunk16 = 2 * hf + ch;                                       // logical chunk (8 samples) of row o
                const uint4 v = *reinterpret_cast<const uint4 *>(bufp(s, 0) + o * 64 + (((chunk16 ^ (o >> 1)) & 3) << 4));
                const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++) { db = db + __uint_as_float(u[k] << 16); db = db + __uint_as_float(u[k] & 0xffff0000u); }
            }
        }
        __syncthreads();
        if (tid == 0) {
            mz_tc_fence_after();
            for (int i = i0; i < i1; i++) {
                const int s = i % NS;
                const uint64_t ad = mz_lr_kdesc(mz_smem_u32(bufp(s, 0))), bd = mz_lr_kdesc(mz_smem_u32(bufp(s, 1)));
                for (int k = 0; k < 2; k++) mz_lr_mma(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), MZ_LR_IDESC_DW, (i > 0 || k > 0) ? 1u : 0u);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mz_smem_u32(&empty[s])) : "memory");
            }
            // refill the stages of the PREVIOUS group (its MMAs have had a whole group to complete), not this one's
            for (int i = i0 - GI; i >= 0 && i < i0; i++)
                if (i + NS < n_inst) { const int ps = i % NS; mz_mbar_wait(&empty[ps], (uint32_t)(i / NS) & 1u); load(i + NS); }
        }
    }
    if (tid == 0) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mz_smem_u32(&done)) : "memory");
    }
    mz_mbar_wait(&done, 0);
    mz_tc_fence_after();
    float *gp = a.gpart + (size_t)chunk * P.total_floats;
    if (n_inst > 0) {
        const uint32_t lane_base = tmem + ((uint32_t)(32 * w) << 16);
#pragma unroll
        for (int cb = 0; cb < 2; cb++) {
            uint32_t v[16];
            mz_tc_ld16x256(lane_base + 32u * (uint32_t)cb, v);
            mz_tc_wait_ld();
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int o = 16 * w + (t >> 2) + 8 * half;
#pragma unroll
                for (int qq = 0; qq < 4; qq++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int k = 32 * cb + 8 * qq + 2 * (t & 3) + e;
                        if (o < l.out && k < l.in) gp[l.w_off + k * l.out_pad + o] = __uint_as_float(v[4 * qq + 2 * half + e]);
                    }
            }
        }
    } else {
        for (int i = tid; i < l.in * l.out_pad; i += 128) gp[l.w_off + i] = 0.0f;
    }
    {
        const float other = __shfl_xor_sync(0xffffffffu, db, 1);
        const int o = tid >> 1;
        if ((tid & 1) == 0 && o < l.out) gp[l.b_off + o] = db + other;
    }
    mz_tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}
