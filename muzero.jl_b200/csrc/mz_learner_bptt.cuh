// mz_learner_bptt.cuh -- grad_mode = MZ_GRAD_BPTT: the K-step unroll forward (src/Learning.jl:347-370) followed by the
// backward pass through it, fused in one kernel.  One CTA = 32 samples (one "tile"), two 128-thread groups:
//
//   forward   group 0: representation, then prediction(h_e); group 1: dynamics(h_e, a_e) -- concurrently, exactly the
//             arithmetic of mz_k_learn_forward (same dense tile), every layer's output also streamed to the tile's
//             activation block in HBM/L2 (k-major [feature][32 samples], 128-byte rows);
//   backward  a host-built program of layer applications (mzh::build_bptt): per stage the layer's {W,b} block and its
//             saved input activations are staged by TMA bulk copies (double-buffered, one mbarrier per slot), then
//               dW[k][o] = sum_rows X[k][row] * dZ[o][row]      -> the tile's partial-gradient block (plain adds: each
//               db[o]    = sum_rows dZ[o][row]                      address has one owner thread, so sums are deterministic)
//               dX[k][row] = act'(X[k][row]) * sum_o W[k][o] * dZ[o][row]   -> shared memory, the next stage's dZ
//             group 0 walks prediction rows K..1 (+ row 0 merged into row 1), group 1 the dynamics steps K..1; after each
//             step d loss / d h_{i-1} = (prediction part) + 2 * (dynamics part) is combined across the groups.
//   mz_k_grad_reduce then sums the tiles' partial blocks in tile order and adds the L2 term 2*theta (Learning.jl:287).
//
// The loss whose gradient this is: Learning.jl:261-288 as written (Q21: policy term mean_j(S_j) * mean_i(1/g_i),
// logitcrossentropy applied to the already-softmaxed policy).  Checked against the oracle's Float64 backward.
#pragma once
#include "mz_learner.cuh"

struct mz_bptt_args {
    mz_learn_args f;            // forward arguments (weights, batch, prediction outputs)
    float *act;                 // [tiles][plan.tile_floats] activation blocks
    float *gpart;               // [tiles][P.total_floats] partial gradients (device weight layout)
    const mz_bstage *stages[2]; // backward programs of the two groups
    int32_t dim_wide;           // rows of in0 and of group 0's staged-input slots: the widest layer input (the observation stack)
    int32_t dim_narrow;         // rows of every other activation buffer: widest layer output / input of a layer other than the representation's first
    int32_t wfloats[2];         // largest {W,b} block of the layers group g runs (0: representation + prediction, 1: dynamics)
};

__device__ __forceinline__ void mz_fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// forward layer with activation saving (same pipeline protocol as mz_nn_layer)
__device__ __forceinline__ void mz_nn_layer_save(mz_nn_pipe &s, const mz_params &P, int layer, int next, const float *src, float *dst, float *gsave) {
    if (s.gtid == 0 && next >= 0) mz_nn_issue(s, P, next, (s.q + 1) & 1u);
    mz_mbar_wait(&s.mbar[s.q & 1u], (s.q >> 1) & 1u);
    const mz_layer &L = P.layers[layer];
    mz_dense_tile_save(L.in, L.out_pad, L.act, mz_smem_u32(s.wbuf[s.q & 1u]), mz_smem_u32(src), mz_smem_u32(dst), s.gtid, gsave);
    mz_group_sync(s.grp);
    s.q++;
}
__device__ __forceinline__ void mz_nn_chain_save(mz_nn_pipe &s, const mz_params &P, const mz_bptt_plan &pl, int first, int n, int after, const float *src,
                                                 float *dst, float *t0, float *t1, float *gblock) {
    const float *cur = src;
    for (int i = 0; i < n; i++) {
        float *d = (i == n - 1) ? dst : ((i & 1) ? t1 : t0);
        mz_nn_layer_save(s, P, first + i, (i == n - 1) ? after : first + i + 1, cur, d, gblock + pl.y_off[first + i]);
        cur = d;
    }
}
__device__ __noinline__ void mz_nn_net_save(mz_nn_pipe &s, const mz_params &P, const mz_bptt_plan &pl, int net, int after, const float *src, float *bufT,
                                            float *h1dst, float *h2dst, float *t0, float *t1, float *gblock) {
    const mz_net &N = P.nets[net];
    int f = N.first;
    if (N.n_h1 == 0) { mz_nn_chain_save(s, P, pl, f, N.n_trunk, after, src, h1dst, t0, t1, gblock); return; }
    mz_nn_chain_save(s, P, pl, f, N.n_trunk, f + N.n_trunk, src, bufT, t0, t1, gblock);
    mz_nn_chain_save(s, P, pl, f + N.n_trunk, N.n_h1, f + N.n_trunk + N.n_h1, bufT, h1dst, t0, t1, gblock);
    mz_nn_chain_save(s, P, pl, f + N.n_trunk + N.n_h1, N.n_h2, after, bufT, h2dst, t0, t1, gblock);
}

// ---- backward tile: one layer application for one 128-thread group -------------------------------------------------
// dW / db into the tile's partial block.  Tiles of 4 k x 4 o, reduction over the 32 rows in 4-row chunks; every lane
// starts at its own chunk (rotation) so the quarter-warps of the 128-bit shared loads never collide on a bank.
__device__ __forceinline__ void mz_bwd_dw(const mz_layer &L, bool first, uint32_t x_smem, uint32_t dz_smem, const float *dz, float *gpart, int gtid) {
    const int lane = gtid & 31, warp = gtid >> 5;
    const int kq = (L.in + 3) >> 2, oq = L.out_pad >> 2, tiles = kq * oq;
    const uint32_t rot = (uint32_t)(lane & 7) * 16u;
    for (int t = gtid; t < tiles; t += MZ_GROUP) {
        const int kg = t / oq, og = t - kg * oq;
        unsigned long long acc2[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc2[i][j] = 0ull;
        // earlier contributions of this tile to the layer (read-modify-write): fetched now, needed after the reduction
        float4 old[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = 4 * kg + i;
            old[i] = (!first && k < L.in) ? *reinterpret_cast<const float4 *>(gpart + L.w_off + k * L.out_pad + 4 * og) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        const uint32_t xa = x_smem + (uint32_t)kg * (4u * MZ_ROWS * 4u), za = dz_smem + (uint32_t)og * (4u * MZ_ROWS * 4u);
#pragma unroll 2
        for (int c = 0; c < 8; c++) {
            const uint32_t ro = (rot + (uint32_t)c * 16u) & 127u;
            float4 xv[4], zv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { xv[i] = mz_lds128(xa + (uint32_t)i * (MZ_ROWS * 4u) + ro); zv[i] = mz_lds128(za + (uint32_t)i * (MZ_ROWS * 4u) + ro); }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                unsigned long long x01, x23;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x01) : "f"(xv[i].x), "f"(xv[i].y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(x23) : "f"(xv[i].z), "f"(xv[i].w));
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    unsigned long long z01, z23;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(z01) : "f"(zv[j].x), "f"(zv[j].y));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(z23) : "f"(zv[j].z), "f"(zv[j].w));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i][j]) : "l"(x01), "l"(z01));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i][j]) : "l"(x23), "l"(z23));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = 4 * kg + i;
            if (k >= L.in) break;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc2[i][j])); v[j] = a + b; }
            float4 *dst = reinterpret_cast<float4 *>(gpart + L.w_off + k * L.out_pad + 4 * og);
            float4 r; r.x = old[i].x + v[0]; r.y = old[i].y + v[1]; r.z = old[i].z + v[2]; r.w = old[i].w + v[3];
            *dst = r;
        }
    }
    // bias gradient: warp w owns outputs 16w .. 16w+15, lanes = rows, butterfly sum; lane j keeps output 16w+j and the
    // 16 sums go out in one coalesced read-modify-write
    {
        float *dst = gpart + L.b_off + warp * 16 + lane;
        const bool mine_ok = lane < 16 && warp * 16 + lane < L.out_pad;
        const float oldb = (mine_ok && !first) ? *dst : 0.0f;
        float mine = 0.0f;
        for (int j = 0; j < 16; j++) {
            const int o = warp * 16 + j;
            if (o >= L.out_pad) break;
            float v = dz[o * MZ_ROWS + lane];
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, m);
            if (lane == j) mine = v;
        }
        if (mine_ok) *dst = oldb + mine;
    }
}
// dX[k][row] = act'(X[k][row]) * scale * sum_o W[k][o] * dZ[o][row]   (rows k >= in are written as zeros: they are the
// pad rows of the consumer's dZ)
__device__ __forceinline__ void mz_bwd_dx(const mz_layer &L, int mode, int prev_act, float scale, uint32_t w_smem, uint32_t x_smem, uint32_t dz_smem,
                                          uint32_t dx_smem, int gtid) {
    const int lane = gtid & 31, warp = gtid >> 5;
    const int rg = lane & 7, kg = (warp << 2) | (lane >> 3);
    const int kq = (L.in + 3) >> 2;
    if (kg >= kq) return;
    unsigned long long acc2[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) { acc2[i][0] = 0ull; acc2[i][1] = 0ull; }
    const uint32_t wstride = (uint32_t)L.out_pad * 4u;
    uint32_t wa[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { int k = 4 * kg + i; k = k > L.in ? L.in : k; wa[i] = w_smem + (uint32_t)k * wstride; }
    uint32_t za = dz_smem + (uint32_t)rg * 16u;
    for (int o = 0; o < L.out_pad; o += 4) {
        float4 wv[4], zv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { wv[i] = mz_lds128(wa[i] + (uint32_t)o * 4u); zv[i] = mz_lds128(za + (uint32_t)(o + i) * (MZ_ROWS * 4u)); }
        float4 c;
        c.x = wv[0].x; c.y = wv[1].x; c.z = wv[2].x; c.w = wv[3].x; mz_fma_step(acc2, zv[0], c);
        c.x = wv[0].y; c.y = wv[1].y; c.z = wv[2].y; c.w = wv[3].y; mz_fma_step(acc2, zv[1], c);
        c.x = wv[0].z; c.y = wv[1].z; c.z = wv[2].z; c.w = wv[3].z; mz_fma_step(acc2, zv[2], c);
        c.x = wv[0].w; c.y = wv[1].w; c.z = wv[2].w; c.w = wv[3].w; mz_fma_step(acc2, zv[3], c);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int k = 4 * kg + i;
        float4 r;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(acc2[i][0]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(r.z), "=f"(r.w) : "l"(acc2[i][1]));
        const uint32_t off = (uint32_t)(k * MZ_ROWS * 4) + (uint32_t)rg * 16u;
        if (k >= L.in) { r.x = r.y = r.z = r.w = 0.0f; }
        else {
            if (prev_act == MZ_ACT_RELU) {
                const float4 x = mz_lds128(x_smem + off);
                r.x = x.x > 0.0f ? r.x : 0.0f; r.y = x.y > 0.0f ? r.y : 0.0f; r.z = x.z > 0.0f ? r.z : 0.0f; r.w = x.w > 0.0f ? r.w : 0.0f;
            } else if (prev_act == MZ_ACT_TANH) {
                const float4 x = mz_lds128(x_smem + off);
                r.x = r.x * (1.0f - x.x * x.x); r.y = r.y * (1.0f - x.y * x.y); r.z = r.z * (1.0f - x.z * x.z); r.w = r.w * (1.0f - x.w * x.w);
            }
            if (scale != 1.0f) { r.x = r.x * scale; r.y = r.y * scale; r.z = r.z * scale; r.w = r.w * scale; }
            if (mode == MZ_DX_ACCUM) { const float4 o = mz_lds128(dx_smem + off); r.x = o.x + r.x; r.y = o.y + r.y; r.z = o.z + r.z; r.w = o.w + r.w; }
        }
        mz_sts128(dx_smem + off, r);
    }
}
__device__ __noinline__ void mz_bwd_tile(const mz_layer &L, int dx_mode, int prev_act, float scale, bool first, uint32_t w_smem, uint32_t x_smem,
                                         const float *dz, uint32_t dx_smem, float *gpart, int gtid) {
    const uint32_t dz_smem = mz_smem_u32(dz);
    mz_bwd_dw(L, first, x_smem, dz_smem, dz, gpart, gtid);
    if (dx_mode != MZ_DX_NONE) mz_bwd_dx(L, dx_mode, prev_act, scale, w_smem, x_smem, dz_smem, dx_smem, gtid);
}

// d loss / d (pre-activation of a head's last layer) for the 32 samples of the tile, written k-major into `dz`
// (rows >= out are zeroed).  Thread `row` < 32 of the group handles one sample.  up_pol = mean_i(1/g_i) / B.
__device__ __forceinline__ void mz_bwd_loss_grad(const mz_params &P, const mz_bptt_args &a, int pre, int row_idx, bool merge0, int64_t g, bool ok,
                                                 float invB, float up_pol, float *dz, int row) {
    const int K1 = P.K + 1;
    if (pre == MZ_PRE_POLICY) {
        float acc[MZ_MAX_A];
        for (int k = 0; k < P.A; k++) acc[k] = 0.0f;
        if (ok) {
            for (int rr = merge0 ? 0 : row_idx; rr <= row_idx; rr++) {      // rows 0 and 1 share prediction(h_0)
                const float *p = a.f.pred_policies + (g * K1 + rr) * P.A, *y = a.f.batch.policies + (g * K1 + rr) * P.A;
                float q[MZ_MAX_A], dp[MZ_MAX_A];
                mz_softmax(p, P.A, q);                                       // logsoftmax's softmax of the already-softmaxed policy (Q21)
                float ys = 0.0f, dot = 0.0f;
                for (int k = 0; k < P.A; k++) ys = ys + y[k];
                for (int k = 0; k < P.A; k++) { dp[k] = up_pol * (q[k] * ys - y[k]); dot = fmaf(p[k], dp[k], dot); }
                for (int k = 0; k < P.A; k++) acc[k] = acc[k] + p[k] * (dp[k] - dot);
            }
        }
        const int pad = (P.A + 3) & ~3;
        for (int k = 0; k < pad; k++) dz[k * MZ_ROWS + row] = k < P.A ? acc[k] : 0.0f;
    } else {
        float d = 0.0f;
        if (ok) {
            const float gs = (P.per && a.f.batch.weights) ? a.f.batch.gscale[g] / a.f.batch.weights[g] : a.f.batch.gscale[g];   // (sum / g) * w
            for (int rr = merge0 ? 0 : row_idx; rr <= row_idx; rr++) {
                const float y = pre == MZ_PRE_VALUE ? a.f.pred_values[g * K1 + rr] : a.f.pred_rewards[g * K1 + rr];
                const float t = pre == MZ_PRE_VALUE ? a.f.batch.values[g * K1 + rr] : a.f.batch.rewards[g * K1 + rr];
                float u = 2.0f * (y - t) / gs * invB;                        // d mean_b(sum_k (y - t)^2 / g_b) / dy
                const int act = pre == MZ_PRE_VALUE ? MZ_ACT_TANH : P.layers[P.nets[2].first + P.nets[2].n_trunk + P.nets[2].n_h1 + P.nets[2].n_h2 - 1].act;
                if (act == MZ_ACT_TANH) u = u * (1.0f - y * y);
                d = d + u;
            }
        }
        dz[row] = d; dz[MZ_ROWS + row] = 0.0f; dz[2 * MZ_ROWS + row] = 0.0f; dz[3 * MZ_ROWS + row] = 0.0f;
    }
}

// Shared memory of the kernel.  Only the representation's first layer sees the observation stack (99 inputs at stacked_observations = 2), and
// only group 0 runs it: in0 and group 0's two staged-input slots have `wide` rows, the other ten activation buffers `narrow` rows, and each
// group's weight slots are as large as ITS largest layer (the selection tables of the search kernels' carve-up are not needed here).
struct mz_bptt_smem { mz_smem_plan sp; float *xbuf[2][2]; float *exch; size_t bytes; };
__host__ __device__ inline mz_bptt_smem mz_bptt_carve(unsigned char *base, int wide, int narrow, int wf0, int wf1, int hidden_pad) {
    mz_bptt_smem r;
    const size_t w[2] = {((size_t)wf0 * 4 + 127) & ~(size_t)127, ((size_t)wf1 * 4 + 127) & ~(size_t)127};
    const size_t bw = (size_t)wide * MZ_ROWS * 4, bn = (size_t)narrow * MZ_ROWS * 4;
    unsigned char *c = base;
    for (int g = 0; g < 2; g++) for (int i = 0; i < 2; i++) { r.sp.wbuf[g][i] = (float *)c; c += w[g]; }
    r.sp.mbar[0] = (uint64_t *)c; r.sp.mbar[1] = (uint64_t *)(c + 32); c += 128;
    r.sp.in0 = (float *)c; c += bw;              // in0, in1, bufT, t0, t1 are contiguous (zeroed together)
    r.sp.in1 = (float *)c; c += bn;
    for (int g = 0; g < 2; g++) { r.sp.bufT[g] = (float *)c; c += bn; r.sp.t0[g] = (float *)c; c += bn; r.sp.t1[g] = (float *)c; c += bn; }
    r.sp.outV = (float *)c; c += 4 * MZ_ROWS * 4;
    r.sp.outL = (float *)c; c += 16 * MZ_ROWS * 4;
    r.sp.outR = (float *)c; c += 4 * MZ_ROWS * 4;
    r.sp.outH = (float *)c; c += (size_t)hidden_pad * MZ_ROWS * 4;
    r.sp.pbc0 = nullptr; r.sp.sqrtN = nullptr; r.sp.path = nullptr;
    for (int i = 0; i < 2; i++) { r.xbuf[0][i] = (float *)c; c += bw; }
    for (int i = 0; i < 2; i++) { r.xbuf[1][i] = (float *)c; c += bn; }
    r.exch = (float *)c; c += bn;
    r.bytes = (size_t)(c - base) + 128;
    return r;
}
__host__ inline size_t mz_bptt_smem_bytes(int wide, int narrow, int wf0, int wf1, int hidden_pad) { return mz_bptt_carve(nullptr, wide, narrow, wf0, wf1, hidden_pad).bytes; }

__global__ void __launch_bounds__(MZ_THREADS) mz_k_learn_bptt(const __grid_constant__ mz_params P, const __grid_constant__ mz_bptt_plan pl, const mz_bptt_args a) {
    extern __shared__ __align__(128) unsigned char mz_smem[];
    const mz_bptt_smem bs = mz_bptt_carve(mz_smem, a.dim_wide, a.dim_narrow, a.wfloats[0], a.wfloats[1], P.hidden_pad);
    const mz_smem_plan &sp = bs.sp;
    __shared__ double s_red[MZ_THREADS];
    const int tid = threadIdx.x, K1 = P.K + 1;
    const int64_t g = (int64_t)blockIdx.x * MZ_ROWS + tid;
    const bool row_ok = tid < MZ_ROWS && g < a.f.B;
    float *act = a.act + (size_t)blockIdx.x * pl.tile_floats;
    float *gpart = a.gpart + (size_t)blockIdx.x * P.total_floats;
    mz_nn_pipe pipe;
    mz_pipe_init(pipe, sp, a.f.wglob);
    for (int i = threadIdx.x; i < (a.dim_wide + 7 * a.dim_narrow) * MZ_ROWS; i += MZ_THREADS) sp.in0[i] = 0.0f;
    // mean_i(1/g_i) over the whole batch (Q21's broadcast couples every sample's policy term to it): same fixed-order sum in every CTA
    {
        double s = 0.0;
        const bool per = P.per && a.f.batch.weights;
        for (int i = tid; i < a.f.B; i += MZ_THREADS) s += (per ? (double)a.f.batch.weights[i] : 1.0) / (double)a.f.batch.gscale[i];
        s_red[tid] = s;
    }
    // layers without any data gradient: their partial sums are zero
    for (int l = 0; l < P.n_layers; l++)
        if ((pl.dead_layers >> l) & 1ull) for (int i = tid; i < P.layers[l].floats; i += MZ_THREADS) gpart[P.layers[l].w_off + i] = 0.0f;
    __syncthreads();
    for (int s = MZ_THREADS / 2; s > 0; s >>= 1) { if (tid < s) s_red[tid] += s_red[tid + s]; __syncthreads(); }
    const float invB = 1.0f / (float)a.f.B;
    const float up_pol = (float)(s_red[0] / (double)a.f.B / (double)a.f.B);
    const int pred_first = P.nets[1].first, dyn_first = P.nets[2].first;
    if (tid == 0) mz_nn_issue(pipe, P, P.nets[0].first, 0);
    if (tid == MZ_GROUP && P.K > 0) mz_nn_issue(pipe, P, dyn_first, 0);

    // ================= forward (Learning.jl:347-370) =================
    {
        const int in_pad = (P.stack_size + 3) & ~3;
        for (int i = tid; i < MZ_ROWS * in_pad; i += MZ_THREADS) {
            int k = i / MZ_ROWS, r = i % MZ_ROWS;
            int64_t gg = (int64_t)blockIdx.x * MZ_ROWS + r;
            float v = (k < P.stack_size && gg < a.f.B) ? a.f.batch.obs[gg * P.stack_size + k] : 0.0f;
            sp.in0[k * MZ_ROWS + r] = v;
            act[pl.net_base[0] + i] = v;
        }
    }
    __syncthreads();
    if (pipe.grp == 0) mz_nn_net_save(pipe, P, pl, 0, pred_first, sp.in0, sp.bufT[0], sp.in1, nullptr, sp.t0[0], sp.t1[0], act + pl.net_base[0]);   // h_0 -> in1
    __syncthreads();
    for (int e = 0; e < pl.n_pred_evals; e++) {
        float *pblock = act + pl.net_base[1] + (size_t)e * pl.net_block[1];
        float *dblock = act + pl.net_base[2] + (size_t)e * pl.net_block[2];
        if (P.K > 0) {   // make_dynamics_input (:293-304): state * 2 (copy), action plane = Float32(a) / A
            if (row_ok || tid < MZ_ROWS) {
                const float plane = row_ok ? a.f.batch.actions[g * K1 + e] / (float)P.A : 0.0f;
                for (int k = 0; k < P.hidden; k++) sp.in0[k * MZ_ROWS + tid] = sp.in1[k * MZ_ROWS + tid] * 2.0f;
                for (int k = P.obs_size; k < P.sa_size; k++) sp.in0[k * MZ_ROWS + tid] = plane;
            }
            __syncthreads();
            const int sa_pad = (P.sa_size + 3) & ~3;
            for (int i = tid; i < sa_pad * MZ_ROWS; i += MZ_THREADS) dblock[i] = i < P.sa_size * MZ_ROWS ? sp.in0[i] : 0.0f;
        }
        for (int i = tid; i < P.hidden_pad * MZ_ROWS; i += MZ_THREADS) pblock[i] = i < P.hidden * MZ_ROWS ? sp.in1[i] : 0.0f;
        const bool more = e + 1 < pl.n_pred_evals;
        if (pipe.grp == 0) mz_nn_net_save(pipe, P, pl, 1, more ? pred_first : -1, sp.in1, sp.bufT[0], sp.outV, sp.outL, sp.t0[0], sp.t1[0], pblock);
        else if (P.K > 0)  mz_nn_net_save(pipe, P, pl, 2, more ? dyn_first : -1, sp.in0, sp.bufT[1], sp.outH, sp.outR, sp.t0[1], sp.t1[1], dblock);
        __syncthreads();
        if (row_ok) {   // rows: evaluation e is row e+1, and also row 0 when e == 0 (Q19); rewards row 0 = 0 (:352)
            float logits[MZ_MAX_A], policy[MZ_MAX_A];
            for (int k = 0; k < P.A; k++) logits[k] = sp.outL[k * MZ_ROWS + tid];
            mz_softmax(logits, P.A, policy);
            for (int rr = (e == 0 ? 0 : e + 1); rr <= (P.K > 0 ? e + 1 : 0); rr++) {
                a.f.pred_values[g * K1 + rr] = sp.outV[tid];
                for (int k = 0; k < P.A; k++) a.f.pred_policies[(g * K1 + rr) * P.A + k] = policy[k];
                a.f.pred_rewards[g * K1 + rr] = rr == 0 ? 0.0f : sp.outR[tid];
            }
        }
        if (P.K > 0) for (int k = tid; k < P.hidden * MZ_ROWS; k += MZ_THREADS) sp.in1[k] = sp.outH[k];   // h_{e+1}
        __syncthreads();
    }

    // ================= backward =================
    // the activation blocks were written through the generic proxy and are read back by TMA (async proxy)
    __threadfence();
    mz_fence_proxy_async_all();
    __syncthreads();
    float *bufs[6] = {sp.bufT[pipe.grp], sp.t0[pipe.grp], sp.t1[pipe.grp], sp.in0, sp.in1, bs.exch};
    float *xbuf[2] = {bs.xbuf[pipe.grp][0], bs.xbuf[pipe.grp][1]};
    const mz_bstage *prog = a.stages[pipe.grp];
    const int nst = pl.n_stages[pipe.grp];
    auto issue = [&](int idx, uint32_t slot) {
        const mz_bstage st = prog[idx];
        const mz_layer &L = P.layers[st.layer];
        const uint32_t wbytes = (uint32_t)L.floats * 4u, xbytes = (uint32_t)((L.in + 3) & ~3) * (MZ_ROWS * 4u);
        mz_mbar_expect_tx(&pipe.mbar[slot], wbytes + xbytes);
        mz_bulk_g2s(pipe.wbuf[slot], pipe.wglob + L.w_off, wbytes, &pipe.mbar[slot]);
        mz_bulk_g2s(xbuf[slot], act + st.x_off, xbytes, &pipe.mbar[slot]);
    };
    if (pipe.gtid == 0 && nst > 0) issue(0, pipe.q & 1u);
    int idx = 0;
    for (int step = 0; step < pl.n_steps; step++) {
        const int end = pl.step_end[pipe.grp][step];
        for (; idx < end; idx++) {
            const mz_bstage st = prog[idx];
            const mz_layer &L = P.layers[st.layer];
            if (pipe.gtid == 0 && idx + 1 < nst) issue(idx + 1, (pipe.q + 1) & 1u);
            if (st.pre != MZ_PRE_NONE) {
                if (pipe.gtid < MZ_ROWS) {
                    const int64_t gg = (int64_t)blockIdx.x * MZ_ROWS + pipe.gtid;
                    mz_bwd_loss_grad(P, a, st.pre, st.row, st.merge0 != 0, gg, gg < a.f.B, invB, up_pol, bufs[st.dz_buf], pipe.gtid);
                }
                mz_group_sync(pipe.grp);
            }
            mz_mbar_wait(&pipe.mbar[pipe.q & 1u], (pipe.q >> 1) & 1u);
            mz_bwd_tile(L, st.dx_mode, st.prev_act, st.dx_scale, st.first != 0, mz_smem_u32(pipe.wbuf[pipe.q & 1u]), mz_smem_u32(xbuf[pipe.q & 1u]),
                        bufs[st.dz_buf], mz_smem_u32(bufs[st.dx_buf]), gpart, pipe.gtid);
            mz_group_sync(pipe.grp);
            pipe.q++;
        }
        __syncthreads();
        if (step + 1 < pl.n_steps) {   // d loss / d h_{i-1} = prediction part + 2 * dynamics part (the factor is applied by the stage)
            const bool dv = pl.dhd_valid[step] != 0;
            for (int i = tid; i < P.hidden_pad * MZ_ROWS; i += MZ_THREADS) {
                float v = 0.0f;
                if (i < P.hidden * MZ_ROWS) { v = bufs[MZ_BUF_DHP][i]; if (dv) v = v + bufs[MZ_BUF_DHD][i]; }
                bufs[MZ_BUF_DH][i] = v;
            }
            __syncthreads();
        }
    }
}

// grad[i] = sum over tiles (in tile order) of the partial gradients + 2 * theta[i]   (loss + sum(sqnorm, params), Learning.jl:287)
__global__ void mz_k_grad_reduce(int n, int tiles, const float *gpart, const float *theta, float *grad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.0f;
    for (int t = 0; t < tiles; t++) s = s + gpart[(size_t)t * n + i];
    grad[i] = s + (theta[i] + theta[i]);
}

// ---- FeedForwardHP.use_batch_norm (Learning.jl:70-79) in grad_mode = MZ_GRAD_BPTT ---------------------------------------------------
// BatchNorm runs in test mode, so Dense + BatchNorm(relu) is the Dense layer W' = diag(s) W, b' = s (b - mu) + beta with
// s = gamma / sqrt(sigma2 + 1f-5): mz_k_learn_bptt runs unchanged on a folded copy of the weights (mz_k_bn_fold) and mz_k_grad_reduce_bn maps
// its dW', db' back by the chain rule: dW = s dW', db = s db', dbeta = db', dgamma = (sum_k dW'[k][o] W[k][o] + db'[o] (b[o] - mu[o])) / sqrt(sigma2 + 1f-5);
// mu and sigma2 are not Flux parameters (gradient 0, no L2 term).  Grid = (layers, slices).
__global__ void __launch_bounds__(256) mz_k_bn_fold(const __grid_constant__ mz_params P, const float *w, float *wf) {
    const mz_layer &l = P.layers[blockIdx.x];
    const int wn = l.in * l.out_pad;
    const float *bnp = w + l.b_off + l.out_pad;                          // beta | gamma | mu | sigma2
    for (int i = blockIdx.y * 256 + threadIdx.x; i < l.floats; i += 256 * gridDim.y) {
        float v = w[l.w_off + i];
        if (l.bn && i < wn + l.out_pad) {
            const int o = i < wn ? i % l.out_pad : i - wn;
            const float s = bnp[l.out_pad + o] / sqrtf(bnp[3 * l.out_pad + o] + 1e-5f);
            v = i < wn ? v * s : s * (v - bnp[2 * l.out_pad + o]) + bnp[o];
        }
        wf[l.w_off + i] = v;
    }
}
__global__ void __launch_bounds__(256) mz_k_grad_reduce_bn(const __grid_constant__ mz_params P, int tiles, const float *gpart, const float *theta, float *grad) {
    const mz_layer &l = P.layers[blockIdx.x];
    const int wn = l.in * l.out_pad, n = P.total_floats;
    const float *bnp = theta + l.b_off + l.out_pad;
    auto tsum = [&](int idx) { float s = 0.0f; for (int t = 0; t < tiles; t++) s = s + gpart[(size_t)t * n + idx]; return s; };
    for (int i = blockIdx.y * 256 + threadIdx.x; i < l.floats; i += 256 * gridDim.y) {
        const int idx = l.w_off + i;
        const float th = theta[idx];
        float g;
        if (!l.bn) g = tsum(idx) + (th + th);
        else {
            const int part = i < wn ? 0 : 1 + (i - wn) / l.out_pad, o = i < wn ? i % l.out_pad : (i - wn) % l.out_pad;
            const float sd = sqrtf(bnp[3 * l.out_pad + o] + 1e-5f);
            if (part <= 1) g = (bnp[l.out_pad + o] / sd) * tsum(idx) + (th + th);                       // W, b
            else if (part == 2) g = tsum(l.b_off + o) + (th + th);                                      // beta
            else if (part == 3) {                                                                       // gamma
                float acc = tsum(l.b_off + o) * (theta[l.b_off + o] - bnp[2 * l.out_pad + o]);
                for (int k = 0; k < l.in; k++) acc = acc + tsum(l.w_off + k * l.out_pad + o) * theta[l.w_off + k * l.out_pad + o];
                g = acc / sd + (th + th);
            } else g = 0.0f;                                                                            // mu, sigma2
        }
        grad[idx] = g;
    }
}
