// mz_rn_host.h -- host side of the ResNet networks (net_type = MZ_NET_RESNET): unit list in blob order, Glorot /
// BatchNorm-default initialisation, and the compilation of the three networks into the step program + bf16 weight
// image that mz_k_search_rn / mz_k_nn_forward_rn execute on the tcgen05 tensor cores.  Plain C++ (no CUDA).
//
// The networks are the REPAIRED form of src/Learning.jl:148-255 (the reference's constructors read undefined names and
// never ran): see DESIGN.md "ResNet" and the independent restatement in oracle/mz_oracle.c.
//   ConvBN(k, cin => cout, act) = Conv((k,k), cin => cout, pad = k/2) ; BatchNorm(cout, act)   (test mode: stored statistics)
//   block(k, n) = relu.(x + BN(Conv(relu(BN(Conv(x))))))
//   representation: ConvBN(K, planes => nf, relu); blocks(K)                                   -> hidden (W,H,nf)
//   prediction    : ConvBN(1, nf => nf, relu); blocks(1); value head ConvBN(1, nf => nvf, relu), flatten, Dense(=> hs, relu),
//                   depth_value x Dense(hs, hs, relu), Dense(hs => 1, tanh); policy head ConvBN(1, nf => npf, relu), flatten,
//                   Dense(=> hs), depth_value x Dense(relu), Dense(hs => A), softmax
//   dynamics      : ConvBN(1, nf+1 => nf, relu); blocks(1); state head ConvBN(1, nf => nf, relu), blocks(1); reward head
//                   like the value head
// Blob: units in construction order; ConvBN = W (k,k,cin,cout) column-major, b, beta, gamma, mu, var; Dense = W (out,in), b.
#pragma once
#include <string>
#include <vector>
#include "mz_host.h"

namespace mzh {

struct rn_unit { int kind /* 0 conv, 1 dense */, k, cin, cout, act; int w_off, b_off, beta_off, gamma_off, mu_off, var_off; };
struct rn_wref {             // one B-operand image (+ its parameter block) inside a step's staged weight block
    int unit, tap;           // tap >= 0: the [cout][cin] slice of tap (ka + k*kb); -1: all taps as one im2col / 1x1 image
    int unit2;               // second unit stacked below the first in the same image (head convs: value + policy filters), or -1
    int w_sub, p_sub, n16, kblocks;
    float mul;               // accumulator multiplier folded into S (2 for the dynamics input, make_state_action's state * 2)
    int plane;               // 1: the unit's last input channel is the action plane -> E
    int params;              // 1: emit the {S, T, E} block
};
struct rn_model {
    mz_rn_params R;
    std::vector<rn_unit> units[3];
    int n_params[3], base[3];
    std::vector<mz_rn_step> steps;
    std::vector<std::vector<rn_wref>> wrefs;   // per step
    int image_bytes;
};

inline int rn_total_params(const rn_model &M) { return M.n_params[0] + M.n_params[1] + M.n_params[2]; }

inline void rn_units_build(const mz_config &c, rn_model &M) {
    int off = 0;
    const int nf = c.rn_num_filters, nb = c.rn_num_blocks, cells = c.W * c.H, hs = c.width_hidden;
    const int planes = c.C * (c.stacked_observations + 1) + c.stacked_observations;
    auto conv = [&](std::vector<rn_unit> &v, int k, int cin, int cout, int act) {
        rn_unit u{}; u.kind = 0; u.k = k; u.cin = cin; u.cout = cout; u.act = act;
        u.w_off = off; off += k * k * cin * cout; u.b_off = off; off += cout; u.beta_off = off; off += cout; u.gamma_off = off; off += cout;
        u.mu_off = off; off += cout; u.var_off = off; off += cout; v.push_back(u);
    };
    auto dense = [&](std::vector<rn_unit> &v, int in, int out, int act) {
        rn_unit u{}; u.kind = 1; u.k = 1; u.cin = in; u.cout = out; u.act = act; u.w_off = off; off += in * out; u.b_off = off; off += out; v.push_back(u);
    };
    auto tower = [&](std::vector<rn_unit> &v, int k, int cin) {
        conv(v, k, cin, nf, MZ_ACT_RELU);
        for (int i = 0; i < nb; i++) { conv(v, k, nf, nf, MZ_ACT_RELU); conv(v, k, nf, nf, MZ_ACT_ID); }
    };
    auto head = [&](std::vector<rn_unit> &v, int f, int first_act, int out, int out_act) {
        conv(v, 1, nf, f, MZ_ACT_RELU); dense(v, cells * f, hs, first_act);
        for (int i = 0; i < c.depth_value; i++) dense(v, hs, hs, MZ_ACT_RELU);
        dense(v, hs, out, out_act);
    };
    for (int n = 0; n < 3; n++) M.units[n].clear();
    M.base[0] = off; tower(M.units[0], c.rn_kernel, planes); M.n_params[0] = off - M.base[0];
    M.base[1] = off; tower(M.units[1], 1, nf); head(M.units[1], c.rn_first_head_filters, MZ_ACT_RELU, 1, MZ_ACT_TANH);
    head(M.units[1], c.rn_second_head_filters, MZ_ACT_ID, c.A, MZ_ACT_ID); M.n_params[1] = off - M.base[1];
    M.base[2] = off; tower(M.units[2], 1, nf + 1); tower(M.units[2], 1, nf); head(M.units[2], c.rn_first_head_filters, MZ_ACT_RELU, 1, MZ_ACT_TANH);
    M.n_params[2] = off - M.base[2];
}

inline const char *rn_validate(const mz_config &c) {
    if (c.nn_mode != MZ_NN_BF16_TC) return "the ResNet networks run on the tensor cores only: set nn_mode = MZ_NN_BF16_TC";
    if (c.rn_num_filters < 8 || c.rn_num_filters > 64 || c.rn_num_filters % 8) return "rn_num_filters must be a multiple of 8 in 8..64";
    if (c.rn_num_blocks < 0 || c.rn_num_blocks > 8) return "rn_num_blocks must be in 0..8";
    if (c.rn_kernel != 1 && c.rn_kernel != 3) return "rn_kernel must be 1 or 3";
    if (c.width_hidden != 64) return "the ResNet heads need width_hidden = 64";
    if (c.W * c.H > 128) return "board too large: W*H must be <= 128";
    if (c.hidden_state_size != c.W * c.H * c.rn_num_filters) return "hidden_state_size must equal W*H*rn_num_filters for the ResNet networks";
    const int planes = c.C * (c.stacked_observations + 1) + c.stacked_observations;
    if (c.rn_kernel * c.rn_kernel * planes > 64) return "k*k*planes of the first representation convolution must be <= 64";
    if (c.rn_first_head_filters < 1 || c.rn_second_head_filters < 1 || c.rn_first_head_filters + c.rn_second_head_filters > 16) return "head filters out of range";
    if (c.W * c.H * c.rn_first_head_filters > 64 || c.W * c.H * c.rn_second_head_filters > 128) return "flattened head inputs must fit 64 (value/reward) and 128 (policy) features";
    if (c.depth_value < 0 || c.depth_value > 4) return "depth_value out of range";
    if (c.A > 16) return "action space too large";
    return nullptr;
}

// ---- program construction ----------------------------------------------------------------------------------------
struct rn_builder {
    rn_model &M; const mz_config &c; int image_off = 0;
    int add_wref(std::vector<rn_wref> &v, int &blk, int net, int unit, int tap, int unit2, int n16, int kblocks, float mul, int plane, int params) {
        rn_wref w{}; w.unit = unit + (net << 16); w.tap = tap; w.unit2 = unit2 >= 0 ? unit2 + (net << 16) : -1; w.n16 = n16; w.kblocks = kblocks; w.mul = mul; w.plane = plane; w.params = params;
        w.w_sub = blk; blk += n16 * 2048 * kblocks;
        w.p_sub = -1;
        v.push_back(w);
        return (int)v.size() - 1;
    }
    void finish(mz_rn_step &st, std::vector<rn_wref> &v, int blk) {
        for (auto &w : v) if (w.params) { w.p_sub = blk; blk += 768; }
        blk = (blk + 127) & ~127;
        st.w_off = image_off; st.w_bytes = blk; image_off += blk;
        for (int j = 0; j < st.njobs; j++) { const rn_wref &w = v[st.jobs[j].wref]; st.jobs[j].w_sub = w.w_sub; st.jobs[j].p_sub = w.p_sub; }
        M.steps.push_back(st); M.wrefs.push_back(v);
    }
    // a ConvBN unit applied to the four (tree,cell) tiles; k x k units become k*k tap-steps over shifted copies of the source
    void conv_steps(int net, int ui, int src, int dst, int skip, int act, float mul, int plane, bool to_pool) {
        const rn_unit &u = M.units[net][ui];
        const int ntaps = (u.k > 1 && ui > 0) ? u.k * u.k : 1;      // unit 0 of the representation is one im2col image
        for (int t = 0; t < ntaps; t++) {
            mz_rn_step st{}; std::vector<rn_wref> v; int blk = 0;
            const bool last = t == ntaps - 1;
            st.njobs = MZ_RN_TILES; st.ntaps = (uint8_t)ntaps; st.tap = (uint8_t)t; st.last = last ? 1 : 0; st.accumulate = t > 0 ? 1 : 0;
            const int pad = u.k / 2, ka = t % u.k, kb = t / u.k;
            st.dx = (int8_t)(ntaps > 1 ? pad - ka : 0); st.dy = (int8_t)(ntaps > 1 ? pad - kb : 0);
            int wi = add_wref(v, blk, net, ui, ntaps > 1 ? t : -1, -1, 4, 1, mul, plane, last ? 1 : 0);
            for (int j = 0; j < MZ_RN_TILES; j++) {
                mz_rn_job &J = st.jobs[j];
                J.a_buf = (uint8_t)(src * MZ_RN_TILES + j); J.dst_buf = (uint8_t)(dst * MZ_RN_TILES + j); J.skip_buf = skip >= 0 ? (uint8_t)(skip * MZ_RN_TILES + j) : 0xff;
                J.epi = MZ_RN_EPI_TILE; J.n16 = 4; J.kblocks = 1; J.act = (uint8_t)act; J.wg = (uint8_t)j; J.acc = (uint8_t)j;
                J.flags = (uint8_t)((plane ? MZ_RN_F_PLANE : 0) | (to_pool ? MZ_RN_F_POOL : 0)); J.wref = wi;
            }
            finish(st, v, blk);
        }
    }
    void tower(int net, int first, int src, int other, bool to_pool, float mul, int plane) {   // input in `src`; output ends in `other` (nb > 0 or not)
        const int nb = c.rn_num_blocks;
        conv_steps(net, first, src, other, -1, MZ_ACT_RELU, mul, plane, to_pool && nb == 0);
        for (int b = 0; b < nb; b++) {
            conv_steps(net, first + 1 + 2 * b, other, src, -1, MZ_ACT_RELU, 1.0f, 0, false);
            conv_steps(net, first + 2 + 2 * b, src, other, other, MZ_ACT_RELU, 1.0f, 0, to_pool && b == nb - 1);   // relu.(x + layers(x)), in place
        }
    }
    // head convolution(s) reading the trunk tiles: filters of unit ua -> head tile ha, of unit ub (optional) -> head tile hb
    void head_conv(int net, int src, int ua, int ha, int ub, int hb) {
        mz_rn_step st{}; std::vector<rn_wref> v; int blk = 0;
        st.njobs = MZ_RN_TILES; st.ntaps = 1; st.last = 1;
        int wi = add_wref(v, blk, net, ua, -1, ub, 1, 1, 1.0f, 0, 1);
        for (int j = 0; j < MZ_RN_TILES; j++) {
            mz_rn_job &J = st.jobs[j];
            J.a_buf = (uint8_t)(src * MZ_RN_TILES + j); J.dst_buf = (uint8_t)ha; J.dst2_buf = (uint8_t)(ub >= 0 ? hb : 0xff); J.skip_buf = 0xff;
            J.epi = MZ_RN_EPI_HEAD; J.n16 = 1; J.kblocks = 1; J.act = MZ_ACT_RELU; J.wg = (uint8_t)j; J.acc = (uint8_t)j;
            J.nfa = (uint8_t)M.units[net][ua].cout; J.nfb = (uint8_t)(ub >= 0 ? M.units[net][ub].cout : 0); J.wref = wi;
        }
        finish(st, v, blk);
    }
    // dense chains of one or two heads in lock-step (job 0 -> warpgroup 0, job 1 -> warpgroup 1); rows = trees
    void dense_chains(int net, int n_chains, const int *first_unit, const int *in_buf, const int *out_id) {
        const int nd = c.depth_value + 2;
        bool lock = true;
        for (int h = 0; h < n_chains; h++) if (M.units[net][first_unit[h]].cin > 64) lock = false;
        for (int pass = 0; pass < (lock ? 1 : n_chains); pass++)
            for (int i = 0; i < nd; i++) {
                mz_rn_step st{}; std::vector<rn_wref> v; int blk = 0;
                st.ntaps = 1; st.last = 1;
                for (int h = (lock ? 0 : pass); h < (lock ? n_chains : pass + 1); h++) {
                    const rn_unit &u = M.units[net][first_unit[h] + i];
                    const bool fin = i == nd - 1;
                    const int kblocks = (u.cin + 63) / 64, n16 = fin ? 1 : 4;
                    int wi = add_wref(v, blk, net, first_unit[h] + i, -1, -1, n16, kblocks, 1.0f, 0, 1);
                    mz_rn_job &J = st.jobs[st.njobs];
                    const int ping = MZ_RN_BUF_X0 + h, pong = MZ_RN_BUF_T0 + h;          // tile h of X / T: free once the head convolutions are done
                    J.a_buf = (uint8_t)(i == 0 ? in_buf[h] : ((i & 1) ? ping : pong)); J.dst_buf = (uint8_t)((i & 1) ? pong : ping); J.skip_buf = 0xff; J.dst2_buf = 0xff;
                    J.epi = (uint8_t)(fin ? MZ_RN_EPI_F32 : MZ_RN_EPI_TILE); J.n16 = (uint8_t)n16; J.kblocks = (uint8_t)kblocks; J.act = (uint8_t)u.act;
                    J.wg = (uint8_t)h; J.acc = (uint8_t)st.njobs; J.flags = MZ_RN_F_TREES; J.out = (uint8_t)u.cout; J.out_id = (uint8_t)out_id[h]; J.wref = wi;
                    st.njobs++;
                }
                finish(st, v, blk);
            }
    }
};

inline const char *rn_build(const mz_config &c, const mz_params &P, rn_model &M) {
    if (const char *e = rn_validate(c)) return e;
    rn_units_build(c, M);
    M.steps.clear(); M.wrefs.clear();
    mz_rn_params &R = M.R; memset(&R, 0, sizeof(R));
    R.cells = c.W * c.H; R.nf = c.rn_num_filters; R.tpt = 128 / R.cells; R.ntrees = MZ_RN_TILES * R.tpt; R.rows_valid = R.tpt * R.cells;
    if (R.ntrees > 64) return "internal: more than 64 trees per CTA";
    R.node_bytes = R.cells * 128;
    R.planes = c.C * (c.stacked_observations + 1) + c.stacked_observations; R.ksize = c.rn_kernel;
    R.nvf = c.rn_first_head_filters; R.npf = c.rn_second_head_filters;
    rn_builder B{M, c};
    const int nt = 1 + 2 * c.rn_num_blocks, nh = c.depth_value + 3;
    // representation: the kernel builds the im2col tiles of the first convolution in X; the hidden state goes to the pool
    R.prog_repr[0] = (int)M.steps.size();
    B.tower(0, 0, 0 /* X */, 1 /* T */, true, 1.0f, 0);
    R.prog_repr[1] = (int)M.steps.size();
    // prediction: H staged in X
    R.prog_pred[0] = (int)M.steps.size();
    B.tower(1, 0, 0, 1, false, 1.0f, 0);
    B.head_conv(1, 1, nt, MZ_RN_BUF_HV, nt + nh, MZ_RN_BUF_HP);
    { const int fu[2] = {nt + 1, nt + nh + 1}, ib[2] = {MZ_RN_BUF_HV, MZ_RN_BUF_HP}, oi[2] = {MZ_RN_OUT_V, MZ_RN_OUT_L}; B.dense_chains(1, 2, fu, ib, oi); }
    R.prog_pred[1] = (int)M.steps.size();
    // dynamics: H staged in X (the x2 of make_state_action is folded into the first convolution's scale, the action plane into E)
    R.prog_dyn[0] = (int)M.steps.size();
    B.tower(2, 0, 0, 1, false, 2.0f, 1);
    B.head_conv(2, 1, 2 * nt, MZ_RN_BUF_HV, -1, 0xff);
    B.tower(2, nt, 1 /* trunk in T */, 0 /* next state ends in X */, true, 1.0f, 0);
    { const int fu[1] = {2 * nt + 1}, ib[1] = {MZ_RN_BUF_HV}, oi[1] = {MZ_RN_OUT_R}; B.dense_chains(2, 1, fu, ib, oi); }
    R.prog_dyn[1] = (int)M.steps.size();
    R.n_steps = (int)M.steps.size(); R.smem_first = R.prog_pred[0];
    // row-local steps: single-tap convolutions whose job j is warpgroup j's own tile; inside a run of them the warpgroups run unsynchronised
    {
        auto rl = [&](int i) {
            const mz_rn_step &s = M.steps[(size_t)i];
            if (s.ntaps != 1 || s.njobs != MZ_RN_TILES || !s.last) return false;
            for (int j = 0; j < s.njobs; j++) if (s.jobs[j].epi != MZ_RN_EPI_TILE || (s.jobs[j].flags & MZ_RN_F_TREES) || s.jobs[j].wg != j || s.jobs[j].acc != j) return false;
            return true;
        };
        const int ends[3] = {R.prog_repr[1], R.prog_pred[1], R.prog_dyn[1]};
        for (auto &st : M.steps) {
            for (int w = 0; w < 4; w++) st.wgjob[w] = -1;
            for (int j = 0; j < st.njobs; j++) {
                if (st.jobs[j].wg >= 4 || st.wgjob[st.jobs[j].wg] >= 0) return "internal: two jobs of one step on the same warpgroup";
                st.wgjob[st.jobs[j].wg] = (int8_t)j;
                if (st.jobs[j].epi == MZ_RN_EPI_TILE && (st.jobs[j].dst_buf >= 8 || (st.jobs[j].skip_buf != 0xff && st.jobs[j].skip_buf >= 8))) return "internal: tile epilogue outside the X/T tiles";
            }
        }
        for (int i = 0; i < R.n_steps; i++) {
            M.steps[(size_t)i].rowlocal = 0;
            if (!rl(i)) continue;
            const bool range_end = i + 1 == ends[0] || i + 1 == ends[1] || i + 1 == ends[2];
            M.steps[(size_t)i].rowlocal = (!range_end && rl(i + 1)) ? 1 : 2;
        }
    }
    M.image_bytes = B.image_off;
    R.image_bytes = B.image_off;
    int slot = 0;
    for (const auto &s : M.steps) if (s.w_bytes > slot) slot = s.w_bytes;
    R.slot_bytes = (slot + 1023) & ~1023;
    // tree pool geometry: node records, then (S+1) hidden states of cells x 64 bf16
    R.hidden_off_bytes = (P.nodes_per_tree * 16 + 127) & ~127;
    R.tree_stride_bytes = R.hidden_off_bytes + (c.num_iters + 1) * R.node_bytes;
    return nullptr;
}

inline const rn_unit &rn_unit_of(const rn_model &M, int packed) { return M.units[packed >> 16][packed & 0xffff]; }

// blob -> bf16 B-operand images (rows = output channels / features, K = input channels, K-major SWIZZLE_128B) + {S, T, E}
inline void rn_pack(const rn_model &M, const float *blob, std::vector<unsigned char> &image) {
    image.assign((size_t)M.image_bytes + 4096, 0);
    for (size_t si = 0; si < M.steps.size(); si++) {
        const mz_rn_step &st = M.steps[si];
        for (const rn_wref &w : M.wrefs[si]) {
            unsigned char *base = image.data() + st.w_off + w.w_sub;
            int row0 = 0;
            for (int part = 0; part < 2; part++) {
                const int pu = part == 0 ? w.unit : w.unit2;
                if (pu < 0) continue;
                const rn_unit &u = rn_unit_of(M, pu);
                const int kin = u.kind == 0 ? (w.plane ? u.cin - 1 : u.cin) : u.cin;
                for (int o = 0; o < u.cout; o++) {
                    // BatchNorm scale (and the x2 of the dynamics input) folded into the weight before the bf16 rounding
                    const float fold = u.kind == 0 ? w.mul * (blob[u.gamma_off + o] / sqrtf(blob[u.var_off + o] + 1e-5f)) : 1.0f;
                    if (u.kind == 1) {
                        for (int k = 0; k < kin; k++) {
                            uint16_t h = f2bf16(blob[u.w_off + o + u.cout * k]);
                            memcpy(base + (k >> 6) * (w.n16 * 2048) + tc_tile_offset(row0 + o, k & 63), &h, 2);
                        }
                    } else if (w.tap >= 0) {
                        const int ka = w.tap % u.k, kb = w.tap / u.k;
                        for (int ci = 0; ci < kin; ci++) { uint16_t h = f2bf16(blob[u.w_off + ka + u.k * (kb + u.k * (ci + u.cin * o))] * fold); memcpy(base + tc_tile_offset(row0 + o, ci), &h, 2); }
                    } else {   // all taps in one image: k index = tap * cin + ci (im2col); 1x1: tap = 0
                        for (int t = 0; t < u.k * u.k; t++) for (int ci = 0; ci < kin; ci++) {
                            uint16_t h = f2bf16(blob[u.w_off + (t % u.k) + u.k * ((t / u.k) + u.k * (ci + u.cin * o))] * fold);
                            memcpy(base + tc_tile_offset(row0 + o, t * kin + ci), &h, 2);
                        }
                    }
                }
                row0 += u.cout;
            }
            if (w.params) {
                float *S = reinterpret_cast<float *>(image.data() + st.w_off + w.p_sub), *T = S + 64, *E = S + 128;
                int c0 = 0;
                for (int part = 0; part < 2; part++) {
                    const int pu = part == 0 ? w.unit : w.unit2;
                    if (pu < 0) continue;
                    const rn_unit &u = rn_unit_of(M, pu);
                    for (int o = 0; o < u.cout; o++) {
                        if (u.kind == 1) { S[c0 + o] = 1.0f; T[c0 + o] = blob[u.b_off + o]; E[c0 + o] = 0.0f; continue; }
                        // BatchNorm (test mode): the scale s = gamma / sqrt(var + eps) lives in the weights; y = acc + plane * (w_plane * s) + ((b - mu) * s + beta)
                        const float den = sqrtf(blob[u.var_off + o] + 1e-5f), s = blob[u.gamma_off + o] / den;
                        S[c0 + o] = w.mul * s;
                        T[c0 + o] = fmaf(blob[u.b_off + o] - blob[u.mu_off + o], s, blob[u.beta_off + o]);
                        E[c0 + o] = w.plane ? blob[u.w_off + 0 + u.k * (0 + u.k * ((u.cin - 1) + u.cin * o))] * s : 0.0f;
                    }
                    c0 += u.cout;
                }
            }
        }
    }
}

// Flux.glorot_uniform with nfan = (k*k*cin, k*k*cout); conv bias 0; BatchNorm beta 0, gamma 1, mu 0, var 1 (Flux defaults)
inline void rn_init_weights(const rn_model &M, uint64_t seed, float *blob) {
    for (int n = 0; n < 3; n++) for (size_t ui = 0; ui < M.units[n].size(); ui++) {
        const rn_unit &u = M.units[n][ui];
        const int nw = u.k * u.k * u.cin * u.cout;
        const float scale = sqrtf(24.0f / (float)(u.k * u.k * (u.cin + u.cout)));
        for (int i = 0; i < nw; i += 4) {
            mz_u4 r = mz_philox(seed, MZ_STREAM_INIT, (uint32_t)n, (uint32_t)ui, (uint32_t)(i / 4), 0);
            uint32_t rr[4] = {r.x, r.y, r.z, r.w};
            for (int j = 0; j < 4 && i + j < nw; j++) blob[u.w_off + i + j] = (mz_u32_to_unit(rr[j]) - 0.5f) * scale;
        }
        for (int o = 0; o < u.cout; o++) {
            blob[u.b_off + o] = 0.0f;
            if (u.kind == 0) { blob[u.beta_off + o] = 0.0f; blob[u.gamma_off + o] = 1.0f; blob[u.mu_off + o] = 0.0f; blob[u.var_off + o] = 1.0f; }
        }
    }
}

}  // namespace mzh
