// host_harness.cpp -- TEST ONLY.  Compiles the product's scalar device building blocks (csrc/mz_common.h,
// csrc/mz_host.h) with g++ and replays, on the CPU, exactly what one thread of the CUDA kernels does
// (mz_k_search / mz_k_replay_gather).  tests/test_host_harness.py compares it with the oracle, so the tree logic,
// weight packing, tables and RNG keying are validated before any GPU time is spent.  It is never shipped or
// linked into libmuzero_b200.so; the CUDA kernels remain the only product compute path.
#include <cstdio>
#include <vector>
#include "../muzero.jl_b200/csrc/mz_host.h"
#include "../muzero.jl_b200/csrc/mz_rn_host.h"

namespace {
struct net_runner {
    const mz_params &P; std::vector<float> dev;
    net_runner(const mz_params &P_, const float *src) : P(P_), dev((size_t)P_.total_floats) { mzh::pack_weights(P, src, dev.data()); }
    void dense(int layer, const float *x, float *y) const {   // same arithmetic as mz_dense_tile, one row
        const mz_layer &L = P.layers[layer];
        for (int o = 0; o < L.out_pad; o++) {
            float acc = 0.0f;
            for (int k = 0; k < L.in; k++) acc = fmaf(dev[(size_t)L.w_off + (size_t)k * L.out_pad + o], x[k], acc);
            float v = acc + dev[(size_t)L.b_off + o];
            if (L.bn) v = mz_batchnorm(v, dev[(size_t)L.b_off + L.out_pad + o], dev[(size_t)L.b_off + 2 * L.out_pad + o], dev[(size_t)L.b_off + 3 * L.out_pad + o], dev[(size_t)L.b_off + 4 * L.out_pad + o]);
            y[o] = mz_activate(v, L.act);
        }
    }
    void chain(int first, int n, const float *x, float *y) const {
        std::vector<float> a(512), b(512);
        const float *cur = x;
        for (int i = 0; i < n; i++) { float *d = (i == n - 1) ? y : ((i & 1) ? b.data() : a.data()); dense(first + i, cur, d); cur = d; }
    }
    void net(int n, const float *x, float *h1, float *h2) const {
        const mz_net &N = P.nets[n];
        if (N.n_h1 == 0) { chain(N.first, N.n_trunk, x, h1); return; }
        std::vector<float> t(512);
        chain(N.first, N.n_trunk, x, t.data());
        chain(N.first + N.n_trunk, N.n_h1, t.data(), h1);
        chain(N.first + N.n_trunk + N.n_h1, N.n_h2, t.data(), h2);
    }
};

struct search_out { int32_t vc[MZ_MAX_A]; float rv; float priors[MZ_MAX_A]; double depth_sum; };

// one thread of mz_k_search
void search_one(const mzh::model &M, const net_runner &nn, std::vector<char> &pool, const float *stacked, uint32_t legal, int to_play,
                int exploration, uint32_t game, uint32_t move, search_out &out) {
    const mz_params &P = M.P;
    mz_tree tree = mz_tree_at(P, pool.data(), 0);
    std::vector<float> h(512), in1(512), in0(512), outV(8), outL(32), outH(512), outR(8);
    nn.net(0, stacked, outH.data(), nullptr);
    nn.net(1, outH.data(), outV.data(), outL.data());
    float logits[MZ_MAX_A], policy[MZ_MAX_A];
    mz_minmax mm; mm.mn = INFINITY; mm.mx = -INFINITY;
    for (int k = 0; k < P.hidden; k++) tree.hidden[k] = outH[(size_t)k];
    for (int i = 0; i < P.A; i++) logits[i] = outL[(size_t)i];
    mz_softmax(logits, P.A, policy);
    mz_f4 root; root.x = mz_bits2f(mz_nx_pack(0, -1, 0)); root.y = 0.0f; root.z = 0.0f; root.w = 0.0f;
    tree.A[0] = root;
    mz_tree_expand(P, tree, 0, 0, legal, policy, 0.0f, 0.0f);
    if (exploration && P.exploration_eps != 0.0f) mz_tree_add_noise(P, tree, legal, game, move);
    out.depth_sum = 0;
    std::vector<uint16_t> path((size_t)P.S + 2);
    for (int sim = 1; sim <= P.S; sim++) {
        mz_leaf leaf = mz_tree_select(P, tree, M.pbc0.data(), M.sqrtN.data(), legal, mm, game, move, (uint32_t)sim, path.data());
        out.depth_sum += leaf.depth;
        int pe = mz_nx_exp(leaf.parent_x), dbl = mz_nx_dbl(leaf.parent_x);
        const float *hp = tree.hidden + (size_t)pe * P.hidden_pad;
        float sc = mz_bits2f((uint32_t)(127 + dbl) << 23);
        for (int k = 0; k < P.hidden; k++) { float v = hp[k] * sc; in1[(size_t)k] = v; in0[(size_t)k] = v * 2.0f; }
        float plane = P.act_plane_play[leaf.action];
        for (int k = P.obs_size; k < P.sa_size; k++) in0[(size_t)k] = plane;
        { mz_f4 pr = tree.A[leaf.parent]; pr.x = mz_bits2f(mz_f2bits(pr.x) + (1u << 24)); tree.A[leaf.parent] = pr; }   // one more in-place doubling
        nn.net(1, in1.data(), outV.data(), outL.data());
        nn.net(2, in0.data(), outH.data(), outR.data());
        float *nh = tree.hidden + (size_t)sim * P.hidden_pad;
        for (int k = 0; k < P.hidden; k++) nh[k] = outH[(size_t)k];
        for (int i = 0; i < P.A; i++) logits[i] = outL[(size_t)i];
        mz_softmax(logits, P.A, policy);
        mz_tree_expand(P, tree, leaf.node, sim, legal, policy, outR[0], leaf.prior);
        mz_tree_backup(P, tree, path.data(), leaf.depth, outV[0], mm);
    }
    for (int i = 0; i < P.A; i++) {
        out.vc[i] = ((legal >> i) & 1u) ? (int32_t)mz_nx_visit(mz_f2bits(tree.A[1 + i].x)) : 0;
        out.priors[i] = ((legal >> i) & 1u) ? tree.A[1 + i].z : 0.0f;
    }
    mz_f4 r = tree.A[0]; int rvc = mz_nx_visit(mz_f2bits(r.x));
    out.rv = rvc == 0 ? 0.0f : r.y / (float)rvc;
}
}  // namespace

extern "C" {

int hh_default_config(mz_config *c) { mzh::default_config(c); return 0; }
int hh_num_params(const mz_config *c) { mzh::model M; if (mzh::build_model(*c, M)) return -1; return M.P.n_params; }
int hh_init_weights(const mz_config *c, uint64_t seed, float *src) { mzh::model M; if (mzh::build_model(*c, M)) return -1; mzh::init_weights(M.P, seed, src); return 0; }

int hh_nn(const mz_config *c, const float *src, int net, const float *in, float *o1, float *o2) {
    mzh::model M; if (const char *e = mzh::build_model(*c, M)) { fprintf(stderr, "%s\n", e); return -1; }
    net_runner nn(M.P, src);
    std::vector<float> a(512), b(512);
    nn.net(net, in, a.data(), b.data());
    if (net == 1) { o1[0] = a[0]; float pol[MZ_MAX_A]; mz_softmax(b.data(), M.P.A, pol); for (int i = 0; i < M.P.A; i++) o2[i] = pol[i]; }
    else { for (int k = 0; k < M.P.hidden; k++) o1[k] = a[(size_t)k]; if (net == 2) o2[0] = b[0]; }
    return 0;
}

int hh_run_mcts(const mz_config *c, const float *src, const float *stacked, uint32_t legal, int to_play, int exploration, uint64_t game, int move,
                int32_t *vc, float *rv, float *priors) {
    mzh::model M; if (const char *e = mzh::build_model(*c, M)) { fprintf(stderr, "%s\n", e); return -1; }
    net_runner nn(M.P, src);
    std::vector<char> pool((size_t)M.P.tree_stride_bytes + 256);
    search_out o;
    search_one(M, nn, pool, stacked, legal, to_play, exploration, (uint32_t)game, (uint32_t)move, o);
    for (int i = 0; i < M.P.A; i++) { vc[i] = o.vc[i]; if (priors) priors[i] = o.priors[i]; }
    *rv = o.rv;
    return 0;
}

// MODE_SLOTS thread of mz_k_search looped over the moves of each game, outputs laid out like the oracle's mzo_self_play
int64_t hh_self_play(const mz_config *c, const float *src, uint64_t first_game, int n_games, float temperature, int32_t *T_out, float *obs,
                     int32_t *actions, float *rewards, int32_t *to_play, float *child_visits, float *root_values) {
    mzh::model M; if (const char *e = mzh::build_model(*c, M)) { fprintf(stderr, "%s\n", e); return -1; }
    const mz_params &P = M.P;
    net_runner nn(P, src);
    std::vector<char> pool((size_t)P.tree_stride_bytes + 256);
    int64_t sims = 0;
    for (int gi = 0; gi < n_games; gi++) {
        std::vector<uint64_t> h1((size_t)P.Tmax, 0), h2((size_t)P.Tmax, 0); std::vector<int32_t> ha((size_t)P.Tmax, 0);
        mz_board b; mz_env_reset_b(P, b);
        int T = 0; bool finished = false;
        uint32_t game = (uint32_t)(first_game + (uint64_t)gi);
        while (!finished) {
            uint32_t legal = mz_env_legal_b(P, b);
            std::vector<float> stacked((size_t)P.stack_size);
            for (int k = 0; k < P.stack_size; k++) stacked[(size_t)k] = mz_stacked_value(P, h1.data(), h2.data(), ha.data(), T + 1, k);
            search_out o;
            search_one(M, nn, pool, stacked.data(), legal, b.player, 1, game, (uint32_t)T + 1u, o);
            sims += P.S;
            int sum_visits = 0; for (int i = 0; i < P.A; i++) sum_visits += o.vc[i];
            int action = mz_select_action_counts(P, o.vc, legal, mz_play_temperature(P, T, temperature), game, (uint32_t)T + 1u);
            int p = b.player;
            size_t oo = (size_t)gi * P.Tmax + T;
            for (int k = 0; k < P.obs_size; k++) obs[oo * P.obs_size + k] = mz_env_obs_value(P, b, k / P.cells, k % P.cells);
            mz_env_step_b(P, b, action);
            float reward = (float)mz_env_reward_b(P, b, p);
            bool done = mz_env_terminated_b(P, b);
            for (int i = 0; i < P.A; i++) child_visits[oo * P.A + i] = ((legal >> i) & 1u) ? (float)((double)o.vc[i] / (double)sum_visits) : 0.0f;
            root_values[oo] = o.rv; actions[oo] = action; rewards[oo] = reward; to_play[oo] = p;
            ha[(size_t)T] = action;
            T += 1;
            if (T < P.Tmax) { h1[(size_t)T] = b.p1; h2[(size_t)T] = b.p2; }
            if (done || T > P.max_moves) finished = true;
        }
        T_out[gi] = T;
    }
    return sims;
}

// thread 0 + gather threads of mz_k_replay_gather; the buffer is given like the oracle's (float observations)
int hh_get_batch(const mz_config *c, int n_games, int64_t first_key, const int32_t *T, const float *obs, const int32_t *actions, const float *rewards,
                 const int32_t *to_play, const float *child_visits, const float *root_values, uint64_t step, int32_t *index_batch, float *obs_batch,
                 float *action_batch, float *value_batch, float *reward_batch, float *policy_batch, float *gscale) {
    mzh::model M; if (const char *e = mzh::build_model(*c, M)) { fprintf(stderr, "%s\n", e); return -1; }
    const mz_params &P = M.P; const int K1 = P.K + 1;
    for (int b = 0; b < c->batch_size; b++) {
        mz_u4 q = mz_philox(P.seed, MZ_STREAM_REPLAY, (uint32_t)step, (uint32_t)b, 0, 0);
        int gi = (int)mz_u32_below(q.x, (uint32_t)n_games);
        int Tg = T[gi];
        int pos = 1 + (int)mz_u32_below(q.y, (uint32_t)Tg);
        index_batch[2 * b] = (int32_t)(first_key + gi); index_batch[2 * b + 1] = pos;
        std::vector<uint64_t> h1((size_t)P.Tmax, 0), h2((size_t)P.Tmax, 0); std::vector<uint8_t> tp((size_t)P.Tmax, 0);
        for (int t = 0; t < P.Tmax; t++) {
            const float *o = obs + ((size_t)gi * P.Tmax + t) * P.obs_size;
            for (int cell = 0; cell < P.cells; cell++) { if (o[cell] != 0.0f) h1[(size_t)t] |= 1ull << cell; if (o[P.cells + cell] != 0.0f) h2[(size_t)t] |= 1ull << cell; }
            tp[(size_t)t] = (uint8_t)to_play[(size_t)gi * P.Tmax + t];
        }
        const float *rew = rewards + (size_t)gi * P.Tmax, *rv = root_values + (size_t)gi * P.Tmax; const int32_t *act = actions + (size_t)gi * P.Tmax;
        for (int k = 0; k < K1; k++) {
            int ci = pos + k; float tv, tr; int a;
            if (ci < Tg) { tv = mz_target_value(P, Tg, rew, tp.data(), rv, ci); tr = rew[ci - 1]; a = act[ci - 1]; }
            else if (ci == Tg) { tv = 0.0f; tr = rew[ci - 1]; a = act[ci - 1]; }
            else { tv = 0.0f; tr = 0.0f; a = 1 + (int)mz_u32_below(mz_philox(P.seed, MZ_STREAM_ABSORB, (uint32_t)step, (uint32_t)b, (uint32_t)k, 0).x, (uint32_t)P.A); }
            value_batch[(size_t)b * K1 + k] = tv; reward_batch[(size_t)b * K1 + k] = tr; action_batch[(size_t)b * K1 + k] = (float)a;
            for (int i = 0; i < P.A; i++) policy_batch[((size_t)b * K1 + k) * P.A + i] = ci < Tg ? child_visits[((size_t)gi * P.Tmax + ci - 1) * P.A + i] : 1.0f / (float)P.A;
        }
        int gs = Tg + 1 - pos; if (P.K < gs) gs = P.K;
        gscale[b] = (float)gs;
        for (int k = 0; k < P.stack_size; k++) obs_batch[(size_t)b * P.stack_size + k] = mz_stacked_value(P, h1.data(), h2.data(), act, pos, k);
    }
    return 0;
}


}  // extern "C"

// ---- ResNet step program replayed on the CPU ------------------------------------------------------------------------
// A scalar interpreter of the host-built program (mzh::rn_build) and weight image (mzh::rn_pack): exactly the data flow of
// mz_k_search_rn / mz_k_rn_forward -- bf16 activation tiles [128 rows][64 channels], B images read back through the swizzle,
// fp32 accumulation, folded affine epilogue, residual, head scatter, dense heads -- so that the program builder and the packer
// are validated against the oracle without a GPU.  net: 0 representation(stacked), 1 prediction(hidden), 2 dynamics(sa).
namespace {
inline float bf16f(float v) { uint16_t h = mzh::f2bf16(v); uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
inline float img_bf16(const unsigned char *img, int row, int k) { uint16_t h; memcpy(&h, img + mzh::tc_tile_offset(row, k), 2); uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
struct rn_cpu {
    const mzh::rn_model &M; const mz_params &P; const std::vector<unsigned char> &image;
    std::vector<float> tile[8];               // X0..3, T0..3: [128][64]
    std::vector<float> hv, hp;                // head tiles [64 trees][64], [64][128]
    std::vector<float> out, plane; std::vector<float> pool;   // out: V[4][64] | L[16][64] | R[4][64]; pool: [ntrees][cells][64]
    std::vector<double> acc[4];
    rn_cpu(const mzh::rn_model &M_, const mz_params &P_, const std::vector<unsigned char> &img) : M(M_), P(P_), image(img) {
        for (auto &t : tile) t.assign(128 * 64, 0.0f);
        hv.assign(64 * 64, 0.0f); hp.assign(64 * 128, 0.0f); out.assign(24 * 64, 0.0f); plane.assign(64, 0.0f);
        pool.assign((size_t)64 * M.R.cells * 64, 0.0f);
        for (auto &a : acc) a.assign(128 * 64, 0.0);
    }
    float *buf(int id, int kb, int row) { if (id < 8) return &tile[id][(size_t)row * 64]; if (id == MZ_RN_BUF_HV) return &hv[(size_t)row * 64]; return &hp[(size_t)row * 128 + 64 * kb]; }
    void run(int first, int last, int ntrees) {
        const mz_rn_params &R = M.R;
        for (int si = first; si < last; si++) {
            const mz_rn_step &st = M.steps[(size_t)si];
            const unsigned char *blk = image.data() + st.w_off;
            for (int j = 0; j < st.njobs; j++) {
                const mz_rn_job &J = st.jobs[j];
                const int N = 16 * J.n16;
                const bool trees = (J.flags & MZ_RN_F_TREES) != 0;
                std::vector<double> &A = acc[J.acc];
                if (!st.accumulate) std::fill(A.begin(), A.end(), 0.0);
                for (int row = 0; row < 128; row++) {
                    if (trees ? row >= 64 : row >= R.rows_valid) continue;
                    for (int kb = 0; kb < J.kblocks; kb++) {
                        const float *a;
                        std::vector<float> shifted(64, 0.0f);
                        if (st.ntaps > 1) {   // tap-step: source row shifted by (dx, dy) cells
                            const int t = row / R.cells, cell = row % R.cells, x = cell % P.W + st.dx, y = cell / P.W + st.dy;
                            if (x >= 0 && x < P.W && y >= 0 && y < P.H) { const float *src = buf(J.a_buf, 0, t * R.cells + x + P.W * y); for (int k = 0; k < 64; k++) shifted[(size_t)k] = src[k]; }
                            a = shifted.data();
                        } else a = buf(J.a_buf, kb, row);
                        const unsigned char *img = blk + J.w_sub + kb * (J.n16 * 2048);
                        for (int n = 0; n < N; n++) { double s = 0.0; for (int k = 0; k < 64; k++) s += (double)a[k] * (double)img_bf16(img, n, k); A[(size_t)row * 64 + n] += s; }
                    }
                }
                if (!st.last) continue;
                const float *S = reinterpret_cast<const float *>(blk + J.p_sub), *T = S + 64, *E = S + 128;
                for (int row = 0; row < 128; row++) {
                    int tree, cell;
                    if (trees) { tree = row; cell = 0; if (row >= 64) continue; } else { if (row >= R.rows_valid) continue; tree = J.acc * R.tpt + row / R.cells; cell = row % R.cells; }
                    const bool valid = tree < ntrees;
                    if (J.epi == MZ_RN_EPI_TILE) {
                        float *dst = buf(J.dst_buf, 0, row); const float *skp = J.skip_buf != 0xff ? buf(J.skip_buf, 0, row) : nullptr;
                        float y[64];
                        for (int c = 0; c < 64; c++) {
                            float v = (float)A[(size_t)row * 64 + c] + ((J.flags & MZ_RN_F_PLANE) ? fmaf(plane[(size_t)tree], E[c], T[c]) : T[c]);
                            if (skp) v = v + skp[c];
                            if (J.act == MZ_ACT_RELU) v = v > 0.0f ? v : 0.0f;
                            y[c] = valid ? bf16f(v) : 0.0f;
                        }
                        for (int c = 0; c < 64; c++) dst[c] = y[c];
                        if ((J.flags & MZ_RN_F_POOL) && valid) for (int c = 0; c < 64; c++) pool[((size_t)tree * R.cells + cell) * 64 + c] = y[c];
                    } else if (J.epi == MZ_RN_EPI_HEAD) {
                        if (!valid) continue;
                        for (int f = 0; f < J.nfa + J.nfb; f++) {
                            float v = (float)A[(size_t)row * 64 + f] + T[f]; v = v > 0.0f ? v : 0.0f;
                            const bool second = f >= J.nfa; const int k = cell + R.cells * (second ? f - J.nfa : f);
                            buf(second ? J.dst2_buf : J.dst_buf, k >> 6, tree)[k & 63] = bf16f(v);
                        }
                    } else {
                        if (row >= R.ntrees) continue;
                        const int base = J.out_id == MZ_RN_OUT_V ? 0 : J.out_id == MZ_RN_OUT_L ? 4 * 64 : 20 * 64;
                        for (int k = 0; k < J.out; k++) { float v = (float)A[(size_t)row * 64 + k] + T[k]; out[(size_t)base + k * 64 + row] = mz_activate(v, J.act); }
                    }
                }
            }
        }
    }
};
}  // namespace

extern "C" {
int hh_rn_num_params(const mz_config *c) { mzh::rn_model M; mzh::rn_units_build(*c, M); return mzh::rn_total_params(M); }
int hh_rn_program_info(const mz_config *c, int32_t *out /* [8]: n_steps, repr steps, pred steps, dyn steps, image bytes, slot bytes, ntrees, tree stride */) {
    mzh::model Mm; if (const char *e = mzh::build_model(*c, Mm)) { fprintf(stderr, "%s\n", e); return -1; }
    mzh::rn_model M; if (const char *e = mzh::rn_build(*c, Mm.P, M)) { fprintf(stderr, "%s\n", e); return -1; }
    const mz_rn_params &R = M.R;
    out[0] = R.n_steps; out[1] = R.prog_repr[1] - R.prog_repr[0]; out[2] = R.prog_pred[1] - R.prog_pred[0]; out[3] = R.prog_dyn[1] - R.prog_dyn[0];
    out[4] = R.image_bytes; out[5] = R.slot_bytes; out[6] = R.ntrees; out[7] = R.tree_stride_bytes;
    for (const auto &s : M.steps) if (s.w_bytes > R.slot_bytes || s.w_off + s.w_bytes > R.image_bytes || s.njobs < 1 || s.njobs > MZ_RN_TILES) return -2;
    return 0;
}
// select_opponent_action / arena outcome: the product's scalar code (the kernels call the same functions)
int hh_opponent_action(const mz_config *c, uint64_t p1, uint64_t p2, int player, int opponent, uint64_t game, int move) {
    mzh::model M; if (mzh::build_model(*c, M)) return -1;
    mz_board b; b.p1 = p1; b.p2 = p2; b.player = player;
    return mz_opponent_action(M.P, b, opponent, (uint32_t)game, (uint32_t)move);
}
int hh_board_after(const mz_config *c, int n, const int32_t *actions, uint64_t *p1, uint64_t *p2, int32_t *player) {
    mzh::model M; if (mzh::build_model(*c, M)) return -1;
    mz_board b; mz_env_reset_b(M.P, b);
    for (int i = 0; i < n; i++) mz_env_step_b(M.P, b, actions[i]);
    *p1 = b.p1; *p2 = b.p2; *player = b.player;
    return 0;
}
int hh_arena_outcome(const mz_config *c, int T, const int32_t *actions, int muzero_player) {
    mzh::model M; if (mzh::build_model(*c, M)) return -99;
    return mz_arena_outcome(M.P, T, actions, muzero_player);
}
// one line per step of the program (debugging aid; also pins the step count the kernels are timed with)
int hh_rn_program_dump(const mz_config *c) {
    mzh::model Mm; if (const char *e = mzh::build_model(*c, Mm)) { fprintf(stderr, "%s\n", e); return -1; }
    mzh::rn_model M; if (const char *e = mzh::rn_build(*c, Mm.P, M)) { fprintf(stderr, "%s\n", e); return -1; }
    const mz_rn_params &R = M.R;
    printf("repr [%d,%d) pred [%d,%d) dyn [%d,%d) smem_first %d slot %d B\n", R.prog_repr[0], R.prog_repr[1], R.prog_pred[0], R.prog_pred[1], R.prog_dyn[0], R.prog_dyn[1], R.smem_first, R.slot_bytes);
    for (size_t i = 0; i < M.steps.size(); i++) {
        const mz_rn_step &s = M.steps[i];
        printf("step %2zu w %6d B njobs %d taps %d/%d last %d :", i, s.w_bytes, s.njobs, s.tap, s.ntaps, s.last);
        for (int j = 0; j < s.njobs; j++) { const mz_rn_job &J = s.jobs[j]; printf(" [a%d->d%d/%d skip %d epi %d n16 %d kb %d act %d wg %d acc %d fl %d]", J.a_buf, J.dst_buf, J.dst2_buf, J.skip_buf, J.epi, J.n16, J.kblocks, J.act, J.wg, J.acc, J.flags); }
        printf("\n");
    }
    return 0;
}
// n <= ntrees inputs through one network; out1 / out2 like the C ABI callables (hidden in Julia (W,H,nf) order)
int hh_rn_forward(const mz_config *c, const float *blob, int net, int n, const float *in, float *out1, float *out2) {
    mzh::model Mm; if (const char *e = mzh::build_model(*c, Mm)) { fprintf(stderr, "%s\n", e); return -1; }
    mzh::rn_model M; if (const char *e = mzh::rn_build(*c, Mm.P, M)) { fprintf(stderr, "%s\n", e); return -1; }
    const mz_params &P = Mm.P; const mz_rn_params &R = M.R;
    if (n > R.ntrees) return -3;
    std::vector<unsigned char> image; mzh::rn_pack(M, blob, image);
    rn_cpu X(M, P, image);
    if (net == 0) {   // im2col tiles of the first convolution
        const int k = R.ksize, pad = k / 2;
        for (int t = 0; t < n; t++) for (int cell = 0; cell < R.cells; cell++) for (int tap = 0; tap < k * k; tap++) for (int pl = 0; pl < R.planes; pl++) {
            const int x = cell % P.W + pad - (tap % k), y = cell / P.W + pad - (tap / k);
            float v = 0.0f;
            if (x >= 0 && x < P.W && y >= 0 && y < P.H) v = in[(size_t)t * P.stack_size + (x + P.W * y) + P.cells * pl];
            X.tile[t / R.tpt][(size_t)((t % R.tpt) * R.cells + cell) * 64 + tap * R.planes + pl] = bf16f(v);
        }
        X.run(R.prog_repr[0], R.prog_repr[1], n);
    } else {
        const int in_dim = net == 2 ? P.sa_size : P.hidden; const float mul = net == 2 ? 0.5f : 1.0f;
        for (int t = 0; t < n; t++) {
            for (int cell = 0; cell < R.cells; cell++) for (int ch = 0; ch < R.nf; ch++)
                X.tile[t / R.tpt][(size_t)((t % R.tpt) * R.cells + cell) * 64 + ch] = bf16f(in[(size_t)t * in_dim + cell + P.cells * ch] * mul);
            if (net == 2) X.plane[(size_t)t] = in[(size_t)t * in_dim + P.hidden];
        }
        if (net == 1) X.run(R.prog_pred[0], R.prog_pred[1], n); else X.run(R.prog_dyn[0], R.prog_dyn[1], n);
    }
    if (net != 1) {
        for (int t = 0; t < n; t++) for (int k = 0; k < P.hidden; k++) out1[(size_t)t * P.hidden + k] = X.pool[((size_t)t * R.cells + k % P.cells) * 64 + k / P.cells];
        if (net == 2) for (int t = 0; t < n; t++) out2[t] = X.out[20 * 64 + t];
    } else {
        for (int t = 0; t < n; t++) {
            float logits[MZ_MAX_A], policy[MZ_MAX_A];
            for (int i = 0; i < P.A; i++) logits[i] = X.out[4 * 64 + i * 64 + t];
            mz_softmax(logits, P.A, policy);
            out1[t] = X.out[t];
            for (int i = 0; i < P.A; i++) out2[(size_t)t * P.A + i] = policy[i];
        }
    }
    return 0;
}

}  // extern "C"
