/*
 * mz_oracle.c -- CPU restatement of deveshjawla/MuZero.jl's hot path.  TEST INFRASTRUCTURE ONLY.
 * See mz_oracle.h for the parity status ("parity unpinned") and the arithmetic contract.
 * All file:line citations are relative to /root/reference.
 *
 * Build: gcc -O3 -ffp-contract=off -fno-math-errno -mavx2 -mfma -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: every fused multiply-add in the contract is an explicit fmaf();
 * every other a*b+c is two roundings, exactly like Julia without @fastmath.
 */
#include "mz_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011).  Replaces the reference's unseeded global RNG / the
 * MersenneTwister(1234) at src/SelfPlay.jl:152, which cannot be reproduced in a batched setting.
 * ------------------------------------------------------------------------------------------ */
enum { STREAM_TIE = 1, STREAM_ACTION = 2, STREAM_DIRICHLET = 3, STREAM_REPLAY = 4, STREAM_ABSORB = 5, STREAM_INIT = 7, STREAM_OPPONENT = 8 };

void mzo_philox(uint64_t seed, uint32_t stream, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ stream;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline float u32_to_unit(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; } /* [0,1) */
static inline uint32_t u32_below(uint32_t x, uint32_t n) { return (uint32_t)(((uint64_t)x * n) >> 32); }

/* ------------------------------------------------------------------------------------------
 * Math contract (stands in for Julia Base exp/log/tanh on Float32; parity unpinned).
 * ------------------------------------------------------------------------------------------ */
static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

float mzo_expf(float x) {
    if (x > 88.0f) x = 88.0f;
    if (x < -87.0f) x = -87.0f;
    float fn = rintf(x * 1.44269504088896341f);
    float r = fmaf(fn, -0.693359375f, x);
    r = fmaf(fn, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r) + 1.0f;
    int n = (int)fn;
    return y * bits2f((uint32_t)(n + 127) << 23);
}

float mzo_logf(float x) {
    /* x > 0, normal */
    uint32_t u = f2bits(x);
    int e = (int)((u >> 23) & 0xff) - 126;
    float m = bits2f((u & 0x007fffffu) | 0x3f000000u); /* [0.5, 1) */
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float y = 7.0376836292e-2f;
    y = fmaf(y, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = (y * m) * z;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(z, -0.5f, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

float mzo_tanhf(float x) {
    float z = fabsf(x);
    if (z > 44.0f) return x > 0.0f ? 1.0f : -1.0f;
    if (z >= 0.625f) {
        float s = mzo_expf(z + z);
        z = 1.0f - 2.0f / (s + 1.0f);
        return x < 0.0f ? -z : z;
    }
    float w = x * x;
    float p = -5.70498872745e-3f;
    p = fmaf(p, w, 2.06390887954e-2f);
    p = fmaf(p, w, -5.37397155531e-2f);
    p = fmaf(p, w, 1.33314422036e-1f);
    p = fmaf(p, w, -3.33332819422e-1f);
    return fmaf(p * w, x, x);
}

/* x^e for select_action's visit_counts .^ (1/temperature) (src/SelfPlay.jl:301): exact repeated
 * product when e is a small integer (T = 1, 0.5, 0.25 give e = 1, 2, 4), else exp(e*log(x)). */
static float pow_contract(float x, float e) {
    if (x == 0.0f) return 0.0f;
    if (e == rintf(e) && e >= 1.0f && e <= 16.0f) {
        float r = x;
        for (int i = 1; i < (int)e; i++) r = r * x;
        return r;
    }
    return mzo_expf(e * mzo_logf(x));
}

/* ------------------------------------------------------------------------------------------
 * Config
 * ------------------------------------------------------------------------------------------ */
/* Julia (<= 1.10) Dict{Int,V} iteration order for keys 1..A inserted ascending into the default
 * 16-slot table: slot = hash_64_64(3|x| + bits(Float64(x))) & 15, linear probing.  SURVEY Q9. */
static uint64_t hash_64_64(uint64_t a) {
    a = ~a + (a << 21);
    a = a ^ (a >> 24);
    a = a + (a << 3) + (a << 8);
    a = a ^ (a >> 14);
    a = a + (a << 2) + (a << 4);
    a = a ^ (a >> 28);
    a = a + (a << 31);
    return a;
}
void mzo_julia_dict_order(int A, int32_t *order) {
    int sz = 16;
    while (A * 3 > sz * 2) sz *= 4; /* Dict grows x4 once count*3 > sz*2 (small tables) */
    int *slots = (int *)calloc((size_t)sz, sizeof(int));
    for (int k = 1; k <= A; k++) {
        double d = (double)k; uint64_t b; memcpy(&b, &d, 8);
        uint64_t h = hash_64_64(3ull * (uint64_t)k + b);
        int i = (int)(h & (uint64_t)(sz - 1));
        while (slots[i]) i = (i + 1) & (sz - 1);
        slots[i] = k;
    }
    int n = 0;
    for (int i = 0; i < sz; i++) if (slots[i]) order[n++] = slots[i];
    free(slots);
}

void mzo_default_config(mzo_config *c) { /* games/tictactoe/params.jl:2-29; src/Constructors.jl:18-52 */
    memset(c, 0, sizeof(*c));
    c->game = MZO_GAME_TICTACTOE;
    c->W = 3; c->H = 3; c->C = 3; c->A = 9; c->num_players = 2;
    c->stacked_observations = 1; c->max_moves = 9; c->num_iters = 10;
    c->num_unroll_steps = 5; c->td_steps = 5; c->batch_size = 32; c->replay_buffer_size = 10000;
    c->pb_c_base = 19652; c->pb_c_init = 1.25f; c->discount = 0.997f;
    c->dirichlet_alpha = 0.25f; c->exploration_eps = 0.25f;
    c->intermediate_rewards = 0; c->tie_mode = MZO_TIE_PHILOX; c->seed = 1337;
    mzo_julia_dict_order(c->A, c->child_order);
    c->width_hidden = 64; c->depth_representation = 3; c->depth_prediction = 3; c->depth_dynamics = 3;
    c->depth_policy = 1; c->depth_value = 1; c->depth_reward = 1; c->depth_state_head = 3;
    c->hidden_state_size = 27; c->reward_activation_tanh = 1;
    c->per = 0; c->per_alpha = 1; c->temperature_threshold = -1;
    c->net_type = 0; c->rn_num_blocks = 2; c->rn_num_filters = 64; c->rn_kernel = 3; c->rn_first_head_filters = 1; c->rn_second_head_filters = 2;
}

/* ------------------------------------------------------------------------------------------
 * Networks: init_representation / init_prediction / init_dynamics (src/Learning.jl:87-142).
 * Blob layout: nets in order (representation, prediction, dynamics); inside a net the Dense layers
 * in Flux.params order (trunk, then Split path 1, then Split path 2); per layer W (out,in) in
 * Julia column-major (W[o + out*k]) followed by b[out].
 * FeedForwardHP.use_batch_norm (Constructors.jl:71): make_dense (Learning.jl:70-79) is Chain(Dense(in, out), BatchNorm(out, relu)); such a
 * layer is followed in the blob by BatchNorm's beta[out], gamma[out] (Flux.trainable(BatchNorm) = (beta, gamma), Flux.params order) and
 * its running statistics mu[out], sigma2[out] (not parameters).  The reference never differentiates a forward pass (Q20) and never calls
 * trainmode!, so BatchNorm always runs in test mode: relu.(gamma .* (x .- mu) ./ sqrt.(sigma2 .+ 1f-5) .+ beta) (Flux 0.12.4).
 * ------------------------------------------------------------------------------------------ */
enum { ACT_ID = 0, ACT_RELU = 1, ACT_TANH = 2 };
typedef struct { int in, out, act, w_off, b_off, bn, beta_off, gamma_off, mu_off, var_off; } layer_t;
typedef struct { int n_trunk, n_h1, n_h2; layer_t trunk[24], h1[24], h2[24]; int base, n_params; } net_t;

static int g_add_bn = 0;   /* build_net: the layer being added comes from make_dense with use_batch_norm */
static int add_layer(layer_t *l, int in, int out, int act, int *off) {
    l->in = in; l->out = out; l->act = act; l->w_off = *off; *off += in * out; l->b_off = *off; *off += out;
    l->bn = g_add_bn && act == ACT_RELU;   /* make_dense layers are exactly the relu layers; the final Dense of every chain has no BatchNorm */
    l->beta_off = l->gamma_off = l->mu_off = l->var_off = 0;
    if (l->bn) { l->beta_off = *off; *off += out; l->gamma_off = *off; *off += out; l->mu_off = *off; *off += out; l->var_off = *off; *off += out; }
    return 1;
}
static int obs_planes(const mzo_config *c) { return c->C * (c->stacked_observations + 1) + c->stacked_observations; }
static int stack_size(const mzo_config *c) { return c->W * c->H * obs_planes(c); }
static int obs_size(const mzo_config *c) { return c->W * c->H * c->C; }
static int sa_size(const mzo_config *c) { return c->W * c->H * (c->C + 1); }

static void build_net(const mzo_config *c, int which, net_t *n, int base) {
    int off = base, w = c->width_hidden;
    memset(n, 0, sizeof(*n)); n->base = base;
    g_add_bn = c->use_batch_norm != 0;
    if (which == 0) { /* Learning.jl:87-98 */
        n->n_trunk += add_layer(&n->trunk[n->n_trunk], stack_size(c), w, ACT_RELU, &off);
        for (int i = 0; i < c->depth_representation; i++) n->n_trunk += add_layer(&n->trunk[n->n_trunk], w, w, ACT_RELU, &off);
        n->n_trunk += add_layer(&n->trunk[n->n_trunk], w, c->hidden_state_size, ACT_ID, &off);
    } else if (which == 1) { /* Learning.jl:100-116 */
        n->n_trunk += add_layer(&n->trunk[n->n_trunk], c->hidden_state_size, w, ACT_RELU, &off);
        for (int i = 0; i < c->depth_prediction; i++) n->n_trunk += add_layer(&n->trunk[n->n_trunk], w, w, ACT_RELU, &off);
        for (int i = 0; i < c->depth_value; i++) n->n_h1 += add_layer(&n->h1[n->n_h1], w, w, ACT_RELU, &off);
        n->n_h1 += add_layer(&n->h1[n->n_h1], w, 1, ACT_TANH, &off);
        for (int i = 0; i < c->depth_policy; i++) n->n_h2 += add_layer(&n->h2[n->n_h2], w, w, ACT_RELU, &off);
        n->n_h2 += add_layer(&n->h2[n->n_h2], w, c->A, ACT_ID, &off); /* softmax applied by caller */
    } else { /* Learning.jl:118-142 */
        n->n_trunk += add_layer(&n->trunk[n->n_trunk], sa_size(c), w, ACT_RELU, &off);
        for (int i = 0; i < c->depth_dynamics; i++) n->n_trunk += add_layer(&n->trunk[n->n_trunk], w, w, ACT_RELU, &off);
        for (int i = 0; i < c->depth_state_head; i++) n->n_h1 += add_layer(&n->h1[n->n_h1], w, w, ACT_RELU, &off);
        n->n_h1 += add_layer(&n->h1[n->n_h1], w, c->hidden_state_size, ACT_ID, &off);
        for (int i = 0; i < c->depth_reward; i++) n->n_h2 += add_layer(&n->h2[n->n_h2], w, w, ACT_RELU, &off);
        n->n_h2 += add_layer(&n->h2[n->n_h2], w, 1, c->reward_activation_tanh ? ACT_TANH : ACT_ID, &off);
    }
    n->n_params = off - base;
}
static void build_nets(const mzo_config *c, net_t nets[3]) {
    build_net(c, 0, &nets[0], 0);
    build_net(c, 1, &nets[1], nets[0].n_params);
    build_net(c, 2, &nets[2], nets[0].n_params + nets[1].n_params);
}
static int rn_num_params(const mzo_config *c, int net);
static void rn_init_weights(const mzo_config *c, uint64_t seed, float *blob);
int mzo_num_params(const mzo_config *c, int net) {
    if (c->net_type == 1) return rn_num_params(c, net);
    net_t n[3]; build_nets(c, n);
    return net < 3 ? n[net].n_params : n[0].n_params + n[1].n_params + n[2].n_params;
}

/* Flux.glorot_uniform: (rand(Float32, out, in) .- 0.5f0) .* sqrt(24f0 / (in + out)); bias zeros.
 * The reference draws from the unseeded global RNG (conf.seed is never read, Constructors.jl:19);
 * contract: element i of layer l of net n = Philox(seed, INIT, n, l, i/4)[i%4]. */
void mzo_init_weights(const mzo_config *c, uint64_t seed, float *blob) {
    if (c->net_type == 1) { rn_init_weights(c, seed, blob); return; }
    net_t nets[3]; build_nets(c, nets);
    for (int n = 0; n < 3; n++) {
        int li = 0;
        for (int part = 0; part < 3; part++) {
            int cnt = part == 0 ? nets[n].n_trunk : part == 1 ? nets[n].n_h1 : nets[n].n_h2;
            layer_t *ls = part == 0 ? nets[n].trunk : part == 1 ? nets[n].h1 : nets[n].h2;
            for (int l = 0; l < cnt; l++, li++) {
                float scale = sqrtf(24.0f / (float)(ls[l].in + ls[l].out));
                int nw = ls[l].in * ls[l].out;
                for (int i = 0; i < nw; i += 4) {
                    uint32_t r[4]; mzo_philox(seed, STREAM_INIT, (uint32_t)n, (uint32_t)li, (uint32_t)(i / 4), 0, r);
                    for (int j = 0; j < 4 && i + j < nw; j++) blob[ls[l].w_off + i + j] = (u32_to_unit(r[j]) - 0.5f) * scale;
                }
                for (int o = 0; o < ls[l].out; o++) blob[ls[l].b_off + o] = 0.0f;
                if (ls[l].bn) for (int o = 0; o < ls[l].out; o++) { blob[ls[l].beta_off + o] = 0.0f; blob[ls[l].gamma_off + o] = 1.0f; blob[ls[l].mu_off + o] = 0.0f; blob[ls[l].var_off + o] = 1.0f; }   /* Flux.BatchNorm(out) */
            }
        }
    }
}

/* bf16 emulation for checking the tensor-core path (MZ_NN_BF16_TC): when enabled, every Dense rounds its input
 * vector and its weights to bfloat16 (round to nearest even) and accumulates in Float32; biases, activations and the
 * final outputs stay Float32.  Off by default (the reference computes in Float32). */
static int g_bf16 = 0;
void mzo_set_bf16(int on) { g_bf16 = on; }
static inline float bf16_round(float f) {
    uint32_t u = f2bits(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return f;
    u += 0x7fffu + ((u >> 16) & 1u);
    return bits2f(u & 0xffff0000u);
}
/* Dense (Flux 0.12.4, un-vendored): y = act.(W*x .+ b).  Contract: sequential-k fmaf, then + b. */
static void dense(const float *blob, const layer_t *l, const float *x, float *y) {
    float acc[256];
    const float *W = blob + l->w_off, *b = blob + l->b_off;
    int out = l->out;
    for (int o = 0; o < out; o++) acc[o] = 0.0f;
    if (g_bf16 >= 2) {
        /* split-precision operands (mzo_set_bf16(2 | 3)): x = xh + xl (+ xm), every part a bfloat16, products exact in Float32, the small
         * cross terms accumulated first and the hi*hi terms on top -- the order the tensor-core kernel issues its MMAs in.
         * 2: hi*lo + lo*hi + hi*hi (16 mantissa bits per operand); 3: all six terms down to 2^-24 (24 bits). */
        float xs[3][256];
        for (int k = 0; k < l->in; k++) {
            float h = bf16_round(x[k]), m = bf16_round(x[k] - h), lo = bf16_round((x[k] - h) - m);
            xs[0][k] = h; xs[1][k] = m; xs[2][k] = lo;
        }
        static const int t2[3][2] = {{0, 1}, {1, 0}, {0, 0}};
        static const int t3[6][2] = {{0, 2}, {2, 0}, {1, 1}, {0, 1}, {1, 0}, {0, 0}};
        const int nt = g_bf16 == 2 ? 3 : 6;
        for (int t = 0; t < nt; t++) {
            const int xi = g_bf16 == 2 ? t2[t][0] : t3[t][0], wi = g_bf16 == 2 ? t2[t][1] : t3[t][1];
            for (int k = 0; k < l->in; k++) {
                const float xk = xs[xi][k];
                const float *wk = W + (size_t)k * out;
                for (int o = 0; o < out; o++) {
                    float h = bf16_round(wk[o]), m = bf16_round(wk[o] - h), lo = bf16_round((wk[o] - h) - m);
                    acc[o] = fmaf(wi == 0 ? h : wi == 1 ? m : lo, xk, acc[o]);
                }
            }
        }
    } else if (g_bf16) {
        for (int k = 0; k < l->in; k++) {
            float xk = bf16_round(x[k]);
            const float *wk = W + (size_t)k * out;
            for (int o = 0; o < out; o++) acc[o] = fmaf(bf16_round(wk[o]), xk, acc[o]);
        }
    } else
    for (int k = 0; k < l->in; k++) {
        float xk = x[k];
        const float *wk = W + (size_t)k * out;
        for (int o = 0; o < out; o++) acc[o] = fmaf(wk[o], xk, acc[o]);
    }
    for (int o = 0; o < out; o++) {
        float v = acc[o] + b[o];
        if (l->bn) v = ((blob[l->gamma_off + o] * (v - blob[l->mu_off + o])) / sqrtf(blob[l->var_off + o] + 1e-5f)) + blob[l->beta_off + o];   /* BatchNorm, test mode */
        if (l->act == ACT_RELU) v = v > 0.0f ? v : 0.0f;       /* NNlib relu(x) = max(0, x) */
        else if (l->act == ACT_TANH) v = mzo_tanhf(v);
        y[o] = v;
    }
}
static void run_layers(const float *blob, const layer_t *ls, int n, const float *x, float *y) {
    float a[256], b[256];
    const float *cur = x;
    for (int i = 0; i < n; i++) {
        float *dst = (i == n - 1) ? y : ((i & 1) ? b : a);
        dense(blob, &ls[i], cur, dst);
        cur = dst;
    }
}
/* NNlib.softmax (un-vendored): exp.(x .- maximum(x)) ./ sum(...), sum in index order. */
static void softmax(const float *x, int n, float *y) {
    float m = x[0];
    for (int i = 1; i < n; i++) m = x[i] > m ? x[i] : m;
    float s = 0.0f;
    for (int i = 0; i < n; i++) { y[i] = mzo_expf(x[i] - m); s = s + y[i]; }
    for (int i = 0; i < n; i++) y[i] = y[i] / s;
}

/* ------------------------------------------------------------------------------------------
 * ResNet networks (net_type = 1): the REPAIRED spec of src/Learning.jl:148-255.
 *
 * The reference's ResNet constructors never ran: they read undefined names (`downsampling`, `size`,
 * `hyper.width_hidden`, `hyper.stacked_actions`; Learning.jl:175,177,203,237) and no ResNetHP exists.
 * Repairs (and nothing else): downsampling = hyper.downsample = false; width_hidden = the Config-level
 * width (64); stacked_actions = 1 (one action plane, as make_state_action builds, SelfPlay.jl:7-14);
 * hlayers(depth, hs, ...) = depth x Dense(hs, hs, relu) (make_dense without batch norm, :70-80).
 *   unit ConvBN(k, cin => cout, act) = Conv((k,k), cin => cout, pad = k / 2) ; BatchNorm(cout, act)
 *   block(k, n) = relu.(x + BN(Conv(relu(BN(Conv(x))))))                                   (:148-158)
 *   representation: ConvBN(K, planes => nf, relu); num_blocks x block(K, nf)               (:160-171) -> (W,H,nf)
 *   prediction: ConvBN(1, nf => nf, relu); blocks(1);                                       (:193-206)
 *       value : ConvBN(1, nf => nvf, relu); flatten; Dense(W*H*nvf => hs, relu); hlayers(depth_value); Dense(hs => 1, tanh)
 *       policy: ConvBN(1, nf => npf, relu); flatten; Dense(W*H*npf => hs) [no activation]; hlayers(depth_value) [sic, :222];
 *               Dense(hs => A); softmax                                                      (:208-226)
 *   dynamics: ConvBN(1, nf + 1 => nf, relu); blocks(1);                                     (:228-241)
 *       state : ConvBN(1, nf => nf, relu); blocks(1)                                         (:243-246)
 *       reward: like the value head                                                           (:248-254)
 * Un-vendored library semantics restated here: NNlib.conv is a TRUE convolution (kernel flipped):
 *   y[x,y,co] = b[co] + sum_{ci,b,a} w[a,b,ci,co] * in[x + pad - a, y + pad - b, ci]   (0-based, zero padding);
 * Flux.BatchNorm outside a gradient context (the reference never differentiates its forward passes, Q20) uses the
 * stored statistics:  act.(gamma .* (x .- mu) ./ sqrt.(var .+ 1f-5) .+ beta);  flatten is column-major (cell fastest).
 * Blob: units in construction order; ConvBN = W (k,k,cin,cout) column-major, b[cout], beta, gamma, mu, var [cout each];
 * Dense = W (out,in) column-major, b.  Init: glorot_uniform with nfan = (k*k*cin, k*k*cout), b = 0, beta = 0, gamma = 1,
 * mu = 0, var = 1 (Flux defaults).
 *
 * bf16 emulation (mzo_set_bf16(1)) mirrors the tensor-core kernel: every stored activation is rounded to bfloat16,
 * weights are rounded to bfloat16 AFTER the BatchNorm scale gamma/sqrt(var+eps) (and the x2 of make_state_action) is folded into them,
 * sums accumulate in Float32, the rest of BatchNorm is a Float32 per-channel shift;
 * the action plane of the dynamics input and the final value / logits / reward stay Float32.
 * ------------------------------------------------------------------------------------------ */
enum { RN_CONV = 0, RN_DENSE = 1 };
typedef struct { int kind, k, cin, cout, act; int w_off, b_off, beta_off, gamma_off, mu_off, var_off; } rn_unit_t;
typedef struct { int n; rn_unit_t u[64]; int base, n_params; } rn_net_t;
typedef struct { rn_net_t net[3]; int trunk[3], blocks; int head1_first[3], head2_first[3]; } rn_model_t;

static void rn_add_conv(rn_net_t *n, int k, int cin, int cout, int act, int *off) {
    rn_unit_t *u = &n->u[n->n++]; memset(u, 0, sizeof(*u));
    u->kind = RN_CONV; u->k = k; u->cin = cin; u->cout = cout; u->act = act;
    u->w_off = *off; *off += k * k * cin * cout; u->b_off = *off; *off += cout;
    u->beta_off = *off; *off += cout; u->gamma_off = *off; *off += cout; u->mu_off = *off; *off += cout; u->var_off = *off; *off += cout;
}
static void rn_add_dense(rn_net_t *n, int in, int out, int act, int *off) {
    rn_unit_t *u = &n->u[n->n++]; memset(u, 0, sizeof(*u));
    u->kind = RN_DENSE; u->k = 1; u->cin = in; u->cout = out; u->act = act;
    u->w_off = *off; *off += in * out; u->b_off = *off; *off += out;
}
static void rn_add_blocks(rn_net_t *n, int k, int nf, int nb, int *off) {
    for (int i = 0; i < nb; i++) { rn_add_conv(n, k, nf, nf, ACT_RELU, off); rn_add_conv(n, k, nf, nf, ACT_ID, off); }
}
static void rn_add_scalar_head(rn_net_t *n, const mzo_config *c, int f, int first_act, int out, int out_act, int *off) {
    rn_add_conv(n, 1, c->rn_num_filters, f, ACT_RELU, off);
    rn_add_dense(n, c->W * c->H * f, c->width_hidden, first_act, off);
    for (int i = 0; i < c->depth_value; i++) rn_add_dense(n, c->width_hidden, c->width_hidden, ACT_RELU, off);
    rn_add_dense(n, c->width_hidden, out, out_act, off);
}
static void rn_build(const mzo_config *c, rn_model_t *m) {
    int off = 0, nf = c->rn_num_filters, nb = c->rn_num_blocks;
    memset(m, 0, sizeof(*m)); m->blocks = nb;
    rn_net_t *r = &m->net[0]; r->base = off;
    rn_add_conv(r, c->rn_kernel, obs_planes(c), nf, ACT_RELU, &off); rn_add_blocks(r, c->rn_kernel, nf, nb, &off);
    r->n_params = off - r->base; m->trunk[0] = r->n;
    rn_net_t *p = &m->net[1]; p->base = off;
    rn_add_conv(p, 1, nf, nf, ACT_RELU, &off); rn_add_blocks(p, 1, nf, nb, &off); m->trunk[1] = p->n;
    m->head1_first[1] = p->n; rn_add_scalar_head(p, c, c->rn_first_head_filters, ACT_RELU, 1, ACT_TANH, &off);
    m->head2_first[1] = p->n; rn_add_scalar_head(p, c, c->rn_second_head_filters, ACT_ID, c->A, ACT_ID, &off);
    p->n_params = off - p->base;
    rn_net_t *d = &m->net[2]; d->base = off;
    rn_add_conv(d, 1, nf + 1, nf, ACT_RELU, &off); rn_add_blocks(d, 1, nf, nb, &off); m->trunk[2] = d->n;
    m->head1_first[2] = d->n; rn_add_conv(d, 1, nf, nf, ACT_RELU, &off); rn_add_blocks(d, 1, nf, nb, &off);
    m->head2_first[2] = d->n; rn_add_scalar_head(d, c, c->rn_first_head_filters, ACT_RELU, 1, ACT_TANH, &off);
    d->n_params = off - d->base;
}
static int rn_num_params(const mzo_config *c, int net) {
    rn_model_t m; rn_build(c, &m);
    return net < 3 ? m.net[net].n_params : m.net[0].n_params + m.net[1].n_params + m.net[2].n_params;
}
static void rn_init_weights(const mzo_config *c, uint64_t seed, float *blob) {
    rn_model_t m; rn_build(c, &m);
    for (int n = 0; n < 3; n++) for (int ui = 0; ui < m.net[n].n; ui++) {
        const rn_unit_t *u = &m.net[n].u[ui];
        int nw = u->k * u->k * u->cin * u->cout;
        float scale = sqrtf(24.0f / (float)(u->k * u->k * (u->cin + u->cout)));
        for (int i = 0; i < nw; i += 4) {
            uint32_t r[4]; mzo_philox(seed, STREAM_INIT, (uint32_t)n, (uint32_t)ui, (uint32_t)(i / 4), 0, r);
            for (int j = 0; j < 4 && i + j < nw; j++) blob[u->w_off + i + j] = (u32_to_unit(r[j]) - 0.5f) * scale;
        }
        for (int o = 0; o < u->cout; o++) {
            blob[u->b_off + o] = 0.0f;
            if (u->kind == RN_CONV) { blob[u->beta_off + o] = 0.0f; blob[u->gamma_off + o] = 1.0f; blob[u->mu_off + o] = 0.0f; blob[u->var_off + o] = 1.0f; }
        }
    }
}
static inline float rn_store(float v) { return g_bf16 ? bf16_round(v) : v; }   /* a stored activation */
static inline float rn_act(float v, int act) { return act == ACT_RELU ? (v > 0.0f ? v : 0.0f) : act == ACT_TANH ? mzo_tanhf(v) : v; }
/* ConvBN on a (W,H,cin) array (cell fastest).  skip (may be NULL) is added after the BatchNorm, before `act`
 * (SkipConnection(layers, +) then relu, :157-158).  extra_plane >= 0: the LAST input channel is a constant plane whose
 * value is rowval (the dynamics action plane) and `in` holds only cin-1 channels, scaled by in_mul (state * 2). */
static void rn_conv(const mzo_config *c, const float *blob, const rn_unit_t *u, const float *in, float in_mul, int has_plane, float plane,
                    const float *skip, int act, float *out) {
    int W = c->W, H = c->H, cells = W * H, k = u->k, pad = k / 2, cin = u->cin, cout = u->cout;
    const float *w = blob + u->w_off;
    for (int co = 0; co < cout; co++) {
        float den = sqrtf(blob[u->var_off + co] + 1e-5f), gamma = blob[u->gamma_off + co], beta = blob[u->beta_off + co], mu = blob[u->mu_off + co], b = blob[u->b_off + co];
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            float acc = 0.0f, v;
            int cmain = has_plane ? cin - 1 : cin;
            for (int ci = 0; ci < cmain; ci++) for (int kb = 0; kb < k; kb++) for (int ka = 0; ka < k; ka++) {
                int sx = x + pad - ka, sy = y + pad - kb;
                if (sx < 0 || sx >= W || sy < 0 || sy >= H) continue;
                float wv = w[ka + k * (kb + k * (ci + cin * co))], xv = in[sx + W * sy + cells * ci];
                if (g_bf16) acc = fmaf(bf16_round(wv * (in_mul * (gamma / den))), xv, acc);   /* BatchNorm scale (and the x2 of the dynamics input) folded into the bf16 weight */
                else acc = fmaf(wv, xv * in_mul, acc);
            }
            if (g_bf16) {   /* the kernel's Float32 epilogue: acc + plane * (w_plane * s) + ((b - mu) * s + beta), s = gamma / sqrt(var + eps) */
                float s = gamma / den, T = fmaf(b - mu, s, beta), E = has_plane ? w[0 + k * (0 + k * ((cin - 1) + cin * co))] * s : 0.0f;
                v = acc + fmaf(plane, E, T);
            } else {
                if (has_plane) acc = fmaf(w[0 + k * (0 + k * ((cin - 1) + cin * co))], plane, acc);
                v = acc + b;
                v = ((gamma * (v - mu)) / den) + beta;                       /* Flux BatchNorm, test mode */
            }
            if (skip) v = v + skip[x + W * y + cells * co];
            out[x + W * y + cells * co] = rn_store(rn_act(v, act));
        }
    }
}
static void rn_dense(const float *blob, const rn_unit_t *u, const float *x, int final, float *y) {
    const float *W = blob + u->w_off, *b = blob + u->b_off;
    for (int o = 0; o < u->cout; o++) {
        float acc = 0.0f;
        for (int k = 0; k < u->cin; k++) acc = fmaf(g_bf16 ? bf16_round(W[o + (size_t)u->cout * k]) : W[o + (size_t)u->cout * k], x[k], acc);
        float v = rn_act(acc + b[o], u->act);
        y[o] = final ? v : rn_store(v);
    }
}
/* ConvBN(relu) + num_blocks blocks starting at unit `first`; returns the index after the tower; result in out (may alias in) */
static int rn_tower(const mzo_config *c, const float *blob, const rn_net_t *n, int first, int nb, const float *in, float in_mul, int has_plane, float plane, float *out) {
    static __thread float t1[MZO_MAX_HIDDEN + 64], t2[MZO_MAX_HIDDEN + 64];
    rn_conv(c, blob, &n->u[first], in, in_mul, has_plane, plane, NULL, ACT_RELU, t1);
    int ui = first + 1;
    for (int b = 0; b < nb; b++, ui += 2) {
        rn_conv(c, blob, &n->u[ui], t1, 1.0f, 0, 0.0f, NULL, ACT_RELU, t2);
        rn_conv(c, blob, &n->u[ui + 1], t2, 1.0f, 0, 0.0f, t1, ACT_RELU, out);       /* relu.(x + layers(x)) */
        memcpy(t1, out, sizeof(float) * (size_t)c->W * c->H * n->u[ui + 1].cout);
    }
    if (nb == 0) memcpy(out, t1, sizeof(float) * (size_t)c->W * c->H * n->u[first].cout);
    return ui;
}
static void rn_scalar_head(const mzo_config *c, const float *blob, const rn_net_t *n, int first, const float *trunk, float *out) {
    float f[MZO_MAX_OBS * 2], a[256], b[256];
    rn_conv(c, blob, &n->u[first], trunk, 1.0f, 0, 0.0f, NULL, ACT_RELU, f);   /* flatten: cell fastest, then filter */
    const float *cur = f; int ui = first + 1, nd = c->depth_value + 2;
    for (int i = 0; i < nd; i++, ui++) {
        float *dst = (i == nd - 1) ? out : ((i & 1) ? b : a);
        rn_dense(blob, &n->u[ui], cur, i == nd - 1, dst);
        cur = dst;
    }
}
static void rn_representation(const mzo_config *c, const float *blob, const float *stacked, float *hidden) {
    rn_model_t m; rn_build(c, &m);
    static __thread float in[MZO_MAX_OBS * 4];
    for (int i = 0; i < stack_size(c); i++) in[i] = rn_store(stacked[i]);
    rn_tower(c, blob, &m.net[0], 0, m.blocks, in, 1.0f, 0, 0.0f, hidden);
}
static void softmax(const float *x, int n, float *y);
static void rn_prediction(const mzo_config *c, const float *blob, const float *hidden, float *value, float *policy) {
    rn_model_t m; rn_build(c, &m);
    static __thread float in[MZO_MAX_HIDDEN], t[MZO_MAX_HIDDEN];
    float logits[MZO_MAX_A];
    for (int i = 0; i < c->hidden_state_size; i++) in[i] = rn_store(hidden[i]);
    rn_tower(c, blob, &m.net[1], 0, m.blocks, in, 1.0f, 0, 0.0f, t);
    rn_scalar_head(c, blob, &m.net[1], m.head1_first[1], t, value);
    rn_scalar_head(c, blob, &m.net[1], m.head2_first[1], t, logits);
    softmax(logits, c->A, policy);
}
/* sa = (W,H,nf+1): nf state channels (already doubled by make_state_action) and the constant action plane */
static void rn_dynamics(const mzo_config *c, const float *blob, const float *sa, float *next_hidden, float *reward) {
    rn_model_t m; rn_build(c, &m);
    static __thread float in[MZO_MAX_HIDDEN], t[MZO_MAX_HIDDEN];
    int hs = c->hidden_state_size;
    float plane = sa[hs];
    if (g_bf16) { for (int i = 0; i < hs; i++) in[i] = bf16_round(sa[i] * 0.5f); }   /* the kernel stages h and multiplies the accumulator by 2 */
    else for (int i = 0; i < hs; i++) in[i] = sa[i];
    rn_tower(c, blob, &m.net[2], 0, m.blocks, in, g_bf16 ? 2.0f : 1.0f, 1, plane, t);
    rn_tower(c, blob, &m.net[2], m.head1_first[2], m.blocks, t, 1.0f, 0, 0.0f, next_hidden);
    rn_scalar_head(c, blob, &m.net[2], m.head2_first[2], t, reward);
}

typedef struct { mzo_config cfg; net_t nets[3]; const float *blob; } model_t;
static void model_init(model_t *m, const mzo_config *c, const float *blob) { m->cfg = *c; build_nets(c, m->nets); m->blob = blob; }

static void representation(const model_t *m, const float *stacked, float *hidden) {
    if (m->cfg.net_type == 1) { rn_representation(&m->cfg, m->blob, stacked, hidden); return; }
    run_layers(m->blob, m->nets[0].trunk, m->nets[0].n_trunk, stacked, hidden);
}
static void prediction(const model_t *m, const float *hidden, float *value, float *policy) {
    if (m->cfg.net_type == 1) { rn_prediction(&m->cfg, m->blob, hidden, value, policy); return; }
    float t[256], logits[MZO_MAX_A];
    const net_t *n = &m->nets[1];
    run_layers(m->blob, n->trunk, n->n_trunk, hidden, t);
    run_layers(m->blob, n->h1, n->n_h1, t, value);
    run_layers(m->blob, n->h2, n->n_h2, t, logits);
    softmax(logits, m->cfg.A, policy); /* Learning.jl:114: the policy head ENDS in softmax (Q1) */
}
static void dynamics(const model_t *m, const float *sa, float *next_hidden, float *reward) {
    if (m->cfg.net_type == 1) { rn_dynamics(&m->cfg, m->blob, sa, next_hidden, reward); return; }
    float t[256];
    const net_t *n = &m->nets[2];
    run_layers(m->blob, n->trunk, n->n_trunk, sa, t);
    run_layers(m->blob, n->h1, n->n_h1, t, next_hidden);
    run_layers(m->blob, n->h2, n->n_h2, t, reward);
}
void mzo_representation(const mzo_config *c, const float *blob, const float *s, float *h) { model_t m; model_init(&m, c, blob); representation(&m, s, h); }
void mzo_prediction(const mzo_config *c, const float *blob, const float *h, float *v, float *p) { model_t m; model_init(&m, c, blob); prediction(&m, h, v, p); }
void mzo_dynamics(const mzo_config *c, const float *blob, const float *sa, float *nh, float *r) { model_t m; model_init(&m, c, blob); dynamics(&m, sa, nh, r); }

/* ------------------------------------------------------------------------------------------
 * Environment: games/tictactoe/game.jl.  Board = BitArray (3,3,3): plane 1 = player-1 marks,
 * plane 2 = player-2 marks, plane 3 = empty (game.jl:8-13).  Here p1/p2 are bit masks with
 * bit (a-1) for action a; CartesianIndices((3,3))[a] is column-major so cell index = a-1.
 * ------------------------------------------------------------------------------------------ */
static const uint32_t TTT_LINES[8] = { /* game.jl:106-113, bits = row-1 + 3*(col-1) */
    (1u << 0) | (1u << 3) | (1u << 6), (1u << 1) | (1u << 4) | (1u << 7), (1u << 2) | (1u << 5) | (1u << 8),
    (1u << 0) | (1u << 1) | (1u << 2), (1u << 3) | (1u << 4) | (1u << 5), (1u << 6) | (1u << 7) | (1u << 8),
    (1u << 0) | (1u << 4) | (1u << 8), (1u << 6) | (1u << 4) | (1u << 2)};

void mzo_env_reset(const mzo_config *c, mzo_env *e) { (void)c; e->p1 = 0; e->p2 = 0; e->player = 1; e->moves = 0; } /* game.jl:15-20 */

/* is_win(env, player) IGNORES its argument and tests env.player, the side to move (game.jl:102-104, Q14). */
static int ttt_is_win_side_to_move(const mzo_env *e) {
    uint32_t b = (uint32_t)(e->player == 1 ? e->p1 : e->p2);
    for (int i = 0; i < 8; i++) if ((b & TTT_LINES[i]) == TTT_LINES[i]) return 1;
    return 0;
}
static uint64_t cells_mask(const mzo_config *c) { return (c->W * c->H >= 64) ? ~0ull : ((1ull << (c->W * c->H)) - 1ull); }

/* ---- MZO_GAME_CONNECT: the synthetic larger-board game of BASELINE.json configs[3] (no reference code; SURVEY 8d:
 * "column drop, 4-in-row, clean termination, reward to last mover").  The observation is (W,H,3) with W rows and H columns,
 * cell = row + W*column, row 0 is the bottom; action a in 1..H drops a mark into column a.  The game ends when the LAST MOVER
 * has four in a row (horizontal, vertical or diagonal) or the board is full; reward(env, p) = +1 / -1 for the last mover /
 * the other player after a win, 0 otherwise.  None of TicTacToe's quirks (Q14-Q16) apply. ---- */
static int cn_at(const mzo_config *c, uint64_t b, int r, int col) { return (r >= 0 && r < c->W && col >= 0 && col < c->H) ? (int)((b >> (r + c->W * col)) & 1ull) : 0; }
static int cn_height(const mzo_config *c, const mzo_env *e, int col) { int h = 0; while (h < c->W && cn_at(c, e->p1 | e->p2, h, col)) h++; return h; }
static int cn_has4(const mzo_config *c, uint64_t b) {
    static const int dr[4] = {1, 0, 1, 1}, dc[4] = {0, 1, 1, -1};
    for (int col = 0; col < c->H; col++) for (int r = 0; r < c->W; r++) for (int d = 0; d < 4; d++) {
        int n = 0;
        while (n < 4 && cn_at(c, b, r + n * dr[d], col + n * dc[d])) n++;
        if (n == 4) return 1;
    }
    return 0;
}
static int cn_last_mover_won(const mzo_config *c, const mzo_env *e) { return cn_has4(c, e->player == 1 ? e->p2 : e->p1); }
static int cn_full(const mzo_config *c, const mzo_env *e) { for (int col = 0; col < c->H; col++) if (cn_height(c, e, col) < c->W) return 0; return 1; }

uint32_t mzo_env_legal_mask(const mzo_config *c, const mzo_env *e) { /* game.jl:35-43 */
    if (c->game == MZO_GAME_CONNECT) {
        uint32_t m = 0;
        if (cn_last_mover_won(c, e)) return 0;
        for (int col = 0; col < c->H; col++) if (cn_height(c, e, col) < c->W) m |= 1u << col;
        return m;
    }
    if (ttt_is_win_side_to_move(e)) return 0; /* is_win(env,1) || is_win(env,2): both test env.player */
    return (uint32_t)(~(e->p1 | e->p2) & cells_mask(c));
}
void mzo_env_step(const mzo_config *c, mzo_env *e, int action) { /* game.jl:45-52 */
    uint64_t bit = 1ull << (action - 1);
    if (c->game == MZO_GAME_CONNECT) bit = 1ull << (cn_height(c, e, action - 1) + c->W * (action - 1));
    if (e->player == 1) e->p1 |= bit; else e->p2 |= bit; /* board[a,3]=false; board[a,player]=true */
    e->moves += 1;
    e->player = e->player == 1 ? 2 : 1;                  /* mod1(player+1, 2) */
}
/* State table (game.jl:117-147): is_terminated = !(has_empty_pos && isnothing(w)), w = 1 when
 * is_win (of the side to move, Q14-Q15) else nothing; winner is never 2. */
int mzo_env_is_terminated(const mzo_config *c, const mzo_env *e) { /* game.jl:85 */
    if (c->game == MZO_GAME_CONNECT) return cn_full(c, e) || cn_last_mover_won(c, e);
    int has_empty = (~(e->p1 | e->p2) & cells_mask(c)) != 0;
    return !(has_empty && !ttt_is_win_side_to_move(e));
}
int mzo_env_reward(const mzo_config *c, const mzo_env *e, int player) { /* game.jl:87-100 */
    if (c->game == MZO_GAME_CONNECT) return !cn_last_mover_won(c, e) ? 0 : (player == (e->player == 1 ? 2 : 1) ? 1 : -1);
    if (!mzo_env_is_terminated(c, e)) return 0;
    if (!ttt_is_win_side_to_move(e)) return 0; /* winner === nothing */
    return player == 1 ? 1 : -1;               /* winner is always 1 (Q15) */
}
void mzo_env_observation(const mzo_config *c, const mzo_env *e, float *obs) { /* Float32 copy of board, SelfPlay.jl:352 */
    int n = c->W * c->H;
    for (int i = 0; i < n; i++) {
        int a = (int)((e->p1 >> i) & 1), b = (int)((e->p2 >> i) & 1);
        obs[i] = (float)a; obs[n + i] = (float)b; obs[2 * n + i] = (float)(!(a | b));
    }
}

/* RLBase.walk-style census of the reachable graph (KAT-env-3). */
typedef struct { const mzo_config *c; uint8_t *seen; int64_t *out; } census_t;
static uint32_t board_key(const mzo_env *e) { return (uint32_t)(e->p1 | (e->p2 << 9)); }
static void census_walk(census_t *z, mzo_env e, int len, int last_mover) {
    uint32_t key = board_key(&e);
    int term = mzo_env_is_terminated(z->c, &e);
    if (!z->seen[key]) { z->seen[key] = 1; z->out[0]++; if (term) z->out[1]++; }
    if (term) {
        int r = mzo_env_reward(z->c, &e, last_mover);
        z->out[2]++; z->out[3 + len]++;
        z->out[32 + len * 6 + (last_mover - 1) * 3 + (r + 1)]++;
        return;
    }
    uint32_t legal = mzo_env_legal_mask(z->c, &e);
    for (int a = 1; a <= z->c->A; a++) if (legal & (1u << (a - 1))) {
        mzo_env n = e; int mover = e.player;
        mzo_env_step(z->c, &n, a);
        census_walk(z, n, len + 1, mover);
    }
}
void mzo_env_census(const mzo_config *c, int64_t *out) {
    census_t z; z.c = c; z.out = out;
    memset(out, 0, sizeof(int64_t) * (32 + 64 * 6));
    z.seen = (uint8_t *)calloc(1u << 18, 1);
    mzo_env e; mzo_env_reset(c, &e);
    census_walk(&z, e, 0, 0 + 1);
    free(z.seen);
}

/* ------------------------------------------------------------------------------------------
 * get_stacked_observations (src/SelfPlay.jl:128-149).  obs_hist is [T][C][H][W] (Julia (W,H,C,T)).
 * Output planes: current obs (C), then for each past index: action plane (RAW index, Q13) + obs (C).
 * ------------------------------------------------------------------------------------------ */
void mzo_stack_observations(const mzo_config *c, const float *obs_hist, const int32_t *action_hist, int index, float *stacked) {
    int plane = c->W * c->H, on = obs_size(c);
    memcpy(stacked, obs_hist + (size_t)(index - 1) * on, sizeof(float) * on);
    float *dst = stacked + on;
    for (int past = index - 1; past >= index - c->stacked_observations; past--) {
        if (past >= 1) {
            float av = (float)action_hist[past - 1]; /* ones .* action_history[past] */
            for (int i = 0; i < plane; i++) dst[i] = av;
            memcpy(dst + plane, obs_hist + (size_t)(past - 1) * on, sizeof(float) * on);
        } else {
            for (int i = 0; i < plane + on; i++) dst[i] = 0.0f;
        }
        dst += plane + on;
    }
}

/* ------------------------------------------------------------------------------------------
 * MCTS: src/SelfPlay.jl:16-39 (MinMaxStats), 62-96 (Node, expand_node!), 102-109 (noise),
 * 157-184 (select_child / ucb_score), 190-217 (backpropagate!), 230-285 (run_mcts).
 * Pointer-based like the reference; children is the Dict{Int,Node}, iterated in cfg.child_order.
 * ------------------------------------------------------------------------------------------ */
typedef struct node {
    int visit_count, to_play, expanded;
    float prior, value_sum, reward;
    struct node *children[MZO_MAX_A]; /* key a -> children[a-1]; NULL = no key */
    float *hidden_state;              /* the Array{Float32,3} the node references (aliased, Q6) */
} node_t;
typedef struct { float min, max; } minmax_t;

typedef struct {
    const model_t *m;
    node_t *nodes; int n_nodes;
    float *hiddens; int n_hiddens;
    uint64_t game_id; int move_idx;
} tree_t;

static node_t *new_node(tree_t *t, float prior) { /* Node(prior=...) defaults, SelfPlay.jl:62-70 */
    node_t *n = &t->nodes[t->n_nodes++];
    memset(n, 0, sizeof(*n));
    n->to_play = 1; n->prior = prior;
    return n;
}
static float node_value(const node_t *n) { /* SelfPlay.jl:76-82 */
    return n->visit_count == 0 ? 0.0f : n->value_sum / (float)n->visit_count;
}
/* expand_node! (SelfPlay.jl:88-96): softmax AGAIN over the legal subset of the already-softmaxed
 * policy (Q1), in ascending action order; children for the given (root's, Q7) legal set. */
static void expand_node(tree_t *t, node_t *node, uint32_t legal, int to_play, float reward, const float *policy, float *hidden) {
    const mzo_config *c = &t->m->cfg;
    float sub[MZO_MAX_A], pv[MZO_MAX_A]; int acts[MZO_MAX_A], n = 0;
    for (int a = 1; a <= c->A; a++) if (legal & (1u << (a - 1))) { acts[n] = a; sub[n] = policy[a - 1]; n++; }
    softmax(sub, n, pv);
    for (int i = 0; i < n; i++) node->children[acts[i] - 1] = new_node(t, pv[i]);
    node->expanded = 1; node->to_play = to_play; node->reward = reward; node->hidden_state = hidden;
}

/* Gamma(alpha<1) sampler for the Dirichlet noise (Distributions.jl un-vendored; contract):
 * Marsaglia-Tsang on alpha+1 with polar-method normals, then * u^(1/alpha), all Float32, draws
 * from Philox(seed, DIRICHLET, game, move, child j, counter). */
typedef struct { uint64_t seed; uint32_t c0, c1, c2, ctr; uint32_t buf[4]; int have; } rstream_t;
static uint32_t rs_next(rstream_t *s) {
    if (s->have == 0) { mzo_philox(s->seed, STREAM_DIRICHLET, s->c0, s->c1, s->c2, s->ctr++, s->buf); s->have = 4; }
    return s->buf[4 - (s->have--)];
}
static float rs_unit_open(rstream_t *s) { return ((float)(rs_next(s) >> 8) + 0.5f) * 5.9604644775390625e-8f; } /* (0,1) */
static float rs_normal(rstream_t *s) {
    for (;;) {
        float a = 2.0f * rs_unit_open(s) - 1.0f, b = 2.0f * rs_unit_open(s) - 1.0f;
        float q = a * a + b * b;
        if (q >= 1.0f || q == 0.0f) continue;
        return a * sqrtf(-2.0f * mzo_logf(q) / q);
    }
}
static float rs_gamma(rstream_t *s, float alpha) {
    float a1 = alpha < 1.0f ? alpha + 1.0f : alpha;
    float d = a1 - 0.333333343f, cc = 1.0f / sqrtf(9.0f * d), g;
    for (;;) {
        float x = rs_normal(s), v = 1.0f + cc * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        float u = rs_unit_open(s);
        if (mzo_logf(u) < 0.5f * x * x + d - d * v + d * mzo_logf(v)) { g = d * v; break; }
    }
    if (alpha < 1.0f) { float u = rs_unit_open(s); g = g * mzo_expf(mzo_logf(u) / alpha); }
    return g;
}
/* add_exploration_noise! (SelfPlay.jl:102-109): noise[i] pairs with the i-th key in Dict order. */
static void add_exploration_noise(tree_t *t, node_t *root, float alpha, float eps) {
    const mzo_config *c = &t->m->cfg;
    float noise[MZO_MAX_A], sum = 0.0f; int n = 0;
    for (int j = 0; j < c->A; j++) {
        node_t *ch = root->children[c->child_order[j] - 1];
        if (!ch) continue;
        rstream_t s = {c->seed, (uint32_t)t->game_id, (uint32_t)t->move_idx, (uint32_t)n, 0, {0, 0, 0, 0}, 0};
        noise[n] = rs_gamma(&s, alpha); sum = sum + noise[n]; n++;
    }
    n = 0;
    for (int j = 0; j < c->A; j++) {
        node_t *ch = root->children[c->child_order[j] - 1];
        if (!ch) continue;
        float nz = noise[n++] / sum;
        ch->prior = ch->prior * (1.0f - eps) + nz * eps;
    }
}

/* ucb_score (SelfPlay.jl:171-184): Float64 exploration term, Float32 value term, Float32 result (Q3-Q4). */
static float ucb_score(const mzo_config *c, const node_t *parent, const node_t *child, const minmax_t *mm) {
    double pb_c = log2((double)(parent->visit_count + c->pb_c_base + 1) / (double)c->pb_c_base) + (double)c->pb_c_init;
    pb_c *= sqrt((double)parent->visit_count) / (double)(child->visit_count + 1);
    double prior_score = pb_c * (double)child->prior;
    if (child->visit_count > 0) {
        float nv = node_value(child);
        float q = child->reward + c->discount * (c->num_players == 1 ? nv : -nv);
        float vs = (mm->max > mm->min) ? (q - mm->min) / (mm->max - mm->min) : q; /* normalize_tree_value :33-39 */
        return (float)(prior_score + (double)vs);
    }
    return (float)(prior_score + 0.0);
}
/* select_child (SelfPlay.jl:157-166).  Ties: the reference draws rand(max_ucbs) from the unseeded
 * global RNG (Q2); contract: Philox(seed, TIE, game, move, sim, depth) over the tied set in Dict order. */
static node_t *select_child(tree_t *t, node_t *node, const minmax_t *mm, int sim, int depth, int *action) {
    const mzo_config *c = &t->m->cfg;
    float scores[MZO_MAX_A]; int acts[MZO_MAX_A], n = 0;
    for (int j = 0; j < c->A; j++) {
        int a = c->child_order[j]; node_t *ch = node->children[a - 1];
        if (!ch) continue;
        scores[n] = ucb_score(c, node, ch, mm); acts[n] = a; n++;
    }
    float mx = scores[0];
    for (int i = 1; i < n; i++) mx = scores[i] > mx ? scores[i] : mx;
    int tied[MZO_MAX_A], nt = 0;
    for (int i = 0; i < n; i++) if (scores[i] == mx) tied[nt++] = i;
    int pick = tied[0];
    if (nt > 1 && c->tie_mode == MZO_TIE_PHILOX) {
        uint32_t r[4]; mzo_philox(c->seed, STREAM_TIE, (uint32_t)t->game_id, (uint32_t)t->move_idx, (uint32_t)sim, (uint32_t)depth, r);
        pick = tied[u32_below(r[0], (uint32_t)nt)];
    }
    *action = acts[pick];
    return node->children[acts[pick] - 1];
}
static void update_tree(minmax_t *mm, float v) { /* SelfPlay.jl:27-31 */
    mm->min = mm->min < v ? mm->min : v;
    mm->max = mm->max > v ? mm->max : v;
}
/* backpropagate! (SelfPlay.jl:190-217), including the two-player precedence bug (Q8). */
static void backpropagate(const mzo_config *c, node_t **path, int n, float value, int to_play, minmax_t *mm) {
    for (int i = n - 1; i >= 0; i--) {
        node_t *nd = path[i];
        if (c->num_players == 1) {
            nd->value_sum = nd->value_sum + value; nd->visit_count += 1;
            update_tree(mm, nd->reward + c->discount * node_value(nd));
            value = nd->reward + c->discount * value;
        } else {
            if (nd->to_play == to_play) nd->value_sum = nd->value_sum + value; else nd->value_sum = nd->value_sum - value;
            nd->visit_count += 1;
            update_tree(mm, nd->reward + c->discount * node_value(nd));
            if (nd->to_play == to_play) value = -nd->reward; else value = nd->reward + c->discount * value;
        }
    }
}
/* make_state_action (SelfPlay.jl:7-14): doubles the CALLER's state in place (Q6); action plane =
 * Float32(Float64(a) / length(action_space)). */
static void make_state_action(const mzo_config *c, float *state, int action, float *sa) {
    int on = c->hidden_state_size, plane = c->W * c->H;   /* the state is the (W,H,channels) hidden state */
    float av = (float)((double)action / (double)c->A);
    for (int i = 0; i < on; i++) { state[i] = state[i] * 2.0f; sa[i] = state[i]; }
    for (int i = 0; i < plane; i++) sa[on + i] = av;
}

static void run_mcts(tree_t *t, const float *stacked, uint32_t legal, int to_play, int exploration, node_t **root_out, float *trace) {
    const model_t *m = t->m; const mzo_config *c = &m->cfg;
    int hs = c->hidden_state_size;
    t->n_nodes = 0; t->n_hiddens = 0;
    node_t *root = new_node(t, 0.0f);                                   /* :232 */
    float *h0 = t->hiddens + (size_t)(t->n_hiddens++) * hs;
    representation(m, stacked, h0);                                      /* :234 */
    float v0, p0[MZO_MAX_A];
    prediction(m, h0, &v0, p0);                                          /* :239 */
    expand_node(t, root, legal, to_play, 0.0f, p0, h0);                  /* :245 */
    if (exploration) add_exploration_noise(t, root, c->dirichlet_alpha, c->exploration_eps); /* :247-249 */
    minmax_t mm = {INFINITY, -INFINITY};                                 /* :251 */
    node_t *path[MZO_MAX_T * 8];
    for (int iter = 1; iter <= c->num_iters; iter++) {                   /* :254 */
        node_t *node = root; int vtp = to_play, np = 0, action = 0, depth = 0;
        path[np++] = node;
        while (node->expanded) {                                         /* :261-268 */
            depth++;
            node = select_child(t, node, &mm, iter, depth, &action);
            path[np++] = node;
            vtp = vtp % c->num_players + 1;                              /* mod1(vtp+1, nplayers) */
        }
        node_t *parent = path[np - 2];                                   /* :270 */
        float value, policy[MZO_MAX_A], sa[MZO_MAX_HIDDEN + 64], reward;
        prediction(m, parent->hidden_state, &value, policy);             /* :271 -- PARENT state (Q5) */
        make_state_action(c, parent->hidden_state, action, sa);          /* :273 */
        float *nh = t->hiddens + (size_t)(t->n_hiddens++) * hs;
        dynamics(m, sa, nh, &reward);                                    /* :275 */
        expand_node(t, node, legal, vtp, reward, policy, nh);            /* :280 -- root's legal set (Q7) */
        backpropagate(c, path, np, value, vtp, &mm);                     /* :281 */
        if (trace) { float *tr = trace + 5 * (iter - 1); tr[0] = (float)depth; tr[1] = (float)action; tr[2] = (float)((parent->hidden_state - t->hiddens) / hs); tr[3] = value; tr[4] = reward; }
    }
    *root_out = root;
}

static tree_t *tree_alloc(const model_t *m) {
    const mzo_config *c = &m->cfg;
    tree_t *t = (tree_t *)calloc(1, sizeof(tree_t));
    t->m = m;
    t->nodes = (node_t *)malloc(sizeof(node_t) * (size_t)(1 + (c->num_iters + 1) * c->A));
    t->hiddens = (float *)malloc(sizeof(float) * (size_t)(c->num_iters + 1) * c->hidden_state_size);
    return t;
}
static void tree_free(tree_t *t) { free(t->nodes); free(t->hiddens); free(t); }

/* select_action (SelfPlay.jl:293-306): counts and actions in Dict order (Q9); T==0 argmax (first
 * max), T==Inf uniform, else Categorical(counts.^(1/T) normalised) (Q10).  Draw contract:
 * Philox(seed, ACTION, game, move)[0] -> Float32 in [0,1), inverse-CDF like Distributions.jl. */
static int select_action_counts(const mzo_config *c, const int *counts, const int *acts, int n, float temperature, uint64_t game_id, int move_idx) {
    if (temperature == 0.0f) {
        int best = 0;
        for (int i = 1; i < n; i++) if (counts[i] > counts[best]) best = i;
        return acts[best];
    }
    uint32_t r[4]; mzo_philox(c->seed, STREAM_ACTION, (uint32_t)game_id, (uint32_t)move_idx, 0, 0, r);
    if (isinf(temperature)) return acts[u32_below(r[0], (uint32_t)n)];
    float d[MZO_MAX_A], s = 0.0f, e = 1.0f / temperature;
    for (int i = 0; i < n; i++) { d[i] = pow_contract((float)counts[i], e); s = s + d[i]; }
    for (int i = 0; i < n; i++) d[i] = d[i] / s;
    float draw = u32_to_unit(r[0]), cp = d[0]; int i = 0;
    while (cp <= draw && i < n - 1) { i++; cp = cp + d[i]; }
    return acts[i];
}
int mzo_select_action(const mzo_config *c, const int32_t *visit_counts, uint32_t legal, float temperature, uint64_t game_id, int move_idx) {
    int counts[MZO_MAX_A], acts[MZO_MAX_A], n = 0;
    for (int j = 0; j < c->A; j++) { int a = c->child_order[j]; if (legal & (1u << (a - 1))) { counts[n] = visit_counts[a - 1]; acts[n] = a; n++; } }
    return select_action_counts(c, counts, acts, n, temperature, game_id, move_idx);
}

void mzo_run_mcts(const mzo_config *cfg, const float *blob, const float *stacked, uint32_t legal, int to_play, int exploration,
                  uint64_t game_id, int move_idx, int32_t *visit_counts, float *root_value, float *root_priors, float *trace) {
    model_t m; model_init(&m, cfg, blob);
    tree_t *t = tree_alloc(&m); t->game_id = game_id; t->move_idx = move_idx;
    node_t *root;
    run_mcts(t, stacked, legal, to_play, exploration, &root, trace);
    for (int a = 0; a < cfg->A; a++) {
        visit_counts[a] = root->children[a] ? root->children[a]->visit_count : 0;
        if (root_priors) root_priors[a] = root->children[a] ? root->children[a]->prior : 0.0f;
    }
    *root_value = node_value(root);
    tree_free(t);
}

/* ------------------------------------------------------------------------------------------
 * play_game (src/SelfPlay.jl:330-382) and the self_play! actor loop (:384-419) over many games.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const model_t *m; uint64_t first_game; int lo, hi; float temperature;
    int32_t *T; float *obs; int32_t *actions; float *rewards; int32_t *to_play; float *child_visits; float *root_values;
    int64_t sims;
    int opponent, muzero_player; int32_t *outcome;     /* competitive play (SelfPlay.jl:421-435); opponent 0 = "self" */
} sp_job_t;

/* a completed line of one side's marks (the real rule, not is_win's side-to-move test) */
static int side_has_line(const mzo_config *c, uint64_t b) {
    if (c->game == MZO_GAME_CONNECT) return cn_has4(c, b);
    for (int i = 0; i < 8; i++) if (((uint32_t)b & TTT_LINES[i]) == TTT_LINES[i]) return 1;
    return 0;
}
/* select_opponent_action (SelfPlay.jl:311-325).  "random": rand(rng, las) -- uniform over the ascending legal actions, here from
 * the Philox stream (seed, STREAM_OPPONENT, game, move); as written the branch reads `las` before defining it and never ran.
 * "expert": expert_agent() is not defined anywhere in the reference; repaired as one-ply lookahead: the first (ascending) legal
 * action that completes a line for the mover, else the first that would complete one for the other side, else the random action. */
int mzo_opponent_action(const mzo_config *c, const mzo_env *e, int opponent, uint64_t game_id, int move_idx) {
    uint32_t legal = mzo_env_legal_mask(c, e);
    int las[MZO_MAX_A], n = 0;
    for (int a = 1; a <= c->A; a++) if ((legal >> (a - 1)) & 1u) las[n++] = a;
    if (n == 0) return 1;
    if (opponent == MZO_OPP_EXPERT) {
        for (int pass = 0; pass < 2; pass++) for (int i = 0; i < n; i++) {
            mzo_env t = *e;
            if (pass == 1) t.player = t.player == 1 ? 2 : 1;
            int who = t.player;
            mzo_env_step(c, &t, las[i]);
            if (side_has_line(c, who == 1 ? t.p1 : t.p2)) return las[i];
        }
    }
    uint32_t r[4]; mzo_philox(c->seed, STREAM_OPPONENT, (uint32_t)game_id, (uint32_t)move_idx, 0, 0, r);
    return las[u32_below(r[0], (uint32_t)n)];
}
/* +1 / 0 / -1: the side that completed a line first (TicTacToe lets the game run one ply past a win, Q14) is the winner */
int mzo_arena_outcome(const mzo_config *c, int T, const int32_t *actions, int muzero_player) {
    mzo_env e; mzo_env_reset(c, &e);
    for (int i = 0; i < T; i++) {
        int who = e.player;
        mzo_env_step(c, &e, actions[i]);
        if (side_has_line(c, who == 1 ? e.p1 : e.p2)) return who == muzero_player ? 1 : -1;
    }
    return 0;
}

static void play_game(const model_t *m, tree_t *t, uint64_t game_id, float temperature, int32_t *T_out, float *obs_h,
                      int32_t *act_h, float *rew_h, int32_t *tp_h, float *cv_h, float *rv_h, int64_t *sims, int opponent, int muzero_player) {
    const mzo_config *c = &m->cfg;
    int on = obs_size(c), T = 0, done = 0;
    mzo_env env; float stacked[MZO_MAX_OBS * 3];
    mzo_env_reset(c, &env);
    while (!done && T <= c->max_moves) {                                  /* :343 */
        if (c->temperature_threshold >= 0 && T >= c->temperature_threshold) temperature = 0.0f;   /* :344-346 */
        int p = env.player;                                               /* :351 */
        mzo_env_observation(c, &env, obs_h + (size_t)T * on);             /* :352 (obs of the board before the move) */
        mzo_stack_observations(c, obs_h, act_h, T + 1, stacked);          /* :355 */
        uint32_t legal = mzo_env_legal_mask(c, &env);
        t->game_id = game_id; t->move_idx = T + 1;
        if (opponent != MZO_OPP_SELF && p != muzero_player) {             /* :358-363 */
            int action = mzo_opponent_action(c, &env, opponent, game_id, T + 1);
            mzo_env_step(c, &env, action);
            float reward = (float)mzo_env_reward(c, &env, p);
            done = mzo_env_is_terminated(c, &env);
            /* store_search_stats!(history, root, ...) with the root of the previous search (:374); before the first search `root`
             * is the Int 0 and the reference would throw: zeros */
            for (int a = 0; a < c->A; a++) cv_h[(size_t)T * c->A + a] = T > 0 ? cv_h[(size_t)(T - 1) * c->A + a] : 0.0f;
            rv_h[T] = T > 0 ? rv_h[T - 1] : 0.0f;
            act_h[T] = action; rew_h[T] = reward; tp_h[T] = p;
            T++;
            continue;
        }
        node_t *root;
        run_mcts(t, stacked, legal, p, 1, &root, NULL);                   /* :359 exploration hard-coded true (Q11) */
        *sims += c->num_iters;
        int counts[MZO_MAX_A], acts[MZO_MAX_A], n = 0, sum_visits = 0;
        for (int j = 0; j < c->A; j++) { int a = c->child_order[j]; node_t *ch = root->children[a - 1]; if (ch) { counts[n] = ch->visit_count; acts[n] = a; n++; sum_visits += ch->visit_count; } }
        int action = select_action_counts(c, counts, acts, n, temperature, game_id, T + 1); /* :360 */
        mzo_env_step(c, &env, action);                                    /* :366 */
        float reward = (float)mzo_env_reward(c, &env, p);                 /* :367 */
        done = mzo_env_is_terminated(c, &env);                            /* :368 */
        for (int a = 1; a <= c->A; a++) {                                 /* store_search_stats! :115-122 (Q12) */
            node_t *ch = root->children[a - 1];
            cv_h[(size_t)T * c->A + (a - 1)] = ch ? (float)((double)ch->visit_count / (double)sum_visits) : 0.0f;
        }
        rv_h[T] = node_value(root);
        act_h[T] = action; rew_h[T] = reward; tp_h[T] = p;                /* :377-379 */
        T++;
    }
    *T_out = T;
}
static void *sp_worker(void *arg) {
    sp_job_t *j = (sp_job_t *)arg; const mzo_config *c = &j->m->cfg;
    int Tmax = c->max_moves + 1, on = obs_size(c);
    tree_t *t = tree_alloc(j->m);
    for (int g = j->lo; g < j->hi; g++)
    {
        play_game(j->m, t, j->first_game + (uint64_t)g, j->temperature, &j->T[g], j->obs + (size_t)g * Tmax * on,
                  j->actions + (size_t)g * Tmax, j->rewards + (size_t)g * Tmax, j->to_play + (size_t)g * Tmax,
                  j->child_visits + (size_t)g * Tmax * c->A, j->root_values + (size_t)g * Tmax, &j->sims, j->opponent, j->muzero_player);
        if (j->outcome) j->outcome[g] = mzo_arena_outcome(c, j->T[g], j->actions + (size_t)g * Tmax, j->muzero_player);
    }
    tree_free(t);
    return NULL;
}
int64_t mzo_self_play(const mzo_config *cfg, const float *blob, uint64_t first_game, int n_games, float temperature, int nthreads,
                      int32_t *T, float *obs, int32_t *actions, float *rewards, int32_t *to_play, float *child_visits, float *root_values) {
    return mzo_arena(cfg, blob, first_game, n_games, MZO_OPP_SELF, 1, temperature, nthreads, T, obs, actions, rewards, to_play, child_visits, root_values, NULL);
}
/* competitive_play! (SelfPlay.jl:421-435) over many games: play_game with an opponent on the other side */
int64_t mzo_arena(const mzo_config *cfg, const float *blob, uint64_t first_game, int n_games, int opponent, int muzero_player, float temperature, int nthreads,
                  int32_t *T, float *obs, int32_t *actions, float *rewards, int32_t *to_play, float *child_visits, float *root_values, int32_t *outcome) {
    model_t m; model_init(&m, cfg, blob);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n_games) nthreads = n_games > 0 ? n_games : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    sp_job_t *jobs = (sp_job_t *)calloc((size_t)nthreads, sizeof(sp_job_t));
    for (int i = 0; i < nthreads; i++) {
        sp_job_t *j = &jobs[i];
        j->m = &m; j->first_game = first_game; j->temperature = temperature; j->opponent = opponent; j->muzero_player = muzero_player; j->outcome = outcome;
        j->lo = (int)((int64_t)n_games * i / nthreads); j->hi = (int)((int64_t)n_games * (i + 1) / nthreads);
        j->T = T; j->obs = obs; j->actions = actions; j->rewards = rewards; j->to_play = to_play; j->child_visits = child_visits; j->root_values = root_values;
        if (nthreads == 1) sp_worker(j); else pthread_create(&th[i], NULL, sp_worker, j);
    }
    int64_t sims = 0;
    for (int i = 0; i < nthreads; i++) { if (nthreads > 1) pthread_join(th[i], NULL); sims += jobs[i].sims; }
    free(th); free(jobs);
    return sims;
}

/* ------------------------------------------------------------------------------------------
 * Replay / targets: src/ReplayBuffer.jl.
 * ------------------------------------------------------------------------------------------ */
/* conf.discount^i for Float32^Int (Julia Base: i<=3 by products, else llvm.pow.f32). */
static float discount_pow(float g, int i) {
    if (i == 0) return 1.0f;
    if (i == 1) return g;
    if (i == 2) return g * g;
    if (i == 3) return g * g * g;
    return powf(g, (float)i);
}
/* compute_target_value (ReplayBuffer.jl:5-20), Q17.  index is 1-based. */
float mzo_compute_target_value(const mzo_config *c, int T, const float *rewards, const int32_t *to_play, const float *root_values, int index) {
    int bootstrap = index + c->td_steps;
    if (bootstrap < T) {
        float last = to_play[bootstrap - 1] == to_play[index - 1] ? root_values[bootstrap - 1] : -root_values[bootstrap - 1];
        float value = last * discount_pow(c->discount, c->td_steps);
        int i = 1;
        for (int ri = index; ri <= bootstrap; ri++, i++) {               /* enumerate(reward_history[index:bootstrap]) */
            float r = rewards[ri - 1];
            value = value + (to_play[index - 1] == to_play[index + i - 1] ? r : -r) * discount_pow(c->discount, i);
        }
        return value;
    }
    return 0.0f;
}
/* get_batch (ReplayBuffer.jl:188-217) with sample_n_games/sample_position uniform branches (:102-104, :80)
 * and make_target (:25-50, Q18).  RNG contract: Philox(seed, REPLAY, step, b) -> [0]: game, [1]: position;
 * absorbing-state actions Philox(seed, ABSORB, step, b, unroll offset). */
/* ------------------------------------------------------------------------------------------
 * Prioritised replay (conf.PER = true): REPAIRED specification of ReplayBuffer.jl:73-100, 136-145, 168-183, 213-215 and
 * Learning.jl:400-404.  As written the reference cannot run with PER: update_priorities! indexes priority[1:end_index -
 * start_index + 1] with up to K+2 elements of a K+1 vector (ReplayBuffer.jl:176-178, BoundsError), and the normalisations sum
 * Float32 vectors in Dict iteration order.  Contract here:
 *   - priority p = |root_value - compute_target_value|^PER_alpha (Float32, Julia's Float32^Int), quantised to fixed point
 *     q = max(1, round(p * 2^16)) so that sums and prefix sums are exact integers (order independent, scan friendly) and every
 *     stored position stays sampleable; game priority = max over its positions (:145);
 *   - games in ascending key order; game ~ q_game / Q, position ~ q_pos / Q_game, by inverse CDF with 64-bit draws
 *     mulhi(u64(philox.x, philox.z), Q) and mulhi(u64(philox.y, philox.w), Q_game) of Philox(seed, REPLAY, step, b);
 *   - weight = 1 / (total_samples * game_prob * pos_prob), probabilities as Float32(q / Q), then ./= maximum (:213-215);
 *   - update_priorities!: rows k = 0 .. min(K, T - pos) of |predicted_values - target_values|^alpha are written to positions
 *     pos + k, batch elements in order (last write wins), then the game priority is recomputed.
 * ------------------------------------------------------------------------------------------ */
static float pow_int_f32(float x, int i) { /* Float32 ^ Int as Julia evaluates it: products for i <= 3, llvm.pow.f32 beyond */
    if (i == 0) return 1.0f;
    if (i == 1) return x;
    if (i == 2) return x * x;
    if (i == 3) return x * x * x;
    return powf(x, (float)i);
}
uint32_t mzo_per_quantise(float p) {
    float s = p * 65536.0f;
    if (!(s >= 1.0f)) return 1u;
    if (s > 4.0e9f) return 4000000000u;
    return (uint32_t)llrintf(s);
}
void mzo_per_priorities(const mzo_config *c, int T, const float *rewards, const int32_t *to_play, const float *root_values,
                        uint32_t *q_pos, uint32_t *q_game) { /* save_game, ReplayBuffer.jl:136-145 */
    uint32_t mx = 0;
    for (int i = 1; i <= T; i++) {
        float p = pow_int_f32(fabsf(root_values[i - 1] - mzo_compute_target_value(c, T, rewards, to_play, root_values, i)), c->per_alpha);
        q_pos[i - 1] = mzo_per_quantise(p);
        if (q_pos[i - 1] > mx) mx = q_pos[i - 1];
    }
    *q_game = mx;
}
static inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
void mzo_per_update(const mzo_config *c, int B, const int32_t *index_batch, const float *pred_values, const float *target_values,
                    int n_games, int64_t first_key, const int32_t *T, uint32_t *q_pos, uint32_t *q_game) {
    int Tmax = c->max_moves + 1, K1 = c->num_unroll_steps + 1;
    for (int b = 0; b < B; b++) {
        int64_t gi = (int64_t)index_batch[2 * b] - first_key; int pos = index_batch[2 * b + 1];
        if (gi < 0 || gi >= n_games) continue;                             /* the game has left the buffer (:172) */
        int Tg = T[gi];
        for (int k = 0; k < K1 && pos + k <= Tg; k++)
            q_pos[(size_t)gi * Tmax + pos + k - 1] = mzo_per_quantise(pow_int_f32(fabsf(pred_values[(size_t)b * K1 + k] - target_values[(size_t)b * K1 + k]), c->per_alpha));
        uint32_t mx = 0;
        for (int i = 0; i < Tg; i++) if (q_pos[(size_t)gi * Tmax + i] > mx) mx = q_pos[(size_t)gi * Tmax + i];
        q_game[gi] = mx;
    }
}

static void get_batch_impl(const mzo_config *c, int n_games, int64_t first_key, const int32_t *T, const float *obs, const int32_t *actions,
                   const float *rewards, const int32_t *to_play, const float *child_visits, const float *root_values, uint64_t step,
                   int32_t *index_batch, float *obs_batch, float *action_batch, float *value_batch, float *reward_batch,
                   float *policy_batch, float *gscale, const uint32_t *q_pos, const uint32_t *q_game, float *weights) {
    int Tmax = c->max_moves + 1, on = obs_size(c), ss = stack_size(c), K1 = c->num_unroll_steps + 1, A = c->A;
    uint64_t Q = 0; int64_t total_samples = 0;
    if (q_game) for (int g = 0; g < n_games; g++) { Q += q_game[g]; total_samples += T[g]; }
    for (int b = 0; b < c->batch_size; b++) {
        uint32_t r[4]; mzo_philox(c->seed, STREAM_REPLAY, (uint32_t)step, (uint32_t)b, 0, 0, r);
        int gi, pos;
        if (q_game) {   /* prioritised: sample_n_games :90-100, sample_position :75-79 */
            uint64_t t = mulhi64(((uint64_t)r[0] << 32) | r[2], Q), acc = 0;
            for (gi = 0; gi < n_games - 1; gi++) { acc += q_game[gi]; if (acc > t) break; }
            const uint32_t *qp = q_pos + (size_t)gi * Tmax;
            uint64_t Qg = 0; for (int i = 0; i < T[gi]; i++) Qg += qp[i];
            uint64_t t2 = mulhi64(((uint64_t)r[1] << 32) | r[3], Qg), a2 = 0;
            for (pos = 1; pos < T[gi]; pos++) { a2 += qp[pos - 1]; if (a2 > t2) break; }
            float game_prob = (float)((double)q_game[gi] / (double)Q), pos_prob = (float)((double)qp[pos - 1] / (double)Qg);
            weights[b] = 1.0f / (((float)total_samples * game_prob) * pos_prob);             /* :213 */
        } else {
            gi = (int)u32_below(r[0], (uint32_t)n_games);
            pos = 1 + (int)u32_below(r[1], (uint32_t)T[gi]);              /* rand(1:length(root_values)) */
        }
        int Tg = T[gi];
        const float *g_obs = obs + (size_t)gi * Tmax * on; const int32_t *g_act = actions + (size_t)gi * Tmax;
        const float *g_rew = rewards + (size_t)gi * Tmax; const int32_t *g_tp = to_play + (size_t)gi * Tmax;
        const float *g_cv = child_visits + (size_t)gi * Tmax * A; const float *g_rv = root_values + (size_t)gi * Tmax;
        index_batch[2 * b] = (int32_t)(first_key + gi); index_batch[2 * b + 1] = pos;
        for (int k = 0; k < K1; k++) {                                    /* make_target :28-48 */
            int ci = pos + k; float tv, tr; int act;
            float *pol = policy_batch + ((size_t)b * K1 + k) * A;
            if (ci < Tg) {
                tv = mzo_compute_target_value(c, Tg, g_rew, g_tp, g_rv, ci); tr = g_rew[ci - 1];
                for (int a = 0; a < A; a++) pol[a] = g_cv[(size_t)(ci - 1) * A + a];
                act = g_act[ci - 1];
            } else if (ci == Tg) {
                tv = 0.0f; tr = g_rew[ci - 1];
                for (int a = 0; a < A; a++) pol[a] = 1.0f / (float)A;
                act = g_act[ci - 1];
            } else {
                tv = 0.0f; tr = 0.0f;
                for (int a = 0; a < A; a++) pol[a] = 1.0f / (float)A;
                uint32_t q[4]; mzo_philox(c->seed, STREAM_ABSORB, (uint32_t)step, (uint32_t)b, (uint32_t)k, 0, q);
                act = 1 + (int)u32_below(q[0], (uint32_t)A);              /* rand(rng, conf.action_space) */
            }
            value_batch[(size_t)b * K1 + k] = tv; reward_batch[(size_t)b * K1 + k] = tr; action_batch[(size_t)b * K1 + k] = (float)act;
        }
        mzo_stack_observations(c, g_obs, g_act, pos, obs_batch + (size_t)b * ss);   /* :207 */
        int gs = Tg + 1 - pos; if (c->num_unroll_steps < gs) gs = c->num_unroll_steps; /* :212 */
        gscale[b] = (float)gs;
    }
    if (q_game) {   /* weight_batch ./= maximum(weight_batch), :215 */
        float mx = weights[0];
        for (int b = 1; b < c->batch_size; b++) mx = weights[b] > mx ? weights[b] : mx;
        for (int b = 0; b < c->batch_size; b++) weights[b] = weights[b] / mx;
    }
}
void mzo_get_batch(const mzo_config *c, int n_games, int64_t first_key, const int32_t *T, const float *obs, const int32_t *actions,
                   const float *rewards, const int32_t *to_play, const float *child_visits, const float *root_values, uint64_t step,
                   int32_t *index_batch, float *obs_batch, float *action_batch, float *value_batch, float *reward_batch,
                   float *policy_batch, float *gscale) {
    get_batch_impl(c, n_games, first_key, T, obs, actions, rewards, to_play, child_visits, root_values, step, index_batch, obs_batch,
                   action_batch, value_batch, reward_batch, policy_batch, gscale, NULL, NULL, NULL);
}
void mzo_get_batch_per(const mzo_config *c, int n_games, int64_t first_key, const int32_t *T, const float *obs, const int32_t *actions,
                       const float *rewards, const int32_t *to_play, const float *child_visits, const float *root_values,
                       const uint32_t *q_pos, const uint32_t *q_game, uint64_t step, int32_t *index_batch, float *obs_batch,
                       float *action_batch, float *value_batch, float *reward_batch, float *policy_batch, float *gscale, float *weights) {
    get_batch_impl(c, n_games, first_key, T, obs, actions, rewards, to_play, child_visits, root_values, step, index_batch, obs_batch,
                   action_batch, value_batch, reward_batch, policy_batch, gscale, q_pos, q_game, weights);
}

/* ------------------------------------------------------------------------------------------
 * Learner: src/Learning.jl:261-304 (loss, make_dynamics_input), 347-397 (unroll, update).
 * ------------------------------------------------------------------------------------------ */
static float sqnorm_net(const float *blob, const net_t *n) { /* sum(sqnorm, params): per array, then across arrays */
    float total = 0.0f; int first = 1;
    for (int part = 0; part < 3; part++) {
        int cnt = part == 0 ? n->n_trunk : part == 1 ? n->n_h1 : n->n_h2;
        const layer_t *ls = part == 0 ? n->trunk : part == 1 ? n->h1 : n->h2;
        for (int l = 0; l < cnt; l++) {
            float sw = 0.0f, sb = 0.0f;
            for (int i = 0; i < ls[l].in * ls[l].out; i++) sw = sw + blob[ls[l].w_off + i] * blob[ls[l].w_off + i];
            for (int i = 0; i < ls[l].out; i++) sb = sb + blob[ls[l].b_off + i] * blob[ls[l].b_off + i];
            if (first) { total = sw; first = 0; } else total = total + sw;
            total = total + sb;
            if (ls[l].bn) {   /* Flux.params: ..., BatchNorm beta, gamma */
                float s1 = 0.0f, s2 = 0.0f;
                for (int i = 0; i < ls[l].out; i++) s1 = s1 + blob[ls[l].beta_off + i] * blob[ls[l].beta_off + i];
                for (int i = 0; i < ls[l].out; i++) s2 = s2 + blob[ls[l].gamma_off + i] * blob[ls[l].gamma_off + i];
                total = total + s1; total = total + s2;
            }
        }
    }
    return total;
}

/* ResNet: sum(sqnorm, Flux.params(net)) -- Conv weight, bias, BatchNorm beta and gamma, Dense weight and bias; the running statistics
 * mu / var are not parameters (Flux.trainable(BatchNorm) = (beta, gamma)) */
static float rn_sqnorm_net(const mzo_config *c, const float *blob, int net) {
    rn_model_t m; rn_build(c, &m);
    float total = 0.0f; int first = 1;
    for (int ui = 0; ui < m.net[net].n; ui++) {
        const rn_unit_t *u = &m.net[net].u[ui];
        int offs[4] = {u->w_off, u->b_off, u->beta_off, u->gamma_off}, cnt[4] = {u->k * u->k * u->cin * u->cout, u->cout, u->cout, u->cout};
        for (int a = 0; a < (u->kind == RN_CONV ? 4 : 2); a++) {
            float s = 0.0f;
            for (int i = 0; i < cnt[a]; i++) s = s + blob[offs[a] + i] * blob[offs[a] + i];
            if (first) { total = s; first = 0; } else total = total + s;
        }
    }
    return total;
}
/* 1 where the blob entry is a Flux parameter (what the reference_l2 gradient 2*theta and ADAM touch), 0 for the BatchNorm statistics */
void mzo_trainable_mask(const mzo_config *c, unsigned char *mask) {
    int np = mzo_num_params(c, 3);
    memset(mask, 1, (size_t)np);
    if (c->net_type != 1) {
        if (!c->use_batch_norm) return;
        net_t nets[3]; build_nets(c, nets);
        for (int n = 0; n < 3; n++) for (int part = 0; part < 3; part++) {
            int cnt = part == 0 ? nets[n].n_trunk : part == 1 ? nets[n].n_h1 : nets[n].n_h2;
            const layer_t *ls = part == 0 ? nets[n].trunk : part == 1 ? nets[n].h1 : nets[n].h2;
            for (int l = 0; l < cnt; l++) if (ls[l].bn) for (int i = 0; i < ls[l].out; i++) { mask[ls[l].mu_off + i] = 0; mask[ls[l].var_off + i] = 0; }
        }
        return;
    }
    rn_model_t m; rn_build(c, &m);
    for (int n = 0; n < 3; n++) for (int ui = 0; ui < m.net[n].n; ui++) {
        const rn_unit_t *u = &m.net[n].u[ui];
        if (u->kind != RN_CONV) continue;
        for (int i = 0; i < u->cout; i++) { mask[u->mu_off + i] = 0; mask[u->var_off + i] = 0; }
    }
}

static void learn_forward_impl(const mzo_config *c, const float *blob, int B, const float *obs_batch, const float *action_batch,
                       const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale,
                       const float *weights /* PER importance weights (1,B), or NULL: weight_batch = 1.0f0 (:266-268) */,
                       float *pred_values, float *pred_rewards, float *pred_policies, float *losses) {
    model_t m; model_init(&m, c, blob);
    int K = c->num_unroll_steps, K1 = K + 1, A = c->A, ss = stack_size(c), plane = c->W * c->H;
    int on = c->net_type == 1 ? c->hidden_state_size : obs_size(c);   /* ResNet: the state is the (W,H,num_filters) hidden state itself */
    static __thread float h[MZO_MAX_HIDDEN], nh[MZO_MAX_HIDDEN], sa[MZO_MAX_HIDDEN + 256];
    for (int b = 0; b < B; b++) {
        float v, r;
        representation(&m, obs_batch + (size_t)b * ss, h);                           /* :347 */
        prediction(&m, h, &v, pred_policies + ((size_t)b * K1) * A);                 /* :351 */
        pred_values[(size_t)b * K1] = v; pred_rewards[(size_t)b * K1] = 0.0f;        /* :352 zeros */
        for (int i = 1; i <= K; i++) {                                               /* :355-370 (Q19) */
            prediction(&m, h, &v, pred_policies + ((size_t)b * K1 + i) * A);         /* BEFORE stepping dynamics */
            float av = action_batch[(size_t)b * K1 + (i - 1)] / (float)A;            /* make_dynamics_input :294 (Float32 divide) */
            for (int j = 0; j < on; j++) sa[j] = h[j] * 2.0f;                         /* :299 (copy, no aliasing) */
            for (int j = 0; j < plane; j++) sa[on + j] = av;
            dynamics(&m, sa, nh, &r);                                                /* :362 */
            memcpy(h, nh, sizeof(float) * (size_t)c->hidden_state_size);
            pred_values[(size_t)b * K1 + i] = v; pred_rewards[(size_t)b * K1 + i] = r;
        }
    }
    /* loss (:261-288), Q21 */
    float vsum = 0.0f; double rsum = 0.0;
    float *S = (float *)malloc(sizeof(float) * (size_t)B);
    for (int b = 0; b < B; b++) {
        float sv = 0.0f; double sr = 0.0; float sp = 0.0f;
        for (int k = 0; k < K1; k++) {
            float d = pred_values[(size_t)b * K1 + k] - value_batch[(size_t)b * K1 + k];
            sv = sv + d * d;
            double dr = (double)pred_rewards[(size_t)b * K1 + k] - (double)reward_batch[(size_t)b * K1 + k];
            sr = sr + dr * dr;
            /* logitcrossentropy: -sum(y .* logsoftmax(yhat)) with yhat ALREADY softmaxed */
            const float *p = pred_policies + ((size_t)b * K1 + k) * A, *y = policy_batch + ((size_t)b * K1 + k) * A;
            float mx = p[0];
            for (int a = 1; a < A; a++) mx = p[a] > mx ? p[a] : mx;
            float se = 0.0f;
            for (int a = 0; a < A; a++) se = se + mzo_expf(p[a] - mx);
            float lse = mzo_logf(se), acc = 0.0f;
            for (int a = 0; a < A; a++) acc = acc + y[a] * ((p[a] - mx) - lse);
            sp = sp + (-acc);
        }
        vsum = vsum + (weights ? (sv / gscale[b]) * weights[b] : sv / gscale[b]);
        rsum = rsum + (weights ? (sr / (double)gscale[b]) * (double)weights[b] : sr / (double)gscale[b]);
        S[b] = sp;
    }
    float value_loss = vsum / (float)B;
    /* sum(x,dims=2) is (1,1,B), gscale is (1,B): the broadcast is (1,B,B); mean over all B*B entries */
    float psum = 0.0f;
    for (int j = 0; j < B; j++) for (int i = 0; i < B; i++) psum = psum + (weights ? (S[j] / gscale[i]) * weights[i] : S[j] / gscale[i]);
    float policy_loss = psum / (float)(B * B);
    free(S);
    float data_loss;
    if (c->intermediate_rewards) data_loss = (float)(((double)value_loss + rsum / (double)B) + (double)policy_loss);
    else data_loss = (value_loss + 0.0f) + policy_loss;
    for (int n = 0; n < 3; n++) losses[n] = data_loss + (c->net_type == 1 ? rn_sqnorm_net(c, blob, n) : sqnorm_net(blob, &m.nets[n]));
}

void mzo_learn_forward(const mzo_config *c, const float *blob, int B, const float *obs_batch, const float *action_batch,
                       const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale,
                       float *pred_values, float *pred_rewards, float *pred_policies, float *losses) {
    learn_forward_impl(c, blob, B, obs_batch, action_batch, value_batch, reward_batch, policy_batch, gscale, NULL, pred_values, pred_rewards, pred_policies, losses);
}
void mzo_learn_forward_w(const mzo_config *c, const float *blob, int B, const float *obs_batch, const float *action_batch,
                         const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale, const float *weights,
                         float *pred_values, float *pred_rewards, float *pred_policies, float *losses) {
    learn_forward_impl(c, blob, B, obs_batch, action_batch, value_batch, reward_batch, policy_batch, gscale, weights, pred_values, pred_rewards, pred_policies, losses);
}

double mzo_cos_schedule(int t) { /* ParameterSchedulers.Cos(l0=1e-4, l1=1e-1, period=10), Learning.jl:319 (Q22) */
    double l0 = 1e-4, l1 = 1e-1, period = 10.0;
    double g = (1.0 + cos(2.0 * 3.14159265358979323846 * (double)(t - 1) / period)) / 2.0;
    return fabs(l0 - l1) * g + (l0 < l1 ? l0 : l1);
}

/* Flux.ADAM apply! + update! (Flux 0.12.4, un-vendored), Q22.  beta powers = beta^t by repeated product. */
static void adam_update(float *theta, float *m, float *v, const float *grad, int n, double eta, int t) {
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    double bp1 = b1, bp2 = b2;
    for (int i = 1; i < t; i++) { bp1 *= b1; bp2 *= b2; }
    for (int i = 0; i < n; i++) {
        float g = grad[i];
        m[i] = (float)(b1 * (double)m[i] + (1.0 - b1) * (double)g);
        v[i] = (float)(b2 * (double)v[i] + (1.0 - b2) * (double)(g * g));
        float delta = (float)((double)m[i] / (1.0 - bp1) / (sqrt((double)v[i] / (1.0 - bp2)) + eps) * eta);
        theta[i] = theta[i] - delta;
    }
}

/* ------------------------------------------------------------------------------------------
 * grad_mode = MZO_GRAD_BPTT: the gradient of the reference's OWN loss value (Learning.jl:261-288, with
 * Q19's unroll order and Q21's policy-term broadcast) taken through the reference's own forward unroll
 * (:347-374) -- what Zygote.pullback would return had the predictions been computed inside the closure
 * (Q20 explains why the reference's actual gradient degenerates to 2*theta).  Every parameter array gets
 * d(data loss)/d(theta) + 2*theta (each net's loss adds its own sum(abs2, theta), :287).
 *
 * The backward arithmetic is Float64.  The forward activations it linearises around are either the
 * Float32 contract's (fwd64 = 0: identical relu masks to the CUDA path) or an all-Float64 forward
 * (fwd64 = 1: used by the finite-difference check of the formulae, tests/test_oracle_kat.py).
 * ------------------------------------------------------------------------------------------ */
#define BP_MAXW 256
typedef struct { double in[BP_MAXW]; double out[24][BP_MAXW]; } chain_act_t;
typedef struct { chain_act_t trunk, h1, h2; } net_act_t;

static void dense_fwd(const float *blob, const double *wd, const layer_t *l, const double *x, double *y, int fwd64) {
    if (fwd64) {
        for (int o = 0; o < l->out; o++) {
            double acc = 0.0;
            for (int k = 0; k < l->in; k++) acc += wd[l->w_off + o + (size_t)l->out * k] * x[k];
            acc += wd[l->b_off + o];
            if (l->bn) acc = wd[l->gamma_off + o] * (acc - wd[l->mu_off + o]) / sqrt(wd[l->var_off + o] + (double)1e-5f) + wd[l->beta_off + o];   /* BatchNorm, test mode */
            y[o] = l->act == ACT_RELU ? (acc > 0.0 ? acc : 0.0) : l->act == ACT_TANH ? tanh(acc) : acc;
        }
    } else {
        float xf[BP_MAXW], yf[BP_MAXW];
        for (int k = 0; k < l->in; k++) xf[k] = (float)x[k];
        dense(blob, l, xf, yf);
        for (int o = 0; o < l->out; o++) y[o] = (double)yf[o];
    }
}
static void chain_fwd(const float *blob, const double *wd, const layer_t *ls, int n, const double *x, chain_act_t *a, int fwd64) {
    for (int k = 0; k < ls[0].in; k++) a->in[k] = x[k];
    const double *cur = a->in;
    for (int l = 0; l < n; l++) { dense_fwd(blob, wd, &ls[l], cur, a->out[l], fwd64); cur = a->out[l]; }
}
/* dy = gradient w.r.t. the post-activation output of the chain's last layer; G accumulates dW (blob order) and db;
 * dx (may be NULL) receives the gradient w.r.t. the chain input. */
static void chain_bwd(const double *wd, double *G, const layer_t *ls, int n, const chain_act_t *a, const double *dy, double *dx) {
    double cur[BP_MAXW], nxt[BP_MAXW];
    for (int o = 0; o < ls[n - 1].out; o++) cur[o] = dy[o];
    for (int l = n - 1; l >= 0; l--) {
        const layer_t *L = &ls[l];
        const double *y = a->out[l], *x = l ? a->out[l - 1] : a->in;
        for (int o = 0; o < L->out; o++) {
            double d = L->act == ACT_RELU ? (y[o] > 0.0 ? 1.0 : 0.0) : L->act == ACT_TANH ? 1.0 - y[o] * y[o] : 1.0;
            cur[o] *= d;
            if (L->bn) {   /* y = act(gamma * (z - mu) / sd + beta), z = W x + b: beta and gamma are Flux parameters, mu and sigma2 are not (no gradient) */
                double z = wd[L->b_off + o], sd = sqrt(wd[L->var_off + o] + (double)1e-5f);
                for (int k = 0; k < L->in; k++) z += wd[L->w_off + o + (size_t)L->out * k] * x[k];
                G[L->beta_off + o] += cur[o];
                G[L->gamma_off + o] += cur[o] * (z - wd[L->mu_off + o]) / sd;
                cur[o] *= wd[L->gamma_off + o] / sd;
            }
            G[L->b_off + o] += cur[o];
        }
        for (int k = 0; k < L->in; k++) {
            double s = 0.0;
            for (int o = 0; o < L->out; o++) { G[L->w_off + o + (size_t)L->out * k] += cur[o] * x[k]; s += wd[L->w_off + o + (size_t)L->out * k] * cur[o]; }
            nxt[k] = s;
        }
        for (int k = 0; k < L->in; k++) cur[k] = nxt[k];
    }
    if (dx) for (int k = 0; k < ls[0].in; k++) dx[k] = cur[k];
}
static void net_fwd(const float *blob, const double *wd, const net_t *n, const double *x, net_act_t *a, int fwd64) {
    chain_fwd(blob, wd, n->trunk, n->n_trunk, x, &a->trunk, fwd64);
    if (n->n_h1) chain_fwd(blob, wd, n->h1, n->n_h1, a->trunk.out[n->n_trunk - 1], &a->h1, fwd64);
    if (n->n_h2) chain_fwd(blob, wd, n->h2, n->n_h2, a->trunk.out[n->n_trunk - 1], &a->h2, fwd64);
}
/* d1 / d2: gradients w.r.t. the two heads' outputs (NULL = zero); dx: gradient w.r.t. the net input */
static void net_bwd(const double *wd, double *G, const net_t *n, const net_act_t *a, const double *d1, const double *d2, double *dx) {
    double dt[BP_MAXW], tmp[BP_MAXW];
    int w = n->trunk[n->n_trunk - 1].out;
    if (n->n_h1 == 0) { chain_bwd(wd, G, n->trunk, n->n_trunk, &a->trunk, d1, dx); return; }
    for (int k = 0; k < w; k++) dt[k] = 0.0;
    if (d1) { chain_bwd(wd, G, n->h1, n->n_h1, &a->h1, d1, tmp); for (int k = 0; k < w; k++) dt[k] += tmp[k]; }
    if (d2) { chain_bwd(wd, G, n->h2, n->n_h2, &a->h2, d2, tmp); for (int k = 0; k < w; k++) dt[k] += tmp[k]; }
    chain_bwd(wd, G, n->trunk, n->n_trunk, &a->trunk, dt, dx);
}

/* grad[n_params] (blob order, Float64) and the Float64 data loss.  perturb_index >= 0 adds perturb_delta to that
 * parameter of the Float64 weight copy first (finite differences; only meaningful with fwd64 = 1). */
static double learn_gradients_impl(const mzo_config *c, const float *blob, int B, const float *obs_batch, const float *action_batch,
                           const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale, const float *weights,
                           int fwd64, int perturb_index, double perturb_delta, double *grad) {
    model_t m; model_init(&m, c, blob);
    int K = c->num_unroll_steps, K1 = K + 1, A = c->A, ss = stack_size(c), on = obs_size(c), plane = c->W * c->H, hs = c->hidden_state_size;
    int np = mzo_num_params(c, 3);
    double *wd = (double *)malloc(sizeof(double) * (size_t)np);
    for (int i = 0; i < np; i++) wd[i] = (double)blob[i];
    if (perturb_index >= 0) wd[perturb_index] += perturb_delta;
    if (grad) for (int i = 0; i < np; i++) grad[i] = 0.0;
    net_act_t *ar = (net_act_t *)malloc(sizeof(net_act_t)), *ap = (net_act_t *)malloc(sizeof(net_act_t) * (size_t)K1), *ad = (net_act_t *)malloc(sizeof(net_act_t) * (size_t)(K ? K : 1));
    double G = 0.0;                                   /* mean_i(1/g_i): Q21's (1,B,B) broadcast factorises */
    for (int b = 0; b < B; b++) G += (weights ? (double)weights[b] : 1.0) / (double)gscale[b];
    G /= (double)B;
    double vsum = 0.0, rsum = 0.0, ssum = 0.0;
    for (int b = 0; b < B; b++) {
        double x[BP_MAXW], sa[BP_MAXW], pol[MZO_MAX_A * 40], q[MZO_MAX_A];
        for (int k = 0; k < ss; k++) x[k] = (double)obs_batch[(size_t)b * ss + k];
        net_fwd(blob, wd, &m.nets[0], x, ar, fwd64);                                        /* :347 */
        const double *h = ar->trunk.out[m.nets[0].n_trunk - 1];
        double g = (double)gscale[b] / (weights ? (double)weights[b] : 1.0);   /* value / reward terms: (sum / g) * w */
        /* rows: row 0 = prediction(h0) (:351); row i = prediction(h_{i-1}) BEFORE dynamics step i (:356-362, Q19) */
        for (int i = 0; i <= K; i++) {
            net_fwd(blob, wd, &m.nets[1], h, &ap[i], fwd64);
            if (i >= 1) {
                double av = fwd64 ? (double)action_batch[(size_t)b * K1 + (i - 1)] / (double)A
                                  : (double)(action_batch[(size_t)b * K1 + (i - 1)] / (float)A);
                for (int j = 0; j < on; j++) sa[j] = h[j] * 2.0;
                for (int j = 0; j < plane; j++) sa[on + j] = av;
                net_fwd(blob, wd, &m.nets[2], sa, &ad[i - 1], fwd64);
                h = ad[i - 1].h1.out[m.nets[2].n_h1 - 1];
            }
        }
        /* loss terms of this sample + gradients w.r.t. the predictions */
        double dv[40], dr[40], dz[40][MZO_MAX_A];
        for (int i = 0; i <= K; i++) {
            const net_act_t *pa = &ap[i];
            double v = pa->h1.out[m.nets[1].n_h1 - 1][0];
            const double *z = pa->h2.out[m.nets[1].n_h2 - 1];
            double tv = (double)value_batch[(size_t)b * K1 + i];
            vsum += (v - tv) * (v - tv) / g;
            dv[i] = 2.0 * (v - tv) / (g * (double)B);
            double r = i == 0 ? 0.0 : ad[i - 1].h2.out[m.nets[2].n_h2 - 1][0], tr = (double)reward_batch[(size_t)b * K1 + i];
            rsum += (r - tr) * (r - tr) / g;
            dr[i] = (c->intermediate_rewards && i >= 1) ? 2.0 * (r - tr) / (g * (double)B) : 0.0;
            /* policy: p = softmax(z) (:114); logitcrossentropy(p, y) = -sum(y .* logsoftmax(p)) (Q21: softmax twice) */
            double *p = pol + (size_t)i * A;
            if (fwd64) {
                double mx = z[0]; for (int a = 1; a < A; a++) mx = z[a] > mx ? z[a] : mx;
                double s = 0.0; for (int a = 0; a < A; a++) { p[a] = exp(z[a] - mx); s += p[a]; }
                for (int a = 0; a < A; a++) p[a] /= s;
            } else {
                float zf[MZO_MAX_A], pf[MZO_MAX_A];
                for (int a = 0; a < A; a++) zf[a] = (float)z[a];
                softmax(zf, A, pf);
                for (int a = 0; a < A; a++) p[a] = (double)pf[a];
            }
            const float *y = policy_batch + ((size_t)b * K1 + i) * A;
            double mx = p[0]; for (int a = 1; a < A; a++) mx = p[a] > mx ? p[a] : mx;
            double s = 0.0; for (int a = 0; a < A; a++) { q[a] = exp(p[a] - mx); s += q[a]; }
            double lse = log(s), ysum = 0.0, ce = 0.0;
            for (int a = 0; a < A; a++) { q[a] /= s; ysum += (double)y[a]; ce -= (double)y[a] * ((p[a] - mx) - lse); }
            ssum += ce;
            double dp[MZO_MAX_A], dot = 0.0, up = G / (double)B;                         /* d(policy_loss)/d(S_b) = G / B */
            for (int a = 0; a < A; a++) { dp[a] = up * (-(double)y[a] + q[a] * ysum); dot += p[a] * dp[a]; }
            for (int a = 0; a < A; a++) dz[i][a] = p[a] * (dp[a] - dot);
        }
        if (grad) {
            double dh[BP_MAXW], dsa[BP_MAXW], dhp[BP_MAXW];
            for (int k = 0; k < hs; k++) dh[k] = 0.0;                                       /* h_K feeds nothing */
            for (int i = K; i >= 1; i--) {
                net_bwd(wd, grad, &m.nets[2], &ad[i - 1], dh, &dr[i], dsa);                  /* dynamics step i produced h_i, r_i */
                net_bwd(wd, grad, &m.nets[1], &ap[i], &dv[i], dz[i], dhp);                   /* row i consumed h_{i-1} */
                for (int k = 0; k < hs; k++) dh[k] = 2.0 * dsa[k] + dhp[k];                  /* state * 2.0f0 (:299) */
            }
            net_bwd(wd, grad, &m.nets[1], &ap[0], &dv[0], dz[0], dhp);                       /* row 0 consumed h_0 */
            for (int k = 0; k < hs; k++) dh[k] += dhp[k];
            net_bwd(wd, grad, &m.nets[0], ar, dh, NULL, NULL);
        }
    }
    if (grad) {                                                                             /* sum(sqnorm, params), :287: Flux.params leave the BatchNorm statistics out */
        unsigned char *mask = (unsigned char *)malloc((size_t)np);
        mzo_trainable_mask(c, mask);
        for (int i = 0; i < np; i++) grad[i] = mask[i] ? grad[i] + 2.0 * wd[i] : 0.0;
        free(mask);
    }
    double value_loss = vsum / (double)B, policy_loss = (ssum / (double)B) * G;
    double data = value_loss + (c->intermediate_rewards ? rsum / (double)B : 0.0) + policy_loss;
    free(wd); free(ar); free(ap); free(ad);
    return data;
}

double mzo_learn_gradients(const mzo_config *c, const float *blob, int B, const float *obs_batch, const float *action_batch,
                           const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale,
                           int fwd64, int perturb_index, double perturb_delta, double *grad) {
    return learn_gradients_impl(c, blob, B, obs_batch, action_batch, value_batch, reward_batch, policy_batch, gscale, NULL, fwd64, perturb_index, perturb_delta, grad);
}
double mzo_learn_gradients_w(const mzo_config *c, const float *blob, int B, const float *obs_batch, const float *action_batch,
                             const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale, const float *weights,
                             int fwd64, int perturb_index, double perturb_delta, double *grad) {
    return learn_gradients_impl(c, blob, B, obs_batch, action_batch, value_batch, reward_batch, policy_batch, gscale, weights, fwd64, perturb_index, perturb_delta, grad);
}

/* Flux.ADAM on a caller-supplied gradient (lets the tests check the CUDA update bit-for-bit given the CUDA gradient) */
void mzo_adam_apply(float *blob, float *adam_m, float *adam_v, const float *grad, int n, int t) {
    adam_update(blob, adam_m, adam_v, grad, n, mzo_cos_schedule(t), t);
}

void mzo_learn_step(const mzo_config *c, float *blob, float *adam_m, float *adam_v, int t, int grad_mode, int B,
                    const float *obs_batch, const float *action_batch, const float *value_batch, const float *reward_batch,
                    const float *policy_batch, const float *gscale, float *losses) {
    int K1 = c->num_unroll_steps + 1, A = c->A, np = mzo_num_params(c, 3);
    float *pv = (float *)malloc(sizeof(float) * (size_t)B * K1), *pr = (float *)malloc(sizeof(float) * (size_t)B * K1);
    float *pp = (float *)malloc(sizeof(float) * (size_t)B * K1 * A), *grad = (float *)malloc(sizeof(float) * (size_t)np);
    mzo_learn_forward(c, blob, B, obs_batch, action_batch, value_batch, reward_batch, policy_batch, gscale, pv, pr, pp, losses);
    if (grad_mode == MZO_GRAD_BPTT) {
        double *gd = (double *)malloc(sizeof(double) * (size_t)np);
        mzo_learn_gradients(c, blob, B, obs_batch, action_batch, value_batch, reward_batch, policy_batch, gscale, 0, -1, 0.0, gd);
        for (int i = 0; i < np; i++) grad[i] = (float)gd[i];
        free(gd);
    } else {
        /* Q20: predictions are computed OUTSIDE Zygote.pullback (Learning.jl:347-374 vs 385-393), so the only
         * parameter-dependent term is sum(sqnorm, params): grad = 2*theta for every array. */
        unsigned char *mask = (unsigned char *)malloc((size_t)np);
        mzo_trainable_mask(c, mask);
        for (int i = 0; i < np; i++) grad[i] = mask[i] ? blob[i] + blob[i] : 0.0f;     /* ADAM leaves an entry with zero gradient and zero moments untouched */
        free(mask);
    }
    adam_update(blob, adam_m, adam_v, grad, np, mzo_cos_schedule(t), t);
    free(pv); free(pr); free(pp); free(grad);
}
