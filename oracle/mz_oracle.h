/*
 * mz_oracle.h -- CPU restatement of deveshjawla/MuZero.jl's self-play + learner hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library (muzero.jl_b200/csrc) never includes, links or calls this code.
 *
 * PARITY STATUS: "parity unpinned".  The reference is pure Julia (no Julia in this image), ships
 * no tests / golden vectors, and its arithmetic lives in un-vendored packages (Flux 0.12.4,
 * NNlib 0.7.22, Zygote 0.6.14, ReinforcementLearningBase 0.9.5, Distributions 0.25.6,
 * ParameterSchedulers 0.2.3; Manifest.toml).  This oracle follows the reference line by line
 * (every function cites file:line under /root/reference) and is pinned only by the hand-derived
 * known-answer vectors of SURVEY.md section 8c (tests/test_oracle_kat.py) and by the golden
 * fixtures it generated itself (tests/golden/, generator committed).
 *
 * Where the reference leaves arithmetic to a library or to an unseeded RNG, the oracle states
 * its own contract (DESIGN.md "Arithmetic contract"):
 *   - Dense: acc = 0; for k = 0..in-1 in order: acc = fmaf(W[o,k], x[k], acc); y = act(acc + b[o])
 *   - expf / logf / tanhf: fixed fmaf polynomials (Cephes-style), defined in mz_oracle.c
 *   - softmax: subtract max, exp, sum in index order, divide
 *   - all random draws: Philox4x32-10 keyed by (seed, stream) with counter (game, move, sim, depth)
 *   - Julia Dict{Int,Node} iteration order: permutation table cfg.child_order (default = the
 *     Julia<=1.10 hash order, [7,4,9,2,3,5,8,6,1] for keys 1..9)
 */
#ifndef MZ_ORACLE_H
#define MZ_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZO_MAX_A 16      /* max action-space size */
#define MZO_MAX_OBS 192   /* max W*H*C of one observation */
#define MZO_MAX_HIDDEN 4096 /* max hidden_state_size (ResNet: W*H*num_filters) */
#define MZO_MAX_T 64      /* max moves stored per game (max_moves+1) */

enum { MZO_GAME_TICTACTOE = 0, MZO_GAME_CONNECT = 1 };
enum { MZO_TIE_PHILOX = 0, MZO_TIE_FIRST = 1 };
enum { MZO_GRAD_REFERENCE_L2 = 0, MZO_GRAD_BPTT = 1 };

/* POD mirror of Config (src/Constructors.jl:18-52) + FeedForwardHP (:62-75). */
typedef struct mzo_config {
    int32_t game;                 /* MZO_GAME_* */
    int32_t W, H, C;              /* observation_shape (3,3,3) */
    int32_t A;                    /* length(action_space) */
    int32_t num_players;          /* length(players) */
    int32_t stacked_observations; /* 1 */
    int32_t max_moves;            /* 9 */
    int32_t num_iters;            /* S */
    int32_t num_unroll_steps;     /* K */
    int32_t td_steps;
    int32_t batch_size;
    int32_t replay_buffer_size;
    int32_t pb_c_base;            /* 19652 */
    int32_t intermediate_rewards; /* false */
    int32_t tie_mode;             /* MZO_TIE_* */
    float   pb_c_init;            /* 1.25 */
    float   discount;             /* 0.997 */
    float   dirichlet_alpha;      /* 0.25 */
    float   exploration_eps;      /* 0.25; 0 = noise off */
    uint64_t seed;                /* conf.seed = 1337 */
    int32_t child_order[MZO_MAX_A]; /* Julia Dict iteration order of keys 1..A (1-based actions) */
    /* FeedForwardHP */
    int32_t width_hidden, depth_representation, depth_prediction, depth_dynamics;
    int32_t depth_policy, depth_value, depth_reward, depth_state_head;
    int32_t hidden_state_size;
    int32_t reward_activation_tanh; /* 1 = tanh (params.jl:28), 0 = identity */
    /* ResNetHP (src/Constructors.jl:77-90) -- the REPAIRED spec of the residual networks (Learning.jl:148-255 as
     * written reference undefined names and never ran; see the ResNet section of mz_oracle.c) */
    int32_t net_type;               /* 0 = FeedForwardHP networks, 1 = ResNetHP networks */
    int32_t rn_num_blocks;          /* num_blocks */
    int32_t rn_num_filters;         /* num_filters: the hidden state is (W,H,num_filters) */
    int32_t rn_kernel;              /* conv_kernel_size = (k,k), odd; representation only (Learning.jl:195,230 fix (1,1) elsewhere) */
    int32_t rn_first_head_filters;  /* num_first_head_filters = 1 (value and reward heads) */
    int32_t rn_second_head_filters; /* num_second_head_filters = 2 (policy head) */
    int32_t per;                    /* conf.PER (params.jl:11: false) */
    int32_t per_alpha;              /* conf.PER_alpha (Constructors.jl:44: 1) */
    int32_t temperature_threshold;  /* conf.temperature_threshold (Constructors.jl:31): -1 = nothing */
    int32_t use_batch_norm;         /* FeedForwardHP.use_batch_norm (Constructors.jl:71): make_dense = Dense + BatchNorm(relu) (Learning.jl:70-79) */
} mzo_config;

void mzo_default_config(mzo_config *cfg);            /* params.jl:2-29 defaults */
int  mzo_num_params(const mzo_config *cfg, int net); /* net: 0 repr, 1 pred, 2 dyn, 3 total */
void mzo_init_weights(const mzo_config *cfg, uint64_t seed, float *blob); /* glorot_uniform, bias 0 */
void mzo_julia_dict_order(int A, int32_t *order);    /* Dict{Int,...} iteration order of 1..A */

/* bf16-operand emulation of the networks (checks the tensor-core path); 0 = exact Float32 (default) */
void mzo_set_bf16(int on);

/* math contract (exported so tests can pin them) */
float mzo_expf(float x);
float mzo_logf(float x);
float mzo_tanhf(float x);
void  mzo_philox(uint64_t seed, uint32_t stream, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);

/* ---- environment (games/tictactoe/game.jl) ---- */
typedef struct mzo_env { uint64_t p1, p2; int32_t player; int32_t moves; } mzo_env;
void  mzo_env_reset(const mzo_config *cfg, mzo_env *e);
void  mzo_env_step(const mzo_config *cfg, mzo_env *e, int action);     /* action 1..A */
uint32_t mzo_env_legal_mask(const mzo_config *cfg, const mzo_env *e);  /* bit (a-1) set = legal */
int   mzo_env_is_terminated(const mzo_config *cfg, const mzo_env *e);
int   mzo_env_reward(const mzo_config *cfg, const mzo_env *e, int player);
void  mzo_env_observation(const mzo_config *cfg, const mzo_env *e, float *obs /* W*H*C */);
/* census of the reachable state graph: out[0]=boards, [1]=terminal boards, [2]=complete games,
 * [3..3+16) = games by length, then (len,last mover,reward) histogram at out[32 + len*6 + (mover-1)*3 + (reward+1)] */
void  mzo_env_census(const mzo_config *cfg, int64_t *out /* >= 32+64*6 */);

/* ---- networks (src/Learning.jl:87-142) ---- */
void mzo_representation(const mzo_config *cfg, const float *blob, const float *stacked_obs, float *hidden);
void mzo_prediction(const mzo_config *cfg, const float *blob, const float *hidden, float *value, float *policy);
void mzo_dynamics(const mzo_config *cfg, const float *blob, const float *state_action, float *next_hidden, float *reward);

/* ---- stacking (src/SelfPlay.jl:128-149) ---- */
void mzo_stack_observations(const mzo_config *cfg, const float *obs_hist /* [T][obs] */, const int32_t *action_hist,
                            int index /* 1-based */, float *stacked);

/* ---- MCTS (src/SelfPlay.jl:230-285) ----
 * visit_counts / priors indexed by action-1; illegal = 0.  trace (optional, may be NULL):
 * per simulation [depth, leaf action, parent expansion id, value, reward] as floats (5*S). */
void mzo_run_mcts(const mzo_config *cfg, const float *blob, const float *stacked_obs, uint32_t legal_mask,
                  int to_play, int exploration, uint64_t game_id, int move_idx,
                  int32_t *visit_counts, float *root_value, float *root_priors, float *trace);
int  mzo_select_action(const mzo_config *cfg, const int32_t *visit_counts, uint32_t legal_mask, float temperature,
                       uint64_t game_id, int move_idx);

/* ---- self-play (src/SelfPlay.jl:330-419) ----
 * Plays games first_game .. first_game+n-1 on nthreads host threads.  Outputs are laid out like
 * GameHistory (Constructors.jl:6-16), padded to Tmax = max_moves+1:
 *   T[n]; obs[n][Tmax][obs]; actions[n][Tmax]; rewards[n][Tmax]; to_play[n][Tmax];
 *   child_visits[n][Tmax][A]; root_values[n][Tmax].  Returns total simulations run. */
int64_t mzo_self_play(const mzo_config *cfg, const float *blob, uint64_t first_game, int n_games, float temperature,
                      int nthreads, int32_t *T, float *obs, int32_t *actions, float *rewards, int32_t *to_play,
                      float *child_visits, float *root_values);

/* ---- competitive play (src/SelfPlay.jl:421-435, select_opponent_action :311-325) ----
 * play_game with `opponent` moving for the side that is not muzero_player; outcome[n] = +1 / 0 / -1 for MuZero (the side that
 * completed a line first wins).  Histories as in mzo_self_play; opponent plies repeat the previous search statistics. */
enum { MZO_OPP_SELF = 0, MZO_OPP_RANDOM = 1, MZO_OPP_EXPERT = 2 };
int mzo_opponent_action(const mzo_config *cfg, const mzo_env *e, int opponent, uint64_t game_id, int move_idx);
int mzo_arena_outcome(const mzo_config *cfg, int T, const int32_t *actions, int muzero_player);
int64_t mzo_arena(const mzo_config *cfg, const float *blob, uint64_t first_game, int n_games, int opponent, int muzero_player,
                  float temperature, int nthreads, int32_t *T, float *obs, int32_t *actions, float *rewards, int32_t *to_play,
                  float *child_visits, float *root_values, int32_t *outcome);

/* ---- replay / targets (src/ReplayBuffer.jl:5-50,188-217) ----
 * The buffer is the same padded layout holding n_games games whose keys (game numbers) are
 * first_key .. first_key+n_games-1.  Outputs follow get_batch's tuple (ReplayBuffer.jl:216):
 *   index_batch[B][2] (game key, position 1-based); obs[B][stack]; actions[B][K+1]; values[B][K+1];
 *   rewards[B][K+1]; policies[B][K+1][A]; gscale[B]. */
float mzo_compute_target_value(const mzo_config *cfg, int T, const float *rewards, const int32_t *to_play,
                               const float *root_values, int index /* 1-based */);
void mzo_get_batch(const mzo_config *cfg, int n_games, int64_t first_key, const int32_t *T, const float *obs,
                   const int32_t *actions, const float *rewards, const int32_t *to_play, const float *child_visits,
                   const float *root_values, uint64_t step, int32_t *index_batch, float *obs_batch,
                   float *action_batch, float *value_batch, float *reward_batch, float *policy_batch, float *gscale);

/* ---- learner (src/Learning.jl:261-304,347-397) ----
 * forward unroll (Q19): pred_values[B][K+1], pred_rewards[B][K+1], pred_policies[B][K+1][A]; losses[3] =
 * (representation, prediction, dynamics) = data loss + that net's own sum(theta^2) (Q21). */
void mzo_learn_forward(const mzo_config *cfg, const float *blob, int B, const float *obs_batch, const float *action_batch,
                       const float *value_batch, const float *reward_batch, const float *policy_batch,
                       const float *gscale, float *pred_values, float *pred_rewards, float *pred_policies,
                       float *losses);
/* one learning! iteration on a given batch: eta from the Cos schedule at step t (1-based), gradients
 * per grad_mode, shared ADAM state (m, v float[n_params]; beta powers kept as doubles per net array). */
double mzo_cos_schedule(int t);
/* 1 per blob entry that is a Flux parameter, 0 for the BatchNorm running statistics of the ResNet networks (length mzo_num_params(cfg, 3)) */
void mzo_trainable_mask(const mzo_config *cfg, unsigned char *mask);
void mzo_learn_step(const mzo_config *cfg, float *blob, float *adam_m, float *adam_v, int t, int grad_mode, int B,
                    const float *obs_batch, const float *action_batch, const float *value_batch,
                    const float *reward_batch, const float *policy_batch, const float *gscale, float *losses);

/* ---- prioritised replay (conf.PER = true), repaired specification: see the PER section of mz_oracle.c ---- */
uint32_t mzo_per_quantise(float p);
void mzo_per_priorities(const mzo_config *cfg, int T, const float *rewards, const int32_t *to_play, const float *root_values,
                        uint32_t *q_pos /* [T] */, uint32_t *q_game);
void mzo_get_batch_per(const mzo_config *cfg, int n_games, int64_t first_key, const int32_t *T, const float *obs, const int32_t *actions,
                       const float *rewards, const int32_t *to_play, const float *child_visits, const float *root_values,
                       const uint32_t *q_pos /* [n][Tmax] */, const uint32_t *q_game /* [n] */, uint64_t step, int32_t *index_batch,
                       float *obs_batch, float *action_batch, float *value_batch, float *reward_batch, float *policy_batch, float *gscale,
                       float *weights /* [B] */);
void mzo_per_update(const mzo_config *cfg, int B, const int32_t *index_batch, const float *pred_values, const float *target_values,
                    int n_games, int64_t first_key, const int32_t *T, uint32_t *q_pos, uint32_t *q_game);
void mzo_learn_forward_w(const mzo_config *cfg, const float *blob, int B, const float *obs_batch, const float *action_batch,
                         const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale,
                         const float *weights, float *pred_values, float *pred_rewards, float *pred_policies, float *losses);
double mzo_learn_gradients_w(const mzo_config *cfg, const float *blob, int B, const float *obs_batch, const float *action_batch,
                             const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale,
                             const float *weights, int fwd64, int perturb_index, double perturb_delta, double *grad);

/* grad_mode = MZO_GRAD_BPTT: d(loss)/d(theta) through the unroll (Float64 backward; see mz_oracle.c).  Returns the
 * Float64 data loss; grad (may be NULL) = d(data loss)/d(theta) + 2*theta in blob order.  fwd64 = 0 linearises around
 * the Float32-contract forward, 1 around an all-Float64 forward; perturb_index >= 0 adds perturb_delta to that
 * parameter first (finite-difference checks). */
double mzo_learn_gradients(const mzo_config *cfg, const float *blob, int B, const float *obs_batch, const float *action_batch,
                           const float *value_batch, const float *reward_batch, const float *policy_batch, const float *gscale,
                           int fwd64, int perturb_index, double perturb_delta, double *grad);
void mzo_adam_apply(float *blob, float *adam_m, float *adam_v, const float *grad, int n, int t);

#ifdef __cplusplus
}
#endif
#endif
