/* The C ABI from plain C (no Python, no torch): the loop a host-language binding (Julia ccall, cgo, JNI ...) would drive.
 *   self_play! -> learning! -> competitive_play!   (games/tictactoe/main.jl:15-45, src/SelfPlay.jl:384-435, src/Learning.jl:306)
 * build: gcc -O2 -I include -o examples/selfplay examples/selfplay.c -L muzero.jl_b200 -lmuzero_b200 -Wl,-rpath,'$ORIGIN/../muzero.jl_b200'
 * run:   examples/selfplay [games] [simulations per move] [training steps]                      (needs a CUDA device) */
#include <stdio.h>
#include <stdlib.h>
#include "muzero_b200.h"

#define CK(call) do { int rc_ = (call); if (rc_ != MZ_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mz_last_error(NULL)); return 1; } } while (0)

int main(int argc, char **argv) {
    const int64_t games = argc > 1 ? atoll(argv[1]) : 1024;
    const int sims_per_move = argc > 2 ? atoi(argv[2]) : 25, steps = argc > 3 ? atoi(argv[3]) : 20;
    mz_ctx *ctx = NULL;
    mz_config cfg;
    if (mz_abi_version() != MZ_ABI_VERSION) { fprintf(stderr, "header / library ABI mismatch\n"); return 1; }
    CK(mz_default_config(&cfg));                 /* params.jl:2-29 */
    cfg.num_iters = sims_per_move; cfg.num_slots = 1024; cfg.replay_buffer_size = 4096; cfg.batch_size = 256;
    CK(mz_create(&cfg, 0, &ctx));
    CK(mz_init_weights(ctx, cfg.seed));          /* init_representation / prediction / dynamics (hyper) */

    int64_t sims = 0, moves = 0, n_games = 0, first_key = 0, samples = 0;
    CK(mz_self_play(ctx, 0, games, 1.0f, &sims, &moves));                                   /* self_play! + save_game */
    CK(mz_replay_info(ctx, &n_games, &first_key, &samples));
    printf("self-play: %lld games, %lld moves, %lld simulations; buffer holds %lld games / %lld positions\n",
           (long long)games, (long long)moves, (long long)sims, (long long)n_games, (long long)samples);

    float losses[3];
    CK(mz_learn_steps(ctx, 1, steps, MZ_GRAD_BPTT, losses));                                /* learning! */
    printf("learner: %d steps, losses (representation, prediction, dynamics) = %g %g %g\n", steps, losses[0], losses[1], losses[2]);

    /* evaluation on a context of its own, with the trained weights */
    mz_ctx *arena = NULL;
    CK(mz_create(&cfg, 0, &arena));
    const int n = mz_num_params(&cfg, MZ_NET_ALL);
    float *blob = (float *)malloc(sizeof(float) * (size_t)n);
    CK(mz_get_weights(ctx, MZ_NET_ALL, blob, n));
    CK(mz_set_weights(arena, MZ_NET_ALL, blob, n));
    for (int opp = MZ_OPP_RANDOM; opp <= MZ_OPP_EXPERT; opp++) {
        int64_t w = 0, d = 0, l = 0, s = 0;
        CK(mz_arena(arena, 1000000, 500, opp, 1, 0.0f, &w, &d, &l, &s));                    /* competitive_play! */
        printf("arena vs %s (MuZero moves first): %lld wins, %lld draws, %lld losses\n", opp == MZ_OPP_RANDOM ? "random" : "expert",
               (long long)w, (long long)d, (long long)l);
        if (w + d + l != 500) { fprintf(stderr, "arena tallies do not add up\n"); return 1; }
    }
    free(blob);
    mz_destroy(arena);
    mz_destroy(ctx);
    printf("ok\n");
    return 0;
}
