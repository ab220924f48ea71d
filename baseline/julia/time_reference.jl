# time_reference.jl -- times the UNMODIFIED reference (deveshjawla/MuZero.jl) on its own CPU path, for anyone with Julia 1.6 and the
# reference's Manifest.toml instantiated (BASELINE.md section 3, SURVEY.md 8d).  Not runnable in the build image (no Julia): the numbers
# bench.py reports beside the GPU come from the C restatement under oracle/.
#
#   julia --project=/path/to/MuZero.jl -t auto baseline/julia/time_reference.jl /path/to/MuZero.jl [num_iters=50] [games=64] [learner_steps=50]
#
# Prints one JSON line per metric in the shape of bench.py's `cpu_baseline`:
#   mcts_simulations_per_sec  -- play_game (src/SelfPlay.jl:330) on `games` games, one per task over all threads, num_iters simulations/move
#   learner_samples_per_sec   -- get_batch + unroll + loss + gradients + ADAM per step at conf.batch_size (src/Learning.jl:327-397)
using Printf

root = length(ARGS) ≥ 1 ? ARGS[1] : error("usage: time_reference.jl <path to MuZero.jl> [num_iters] [games] [learner_steps]")
S = length(ARGS) ≥ 2 ? parse(Int, ARGS[2]) : 50
G = length(ARGS) ≥ 3 ? parse(Int, ARGS[3]) : 64
L = length(ARGS) ≥ 4 ? parse(Int, ARGS[4]) : 50

include(joinpath(root, "src", "Constructors.jl"))
include(joinpath(root, "src", "RemoteBufferChannel.jl"))
include(joinpath(root, "src", "SelfPlay.jl"))
include(joinpath(root, "src", "ReplayBuffer.jl"))
include(joinpath(root, "src", "Learning.jl"))
include(joinpath(root, "games", "tictactoe", "game.jl"))
# games/tictactoe/params.jl binds `const conf`, `const hyper`; num_iters is a Config field, so rebuild conf with the benchmark's value
const conf = Config(observation_shape=(3, 3, 3), action_space=collect(1:9), players=collect(1:2), stacked_observations=1, max_moves=9,
                    num_unroll_steps=5, td_steps=5, PER=false, training_steps=10000, batch_size=32, num_iters=S)
const hyper = FeedForwardHP(width_hidden=64, depth_representation=3, depth_prediction=3, depth_dynamics=3, depth_policy=1, depth_value=1,
                            depth_reward=1, depth_state_head=3, use_batch_norm=false, batch_norm_momentum=0.6f0, hidden_state_size=27,
                            reward_activation=tanh)

NNs = (representation=init_representation(hyper), prediction=init_prediction(hyper), dynamics=init_dynamics(hyper))

# ---- self-play: the reference's hot loop, one game per task (the reference itself runs ONE self-play actor, main.jl:30) ----
play_game(TicTacToe(), 1.0f0, false, "self", 1, NNs)          # compile
function timed_games(n, nthreads)
    moves = Threads.Atomic{Int}(0)
    t = @elapsed begin
        if nthreads == 1
            for _ in 1:n
                h = play_game(TicTacToe(), 1.0f0, false, "self", 1, NNs); Threads.atomic_add!(moves, length(h.action_history))
            end
        else
            Threads.@threads for _ in 1:n
                h = play_game(TicTacToe(), 1.0f0, false, "self", 1, deepcopy(NNs)); Threads.atomic_add!(moves, length(h.action_history))
            end
        end
    end
    return moves[] * S / t, moves[] * S, t
end
for nt in unique((1, Threads.nthreads()))
    rate, sims, t = timed_games(nt == 1 ? max(4, G ÷ 8) : G, nt)
    @printf("{\\"metric\\": \\"mcts_simulations_per_sec\\", \\"value\\": %.1f, \\"unit\\": \\"simulations/s\\", \\"cores\\": %d, \\"kind\\": \\"reference\\", \\"sample\\": \\"%d simulations in %.2f s, num_iters=%d\\"}\\n",
            rate, nt, sims, t, S)
end

# ---- learner: the body of learning!'s loop (src/Learning.jl:329-397) on a buffer of self-played games ----
buffer = Dict{Int,GameHistory}()
for k in 1:64
    buffer[k] = play_game(TicTacToe(), 1.0f0, false, "self", 1, NNs)
end
optimiser = Flux.ADAMW(); schedule = Stateful(Cos(λ0=1e-4, λ1=1e-1, period=10))   # src/Learning.jl:318-319
representation, prediction, dynamics = deepcopy(NNs.representation), deepcopy(NNs.prediction), deepcopy(NNs.dynamics)
function one_step()   # the body of the `while` loop of learning! (src/Learning.jl:329-397), minus the channels
    index_batch, batch = get_batch(buffer)                                                  # :331
    observation_batch, action_batch, target_values, target_rewards, target_policies, weight_batch, gradient_scale_batch = batch
    gradient_scale_batch = permutedims(gradient_scale_batch)                                # :334
    hidden_state = representation(observation_batch)                                        # :347
    ndims(hidden_state) == 2 && (hidden_state = reshape(hidden_state, (conf.observation_shape..., conf.batch_size)))
    predicted_values, predicted_policies = prediction(hidden_state)                         # :351
    predicted_rewards = zeros((1, conf.batch_size))
    predicted_policies = Flux.unsqueeze(predicted_policies, 2)
    for i = 1:conf.num_unroll_steps                                                         # :355-370
        value, policy_logits = prediction(hidden_state)
        policy_logits = Flux.unsqueeze(policy_logits, 2)
        hidden_state, reward = dynamics(make_dynamics_input(hidden_state, action_batch[i, :], conf))
        ndims(hidden_state) == 2 && (hidden_state = reshape(hidden_state, (conf.observation_shape..., conf.batch_size)))
        predicted_values = vcat(predicted_values, value)
        predicted_rewards = vcat(predicted_rewards, reward)
        predicted_policies = cat(predicted_policies, policy_logits, dims=2)
    end
    targets = (target_values, target_rewards, target_policies)
    predictions = (predicted_values, predicted_rewards, predicted_policies)
    optimiser[1].eta = next!(schedule)                                                      # :382
    for net in (representation, prediction, dynamics)                                       # :385-397
        ps = Flux.params(net)
        l, g = loss_grad(ps) do
            loss(ps, predictions, targets, weight_batch, gradient_scale_batch)
        end
        Flux.update!(optimiser, ps, g)
    end
end
one_step()                                                        # compile (Zygote)
t = @elapsed for _ in 1:L
    one_step()
end
@printf("{\\"metric\\": \\"learner_samples_per_sec\\", \\"value\\": %.1f, \\"unit\\": \\"samples/s\\", \\"cores\\": 1, \\"kind\\": \\"reference\\", \\"sample\\": \\"%d steps of batch %d in %.2f s\\"}\\n",
        conf.batch_size * L / t, L, conf.batch_size, t)
