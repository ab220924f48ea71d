"""Pure-Python restatement of run_mcts (src/SelfPlay.jl:62-96,157-217,230-285), written independently
of oracle/mz_oracle.c (dict-based nodes like the reference) and used only to cross-check the C oracle's
tree logic on small cases.  Network outputs come from the oracle's NN functions; float32 arithmetic is
emulated with numpy scalars."""
import math

import numpy as np

f32 = np.float32


class Node:
    def __init__(self, prior):
        self.visit_count = 0
        self.to_play = 1
        self.prior = f32(prior)
        self.value_sum = f32(0)
        self.children = None
        self.hidden_state = None
        self.reward = f32(0)


def node_value(n):
    return f32(0) if n.visit_count == 0 else f32(n.value_sum / f32(n.visit_count))


def softmax32(x, expf):
    m = max(x)
    e = [f32(expf(float(f32(v - m)))) for v in x]
    s = f32(0)
    for v in e:
        s = f32(s + v)
    return [f32(v / s) for v in e]


def run_mcts(cfg, nn, stacked, legal, to_play, order, tie_pick):
    """nn: object with representation/prediction/dynamics/expf; legal: ascending list of 1-based actions;
    order: Dict iteration order; tie_pick(sim, depth, n_tied) -> index."""
    disc = f32(cfg.discount)

    def expand(node, tp, reward, policy, hidden):
        pv = softmax32([policy[a - 1] for a in legal], nn.expf)
        node.children = {a: Node(pv[i]) for i, a in enumerate(legal)}
        node.to_play, node.reward, node.hidden_state = tp, f32(reward), hidden

    def iter_children(node):
        return [(a, node.children[a]) for a in order if a in node.children]

    mm = [f32(np.inf), f32(-np.inf)]

    def ucb(parent, child):
        pb_c = math.log2((parent.visit_count + cfg.pb_c_base + 1) / cfg.pb_c_base) + float(f32(cfg.pb_c_init))
        pb_c *= math.sqrt(parent.visit_count) / (child.visit_count + 1)
        prior_score = pb_c * float(child.prior)
        if child.visit_count > 0:
            q = f32(child.reward + f32(disc * f32(-node_value(child))))
            if mm[1] > mm[0]:
                q = f32(f32(q - mm[0]) / f32(mm[1] - mm[0]))
            return f32(prior_score + float(q))
        return f32(prior_score)

    root = Node(0.0)
    h0 = nn.representation(stacked)
    v0, p0 = nn.prediction(h0)
    expand(root, to_play, 0.0, p0, h0.copy())
    for it in range(1, cfg.num_iters + 1):
        node, vtp, path, action, depth = root, to_play, [root], 0, 0
        while node.children is not None:
            depth += 1
            ch = iter_children(node)
            scores = [ucb(node, c) for _, c in ch]
            mx = max(scores)
            tied = [i for i, s in enumerate(scores) if s == mx]
            i = tied[tie_pick(it, depth, len(tied))] if len(tied) > 1 else tied[0]
            action, node = ch[i]
            path.append(node)
            vtp = vtp % cfg.num_players + 1
        parent = path[-2]
        value, policy = nn.prediction(parent.hidden_state)
        parent.hidden_state *= f32(2.0)  # make_state_action doubles in place (Q6)
        plane = f32(action / cfg.A)
        sa = np.concatenate([parent.hidden_state, np.full(cfg.W * cfg.H, plane, f32)])
        nh, reward = nn.dynamics(sa)
        expand(node, vtp, reward, policy, nh.copy())
        value = f32(value)
        for nd in reversed(path):  # backpropagate!, two players (Q8)
            if nd.to_play == vtp:
                nd.value_sum = f32(nd.value_sum + value)
            else:
                nd.value_sum = f32(nd.value_sum - value)
            nd.visit_count += 1
            v = f32(nd.reward + f32(disc * node_value(nd)))
            mm[0] = mm[0] if mm[0] < v else v
            mm[1] = mm[1] if mm[1] > v else v
            if nd.to_play == vtp:
                value = f32(-nd.reward)
            else:
                value = f32(nd.reward + f32(disc * value))
    counts = np.zeros(cfg.A, np.int32)
    for a, c in root.children.items():
        counts[a - 1] = c.visit_count
    return counts, node_value(root), root
