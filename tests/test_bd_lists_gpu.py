"""gc_bd_update_lists (the fused likelihood + posterior update over per-env hypothesis LISTS that three and four
agents use) against the two-kernel path gc_bd_likelihood_rows -> gc_bd_posterior, which the reference dumps pin
(tests/test_bd_gpu.py, tests/test_bd_rows_gpu.py): same inputs, the list expanded to the per-env table the
two-kernel path wants."""
import numpy as np
import pytest
import torch

from gym_cooking_b200 import planning

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_agents,observer,H,W,dtype", [(3, 0, 300, 40, torch.float64), (4, 2, 5000, 128, torch.float64),
                                                          (2, 1, 60, 7, torch.float64), (4, 3, 900, 33, torch.float32)])
def test_list_update_equals_rows_then_posterior(n_agents, observer, H, W, dtype):
    rng = np.random.RandomState(H + W)
    dev = torch.device("cuda:0")
    n, n_pairs, states = 3000, 40, 500
    # likelihood rows: a None row per agent, single rows, joint rows with and without the observer
    kinds, agents, agents2, pairs = [], [], [], []
    for a in range(n_agents):
        kinds.append(0), agents.append(a), agents2.append(a), pairs.append(0)
    for _ in range(30):
        a = rng.randint(n_agents)
        if n_agents > 1 and rng.rand() < 0.5:
            b = (a + 1 + rng.randint(n_agents - 1)) % n_agents
            a, b = min(a, b), max(a, b)
            kinds.append(2 if observer in (a, b) else 3)
            agents.append(a), agents2.append(b)
        else:
            kinds.append(1), agents.append(a), agents2.append(a)
        pairs.append(rng.randint(n_pairs))
    P = len(kinds)
    E = n_agents
    q = rng.rand(states, n_pairs, 25).astype(np.float32) * 20
    q[rng.rand(*q.shape) < 0.3] = np.nan
    q[rng.rand(*q.shape) < 0.05] = np.inf
    q_table = torch.from_numpy(q).to(dev)
    q_row = torch.from_numpy(rng.randint(states, size=n)).to(dev)
    executed = torch.from_numpy(rng.randint(5, size=(n, n_agents)).astype(np.uint8)).to(dev)
    n_moves = torch.from_numpy(rng.randint(5, size=n).astype(np.uint8)).to(dev)
    hyp = rng.randint(P, size=(H, E)).astype(np.uint8)
    hyp[rng.rand(H, E) < 0.3] = 255
    hyp_pair = torch.from_numpy(hyp).to(dev)
    pair_w = rng.randint(1, 3, size=P).astype(np.uint8)
    rid = torch.from_numpy(rng.randint(H + 1, size=(n, W))).to(dev)  # H = padding row
    alive = torch.from_numpy(rng.rand(n, W) < 0.7).to(dev) & (rid < H)
    alive[:5] = False  # envs without any hypothesis
    probs = torch.from_numpy(rng.rand(n, W)).to(dev).to(dtype)
    probs[5:10] = 0  # total == 0 -> uniform over the alive ones
    got = planning.bd_update_lists(probs.clone(), alive.view(torch.uint8), rid, hyp_pair, pair_w, q_table, q_row, pairs,
                                   kinds, agents, agents2, executed, n_moves, observer, 0.5, 1.3)
    qdiff, n_valid, act_idx = planning.bd_likelihood_rows(q_table, q_row, pairs, kinds, agents, agents2, executed, n_moves,
                                                          observer, 0.5, dtype=dtype)
    table = torch.cat([hyp_pair, torch.full((1, E), 255, dtype=torch.uint8, device=dev)])
    want = (probs * alive).contiguous()
    planning.bd_posterior(want, alive.to(torch.uint8).contiguous(), table[rid].contiguous(),
                          torch.from_numpy(pair_w).to(dev)[None].expand(n, P).contiguous(), qdiff, n_valid, act_idx, 1.3)
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    assert float((got - want).abs().max()) <= tol
    assert bool((got[~alive] == 0).all())
    s = got.sum(1)[alive.any(1)]
    assert float((s - 1).abs().max()) <= 1e-5
