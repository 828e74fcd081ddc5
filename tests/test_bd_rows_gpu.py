"""Row construction of prob_nav_actions (bayesian_delegator.py:618-689) against the reference: for the
64 reference bayes_update calls of tests/golden/bd_rows.npz (oracle/gen_golden_plan.py records, per
likelihood row, the planner's valid-action list after the partner filter of :677-679, the index of the
taken action, the planning level and obs_tm1), (A) the offered actions of gc_subtask_q / gc_joint_q at
obs_tm1 must be the reference's list, in order, and (B) gc_bd_likelihood_rows must produce the same n_valid
and act_idx (None rows included: [p, (1-p)/k, ...] with k = the observer's own move count, :618-641)."""
import os

import numpy as np
import pytest
import torch

import gym_cooking_b200 as gcb
from gym_cooking_b200 import planning

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rows(golden_dir):
    return np.load(os.path.join(golden_dir, "bd_rows.npz"))


def _job_inputs(g, r):
    """KitchenBatch at obs_tm1 carrying the reference's own subtask list, and the job's likelihood rows"""
    n_agents, ns = int(g["n_agents"][r]), int(g["n_subtasks"][r])
    kb = gcb.KitchenBatch(str(g["levels"][g["level"][r]]), n_agents, 1, 100)
    kb.set_subtask_masks([tuple(int(v) for v in m) for m in g["subtasks"][r, :ns]])
    kb.state.copy_(torch.from_numpy(g["state"][r:r + 1].view(np.int32)).to(kb.device))
    subtasks = [tuple(int(v) for v in m) for m in g["subtasks"][r, :ns]]
    out = []
    for p in range(g["row_kind"].shape[1]):
        nv = int(g["row_n_valid"][r, p])
        if nv == 0:
            continue
        kind = int(g["row_kind"][r, p])
        row = dict(kind=kind, i=int(g["row_i"][r, p]), j=int(g["row_j"][r, p]), n_valid=nv,
                   act_idx=int(g["row_act_idx"][r, p]), valid=[int(v) for v in g["row_valid"][r, p, :nv]],
                   level1=bool(g["row_level1"][r, p]))
        if kind:
            row["subtask"] = subtasks.index(tuple(int(v) for v in g["row_masks"][r, p]))
        out.append(row)
    return kb, out


def test_offered_actions_equal_the_reference_lists(rows):
    g = rows
    compared = joint = 0
    for r in range(g["state"].shape[0]):
        kb, job = _job_inputs(g, r)
        observer, ex = int(g["observer"][r]), [int(a) for a in g["executed"][r]]
        plan = [row for row in job if row["kind"]]
        if not plan:
            continue
        pairs = [(row["subtask"], row["i"], None if row["j"] == 255 else row["j"], row["level1"]) for row in plan]
        _, q, _ = planning.subtask_q(kb, pairs)
        q = q[0].cpu().numpy()
        for p, row in enumerate(plan):
            if row["kind"] == 1:
                offered = [a for a in range(5) if not np.isnan(q[p, a])]
                taken = ex[row["i"]]
            else:
                offered = [a for a in range(25) if not np.isnan(q[p, a])]
                taken = 5 * ex[row["i"]] + ex[row["j"]]
                if observer in (row["i"], row["j"]):  # bd:677-679: joint actions that match the partner's move
                    if observer == row["i"]:
                        offered = [a for a in offered if a % 5 == ex[row["j"]]]
                    else:
                        offered = [a for a in offered if a // 5 == ex[row["i"]]]
                joint += 1
            assert offered == row["valid"], (r, p, row)
            assert offered.index(taken) == row["act_idx"], (r, p, row)
            compared += 1
    assert compared > 150 and joint > 40


def test_likelihood_row_kernel_equals_the_reference_rows(rows):
    """gc_bd_likelihood_rows_f64 on the two-agent jobs (the kernel's envelope: five entries per row)"""
    g = rows
    compared = none_rows = 0
    for r in range(g["state"].shape[0]):
        if int(g["n_agents"][r]) != 2:
            continue
        kb, job = _job_inputs(g, r)
        observer = int(g["observer"][r])
        # pair table: the planning rows, plus the observer alone at level 1 = the real env with everybody in
        # it, whose offered moves are get_single_actions(obs_tm1, observer) (bd:621-623)
        plan = [row for row in job if row["kind"]]
        pairs = [(row["subtask"], row["i"], None if row["j"] == 255 else row["j"], row["level1"]) for row in plan]
        pairs.append((0, observer, None, True))
        _, q, _ = planning.subtask_q(kb, pairs)
        n_moves = (~torch.isnan(q[:, len(pairs) - 1, :4])).sum(-1).to(torch.uint8)
        row_pair, k = [], 0
        for row in job:
            row_pair.append(k if row["kind"] else 0)
            k += 1 if row["kind"] else 0
        executed = torch.from_numpy(g["executed"][r:r + 1, :2].copy()).to(kb.device)
        qd, nv, ai = planning.bd_likelihood_rows(
            q.contiguous(), torch.zeros(1, dtype=torch.int64, device=kb.device), row_pair, [row["kind"] for row in job],
            [row["i"] for row in job], [0 if row["j"] == 255 else row["j"] for row in job], executed, n_moves, observer, 0.5)
        for p, row in enumerate(job):
            assert int(nv[0, p]) == row["n_valid"], (r, p, row, nv[0].tolist())
            assert int(ai[0, p]) == row["act_idx"], (r, p, row, ai[0].tolist())
            if row["kind"]:
                assert float(qd[0, p, row["act_idx"]]) == 0.0  # Q(taken) - Q(taken)
            else:  # [p_none, (1 - p_none) / k, ...] (bd:624-626)
                assert float(qd[0, p, 0]) == 0.5 and abs(float(qd[0, p, 1]) - 0.5 / (row["n_valid"] - 1)) < 1e-15
            none_rows += row["kind"] == 0
            compared += 1
    assert compared > 100 and none_rows > 5
