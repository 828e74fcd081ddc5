"""Batched planner entry points (paths B and C) over CUDA tensors - thin marshalling of
gc_bd_posterior_*, gc_lower_bound and gc_subtask_q (include/gymcook.h)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

MAX_PAIRS = 128  # GC_MAX_PAIRS
PERIMETER_NOT_DOABLE = 28.0  # a 7x7 kitchen: lower bound >= world.perimeter means "not doable" (bd:156)


def bd_posterior(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta):
    """One Bayesian-Delegation posterior update per row, in place on `probs` ([n][H] float32/64).
    Restates BayesianDelegator.bayes_update (bd:1045-1072) on dumped inputs - see gymcook.h."""
    lib = _lib.load()
    n, H = probs.shape
    P, A = qdiff.shape[1], qdiff.shape[2]
    E = hyp_pair.shape[2]
    if qdiff.dtype != probs.dtype:
        raise _lib.GcError("probs and qdiff must share a dtype")
    fn = {torch.float32: lib.gc_bd_posterior_f32, torch.float64: lib.gc_bd_posterior_f64}[probs.dtype]
    u8 = torch.uint8
    with torch.cuda.device(probs.device):
        _lib.check(fn(_lib.ptr(probs), _lib.ptr(alive, u8), _lib.ptr(hyp_pair, u8), _lib.ptr(pair_w, u8),
                      _lib.ptr(qdiff), _lib.ptr(n_valid, u8), _lib.ptr(act_idx, u8), float(beta), n, H, P, A, E,
                      _lib.stream_ptr(probs.device)))
    return probs


def bd_likelihood_rows(q_table, q_row, row_pair, row_kind, row_agent, row_agent2, executed, n_moves, observer,
                       none_action_prob, q_cap=100.0, dtype=torch.float64):
    """prob_nav_actions' softmax inputs (bd:461-689) for every (env, likelihood row): returns
    (qdiff [n][P][A], n_valid [n][P], act_idx [n][P]) ready for bd_posterior - see gymcook.h.  A = 5, or 25
    when the table holds joint rows the observer is not part of (kind 3: three or more agents)."""
    lib = _lib.load()
    n, P = q_row.shape[0], len(row_pair)
    dev = q_table.device
    rp = np.ascontiguousarray(np.asarray(row_pair, dtype=np.int32))
    rk, ra, ra2 = (np.ascontiguousarray(np.asarray(a, dtype=np.uint8)) for a in (row_kind, row_agent, row_agent2))
    fn = {torch.float32: lib.gc_bd_likelihood_rows_f32, torch.float64: lib.gc_bd_likelihood_rows_f64}[dtype]
    with torch.cuda.device(dev):
        A = 25 if bool((rk == 3).any()) else 5
        qdiff = torch.empty((n, P, A), dtype=dtype, device=dev)
        n_valid = torch.empty((n, P), dtype=torch.uint8, device=dev)
        act_idx = torch.empty((n, P), dtype=torch.uint8, device=dev)
        _lib.check(fn(_lib.ptr(q_table, torch.float32), _lib.ptr(q_row, torch.int64), q_table.shape[1],
                      rp.ctypes.data_as(C.c_void_p), rk.ctypes.data_as(C.c_void_p), ra.ctypes.data_as(C.c_void_p),
                      ra2.ctypes.data_as(C.c_void_p), _lib.ptr(executed, torch.uint8), _lib.ptr(n_moves, torch.uint8),
                      int(observer), float(none_action_prob), float(q_cap), _lib.ptr(qdiff), _lib.ptr(n_valid),
                      _lib.ptr(act_idx), n, P, executed.shape[1], A, _lib.stream_ptr(dev)))
    return qdiff, n_valid, act_idx


def bd_update_lists(probs, alive, rid, hyp_pair, pair_w, q_table, q_row, row_pair, row_kind, row_agent, row_agent2,
                    executed, n_moves, observer, none_action_prob, beta, q_cap=100.0):
    """bayes_update (bd:1026-1072) in place on per-env hypothesis LISTS: probs/alive/rid [n][W], rid naming rows of
    the shared table hyp_pair [H][E]; the likelihood rows are built inside the kernel - see gymcook.h."""
    lib = _lib.load()
    n, W = probs.shape
    H, E = hyp_pair.shape
    P = len(row_pair)
    dev = probs.device
    rp = np.ascontiguousarray(np.asarray(row_pair, dtype=np.int32))
    rk, ra, ra2, pw = (np.ascontiguousarray(np.asarray(a, dtype=np.uint8)) for a in (row_kind, row_agent, row_agent2, pair_w))
    fn = {torch.float32: lib.gc_bd_update_lists_f32, torch.float64: lib.gc_bd_update_lists_f64}[probs.dtype]
    u8 = torch.uint8
    if alive.shape != probs.shape or rid.shape != probs.shape or len(pw) != P:
        raise _lib.GcError("bd_update_lists: probs, alive and rid must share [n][W]; pair_w has one weight per row")
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.ptr(probs), _lib.ptr(alive, u8), _lib.ptr(rid, torch.int64), W, _lib.ptr(hyp_pair, u8), H, E,
                      pw.ctypes.data_as(C.c_void_p), _lib.ptr(q_table, torch.float32), _lib.ptr(q_row, torch.int64),
                      q_table.shape[1], rp.ctypes.data_as(C.c_void_p), rk.ctypes.data_as(C.c_void_p),
                      ra.ctypes.data_as(C.c_void_p), ra2.ctypes.data_as(C.c_void_p), _lib.ptr(executed, u8),
                      _lib.ptr(n_moves, u8), int(observer), float(none_action_prob), float(q_cap), float(beta), n, P,
                      executed.shape[1], _lib.stream_ptr(dev)))
    return probs


def offered_actions(batch):
    """nav_utils.get_single_actions (navigation_planner/utils.py:55-90) for every env and agent of `batch`:
    uint8[N][n_agents], bit a = move a is offered (staying always is)."""
    lib = _lib.load()
    with torch.cuda.device(batch.device):
        out = torch.empty((batch.num_envs, batch.num_agents), dtype=torch.uint8, device=batch.device)
        _lib.check(lib.gc_offered_actions(batch._lv(), batch.n_levels, _lib.ptr(batch.level_id), _lib.ptr(batch.state),
                                          _lib.ptr(out), batch.num_envs, batch.num_agents, batch._stream()))
    return out


def subtasks_completed(batch, before, subtask):
    """RealAgent.def_subtask_completion (utils/agent.py:286-368): uint8[N][n_agents], 1 where batch.state holds more
    goal objects of subtask[env][agent] (uint8, >= the subtask count = none) than `before` (int32[N][4])."""
    lib = _lib.load()
    with torch.cuda.device(batch.device):
        out = torch.empty((batch.num_envs, batch.num_agents), dtype=torch.uint8, device=batch.device)
        _lib.check(lib.gc_subtasks_completed(batch._lv(), batch.n_levels, _lib.ptr(batch.level_id), _lib.ptr(before),
                                             _lib.ptr(batch.state), _lib.ptr(subtask, torch.uint8), _lib.ptr(out),
                                             batch.num_envs, batch.num_agents, batch._stream()))
    return out


def _pairs_array(pairs):
    """[(subtask index, agent i, agent j or None[, level1])] -> host uint8[n_pairs][3]; a truthy
    fourth element selects the level-1 planning world (bit 7 of the subtask byte, gymcook.h)."""
    rows = []
    for pr in pairs:
        s, i, j = pr[0], pr[1], pr[2]
        level1 = len(pr) > 3 and pr[3]
        rows.append([s | (0x80 if level1 else 0), i, 255 if j is None else j])
    return np.ascontiguousarray(np.array(rows, dtype=np.uint8).reshape(-1, 3))


def lower_bound(batch, pairs, out=None):
    """env.get_lower_bound_for_subtask_given_objs (env:594-664) for every (env, pair) of a
    KitchenBatch -> float32[N][n_pairs].  The levels of `batch` must carry their subtasks
    (KitchenBatch.set_subtasks)."""
    lib = _lib.load()
    if len(pairs) > MAX_PAIRS:  # GC_MAX_PAIRS per call: larger lists go in chunks
        parts = [lower_bound(batch, pairs[k:k + MAX_PAIRS]) for k in range(0, len(pairs), MAX_PAIRS)]
        res = torch.cat(parts, dim=1)
        if out is not None:
            out.copy_(res)
            return out
        return res
    arr = _pairs_array(pairs)
    with torch.cuda.device(batch.device):
        if out is None:
            out = torch.empty((batch.num_envs, len(pairs)), dtype=torch.float32, device=batch.device)
        _lib.check(lib.gc_lower_bound(batch._lv(), batch.n_levels, _lib.ptr(batch.level_id), _lib.ptr(batch.state),
                                      arr.ctypes.data_as(C.c_void_p), len(pairs), _lib.ptr(out), batch.num_envs,
                                      batch.num_agents, batch._stream()))
    return out


def subtask_q(batch, pairs, want_q=True):
    """Exact level-0 V* / Q(start, .) for every (env, pair): returns (v[N][P], q[N][P][25] or None,
    status uint8[N][P]).  Single-agent pairs: gc_subtask_q (interaction-level IDA*); joint pairs:
    gc_joint_q (budgeted uniform-cost search, scratch arena allocated here as a torch tensor)."""
    lib = _lib.load()
    if len(pairs) > MAX_PAIRS:
        parts = [subtask_q(batch, pairs[k:k + MAX_PAIRS], want_q) for k in range(0, len(pairs), MAX_PAIRS)]
        return (torch.cat([p[0] for p in parts], dim=1),
                torch.cat([p[1] for p in parts], dim=1) if want_q else None,
                torch.cat([p[2] for p in parts], dim=1))
    arr = _pairs_array(pairs)
    has_joint = bool((arr[:, 2] != 255).any())
    with torch.cuda.device(batch.device):
        v = torch.empty((batch.num_envs, len(pairs)), dtype=torch.float32, device=batch.device)
        q = (torch.empty((batch.num_envs, len(pairs), 25), dtype=torch.float32, device=batch.device)
             if (want_q or has_joint) else None)
        status = torch.empty((batch.num_envs, len(pairs)), dtype=torch.uint8, device=batch.device)
        _lib.check(lib.gc_subtask_q(batch._lv(), batch.n_levels, _lib.ptr(batch.level_id), _lib.ptr(batch.state),
                                    arr.ctypes.data_as(C.c_void_p), len(pairs), _lib.ptr(v), _lib.ptr(q),
                                    _lib.ptr(status), batch.num_envs, batch.num_agents, batch._stream()))
        if has_joint:
            need = lib.gc_joint_q_scratch_bytes(batch.num_envs, len(pairs), None)
            owner = getattr(batch, "_batch", batch)  # a state view keeps the arena on its KitchenBatch
            scratch = getattr(owner, "_joint_scratch", None)
            if scratch is None or scratch.numel() < need:
                scratch = owner._joint_scratch = torch.empty(need, dtype=torch.uint8, device=batch.device)
            _lib.check(lib.gc_joint_q(batch._lv(), batch.n_levels, _lib.ptr(batch.level_id), _lib.ptr(batch.state),
                                      arr.ctypes.data_as(C.c_void_p), len(pairs), _lib.ptr(v), _lib.ptr(q),
                                      _lib.ptr(status), _lib.ptr(scratch), scratch.numel(), batch.num_envs,
                                      batch.num_agents, batch._stream()))
    return v, (q if want_q else None), status


_T_MASK = ~0xFF000000  # clears t and done of word 0 inside the low int64 of a packed state


class _StateView:
    """What the planner entry points need from a KitchenBatch, over another state tensor of the same level(s)."""

    def __init__(self, batch, state, level_id=None):
        self._batch = batch
        self.num_agents, self.n_levels, self.device = batch.num_agents, batch.n_levels, batch.device
        self.state, self.num_envs, self.level_id = state, state.shape[0], level_id

    def _lv(self):
        return self._batch._lv()

    def _stream(self):
        return self._batch._stream()


def subtask_q_unique(batch, pairs, want_q=True, chunk=1 << 16):
    """subtask_q with every distinct PLANNING state of the batch solved once: the value of a (subtask, agents)
    pair does not depend on t or the done bit, and in a large batch most envs share their planning state
    with others (the batched form of the reference memoising v_l / v_u by state repr, e2e:216-352).
    Returns (v[N][P], q[N][P][25] or None, status[N][P], n_unique).  Single-level batches."""
    if batch.n_levels != 1:
        raise _lib.GcError("subtask_q_unique needs a single-level batch")
    key = batch.state.contiguous().view(torch.int64).clone()
    key[:, 0] &= _T_MASK
    uk, inv = torch.unique(key, dim=0, return_inverse=True)
    U = uk.shape[0]
    ustate = uk.contiguous().view(torch.int32)
    vs, qs, ss = [], [], []
    for lo in range(0, U, chunk):  # bounded scratch arena of the joint solver
        v, q, st = subtask_q(_StateView(batch, ustate[lo:lo + chunk].contiguous()), pairs, want_q)
        vs.append(v)
        qs.append(q)
        ss.append(st)
    v, st = torch.cat(vs), torch.cat(ss)
    q = torch.cat(qs) if want_q else None
    return v[inv], (q[inv] if want_q else None), st[inv], U
