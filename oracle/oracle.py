"""ctypes binding of the CPU oracle (oracle/gc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
`--impl reference` legs of bench.py - never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgcoracle.so")
_lib = None

MAX_AGENTS, MAX_OBJS, MAX_GOALS = 4, 6, 4


class Level(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("type", (C.c_int * 8) * 8),
        ("n_agent_starts", C.c_int),
        ("agent_x", C.c_int * MAX_AGENTS), ("agent_y", C.c_int * MAX_AGENTS),
        ("n_objs", C.c_int),
        ("obj_mask", C.c_int * MAX_OBJS), ("obj_x", C.c_int * MAX_OBJS), ("obj_y", C.c_int * MAX_OBJS),
        ("n_goals", C.c_int),
        ("goal_mask", C.c_int * MAX_GOALS),
        ("delivery_x", C.c_int), ("delivery_y", C.c_int),
        ("max_timesteps", C.c_int),
    ]


class Agent(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("hold", C.c_int)]


class Obj(C.Structure):
    _fields_ = [("alive", C.c_int), ("mask", C.c_int), ("x", C.c_int), ("y", C.c_int), ("held_by", C.c_int)]


class Env(C.Structure):
    _fields_ = [("t", C.c_int), ("done", C.c_int), ("successful", C.c_int),
                ("n_agents", C.c_int), ("n_objs", C.c_int),
                ("ag", Agent * MAX_AGENTS), ("ob", Obj * MAX_OBJS)]


class Subtask(C.Structure):
    _fields_ = [("kind", C.c_int), ("a", C.c_int), ("b", C.c_int), ("goal", C.c_int)]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("gc_oracle.c", "gc_oracle_nav.c", "gc_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    u32p, u8p, u64p, f64p = (C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64),
                             C.POINTER(C.c_double))
    L.gco_level_parse.argtypes = [C.c_char_p, C.c_int, C.POINTER(Level)]
    L.gco_level_parse.restype = C.c_int
    L.gco_reset.argtypes = [C.POINTER(Level), C.c_int, C.POINTER(Env)]
    L.gco_reset.restype = None
    L.gco_step.argtypes = [C.POINTER(Level), C.POINTER(Env), u8p, u8p]
    L.gco_step.restype = C.c_int
    L.gco_pack.argtypes = [C.POINTER(Env), u32p]
    L.gco_unpack.argtypes = [u32p, C.c_int, C.POINTER(Env)]
    L.gco_hash_packed.argtypes = [u32p, C.c_int]
    L.gco_hash_packed.restype = C.c_uint64
    L.gco_canonical_keys.argtypes = [u32p, C.POINTER(C.c_uint16)]
    L.gco_canonical_keys.restype = C.c_int
    L.gco_philox_actions.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, u8p]
    L.gco_step_batch.argtypes = [C.POINTER(Level), u32p, u8p, u8p, u32p, C.c_int64, C.c_int, C.c_int]
    L.gco_rollout_batch.argtypes = [C.POINTER(Level), u32p, u8p, u64p, u32p, C.c_int64, C.c_int,
                                    C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_int]
    L.gco_bd_posterior.argtypes = [f64p, u8p, u8p, u8p, f64p, u8p, u8p, C.c_double, C.c_int64,
                                   C.c_int, C.c_int, C.c_int, C.c_int]
    L.gco_replay.argtypes = [C.POINTER(Level), C.c_int, u8p, C.c_int, u32p, u8p, u8p, u8p]
    L.gco_replay.restype = None
    if hasattr(L, "gco_lower_bound"):
        L.gco_lower_bound.argtypes = [C.POINTER(Level), C.POINTER(Env), C.POINTER(Subtask), C.c_int, C.c_int]
        L.gco_lower_bound.restype = C.c_double
    if hasattr(L, "gco_subtask_q"):
        L.gco_subtask_q.argtypes = [C.POINTER(Level), C.POINTER(Env), C.POINTER(Subtask), C.c_int, C.c_int,
                                    f64p, f64p, C.c_int]
        L.gco_subtask_q.restype = C.c_int
        L.gco_set_planner_level.argtypes = [C.c_int]
    _lib = L
    return L


def _p(arr, ctype):
    return None if arr is None else arr.ctypes.data_as(C.POINTER(ctype))


def parse_level(text, max_timesteps=100):
    lv = Level()
    rc = lib().gco_level_parse(text.encode(), max_timesteps, C.byref(lv))
    if rc != 0:
        raise ValueError("oracle level parse failed: %d" % rc)
    return lv


def reset_state(lv, n_agents, n=1):
    """uint32[n][4] packed initial states."""
    e = Env()
    lib().gco_reset(C.byref(lv), n_agents, C.byref(e))
    w = (C.c_uint32 * 4)()
    lib().gco_pack(C.byref(e), w)
    return np.tile(np.array(list(w), dtype=np.uint32), (n, 1))


def step_batch(lv, state, actions, n_agents, n_threads=1, want_collisions=True):
    """One step for all envs, in place on `state` (uint32[n][4]).  Returns (reward_done, collisions)."""
    n = state.shape[0]
    assert state.dtype == np.uint32 and state.flags.c_contiguous
    actions = np.ascontiguousarray(actions, dtype=np.uint8)
    rd = np.zeros(n, dtype=np.uint8)
    coll = np.zeros(n, dtype=np.uint32) if want_collisions else None
    lib().gco_step_batch(C.byref(lv), _p(state, C.c_uint32), _p(actions, C.c_uint8), _p(rd, C.c_uint8),
                         _p(coll, C.c_uint32), n, n_agents, n_threads)
    return rd, coll


def rollout_batch(lv, state, n_agents, n_steps, t0=0, env0=0, seed=1234, n_threads=1, want_hash=False):
    n = state.shape[0]
    rd = np.zeros(n, dtype=np.uint8)
    coll = np.zeros(n, dtype=np.uint32)
    ht = np.zeros((n_steps, n), dtype=np.uint64) if want_hash else None
    lib().gco_rollout_batch(C.byref(lv), _p(state, C.c_uint32), _p(rd, C.c_uint8), _p(ht, C.c_uint64),
                            _p(coll, C.c_uint32), n, n_agents, n_steps, t0, env0, seed, n_threads)
    return rd, coll, ht


def fill_actions(n, n_agents, n_steps, t0=0, env0=0, seed=1234):
    """uint8[n_steps][n][n_agents]: the philox action stream of cfg-2, materialised on the CPU."""
    out = np.empty((n_steps, n, n_agents), dtype=np.uint8)
    L = lib()
    L.gco_fill_actions.argtypes = [C.POINTER(C.c_uint8), C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_uint64]
    L.gco_fill_actions.restype = None
    L.gco_fill_actions(_p(out, C.c_uint8), n, n_agents, n_steps, t0, env0, seed)
    return out


def replay(lv, n_agents, actions):
    """actions uint8[T][4] -> (states uint32[T+1][4], reward_done[T+1], ncoll[T+1], executed[T+1][4])."""
    actions = np.ascontiguousarray(actions, dtype=np.uint8)
    T = actions.shape[0]
    states = np.zeros((T + 1, 4), dtype=np.uint32)
    rd = np.zeros(T + 1, dtype=np.uint8)
    nc = np.zeros(T + 1, dtype=np.uint8)
    ex = np.zeros((T + 1, 4), dtype=np.uint8)
    lib().gco_replay(C.byref(lv), n_agents, _p(actions, C.c_uint8), T, _p(states, C.c_uint32),
                     _p(rd, C.c_uint8), _p(nc, C.c_uint8), _p(ex, C.c_uint8))
    return states, rd, nc, ex


def slots_of(state):
    """Packed states uint32[n][4] (byte planes, include/gymcook.h) -> int64[n][6] object slots in the
    16-bit working form mask | cell << 7 | holder << 13 (holder 7 = dead, slot 0xE000)."""
    st = np.asarray(state, dtype=np.uint32).astype(np.int64).reshape(-1, 4)
    place = np.stack([(st[:, 1] >> (8 * k)) & 0xFF for k in range(4)] +
                     [(st[:, 3] >> (8 * k)) & 0xFF for k in range(2)], axis=1)
    mask = np.stack([(st[:, 2] >> (8 * k)) & 0xFF for k in range(4)] +
                    [(st[:, 3] >> (8 * (k + 2))) & 0xFF for k in range(2)], axis=1)
    return np.where(place >= 0x40, (place & 7) << 13, place << 7) | mask


def words_of(w0, slots):
    """Inverse of slots_of for one env: word 0 and six working slots -> the four packed words."""
    slots = list(slots) + [0xE000] * (MAX_OBJS - len(slots))
    place = [(0x40 | (s >> 13)) if (s >> 13) else (s >> 7) & 63 for s in slots]
    mask = [s & 0x7F for s in slots]
    w1 = sum(place[k] << (8 * k) for k in range(4))
    w2 = sum(mask[k] << (8 * k) for k in range(4))
    w3 = place[4] | place[5] << 8 | mask[4] << 16 | mask[5] << 24
    return [int(w0) & 0xFFFFFFFF, w1, w2, w3]


def v1_to_v2(state):
    """ABI-1 packed states (six 16-bit slots in words 1..3) -> the byte-plane form; used once, by
    oracle/convert_golden_v2.py, on the fixtures generated before the layout change."""
    st = np.asarray(state, dtype=np.uint32).reshape(-1, 4)
    out = np.empty_like(st)
    for i in range(st.shape[0]):
        w = [int(x) for x in st[i]]
        out[i] = words_of(w[0], [(w[1 + k // 2] >> (16 * (k % 2))) & 0xFFFF for k in range(MAX_OBJS)])
    return out


def decode_batch(state, n_agents):
    """Vectorised `decode`: state uint32[n][4] -> (t[n], done[n], agents[n][n_agents][2],
    keys uint16[n][6] sorted, 0x3FFF padded)."""
    state = np.asarray(state, dtype=np.uint32)
    w0 = state[:, 0].astype(np.int64)
    t = (w0 >> 24) & 127
    done = (w0 >> 31) & 1
    slots = slots_of(state)
    holder = slots >> 13
    mask = slots & 0x7F
    cells = np.stack([(w0 >> (6 * i)) & 63 for i in range(4)], axis=1)
    agents = np.zeros((state.shape[0], n_agents, 2), dtype=np.int64)
    for i in range(n_agents):
        agents[:, i, 0] = cells[:, i]
        agents[:, i, 1] = np.where(holder == i + 1, mask, 0).max(axis=1)
    held = (holder >= 1) & (holder <= 4)
    hcell = np.take_along_axis(cells, np.clip(holder - 1, 0, 3), axis=1)
    cell = np.where(held, hcell, (slots >> 7) & 63)
    keys = (mask << 7) | (cell << 1) | held.astype(np.int64)
    keys = np.where(holder == 7, 0x3FFF, keys)
    keys = np.sort(keys, axis=1).astype(np.uint16)
    return t, done, agents, keys


def hash_states(state, n_agents):
    out = np.empty(state.shape[0], dtype=np.uint64)
    L = lib()
    for i in range(state.shape[0]):
        out[i] = L.gco_hash_packed(_p(state[i], C.c_uint32), n_agents)
    return out


def canonical_keys(w):
    w = np.ascontiguousarray(w, dtype=np.uint32)
    keys = (C.c_uint16 * MAX_OBJS)()
    lib().gco_canonical_keys(_p(w, C.c_uint32), keys)
    return list(keys)


def philox_actions(seed, t, env):
    out = (C.c_uint8 * 4)()
    lib().gco_philox_actions(seed, t, env, out)
    return list(out)


def decode(w, n_agents):
    """Packed state -> (t, done, [(cell, hold_mask)...], sorted item keys) - the canonical
    tuple the golden fixtures store."""
    w = [int(x) for x in w]
    t, done = (w[0] >> 24) & 127, (w[0] >> 31) & 1
    slots = [int(v) for v in slots_of(np.array([w], dtype=np.uint32))[0]]
    agents = []
    for i in range(n_agents):
        hm = 0
        for s in slots:
            if (s >> 13) == i + 1:
                hm = s & 0x7F
        agents.append(((w[0] >> (6 * i)) & 63, hm))
    keys = []
    for s in slots:
        holder = s >> 13
        if holder == 7:
            continue
        cell, held = (s >> 7) & 63, 0
        if holder:
            cell, held = (w[0] >> (6 * (holder - 1))) & 63, 1
        keys.append(((s & 0x7F) << 7) | (cell << 1) | held)
    keys.sort()
    return t, done, agents, keys


_M64 = (1 << 64) - 1


def _mix64(z):
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z


def hash_canonical(t, agents, keys):
    """Independent Python statement of the canonical hash (third implementation, used to pin
    both the C oracle and the CUDA kernel)."""
    W0 = (t & 127) << 52
    for i, (cell, hm) in enumerate(agents):
        W0 |= (cell | (hm << 6)) << (13 * i)
    ks = list(keys) + [0x3FFF] * (MAX_OBJS - len(keys))
    W1 = ks[0] | (ks[1] << 14) | (ks[2] << 28)
    W2 = ks[3] | (ks[4] << 14) | (ks[5] << 28)
    h = _mix64((W0 + 0x9E3779B97F4A7C15) & _M64)
    h = _mix64(h ^ W1)
    h = _mix64(h ^ W2)
    return h


def bd_posterior(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta):
    probs = np.array(probs, dtype=np.float64, copy=True, order="C")
    n, H = probs.shape
    P, A = qdiff.shape[1], qdiff.shape[2]
    ne = hyp_pair.shape[2]
    alive = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
    hyp_pair = np.ascontiguousarray(hyp_pair, dtype=np.uint8)
    pair_w = np.ascontiguousarray(pair_w, dtype=np.uint8)
    qdiff = np.ascontiguousarray(qdiff, dtype=np.float64)
    n_valid = np.ascontiguousarray(n_valid, dtype=np.uint8)
    act_idx = np.ascontiguousarray(act_idx, dtype=np.uint8)
    lib().gco_bd_posterior(_p(probs, C.c_double), _p(alive, C.c_uint8), _p(hyp_pair, C.c_uint8),
                           _p(pair_w, C.c_uint8), _p(qdiff, C.c_double), _p(n_valid, C.c_uint8),
                           _p(act_idx, C.c_uint8), beta, n, H, P, A, ne)
    return probs
