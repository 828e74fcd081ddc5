"""Pins the CPU oracle (oracle/gc_oracle.c) for path A against fixtures produced by running
the unmodified reference (oracle/gen_golden.py env -> tests/golden/env_traces.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as O
import gym_cooking_b200 as gcb


from conftest import TRACE_FILES, level_source


@pytest.fixture(scope="module", params=TRACE_FILES)
def traces(golden_dir, request):
    return np.load(os.path.join(golden_dir, request.param))


def _replay(traces, r):
    """Yield (step, packed state, rd, ncoll, executed) from the oracle for trace r."""
    lvl_idx, n_agents, max_t, crashed = (int(x) for x in traces["meta"][r])
    lv = O.parse_level(level_source(str(traces["levels"][lvl_idx]))[0], max_t)
    L = O.lib()
    e = O.Env()
    L.gco_reset(C.byref(lv), n_agents, C.byref(e))
    w = (C.c_uint32 * 4)()
    L.gco_pack(C.byref(e), w)
    yield 0, list(w), 0, 0, [4] * n_agents
    for s in range(int(traces["length"][r])):
        a = (C.c_uint8 * 4)(*[int(x) for x in traces["actions"][r, s]])
        ex = (C.c_uint8 * 4)(4, 4, 4, 4)
        nc = L.gco_step(C.byref(lv), C.byref(e), a, ex)
        L.gco_pack(C.byref(e), w)
        rd = (1 if e.done else 0) | (2 if e.successful else 0)
        yield s + 1, list(w), rd, nc, list(ex)[:n_agents]


def test_oracle_matches_reference_traces(traces):
    n = traces["meta"].shape[0]
    steps = 0
    lv_cache = {}
    for r in range(n):
        lvl_idx, n_agents, max_t, crashed = (int(x) for x in traces["meta"][r])
        key = (lvl_idx, max_t)
        if key not in lv_cache:
            lv_cache[key] = O.parse_level(level_source(str(traces["levels"][lvl_idx]))[0], max_t)
        L = int(traces["length"][r])
        states, rd, nc, ex = O.replay(lv_cache[key], n_agents, traces["actions"][r, :L])
        t, done, agents, keys = O.decode_batch(states, n_agents)
        ctx = "trace %d" % r
        assert (t == traces["t"][r, :L + 1]).all(), ctx
        assert (agents == traces["agents"][r, :L + 1, :n_agents]).all(), ctx
        assert (keys == traces["keys"][r, :L + 1]).all(), ctx
        assert ((rd & 1) == traces["done"][r, :L + 1]).all(), ctx
        assert ((rd >> 1) == traces["reward"][r, :L + 1]).all(), ctx
        assert (done == traces["done"][r, :L + 1]).all(), ctx
        assert (nc == traces["ncoll"][r, :L + 1]).all(), ctx
        assert (ex[:, :n_agents] == traces["executed"][r, :L + 1, :n_agents]).all(), ctx
        steps += L + 1
    assert steps > (30000 if n > 100 else 3000)


def test_hash_c_equals_python(traces):
    L = O.lib()
    for r in range(0, traces["meta"].shape[0], 37):
        n_agents = int(traces["meta"][r, 1])
        for s, w, *_ in _replay(traces, r):
            arr = np.array(w, dtype=np.uint32)
            h_c = L.gco_hash_packed(arr.ctypes.data_as(C.POINTER(C.c_uint32)), n_agents)
            t, _, agents, keys = O.decode(w, n_agents)
            assert h_c == O.hash_canonical(t, agents, keys)
            assert O.canonical_keys(arr)[:len(keys)] == keys


def test_pack_unpack_roundtrip(traces):
    L = O.lib()
    for r in range(0, traces["meta"].shape[0], 41):
        n_agents = int(traces["meta"][r, 1])
        for s, w, *_ in _replay(traces, r):
            arr = np.array(w, dtype=np.uint32)
            e = O.Env()
            L.gco_unpack(arr.ctypes.data_as(C.POINTER(C.c_uint32)), n_agents, C.byref(e))
            e.n_objs = O.MAX_OBJS
            w2 = (C.c_uint32 * 4)()
            L.gco_pack(C.byref(e), w2)
            assert list(w2) == w


def test_frozen_after_done(traces):
    """Batched convention: a finished episode no longer mutates and keeps its outcome."""
    lv = O.parse_level(gcb.levels.level_text("open-divider_tomato"), 5)
    st = O.reset_state(lv, 2, n=4)
    rng = np.random.RandomState(0)
    for _ in range(5):
        O.step_batch(lv, st, rng.randint(0, 5, size=(4, 2)), 2)
    snap = st.copy()
    rd, _ = O.step_batch(lv, st, rng.randint(0, 5, size=(4, 2)), 2)
    assert (st == snap).all() and (rd == 1).all()


def test_philox_reference_vector():
    """Random123 known-answer test for philox4x32-10: counter = key = 0 and all-ones."""
    import ctypes
    L = O.lib()
    # action = mulhi(word, 5); check through the raw words by exhaustive identity instead:
    # KAT words (Random123 kat_vectors): ctr 0 key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    words = [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox_actions(0, 0, 0) == [(w * 5) >> 32 for w in words]
    # ctr = ffffffff x4 is not reachable (4th counter word fixed to 0), so check a second
    # point against the pure-Python philox below
    assert O.philox_actions(0x0123456789ABCDEF, 77, (3 << 32) | 9) == _py_philox(0x0123456789ABCDEF, 77, (3 << 32) | 9)


def _py_philox(seed, t, env):
    c = [t, env & 0xFFFFFFFF, env >> 32, 0]
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return [(w * 5) >> 32 for w in c]
