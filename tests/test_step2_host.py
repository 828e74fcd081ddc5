"""The product's transition function - gym-cooking_b200/csrc/gc_step2.cuh, the code the CUDA step kernels
inline - compiled as HOST code (tests/step2_host_shim.cpp, g++) and held against (1) the reference-generated
golden traces and (2) the CPU oracle on random walks.  No GPU needed: this is the CPU suite's check of the
kernel's own source; the GPU suite repeats both through the C-ABI on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import gym_cooking_b200 as gcb
from conftest import TRACE_FILES, level_source

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("step2") / "libstep2host.so")
    subprocess.run(["g++", "-O2", "-w", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "step2_host_shim.cpp")],
                   check=True)
    lib = C.CDLL(out)
    lib.s2h_step.argtypes = [C.POINTER(gcb._lib.Level), C.c_int] + [C.c_void_p] * 5 + [C.c_longlong]
    lib.s2h_step.restype = C.c_int
    lib.s2h_initial_state.argtypes = [C.POINTER(gcb._lib.Level), C.c_int, C.c_void_p]
    return lib


def _step(lib, lv, n_agents, state, actions):
    n = state.shape[0]
    actions = np.ascontiguousarray(actions, dtype=np.uint8)
    rd, ex, nc = np.zeros(n, np.uint8), np.zeros((n, n_agents), np.uint8), np.zeros(n, np.uint32)
    assert lib.s2h_step(C.byref(lv), n_agents, state.ctypes.data, actions.ctypes.data, rd.ctypes.data, ex.ctypes.data,
                        nc.ctypes.data, n) == 0
    return rd, ex, nc


@pytest.mark.parametrize("trace_file", TRACE_FILES)
def test_host_compiled_step_matches_reference_traces(shim, golden_dir, trace_file):
    traces = np.load(os.path.join(golden_dir, trace_file))
    meta, length = traces["meta"], traces["length"]
    groups = {}
    for r in range(meta.shape[0]):
        groups.setdefault((int(meta[r, 0]), int(meta[r, 1]), int(meta[r, 2])), []).append(r)
    checked = 0
    for (lvl, n_agents, max_t), rows in sorted(groups.items()):
        rows = np.array(rows)
        lv = gcb._lib.parse_level(level_source(str(traces["levels"][lvl]))[0], max_t)
        w0 = np.zeros(4, np.uint32)
        shim.s2h_initial_state(C.byref(lv), n_agents, w0.ctypes.data)
        state = np.tile(w0, (len(rows), 1))
        L = length[rows]
        for s in range(1, int(L.max()) + 1):
            rd, ex, nc = _step(shim, lv, n_agents, state, traces["actions"][rows, s - 1, :n_agents])
            live = L >= s
            t, done, agents, keys = O.decode_batch(state, n_agents)
            assert (t[live] == traces["t"][rows, s][live]).all()
            assert (agents[live] == traces["agents"][rows, s, :n_agents][live]).all()
            assert (keys[live] == traces["keys"][rows, s][live]).all()
            assert ((rd & 1)[live] == traces["done"][rows, s][live]).all()
            assert ((rd >> 1)[live] == traces["reward"][rows, s][live]).all()
            assert (nc[live] == traces["ncoll"][rows, s][live]).all()
            assert (ex[live] == traces["executed"][rows, s, :n_agents][live]).all()
            checked += int(live.sum())
    assert checked > (30000 if meta.shape[0] > 100 else 3000)


@pytest.mark.parametrize("level,n_agents", [("partial-divider_tl", 2), ("full-divider_salad", 3), ("open-divider_salad", 4),
                                            ("onion-8x8", 4), ("open-divider_tomato", 1)])
def test_host_compiled_step_matches_oracle_on_random_walks(shim, level, n_agents):
    n, steps = 4000, 60
    text, _ = level_source(level)
    lv = gcb._lib.parse_level(text, 40)
    olv = O.parse_level(text, 40)
    ost = O.reset_state(olv, n_agents, n)
    state = ost.copy()
    rng = np.random.RandomState(5)
    for s in range(steps):
        acts = rng.randint(0, 6, size=(n, n_agents)).astype(np.uint8)  # 5 = out of range: "stay"
        rd, _, nc = _step(shim, lv, n_agents, state, acts)
        ord_, ocoll = O.step_batch(olv, ost, acts, n_agents)
        assert (state == ost).all(), s
        assert (rd == ord_).all() and (nc == ocoll).all(), s
