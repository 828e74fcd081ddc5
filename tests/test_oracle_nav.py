"""Pins the CPU oracle for path B against reference-generated fixtures:
  - the distance heuristic must EQUAL env.get_lower_bound_for_subtask_given_objs on every row;
  - the exact level-0 value must lie inside the reference BRTDP's converged [v_l, v_u] bracket
    (tolerance 1e-4, BASELINE.json north_star)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as O
import gym_cooking_b200 as gcb


def _env_from_words(words, n_agents):
    arr = np.ascontiguousarray(words, dtype=np.uint32)
    e = O.Env()
    O.lib().gco_unpack(arr.ctypes.data_as(C.POINTER(C.c_uint32)), n_agents, C.byref(e))
    e.n_objs = O.MAX_OBJS
    return e


from conftest import level_source


@pytest.mark.parametrize("fixture", ["lower_bounds.npz", "lower_bounds_custom.npz"])
def test_lower_bound_equals_reference(golden_dir, fixture):
    g = np.load(os.path.join(golden_dir, fixture))
    L = O.lib()
    lv = {i: O.parse_level(level_source(str(n))[0], 100) for i, n in enumerate(g["levels"])}
    bad = 0
    for r in range(len(g["lb"])):
        n_agents = int(g["n_agents"][r])
        e = _env_from_words(g["state"][r], n_agents)
        k, a, b, goal = (int(v) for v in g["subtask"][r])
        st = O.Subtask(k, a, b, goal)
        aj = int(g["agent_j"][r])
        got = L.gco_lower_bound(C.byref(lv[int(g["level"][r])]), C.byref(e), C.byref(st), int(g["agent_i"][r]),
                                -1 if aj == 255 else aj)
        if got != float(g["lb"][r]):
            bad += 1
            if bad < 5:
                print("row", r, "level", g["levels"][g["level"][r]], "subtask", (k, a, b, goal),
                      "agents", g["agent_i"][r], aj, "expected", g["lb"][r], "got", got)
    assert bad == 0
    assert len(g["lb"]) > 4000
