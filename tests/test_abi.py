"""CPU-only checks of the drop-in boundary: the library builds, loads, exports every entry
point include/gymcook.h declares, and its host-side level loader agrees with the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gym_cooking_b200 as gcb
from gym_cooking_b200 import _lib
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gymcook.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "libgymcook.so does not export %s" % name
    assert sorted(_lib.ABI_SYMBOLS) == declared
    assert lib.gc_version() == 2


def test_level_loader_matches_oracle():
    for name in gcb.levels.LEVEL_NAMES:
        text = gcb.levels.level_text(name)
        lv = _lib.parse_level(text, 100)
        olv = O.parse_level(text, 100)
        assert (lv.width, lv.height) == (olv.width, olv.height) == (7, 7)
        assert lv.n_objects == olv.n_objs == 4
        assert lv.delivery_cell == olv.delivery_y * 8 + olv.delivery_x
        for y in range(7):
            for x in range(7):
                assert lv.cell_type[y * 8 + x] == olv.type[y][x]
        assert [lv.goal_mask[g] for g in range(lv.n_goals)] == [olv.goal_mask[g] for g in range(olv.n_goals)]
        for k in range(4):
            assert lv.object_init[k] == olv.obj_mask[k] | ((olv.obj_y[k] * 8 + olv.obj_x[k]) << 7)
        assert [lv.agent_cell[i] for i in range(4)] == [olv.agent_y[i] * 8 + olv.agent_x[i] for i in range(4)]


@pytest.mark.parametrize("text,code", [
    ("---\n- -\n---\n\nSalad\n\n1 1\n", -2),             # no delivery square
    ("---\n* -\n---\n\nPizza\n\n1 1\n", -2),             # unknown recipe
    ("---\n*  \n---\n\nSalad\n\n1 1\n", -2),             # floor on the outer ring
    ("-t-t-\n*   -\n-----\n\nSalad\n\n1 1\n", -4),       # two tomatoes
    ("---\n* -\n---\n\nSalad\n\n0 0\n", -2),             # agent on a counter
])
def test_level_loader_rejects(text, code):
    lv = _lib.Level()
    data = text.encode()
    assert _lib.load().gc_level_parse(data, len(data), 100, C.byref(lv)) == code
    assert _lib.load().gc_last_error()


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.gc_device_count() == 0
    lv = _lib.parse_level(gcb.levels.level_text("open-divider_tomato"), 100)
    fake = C.c_void_p(16)
    rc = lib.gc_env_step(C.byref(lv), 1, None, fake, fake, None, None, None, None, 4, 2, None)
    assert rc == -3 and b"no CPU fallback" in lib.gc_last_error()
    with pytest.raises(_lib.GcError):
        gcb.KitchenBatch("open-divider_tomato", 2, 8)
