"""Host recipe planner: subtask sets equal the reference's (modulo the argument order of the
food-food Merge, which the reference picks by set-iteration order, i.e. PYTHONHASHSEED)."""
import pytest

from gym_cooking_b200 import recipe_planner as rp

OBJECTS = ["Tomato", "Lettuce", "Plate", "Plate"]


def _canon(subtasks):
    out = []
    for s in subtasks:
        s = str(s)
        if s.startswith("Merge(") and "Plate" not in s and "-" not in s:
            a, b = s[6:-1].split(", ")
            s = "Merge(%s)" % ", ".join(sorted((a, b)))
        out.append(s)
    return sorted(out)


def test_known_subtask_sets():
    assert [str(s) for s in rp.level_subtasks(["SimpleTomato"], OBJECTS)] == \
        ["Chop(Tomato)", "Merge(Tomato, Plate)", "Deliver(Plate-Tomato)"]
    tl = [str(s) for s in rp.level_subtasks(["SimpleTomato", "SimpleLettuce"], OBJECTS)]
    assert tl == ["Chop(Tomato)", "Merge(Tomato, Plate)", "Deliver(Plate-Tomato)",
                  "Chop(Lettuce)", "Merge(Lettuce, Plate)", "Deliver(Lettuce-Plate)"]
    salad = _canon(rp.level_subtasks(["Salad"], OBJECTS))
    assert salad == sorted(["Chop(Lettuce)", "Chop(Tomato)", "Deliver(Lettuce-Plate-Tomato)", "Merge(Lettuce, Plate)",
                            "Merge(Lettuce, Plate-Tomato)", "Merge(Lettuce, Tomato)", "Merge(Lettuce-Tomato, Plate)",
                            "Merge(Tomato, Lettuce-Plate)", "Merge(Tomato, Plate)"])


def test_subtask_masks():
    st = rp.level_subtasks(["Salad"], OBJECTS)
    masks = {str(s): rp.subtask_masks(s) for s in st}
    assert masks["Chop(Tomato)"] == (1, 0x01, 0, 0x11)
    assert masks["Merge(Tomato, Plate)"] == (2, 0x11, 0x08, 0x19)
    assert masks["Merge(Lettuce, Plate-Tomato)"] == (2, 0x22, 0x19, 0x3B)
    assert masks["Deliver(Lettuce-Plate-Tomato)"] == (3, 0x3B, 0, 0x3B)
    assert rp.subtask_masks(None) == (0, 0, 0, 0)


def test_matches_reference_when_present():
    import ref_harness as H
    if not H.reference_available():
        pytest.skip("reference not mounted")
    for lvl, recipes in (("open-divider_tomato", ["SimpleTomato"]), ("full-divider_tl", ["SimpleTomato", "SimpleLettuce"]),
                         ("partial-divider_salad", ["Salad"])):
        env = H.make_env(lvl, 2)
        assert _canon(env.all_subtasks) == _canon(rp.level_subtasks(recipes, OBJECTS))
    ref = H.load_reference()
    env = H.make_env("open-divider_salad", 2)
    with H.quiet():
        sw = ref["env"].STRIPSWorld(env.world, [ref["recipe"].OnionSalad()])
        sw.initial.add_predicate(ref["recipe_utils"].Fresh("Onion"))
        sub = sw.get_subtasks(max_path_length=14)
    assert _canon(s for p in sub for s in p) == _canon(rp.level_subtasks(["OnionSalad"], OBJECTS[:2] + ["Onion", "Plate", "Plate"]))
