"""Host-side outputs of the reference's main loop: the `Bag` pickle (misc/metrics/metrics_bag.py:5-72)
and the `--record` PNG frames (misc/game/gameimage.py:54-62)."""
import argparse
import pickle
import types

import numpy as np

from gym_cooking_b200.misc.game.gameimage import GameImage, decode_png, encode_png
from gym_cooking_b200.misc.metrics.metrics_bag import Bag

REFERENCE_KEYS = {"level", "num_agents", "profiling", "num_completed_subtasks", "agent-1", "agent-2", "states",
                  "actions", "subtasks", "subtask_agents", "bayes", "holding", "incomplete_subtasks", "all_subtasks",
                  "num_total_subtasks", "collisions", "termination", "was_successful", "num_completed_subtasks_end"}


def test_png_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(560, 560, 3), dtype=np.uint8)
    data = encode_png(img)
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and data[12:16] == b"IHDR"
    assert (decode_png(data) == img).all()
    g = GameImage("lvl_agents2_seed1", lambda: img, record=True, root=str(tmp_path))
    (tmp_path / "lvl_agents2_seed1" / "stale.png").write_bytes(b"x")
    g = GameImage("lvl_agents2_seed1", lambda: img, record=True, root=str(tmp_path))  # clears the folder
    path = g.save_image_obs(7)
    assert path.endswith("t=007.png") and sorted(p.name for p in (tmp_path / "lvl_agents2_seed1").iterdir()) == ["t=007.png"]
    assert (decode_png(open(path, "rb").read()) == img).all()


def test_bag_has_the_reference_keys(tmp_path):
    arglist = argparse.Namespace(level="open-divider_tomato", num_agents=2, model1="bd", model2="up", model3=None,
                                 model4=None)
    bag = Bag(arglist, "open-divider_tomato_agents2_seed1", directory=str(tmp_path))
    subtasks = ["Chop(Tomato)", "Merge(Tomato, Plate)", "Deliver(Plate-Tomato)"]
    bag.set_recipe(subtasks)
    probs = types.SimpleNamespace(get_list=lambda: [((("Chop(Tomato)", ("agent-1",)),), 1.0)])
    agents = [types.SimpleNamespace(name="agent-%d" % (i + 1), location=(1 + i, 2), action=(0, 1), subtask=subtasks[0],
                                    subtask_agent_names=("agent-1",), incomplete_subtasks=subtasks[i:],
                                    get_holding=lambda: "None", delegator=types.SimpleNamespace(probs=probs))
              for i in range(2)]
    bag.add_status(1, agents)
    bag.add_status(2, agents)
    bag.set_collisions([])
    path = bag.set_termination("Terminating because all deliveries were completed", True)
    data = pickle.load(open(path, "rb"))
    assert set(data) == REFERENCE_KEYS
    assert data["num_completed_subtasks"] == [1, 1] and data["num_completed_subtasks_end"] == 1
    assert data["states"]["agent-2"] == [(2, 2), (2, 2)] and list(data["bayes"]["agent-1"]) == [1, 2]
    assert data["agent-1"] == "bd" and data["agent-2"] == "up" and data["was_successful"] is True
