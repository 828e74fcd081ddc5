"""gc_render against an independent numpy compositing of the same rules (misc/game/game.py
geometry; image parity with pygame itself is unpinned - no pygame in this image)."""
import numpy as np
import pytest
import torch

import gym_cooking_b200 as gcb
from gym_cooking_b200 import render as R

pytestmark = pytest.mark.gpu
T = 80


def _over(dst, spr, size, off_x, off_y):
    s = spr[R.SIZES.index(size), :size, :size].astype(np.float32)  # the frame pre-scaled to `size`
    a = s[..., 3:4] / 255.0
    box = dst[off_y:off_y + size, off_x:off_x + size]
    box[...] = s[..., :3] * a + box * (1 - a)


def _object(dst, atlas, mask, size, ox, oy):
    if mask & 8:
        _over(dst, atlas[R.SP_PLATE], size, ox, oy)
        if mask & 7:
            _over(dst, atlas[R.SP_FOOD0 + R.food_code(mask)], int(0.7 * size), ox + int(size * 0.15), oy + int(size * 0.15))
    else:
        _over(dst, atlas[R.SP_FOOD0 + R.food_code(mask)], size, ox, oy)


def numpy_render(level, words, n_agents, atlas):
    W, H = level.width, level.height
    img = np.empty((H * T, W * T, 3), dtype=np.float32)
    img[...] = (245, 230, 210)
    st = gcb.decode_state(words, n_agents)
    for y in range(H):
        for x in range(W):
            ty = level.cell_type[y * 8 + x]
            tile = img[y * T:(y + 1) * T, x * T:(x + 1) * T]
            if ty == 3:
                tile[...] = 96
                _over(img, atlas[R.SP_DELIVERY], T, x * T, y * T)
            elif ty != 0:
                tile[...] = (220, 170, 110)
                tile[0, :] = tile[-1, :] = tile[:, 0] = tile[:, -1] = (114, 93, 51)
                if ty == 2:
                    _over(img, atlas[R.SP_CUTBOARD], T, x * T, y * T)
    lying = {}
    for mask, x, y, holder in st["objects"]:
        if not holder:
            lying[(x, y)] = mask  # later slots win, like the kernel
    for (x, y), mask in lying.items():
        _object(img, atlas, mask, T, x * T, y * T)
    top = {}
    for i, (x, y, hold) in enumerate(st["agents"]):
        top[(x, y)] = (i, hold)
    for (x, y), (i, hold) in top.items():
        _over(img, atlas[R.SP_AGENT0 + i], T, x * T, y * T)
        if hold:
            _object(img, atlas, hold, T // 2, x * T + T // 2, y * T + T // 2)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("level,n_agents", [("open-divider_salad", 4), ("partial-divider_tl", 2), ("full-divider_tomato", 3)])
def test_render_matches_numpy_compositing(level, n_agents):
    n = 64
    kb = gcb.KitchenBatch(level, n_agents, n, 100)
    acts = kb.random_actions(80, seed=11)
    idx = torch.arange(n, device=kb.device)
    for s in range(80):
        a = acts[s].clone()
        a[(idx * 7 % 81) <= s] = 4
        kb.step(a)
    atlas = R.default_atlas()
    img = R.render(kb, atlas).cpu().numpy()
    assert img.shape == (n, 560, 560, 3)
    words = kb.state.cpu().numpy().view(np.uint32)
    for e in range(0, n, 5):
        exp = numpy_render(kb.levels[0], words[e], n_agents, atlas)
        diff = np.abs(img[e].astype(np.int32) - exp.astype(np.int32))
        assert diff.max() <= 1, (e, int(diff.max()), int((diff > 1).sum()))  # rounding of the float blend only
    sub = R.render(kb, atlas, envs=torch.tensor([3, 10], device=kb.device)).cpu().numpy()
    assert (sub[0] == img[3]).all() and (sub[1] == img[10]).all()


def test_facade_image_obs():
    import argparse
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=100, max_num_subtasks=14, seed=1,
                            model1=None, model2=None, model3=None, model4=None, with_image_obs=True, record=False)
    env = gcb.OvercookedEnvironment(ns)
    env.reset()
    _, _, _, info = env.step({"agent-1": (0, 1), "agent-2": (-1, 0)})
    img = info["image_obs"]
    assert img.shape == (560, 560, 3) and img.dtype == np.uint8
    exp = numpy_render(env.level, env.state[0].tolist(), 2, R.default_atlas())
    assert np.abs(img.astype(np.int32) - exp.astype(np.int32)).max() <= 1
    ns.with_image_obs = False
    env2 = gcb.OvercookedEnvironment(ns)
    env2.reset()
    assert env2.step({"agent-1": (0, 0), "agent-2": (0, 0)})[3]["image_obs"] is None


def test_render_matches_the_reference_drawing_code(golden_dir):
    """tests/golden/render.npz: eight reference env states drawn by the reference's own Game.on_render
    (misc/game/game.py:56-185, unmodified) over a PIL-backed pygame stand-in (oracle/gen_golden_render.py), and
    the sprite atlas built from the reference's PNG files.  gc_render must reproduce every image within +-1
    (rounding of the alpha blend): squares, objects, agents, held and plated items, sizes, offsets, colours,
    draw order."""
    import os
    g = np.load(os.path.join(golden_dir, "render.npz"))
    atlas = g["atlas"]
    assert atlas.shape == (R.N_SPRITES, len(R.SIZES), T, T, 4)
    for r in range(g["state"].shape[0]):
        level, n_agents = str(g["levels"][g["level"][r]]), int(g["n_agents"][r])
        kb = gcb.KitchenBatch(level, n_agents, 1, 100)
        kb.state.copy_(torch.from_numpy(g["state"][r:r + 1].view(np.int32)).to(kb.device))
        img = R.render(kb, atlas).cpu().numpy()[0]
        want = g["image"][r]
        assert img.shape == want.shape
        diff = np.abs(img.astype(np.int32) - want.astype(np.int32))
        assert diff.max() <= 1, (r, level, int(diff.max()), int((diff > 1).sum()))
