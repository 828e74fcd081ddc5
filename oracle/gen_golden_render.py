"""Golden images for path A' (image_obs): the UNMODIFIED drawing code of the reference -
misc/game/game.py Game.on_render / draw_gridsquare / draw_object / draw_agent / draw_agent_object and their
geometry helpers (:56-185) - run over reference env states, with a pygame stand-in backed by PIL (this
image has no pygame): Surface = an RGB PIL image, fill / draw.rect = rectangle fills, transform.scale = PIL's
nearest-neighbour resize of the loaded image, blit = alpha compositing ("over") of the RGBA sprite.
What this pins: which squares, objects and agents are drawn, in which order, at which sizes and offsets,
with which colours and which of the reference's own PNG files.  What it cannot pin: the exact source pixel
pygame's scaler would pick and the rounding of SDL's blend - hence the tests' tolerance of +-1 and this
caveat in DESIGN.md.  Also written: the sprite atlas gc_render needs as its input, built from the same PNGs
(render.load_atlas) - test input data, like the level files.

    python oracle/gen_golden_render.py        ->  tests/golden/render.npz   (build container only)
"""
import os
import sys
import types

import numpy as np
from PIL import Image

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
sys.path.insert(0, os.path.join(_HERE, ".."))
GOLDEN = os.path.join(_HERE, "..", "tests", "golden")


class _Surface:
    def __init__(self, size):
        self.img = Image.new("RGB", tuple(int(v) for v in size))

    def fill(self, color):
        self.img.paste(tuple(color), (0, 0) + self.img.size)

    def blit(self, sprite, location):
        x, y = (int(v) for v in location)
        self.img.paste(sprite.img.convert("RGB"), (x, y), sprite.img.split()[3])  # "over" with the sprite's alpha


class _Sprite:
    def __init__(self, img):
        self.img = img

    def convert_alpha(self):
        return self


def _install_pygame_stub(graphics_dir):
    pg = types.ModuleType("pygame")
    pg.Surface = _Surface
    pg.init = lambda: None
    pg.quit = lambda: None
    pg.QUIT = 0
    pg.Rect = lambda x, y, w, h: (int(x), int(y), int(w), int(h))

    def rect(surface, color, r, width=0):
        x, y, w, h = r
        if width == 0:
            surface.img.paste(tuple(color), (x, y, x + w, y + h))
        else:  # outline of `width` px inside the rectangle
            for k in range(width):
                for box in ((x + k, y + k, x + w - k, y + k + 1), (x + k, y + h - k - 1, x + w - k, y + h - k),
                            (x + k, y + k, x + k + 1, y + h - k), (x + w - k - 1, y + k, x + w - k, y + h - k)):
                    surface.img.paste(tuple(color), box)

    pg.draw = types.SimpleNamespace(rect=rect)
    pg.transform = types.SimpleNamespace(
        scale=lambda sprite, size: _Sprite(sprite.img.resize(tuple(int(v) for v in size), Image.NEAREST)))

    def load(path):
        if not os.path.exists(path):  # game.py draws 'Plate' while the file is plate.png (case-insensitive on Windows)
            d, f = os.path.split(path)
            alt = [g for g in os.listdir(d) if g.lower() == f.lower()]
            path = os.path.join(d, alt[0])
        return _Sprite(Image.open(path).convert("RGBA"))

    pg.image = types.SimpleNamespace(load=load)
    pg.display = types.SimpleNamespace(set_mode=lambda size: _Surface(size), flip=lambda: None, update=lambda: None)
    pg.event = types.SimpleNamespace(get=lambda: [])
    for name in ("K_UP", "K_DOWN", "K_LEFT", "K_RIGHT", "K_1", "K_2", "K_3", "K_4", "K_RETURN", "KEYDOWN"):
        setattr(pg, name, 0)
    sys.modules["pygame"] = pg
    return pg


def main():
    import ref_harness as H
    from gen_golden import walker_actions, DELTA
    from gen_golden_plan import pack_env
    gdir = os.path.join(H.REF_ROOT, "misc", "game", "graphics")
    _install_pygame_stub(gdir)
    ref = H.load_reference()
    import misc.game.game as game_mod  # the reference's module, from the scratch copy on sys.path
    levels, rows = [], []
    jobs = [("open-divider_salad", 4, 31, 12), ("open-divider_salad", 4, 32, 30), ("partial-divider_tl", 2, 33, 25),
            ("partial-divider_tl", 2, 34, 45), ("full-divider_tomato", 3, 35, 35), ("open-divider_tomato", 1, 36, 20),
            ("full-divider_salad", 3, 37, 50), ("open-divider_tl", 2, 38, 60)]
    for level, n_agents, seed, n_steps in jobs:
        rng = np.random.RandomState(seed)
        env = H.make_env(level, n_agents, 100)
        names = env.get_agent_names()
        targets = [None] * n_agents
        for _ in range(n_steps):
            acts = walker_actions(env, rng, 0.25, targets)
            try:
                with H.quiet():
                    _, _, done, _ = env.step({names[i]: DELTA[a] for i, a in enumerate(acts)})
            except (AssertionError, AttributeError):
                break
            if done:
                break
        game = game_mod.Game(env.world, env.sim_agents)
        game.on_init()
        game.on_render()
        img = np.asarray(game.screen.img, dtype=np.uint8)
        if level not in levels:
            levels.append(level)
        rows.append((levels.index(level), n_agents, pack_env(env), img))
        print(level, n_agents, "t =", env.t, img.shape)
    from gym_cooking_b200 import render as R  # builds the atlas from the same PNG files
    atlas = R.load_atlas(gdir)
    np.savez_compressed(os.path.join(GOLDEN, "render.npz"), levels=np.array(levels),
                        level=np.array([r[0] for r in rows], dtype=np.uint8),
                        n_agents=np.array([r[1] for r in rows], dtype=np.uint8),
                        state=np.array([r[2] for r in rows], dtype=np.uint32), state_layout=np.array("abi2-byte-planes"),
                        image=np.stack([r[3] for r in rows]), atlas=atlas)
    print("wrote render.npz:", len(rows), "images,", os.path.getsize(os.path.join(GOLDEN, "render.npz")) >> 10, "KB")


if __name__ == "__main__":
    main()
