"""image_obs for batched kitchens (optional path A'): sprite atlas + gc_render marshalling.

The reference renders with pygame from misc/game/graphics/*.png (misc/game/game.py:105-165).
Those files are assets of the reference, not of this repo; `load_atlas(graphics_dir)` reads them
with PIL when a checkout is at hand, `default_atlas()` draws simple procedural sprites so that
rendering works without it.  Atlas layout (include/gymcook.h, gc_render): uint8[71][4][80][80][4] RGBA -
0 delivery, 1 cutboard, 2 plate, 3-6 agents (blue, magenta, yellow, green), 7 + code food sprites
with code = (mask & 7) | ((mask >> 4) & 7) << 3; per sprite four frames, pre-scaled to the sizes the
reference blits at (SIZES = 80, 56, 40, 28 px; game.py:105-108 scales the original image each time).
"""
import os

import numpy as np
import torch

from . import _lib
from .utils.core import mask_names

N_SPRITES, TILE = 71, 80
SIZES = (80, 56, 40, 28)  # tile, plated contents (0.7), held object (0.5), held plated contents (0.35)
SP_DELIVERY, SP_CUTBOARD, SP_PLATE, SP_AGENT0, SP_FOOD0 = 0, 1, 2, 3, 7
AGENT_COLORS = ((0, 0, 255), (255, 0, 255), (255, 255, 0), (0, 160, 0))
FOOD_COLORS = {1: (200, 40, 40), 2: (60, 170, 60), 4: (170, 100, 190)}


def food_code(mask):
    return (mask & 7) | (((mask >> 4) & 7) << 3)


def _disc(rgb, radius, alpha=255, centre=(40, 40)):
    yy, xx = np.mgrid[0:TILE, 0:TILE]
    inside = (xx - centre[0]) ** 2 + (yy - centre[1]) ** 2 <= radius * radius
    out = np.zeros((TILE, TILE, 4), dtype=np.uint8)
    out[inside] = (*rgb, alpha)
    return out


def with_scaled_frames(sprites80):
    """uint8[n][80][80][4] -> uint8[n][4][80][80][4]: frames 1..3 are nearest-neighbour samples of the 80 px
    sprite (source pixel (k * 80) // size), in the top-left corner of their frame"""
    out = np.zeros((sprites80.shape[0], len(SIZES), TILE, TILE, 4), dtype=np.uint8)
    for lvl, size in enumerate(SIZES):
        idx = (np.arange(size) * TILE) // size
        out[:, lvl, :size, :size] = sprites80[:, idx][:, :, idx]
    return out


def default_atlas():
    """Procedural sprites: discs for food (lighter when chopped, one wedge per ingredient), a
    white disc for the plate, colour squares for agents, inset rectangles for cutboard/delivery."""
    return with_scaled_frames(_default_sprites())


def _default_sprites():
    atlas = np.zeros((N_SPRITES, TILE, TILE, 4), dtype=np.uint8)
    atlas[SP_DELIVERY, 20:60, 20:60] = (160, 160, 160, 255)
    atlas[SP_CUTBOARD, 12:68, 18:62] = (150, 110, 60, 255)
    atlas[SP_PLATE] = _disc((245, 245, 245), 34)
    for i, col in enumerate(AGENT_COLORS):
        atlas[SP_AGENT0 + i, 8:72, 8:72] = (*col, 255)
        atlas[SP_AGENT0 + i, 20:34, 20:60] = (255, 255, 255, 255)  # a visor, so orientation of the tile is visible
    for code in range(1, 64):
        present, chopped = code & 7, code >> 3
        if chopped & ~present:
            continue
        kinds = [b for b in (1, 2, 4) if present & b]
        spr = np.zeros((TILE, TILE, 4), dtype=np.uint8)
        for j, b in enumerate(kinds):
            base = np.array(FOOD_COLORS[b], dtype=np.float32)
            col = tuple(int(v) for v in (base + (255 - base) * (0.45 if chopped & b else 0.0)))
            cx = 40 + (j - (len(kinds) - 1) / 2) * 18
            d = _disc(col, 26 - 4 * (len(kinds) - 1), centre=(cx, 40))
            spr = np.where(d[..., 3:4] > 0, d, spr)
        atlas[SP_FOOD0 + code] = spr
    return atlas


def load_atlas(graphics_dir):
    """Atlas from the reference's PNG sprites (misc/game/graphics): every image scaled from its ORIGINAL
    resolution to each of the four blit sizes, as Game.draw does (game.py:105-108: pygame.transform.scale(
    image, size) on the loaded image); PIL's nearest-neighbour resize stands in for pygame's scaler."""
    from PIL import Image

    def load(name):
        path = os.path.join(graphics_dir, name + ".png")
        if not os.path.exists(path):
            alt = [f for f in os.listdir(graphics_dir) if f.lower() == (name + ".png").lower()]  # 'Plate' vs plate.png
            if not alt:
                return None
            path = os.path.join(graphics_dir, alt[0])
        src = Image.open(path).convert("RGBA")
        frames = np.zeros((len(SIZES), TILE, TILE, 4), dtype=np.uint8)
        for lvl, size in enumerate(SIZES):
            frames[lvl, :size, :size] = np.asarray(src.resize((size, size), Image.NEAREST))
        return frames

    atlas = default_atlas()
    for slot, name in ((SP_DELIVERY, "delivery"), (SP_CUTBOARD, "cutboard"), (SP_PLATE, "plate"),
                       (SP_AGENT0, "agent-blue"), (SP_AGENT0 + 1, "agent-magenta"), (SP_AGENT0 + 2, "agent-yellow"),
                       (SP_AGENT0 + 3, "agent-green")):
        img = load(name)
        if img is not None:
            atlas[slot] = img
    for code in range(1, 64):
        present, chopped = code & 7, code >> 3
        if chopped & ~present:
            continue
        img = load(mask_names(present | (chopped << 4))[1])
        if img is not None:
            atlas[SP_FOOD0 + code] = img
    return atlas


def render(batch, atlas, envs=None):
    """uint8[m][H*80][W*80][3] RGB images of `envs` (index tensor; default all) of a KitchenBatch."""
    lib = _lib.load()
    state = batch.state if envs is None else batch.state[envs].contiguous()
    level_id = batch.level_id if (envs is None or batch.level_id is None) else batch.level_id[envs].contiguous()
    m = state.shape[0]
    lv = batch.levels[0]
    with torch.cuda.device(batch.device):
        if not isinstance(atlas, torch.Tensor):
            atlas = torch.from_numpy(np.ascontiguousarray(atlas)).to(batch.device)
        img = torch.empty((m, lv.height * TILE, lv.width * TILE, 3), dtype=torch.uint8, device=batch.device)
        _lib.check(lib.gc_render(batch._lv(), batch.n_levels, _lib.ptr(level_id), _lib.ptr(state),
                                 _lib.ptr(atlas, torch.uint8), _lib.ptr(img), m, batch.num_agents, batch._stream()))
    return img
