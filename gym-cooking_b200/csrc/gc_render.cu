// gc_render.cu - path A': image_obs renderer (optional row of SURVEY.md section 8).
//
// Restates the geometry of misc/game/game.py (Game.on_render :56-82, draw_gridsquare :85-103,
// draw_object :139-150, draw_agent / draw_agent_object :112-137, offsets :167-185) on the packed
// state: 80 px tiles; floor (245,230,210) background; counters (220,170,110) with a 1 px
// (114,93,51) border; delivery (96,96,96) + sprite; cutboard = counter + sprite; lying objects at
// tile size (plated contents 56 px at +12); agents at tile size; the held object 40 px at +40
// (its plated contents 28 px at +46).  Sprites come from a caller-supplied RGBA atlas
// uint8[GC_N_SPRITES][4][80][80][4]: every sprite pre-scaled to the four sizes the reference draws
// at (80, 56, 40, 28 px, each in the top-left corner of its own 80 x 80 frame) - game.py:105-108
// scales the ORIGINAL image to the blit size every time, so sampling an 80 px copy again would pick
// different source pixels (gym_cooking_b200.render builds the atlas procedurally, or from the
// reference's PNG files when they are available); blits are "over" alpha compositing like pygame's
// per-pixel-alpha blit.
// Output is RGB (the reference's get_image_obs writes (g, b, r) of a mapped pixel value,
// gameimage.py:48-50, which depends on the surface format - image parity is unpinned).
//
// One CTA per image row band; a thread produces 4 consecutive pixels (one tile never splits
// them: 80 % 4 == 0) and writes them as three 32-bit words, so a warp stores 384 contiguous
// bytes.  The per-square layer table of the image is staged in shared memory once per CTA.
#include "gc_device.cuh"
#include "gc_host.h"
#include "gc_nav.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 80;
constexpr int kSpriteBytes = kTile * kTile * 4;

// atlas slots
constexpr int SP_DELIVERY = 0, SP_CUTBOARD = 1, SP_PLATE = 2, SP_AGENT0 = 3, SP_FOOD0 = 7;  // + 6-bit food code

struct Square {     // what is drawn on one grid square, bottom to top
  uint8_t type;     // GC_CELL_*
  uint8_t lying;    // content mask of an un-held object lying here (0 = none)
  uint8_t agent;    // agent index + 1 standing here (0 = none); the highest index wins like the draw order
  uint8_t held;     // content mask that agent holds
};

__device__ __forceinline__ int food_sprite(uint32_t mask) { return SP_FOOD0 + (int)((mask & 7u) | (((mask >> 4) & 7u) << 3)); }

// "over" blend of sprite `sp` sampled for a destination box of `size` px at in-tile offset `off`
__device__ __forceinline__ void blit(const uint8_t* __restrict__ atlas, int sp, int size, int off, int px, int py,
                                     float (&rgb)[3]) {
  const int lx = px - off, ly = py - off;
  if (lx < 0 || ly < 0 || lx >= size || ly >= size) return;
  const int level = size == kTile ? 0 : size == 56 ? 1 : size == 40 ? 2 : 3;  // the frame pre-scaled to `size`
  const uchar4 t = *reinterpret_cast<const uchar4*>(atlas + ((size_t)sp * 4 + level) * kSpriteBytes + (ly * kTile + lx) * 4);
  const float a = t.w * (1.0f / 255.0f);
  rgb[0] = t.x * a + rgb[0] * (1.0f - a);
  rgb[1] = t.y * a + rgb[1] * (1.0f - a);
  rgb[2] = t.z * a + rgb[2] * (1.0f - a);
}

// an object (game.py:139-150 / 120-137): the plate first, then the rest at 0.7 scale, centred
__device__ __forceinline__ void blit_object(const uint8_t* __restrict__ atlas, uint32_t mask, int size, int off,
                                            int px, int py, float (&rgb)[3]) {
  if (mask & GC_M_PLATE) {
    blit(atlas, SP_PLATE, size, off, px, py, rgb);
    if (mask & 7u) {
      const int inner = (int)(0.7f * size);                  // container_scale
      const int inner_off = off + (int)(size * 0.15f);       // (1 - 0.7) / 2
      blit(atlas, food_sprite(mask), inner, inner_off, px, py, rgb);
    }
  } else {
    blit(atlas, food_sprite(mask), size, off, px, py, rgb);
  }
}

template <int NA>
__global__ void __launch_bounds__(kThreads)
render_kernel(const __grid_constant__ GcNavLevels levels, const uint8_t* __restrict__ level_id,
              const uint4* __restrict__ state, const uint8_t* __restrict__ atlas, uint8_t* __restrict__ img,
              int width, int height, int bands_per_image) {
  __shared__ Square sq[64];
  const int image = blockIdx.x / bands_per_image, band = blockIdx.x % bands_per_image;
  const GcNavLevel& L = levels.lv[level_id ? level_id[image] : 0];
  const uint4 s = state[image];
  if (threadIdx.x < 64) {
    const uint32_t c = threadIdx.x;
    Square q;
    q.type = ((L.floor_mask >> c) & 1ull) ? GC_CELL_FLOOR
             : ((L.cut_mask >> c) & 1ull) ? GC_CELL_CUTBOARD
             : ((L.deliv_mask >> c) & 1ull) ? GC_CELL_DELIVERY : GC_CELL_COUNTER;
    q.lying = q.agent = q.held = 0;
#pragma unroll
    for (int k = 0; k < GC_MAX_OBJECTS; k++) {
      const uint32_t sl = gcnav::slot_of(s, k);
      if ((sl >> 13) == 0u && ((sl >> 7) & 63u) == c) q.lying = (uint8_t)(sl & 0x7fu);
    }
#pragma unroll
    for (int i = 0; i < NA; i++) {
      if (((s.x >> (6 * i)) & 63u) != c) continue;
      q.agent = (uint8_t)(i + 1);
      q.held = 0;
#pragma unroll
      for (int k = 0; k < GC_MAX_OBJECTS; k++) {
        const uint32_t sl = gcnav::slot_of(s, k);
        if ((sl >> 13) == (uint32_t)(i + 1)) q.held = (uint8_t)(sl & 0x7fu);
      }
    }
    sq[c] = q;
  }
  __syncthreads();
  const int W = width * kTile, H = height * kTile;
  const int rows_per_band = (H + bands_per_image - 1) / bands_per_image;
  const int y0 = band * rows_per_band, y1 = min(H, y0 + rows_per_band);
  const int quads_per_row = W / 4;
  uint8_t* out = img + (size_t)image * W * H * 3;
  for (int q = threadIdx.x; q < (y1 - y0) * quads_per_row; q += kThreads) {
    const int y = y0 + q / quads_per_row, x = (q % quads_per_row) * 4;
    const int tx = x / kTile, ty = y / kTile, py = y - ty * kTile;
    const Square cell = sq[ty * 8 + tx];
    uint32_t packed[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int px = x + k - tx * kTile;
      float rgb[3] = {245.f, 230.f, 210.f};  // Color.FLOOR
      if (cell.type != GC_CELL_FLOOR) {
        const bool border = px == 0 || py == 0 || px == kTile - 1 || py == kTile - 1;
        if (cell.type == GC_CELL_DELIVERY) {
          rgb[0] = rgb[1] = rgb[2] = 96.f;  // Color.DELIVERY (no border, game.py:93-95)
          blit(atlas, SP_DELIVERY, kTile, 0, px, py, rgb);
        } else {
          rgb[0] = border ? 114.f : 220.f;  // Color.COUNTER_BORDER / Color.COUNTER
          rgb[1] = border ? 93.f : 170.f;
          rgb[2] = border ? 51.f : 110.f;
          if (cell.type == GC_CELL_CUTBOARD) blit(atlas, SP_CUTBOARD, kTile, 0, px, py, rgb);
        }
      }
      if (cell.lying) blit_object(atlas, cell.lying, kTile, 0, px, py, rgb);
      if (cell.agent) {
        blit(atlas, SP_AGENT0 + cell.agent - 1, kTile, 0, px, py, rgb);
        if (cell.held) blit_object(atlas, cell.held, kTile / 2, kTile / 2, px, py, rgb);  // holding_scale 0.5
      }
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const uint32_t v = (uint32_t)__float2int_rn(fminf(fmaxf(rgb[ch], 0.f), 255.f));
        const int byte = k * 3 + ch;
        packed[byte >> 2] |= v << (8 * (byte & 3));
      }
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + ((size_t)y * W + x) * 3);  // (y*W + x)*3 is a multiple of 12
    dst[0] = packed[0];
    dst[1] = packed[1];
    dst[2] = packed[2];
  }
}

}  // namespace

extern "C" {

int gc_render(const gc_level* levels, int n_levels, const uint8_t* level_id, const uint32_t* state,
              const uint8_t* sprites, uint8_t* img, int64_t m, int n_agents, void* stream) {
  GcNavLevels lv;
  if (n_agents < 1 || n_agents > GC_MAX_AGENTS) return gc_fail(GC_E_ARG, "n_agents must be 1..4");
  if (int rc = gc_nav_levels_to_dev(levels, n_levels, &lv)) return rc;
  if (!state || !img || m < 0) return gc_fail(GC_E_ARG, "gc_render: null state/img or m < 0");
  if (!sprites) return gc_fail(GC_E_ARG, "gc_render: a sprite atlas uint8[%d][4][80][80][4] is required", 7 + 64);
  if (n_levels > 1 && !level_id) return gc_fail(GC_E_ARG, "gc_render: n_levels > 1 needs level_id");
  for (int l = 1; l < n_levels; l++)
    if (levels[l].width != levels[0].width || levels[l].height != levels[0].height)
      return gc_fail(GC_E_ARG, "gc_render: all levels of a batch must have the same size");
  if (m == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const int bands = 8;
  const unsigned grid = (unsigned)(m * bands);
  auto* s4 = reinterpret_cast<const uint4*>(state);
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* lid = n_levels > 1 ? level_id : nullptr;
  const int W = levels[0].width, H = levels[0].height;
  switch (n_agents) {
    case 1: render_kernel<1><<<grid, kThreads, 0, st>>>(lv, lid, s4, sprites, img, W, H, bands); break;
    case 2: render_kernel<2><<<grid, kThreads, 0, st>>>(lv, lid, s4, sprites, img, W, H, bands); break;
    case 3: render_kernel<3><<<grid, kThreads, 0, st>>>(lv, lid, s4, sprites, img, W, H, bands); break;
    default: render_kernel<4><<<grid, kThreads, 0, st>>>(lv, lid, s4, sprites, img, W, H, bands); break;
  }
  return gc_check_launch("gc_render");
}

}  // extern "C"
