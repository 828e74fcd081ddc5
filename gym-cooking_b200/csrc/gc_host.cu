// gc_host.cu - host half of the C-ABI: errors, device check, level loader and table conversion.
#include "gc_host.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "gc_device.cuh"
#include "gc_step2.cuh"

static_assert(sizeof(gc_level) == 256, "gc_level is part of the ABI: 256 bytes");

namespace {
thread_local char g_err[512] = "";
}

int gc_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int gc_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gc_fail(GC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return GC_OK;
}

int gc_require_device() {
  static bool seen = false;  // a positive answer does not change; a negative one is asked again
  if (seen) return GC_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    cudaGetLastError();
    return gc_fail(GC_E_CUDA, "no usable CUDA device (%s); libgymcook has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  seen = true;
  return GC_OK;
}

int gc_levels_to_dev(const gc_level* levels, int n_levels, int n_agents, GcLevelsDev* out, int* max_objs) {
  if (!levels || n_levels < 1) return gc_fail(GC_E_ARG, "levels: need at least one level");
  if (n_levels > GC_MAX_LEVELS) return gc_fail(GC_E_LIMIT, "levels: at most %d per call", GC_MAX_LEVELS);
  if (n_agents < 1 || n_agents > GC_MAX_AGENTS) return gc_fail(GC_E_ARG, "n_agents must be 1..4, got %d", n_agents);
  memset(out, 0, sizeof(*out));
  *max_objs = 0;
  for (int l = 0; l < n_levels; l++) {
    const gc_level& s = levels[l];
    GcLevelDev& d = out->lv[l];
    if (s.n_agent_starts < n_agents)
      return gc_fail(GC_E_ARG, "level %d has %d agent start lines, %d agents requested", l, s.n_agent_starts, n_agents);
    if (s.n_objects < 0 || s.n_objects > GC_MAX_OBJECTS || s.n_goals < 1 || s.n_goals > GC_MAX_GOALS ||
        s.delivery_cell < 0 || s.delivery_cell >= GC_MAX_CELLS || s.max_timesteps < 0 || s.max_timesteps > 127)
      return gc_fail(GC_E_ARG, "level %d: table out of range (objects %d goals %d delivery %d max_t %d)", l,
                     s.n_objects, s.n_goals, s.delivery_cell, s.max_timesteps);
    for (int c = 0; c < GC_MAX_CELLS; c++) {
      const unsigned long long b = 1ull << c;
      if (s.cell_type[c] == GC_CELL_FLOOR) d.floor_mask |= b;
      if (s.cell_type[c] == GC_CELL_CUTBOARD) d.cut_mask |= b;
      if (s.cell_type[c] == GC_CELL_DELIVERY) d.deliv_mask |= b;
    }
    for (int g = 0; g < GC_MAX_GOALS; g++)
      d.goal_slot[g] = (uint32_t)s.goal_mask[g < s.n_goals ? g : 0] | ((uint32_t)s.delivery_cell << 7);
    d.n_goals = (uint32_t)s.n_goals;
    d.max_t = (uint32_t)s.max_timesteps;
    gcs2::initial_state(s, n_agents, d.init);
    if (s.n_objects > *max_objs) *max_objs = s.n_objects;
  }
  return GC_OK;
}

extern "C" {

int gc_version(void) { return GC_ABI_VERSION; }

const char* gc_last_error(void) { return g_err; }

int gc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// Level text -> tables.  Follows OvercookedEnvironment.load_level (env:130-198): three phases
// separated by blank lines; map characters via RepToClass (utils/core.py:372-381), unknown
// characters become Floor (env:170-173); recipes by class name (env:182); agent "x y" lines.
int gc_level_parse(const char* txt, int len, int max_timesteps, gc_level* out) {
  if (!txt || !out || len < 0) return gc_fail(GC_E_ARG, "gc_level_parse: null argument");
  if (max_timesteps < 0 || max_timesteps > 127)
    return gc_fail(GC_E_LIMIT, "max_num_timesteps %d outside 0..127 (t is a 7-bit field)", max_timesteps);
  memset(out, 0, sizeof(*out));
  memset(out->cell_type, GC_CELL_COUNTER, sizeof(out->cell_type));
  for (int k = 0; k < GC_MAX_OBJECTS; k++) out->object_init[k] = GC_SLOT_DEAD;
  out->delivery_cell = -1;
  out->max_timesteps = max_timesteps;
  int phase = 1, row = 0, width = 0, kinds_seen = 0;
  int pos = 0;
  while (pos < len) {
    int end = pos;
    while (end < len && txt[end] != '\n') end++;
    const char* line = txt + pos;
    const int n = end - pos;
    if (n == 0) {
      phase++;
    } else if (phase == 1) {
      if (row >= GC_GRID_STRIDE || n > GC_GRID_STRIDE)
        return gc_fail(GC_E_LIMIT, "map larger than 8x8 (row %d has %d squares)", row, n);
      if (row > 0 && n != width) return gc_fail(GC_E_PARSE, "map row %d has %d squares, expected %d", row, n, width);
      width = n;
      for (int x = 0; x < n; x++) {
        const int cell = row * GC_GRID_STRIDE + x;
        uint8_t mask = 0;
        switch (line[x]) {
          case 't': mask = GC_M_TOMATO; break;
          case 'l': mask = GC_M_LETTUCE; break;
          case 'o': mask = GC_M_ONION; break;
          case 'p': mask = GC_M_PLATE; break;
          case '-': out->cell_type[cell] = GC_CELL_COUNTER; break;
          case '/': out->cell_type[cell] = GC_CELL_CUTBOARD; break;
          case '*':
            out->cell_type[cell] = GC_CELL_DELIVERY;
            if (out->delivery_cell < 0) out->delivery_cell = cell;  // env.done uses the first one (env:349)
            break;
          default: out->cell_type[cell] = GC_CELL_FLOOR; break;
        }
        if (mask) {  // an object lying on a Counter (env:158-165)
          out->cell_type[cell] = GC_CELL_COUNTER;
          if (out->n_objects >= GC_MAX_OBJECTS) return gc_fail(GC_E_LIMIT, "more than %d objects", GC_MAX_OBJECTS);
          if (mask != GC_M_PLATE) {
            if (kinds_seen & mask) return gc_fail(GC_E_LIMIT, "more than one food of the same kind is not supported");
            kinds_seen |= mask;
          }
          out->object_init[out->n_objects++] = (uint16_t)(mask | (cell << 7));
        }
      }
      row++;
    } else if (phase == 2) {
      struct { const char* name; uint8_t goal; uint8_t code; } static const recipes[] = {
          {"SimpleTomato", GC_M_PLATE | GC_M_TOMATO | GC_M_CHOP_T, 1},
          {"SimpleLettuce", GC_M_PLATE | GC_M_LETTUCE | GC_M_CHOP_L, 2},
          {"Salad", GC_M_PLATE | GC_M_TOMATO | GC_M_LETTUCE | GC_M_CHOP_T | GC_M_CHOP_L, 3},
          {"OnionSalad", 0x7f, 4},
      };
      int found = -1;
      for (int r = 0; r < 4; r++)
        if ((int)strlen(recipes[r].name) == n && !strncmp(recipes[r].name, line, n)) found = r;
      if (found < 0) return gc_fail(GC_E_PARSE, "unknown recipe class '%.*s'", n, line);
      if (out->n_goals >= GC_MAX_GOALS) return gc_fail(GC_E_LIMIT, "more than %d recipes", GC_MAX_GOALS);
      out->goal_mask[out->n_goals] = recipes[found].goal;
      out->recipe_code[out->n_goals] = recipes[found].code;
      out->n_goals++;
    } else if (phase == 3) {
      if (out->n_agent_starts < GC_MAX_AGENTS) {
        int x = -1, y = -1;
        char buf[32];
        const int m = n < 31 ? n : 31;
        memcpy(buf, line, m);
        buf[m] = 0;
        if (sscanf(buf, "%d %d", &x, &y) != 2) return gc_fail(GC_E_PARSE, "bad agent line '%s'", buf);
        if (x < 0 || y < 0 || x >= GC_GRID_STRIDE || y >= GC_GRID_STRIDE)
          return gc_fail(GC_E_PARSE, "agent start (%d,%d) outside the map", x, y);
        out->agent_cell[out->n_agent_starts++] = (uint8_t)(y * GC_GRID_STRIDE + x);
      }
    }
    pos = end + 1;
  }
  out->width = width;
  out->height = row;
  if (width < 3 || row < 3) return gc_fail(GC_E_PARSE, "map is %dx%d, need at least 3x3", width, row);
  if (out->n_goals < 1) return gc_fail(GC_E_PARSE, "no recipe: done() needs a Deliver subtask (env:338)");
  if (out->delivery_cell < 0) return gc_fail(GC_E_PARSE, "no Delivery square '*' (env:349)");
  // supported envelope: the outer ring has no floor, so agents never leave the map
  for (int y = 0; y < row; y++)
    for (int x = 0; x < width; x++)
      if ((x == 0 || y == 0 || x == width - 1 || y == row - 1) &&
          out->cell_type[y * GC_GRID_STRIDE + x] == GC_CELL_FLOOR)
        return gc_fail(GC_E_PARSE, "floor on the outer ring at (%d,%d) is not supported", x, y);
  for (int i = 0; i < out->n_agent_starts; i++) {
    const int c = out->agent_cell[i];
    if ((c & 7) >= width || (c >> 3) >= row || out->cell_type[c] != GC_CELL_FLOOR)
      return gc_fail(GC_E_PARSE, "agent %d does not start on a floor square", i + 1);
  }
  return GC_OK;
}

int gc_level_set_subtasks(gc_level* lvl, const gc_subtask* subtasks, int n) {
  if (!lvl || (!subtasks && n > 0) || n < 0) return gc_fail(GC_E_ARG, "gc_level_set_subtasks: bad arguments");
  if (n > GC_MAX_SUBTASKS) return gc_fail(GC_E_LIMIT, "more than %d subtasks", GC_MAX_SUBTASKS);
  memset(lvl->subtask, 0, sizeof(lvl->subtask));
  for (int k = 0; k < n; k++) lvl->subtask[k] = subtasks[k];
  lvl->n_subtasks = n;
  return GC_OK;
}

}  // extern "C"
