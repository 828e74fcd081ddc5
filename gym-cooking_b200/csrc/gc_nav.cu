// gc_nav.cu - path B, part 1: the navigation planner's distance heuristic.
//
// gc_lower_bound restates env.get_lower_bound_for_subtask_given_objs (envs/
// overcooked_environment.py:594-664) = holding penalty + World.get_lower_bound_between
// (utils/world.py:115-283) on the packed state.  The reference answers every distance with a
// networkx bidirectional BFS on World.reachability_graph (world.py:67-107); that graph is
// static per level (it is never rebuilt, world.py:38), so each CTA first fills an all-pairs
// floor-distance table in shared memory with 64-bit bitboard BFS (one source square per
// thread, <= 25 wavefronts of shift/and/or) and every (env, pair) thread then only does
// table look-ups: D(square a approached from floor f, ...) = D_floor(f, f') + 1 per collidable end.
#include "gc_device.cuh"
#include "gc_host.h"
#include "gc_nav.cuh"

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------
// one thread per (env, pair)
// ---------------------------------------------------------------------------------------
template <int NA, bool MULTI>
__global__ void __launch_bounds__(kThreads)
lower_bound_kernel(const __grid_constant__ GcNavLevels levels, const __grid_constant__ GcPairs pairs,
                   const uint8_t* __restrict__ level_id, const uint4* __restrict__ state,
                   float* __restrict__ lb, int64_t n, int n_levels) {
  extern __shared__ uint8_t s_dist[];  // [n_levels][64][64]
  for (int l = 0; l < n_levels; l++) gcnav::fill_floor_distances(levels.lv[l].floor_mask, s_dist + l * 4096);
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= n * pairs.n) return;
  const int64_t env = idx / pairs.n;
  const int p = (int)(idx - env * pairs.n);
  const int lvl = MULTI ? level_id[env] : 0;
  const GcNavLevel& L = levels.lv[lvl];
  const uint4 s = state[env];
  lb[idx] = gcnav::lower_bound<NA>(L, s_dist + lvl * 4096, s, pairs.p[p][0], pairs.p[p][1], pairs.p[p][2]);
}

}  // namespace

int gc_nav_levels_to_dev(const gc_level* levels, int n_levels, GcNavLevels* out) {
  if (!levels || n_levels < 1 || n_levels > GC_MAX_LEVELS) return gc_fail(GC_E_ARG, "levels: need 1..%d", GC_MAX_LEVELS);
  memset(out, 0, sizeof(*out));
  for (int l = 0; l < n_levels; l++) {
    const gc_level& s = levels[l];
    GcNavLevel& d = out->lv[l];
    for (int c = 0; c < GC_MAX_CELLS; c++) {
      const unsigned long long b = 1ull << c;
      if (s.cell_type[c] == GC_CELL_FLOOR) d.floor_mask |= b;
      if (s.cell_type[c] == GC_CELL_CUTBOARD) d.cut_mask |= b;
      if (s.cell_type[c] == GC_CELL_DELIVERY) d.deliv_mask |= b;
    }
    d.perimeter = (uint32_t)(2 * (s.width + s.height));
    if (s.n_subtasks < 0 || s.n_subtasks > GC_MAX_SUBTASKS) return gc_fail(GC_E_ARG, "level %d: bad n_subtasks", l);
    d.n_subtasks = (uint32_t)s.n_subtasks;
    for (int k = 0; k < s.n_subtasks; k++) d.st[k] = s.subtask[k];
  }
  return GC_OK;
}

int gc_pairs_to_dev(const uint8_t* pairs, int n_pairs, int n_agents, GcPairs* out) {
  if (!pairs || n_pairs < 1) return gc_fail(GC_E_ARG, "pairs: need at least one (subtask, agent, agent) triple");
  if (n_pairs > GC_MAX_PAIRS) return gc_fail(GC_E_LIMIT, "pairs: at most %d per call", GC_MAX_PAIRS);
  memset(out, 0, sizeof(*out));
  out->n = n_pairs;
  for (int k = 0; k < n_pairs; k++) {
    const uint8_t s = pairs[3 * k] & 0x7F, i = pairs[3 * k + 1], j = pairs[3 * k + 2];
    out->p[k][3] = pairs[3 * k] >> 7;  // 1 = level-1 planning world (other agents are plain obstacles)
    if (s >= GC_MAX_SUBTASKS || i >= n_agents || (j != 0xFF && (j >= n_agents || j == i)))
      return gc_fail(GC_E_ARG, "pair %d = (%d, %d, %d) is out of range", k, s, i, j);
    out->p[k][0] = s;
    out->p[k][1] = (j != 0xFF && j < i) ? j : i;  // sim_agents order (env:641)
    out->p[k][2] = (j != 0xFF && j < i) ? i : j;
  }
  return GC_OK;
}

extern "C" {

int gc_lower_bound(const gc_level* levels, int n_levels, const uint8_t* level_id, const uint32_t* state,
                   const uint8_t* pairs, int n_pairs, float* lb, int64_t n, int n_agents, void* stream) {
  GcNavLevels lv;
  GcPairs pr;
  if (n_agents < 1 || n_agents > GC_MAX_AGENTS) return gc_fail(GC_E_ARG, "n_agents must be 1..4");
  if (int rc = gc_nav_levels_to_dev(levels, n_levels, &lv)) return rc;
  if (int rc = gc_pairs_to_dev(pairs, n_pairs, n_agents, &pr)) return rc;
  if (!state || !lb || n < 0) return gc_fail(GC_E_ARG, "gc_lower_bound: null state/lb or n < 0");
  if (n_levels > 1 && !level_id) return gc_fail(GC_E_ARG, "gc_lower_bound: n_levels > 1 needs level_id");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const int64_t threads = n * n_pairs;
  const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
  const size_t smem = (size_t)n_levels * 4096;
  auto* s4 = reinterpret_cast<const uint4*>(state);
  cudaStream_t st = (cudaStream_t)stream;
#define GC_LB_LAUNCH(NA_)                                                                                   \
  do {                                                                                                      \
    if (n_levels > 1) {                                                                                     \
      cudaFuncSetAttribute(lower_bound_kernel<NA_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      lower_bound_kernel<NA_, true><<<grid, kThreads, smem, st>>>(lv, pr, level_id, s4, lb, n, n_levels);   \
    } else {                                                                                                \
      lower_bound_kernel<NA_, false><<<grid, kThreads, smem, st>>>(lv, pr, level_id, s4, lb, n, n_levels);  \
    }                                                                                                       \
  } while (0)
  switch (n_agents) {
    case 1: GC_LB_LAUNCH(1); break;
    case 2: GC_LB_LAUNCH(2); break;
    case 3: GC_LB_LAUNCH(3); break;
    default: GC_LB_LAUNCH(4); break;
  }
#undef GC_LB_LAUNCH
  return gc_check_launch("gc_lower_bound");
}

}  // extern "C"
