// gc_agents.cu - what a RealAgent reads off the real env every step (utils/agent.py), for every env and agent.
//
// Two small per-env kernels that replace a few hundred tensor-op launches per step of the batched delegation loop:
//   gc_offered_actions      nav_utils.get_single_actions (navigation_planner/utils.py:55-90) on the real env: which of
//                           the four moves each agent is offered (staying always is): the square faced is free of
//                           agents and is floor / a delivery square, or a counter the agent can put its object on,
//                           pick an object from, or merge with;
//   gc_subtasks_completed   RealAgent.def_subtask_completion (utils/agent.py:286-368): the agent's subtask counts as
//                           completed when the env holds MORE objects equal to the subtask's goal than before the
//                           step (for Deliver: lying on a delivery square).
// One thread per env: 16 (or 32) bytes in, n_agents bytes out.
#include "gc_device.cuh"
#include "gc_host.h"
#include "gc_nav.cuh"

namespace {

constexpr int kThreads = 256;

// core.mergeable on content masks: at most one plate between the two, and every food involved is chopped
__device__ __forceinline__ bool can_merge(uint32_t a, uint32_t b) {
  const uint32_t u = a | b;
  return !(a & b & 8u) && (((u & 7u) & ~(u >> 4)) == 0u);
}

template <int NA>
__global__ void __launch_bounds__(kThreads)
offered_actions_kernel(const __grid_constant__ GcNavLevels levels, const uint8_t* __restrict__ level_id,
                       const uint4* __restrict__ state, uint8_t* __restrict__ offered, int64_t n) {
  const int64_t env = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (env >= n) return;
  const GcNavLevel& L = levels.lv[level_id ? level_id[env] : 0];
  const uint4 s = state[env];
  uint32_t cell[NA], hold[NA];
  unsigned long long agents = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    cell[i] = (s.x >> (6 * i)) & 63u;
    hold[i] = 0;
    agents |= 1ull << cell[i];
  }
  uint32_t slot[GC_MAX_OBJECTS];
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++) {
    slot[k] = gcnav::slot_of(s, k);
    const uint32_t holder = slot[k] >> 13;
#pragma unroll
    for (int i = 0; i < NA; i++)
      if (holder == (uint32_t)(i + 1)) hold[i] += slot[k] & 0x7fu;
  }
#pragma unroll
  for (int i = 0; i < NA; i++) {
    uint32_t bits = 0;
#pragma unroll
    for (uint32_t a = 0; a < 4; a++) {
      const uint32_t tgt = (cell[i] + (uint32_t)gc::action_delta(a)) & 63u;
      if ((agents >> tgt) & 1ull) continue;  // `if new_loc in agent_locs: continue` (:71)
      bool ok = ((L.floor_mask >> tgt) & 1ull) || ((L.deliv_mask >> tgt) & 1ull);
      if (!ok) {
        uint32_t on = 0;
#pragma unroll
        for (int k = 0; k < GC_MAX_OBJECTS; k++)
          if ((slot[k] >> 13) == 0u && ((slot[k] >> 7) & 63u) == tgt) on += slot[k] & 0x7fu;
        ok = (on == 0u && hold[i] != 0u) || (on != 0u && hold[i] == 0u) || (on != 0u && hold[i] != 0u && can_merge(hold[i], on));
      }
      bits |= (ok ? 1u : 0u) << a;
    }
    offered[env * NA + i] = (uint8_t)bits;
  }
}

__device__ __forceinline__ int goal_objects(const GcNavLevel& L, const uint4& s, const gc_subtask& st) {
  int count = 0;
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++) {
    const uint32_t sl = gcnav::slot_of(s, k), holder = sl >> 13;
    if (holder == 7u || (sl & 0x7fu) != st.goal) continue;
    if (st.kind == GC_ST_DELIVER) count += (holder == 0u && ((L.deliv_mask >> ((sl >> 7) & 63u)) & 1ull)) ? 1 : 0;
    else count += 1;
  }
  return count;
}

__global__ void __launch_bounds__(kThreads)
subtasks_completed_kernel(const __grid_constant__ GcNavLevels levels, const uint8_t* __restrict__ level_id,
                          const uint4* __restrict__ before, const uint4* __restrict__ after,
                          const uint8_t* __restrict__ subtask, uint8_t* __restrict__ completed, int64_t n, int n_agents) {
  const int64_t env = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (env >= n) return;
  const GcNavLevel& L = levels.lv[level_id ? level_id[env] : 0];
  const uint4 s0 = before[env], s1 = after[env];
  for (int i = 0; i < n_agents; i++) {
    const uint32_t sub = subtask[env * n_agents + i];
    uint8_t done = 0;
    if (sub < L.n_subtasks) {
      const gc_subtask st = L.st[sub];
      done = goal_objects(L, s1, st) > goal_objects(L, s0, st) ? 1 : 0;
    }
    completed[env * n_agents + i] = done;
  }
}

}  // namespace

extern "C" {

int gc_offered_actions(const gc_level* levels, int n_levels, const uint8_t* level_id, const uint32_t* state,
                       uint8_t* offered, int64_t n, int n_agents, void* stream) {
  GcNavLevels lv;
  if (n_agents < 1 || n_agents > GC_MAX_AGENTS) return gc_fail(GC_E_ARG, "gc_offered_actions: n_agents must be 1..4");
  if (int rc = gc_nav_levels_to_dev(levels, n_levels, &lv)) return rc;
  if (!state || !offered || n < 0) return gc_fail(GC_E_ARG, "gc_offered_actions: null state/offered or n < 0");
  if (n_levels > 1 && !level_id) return gc_fail(GC_E_ARG, "gc_offered_actions: n_levels > 1 needs level_id");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
  const uint8_t* lid = n_levels > 1 ? level_id : nullptr;
  auto* s4 = reinterpret_cast<const uint4*>(state);
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_agents) {
    case 1: offered_actions_kernel<1><<<grid, kThreads, 0, st>>>(lv, lid, s4, offered, n); break;
    case 2: offered_actions_kernel<2><<<grid, kThreads, 0, st>>>(lv, lid, s4, offered, n); break;
    case 3: offered_actions_kernel<3><<<grid, kThreads, 0, st>>>(lv, lid, s4, offered, n); break;
    default: offered_actions_kernel<4><<<grid, kThreads, 0, st>>>(lv, lid, s4, offered, n); break;
  }
  return gc_check_launch("gc_offered_actions");
}

int gc_subtasks_completed(const gc_level* levels, int n_levels, const uint8_t* level_id, const uint32_t* before,
                          const uint32_t* after, const uint8_t* subtask, uint8_t* completed, int64_t n, int n_agents,
                          void* stream) {
  GcNavLevels lv;
  if (n_agents < 1 || n_agents > GC_MAX_AGENTS) return gc_fail(GC_E_ARG, "gc_subtasks_completed: n_agents must be 1..4");
  if (int rc = gc_nav_levels_to_dev(levels, n_levels, &lv)) return rc;
  if (!before || !after || !subtask || !completed || n < 0)
    return gc_fail(GC_E_ARG, "gc_subtasks_completed: null array or n < 0");
  if (n_levels > 1 && !level_id) return gc_fail(GC_E_ARG, "gc_subtasks_completed: n_levels > 1 needs level_id");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
  subtasks_completed_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      lv, n_levels > 1 ? level_id : nullptr, reinterpret_cast<const uint4*>(before), reinterpret_cast<const uint4*>(after),
      subtask, completed, n, n_agents);
  return gc_check_launch("gc_subtasks_completed");
}

}  // extern "C"
