// gc_search.cu - path B, part 2: exact level-0 subtask values V*(s) and Q(s, a).
//
// What the reference computes: E2E_BRTDP (navigation_planner/planners/e2e_brtdp.py) brackets the
// optimal cost of a deterministic shortest-path MDP - state = the whole kitchen, actions =
// get_single_actions (navigation_planner/utils.py:55-90), transitions = interact
// (utils/interact.py), cost 1 + 0.1 per moving agent (:816-826), goal = "one more goal object
// than at planning start" (:435-566), every agent outside the subtask frozen into an
// Agent-Counter and the object it holds deleted (:360-406).  BRTDP's answer depends on its
// exploration order, RNG tie-breaks and caps; this kernel computes the fixed point itself.
//
// How (single-agent pairs): with one mover the world only changes at INTERACTIONS (pick, put,
// chop, merge, deliver); between two interactions an optimal plan walks a shortest path, and
// steps that change nothing are never optimal.  So V* = 1.1 x min over interaction sequences
// of sum(walk + 1), searched by iterative-deepening A* over interaction sequences: depth <= 7,
// walks answered by a 64-bit bitboard BFS from the agent's square over floor minus frozen
// agents, pruning with an admissible bound (exact first leg, static-table later legs).  One
// thread owns one (env, pair); nothing but the 16-byte state is read from HBM and only V, Q[25]
// and a status byte are written - the search tile lives in registers / local memory.
//
// Joint (two-agent) pairs need a search over joint positions with the collision rules; that
// solver is not part of this round and such pairs return status 4.
#include "gc_device.cuh"
#include "gc_host.h"
#include "gc_nav.cuh"

namespace {

constexpr int kThreads = 64;
constexpr int kMaxDepth = 7;          // interactions on one plan
constexpr int kMaxSteps = 60;         // deepest bound tried before declaring the budget exceeded
#ifndef GC_SEARCH_NODE_BUDGET
#define GC_SEARCH_NODE_BUDGET 32768
#endif
// search nodes per (env, pair), all of its solves together.  Round 1 allowed 400 000; on 4 096-env samples of four
// levels (scripts/time_single.py) the SAME 52 problems fail at 25 000, 100 000 and 400 000 nodes with identical results
// everywhere else - they are goals the deepening can never reach, not deep plans - and the launch waits for them:
// 6.1 s at 400 000, 1.5 s at 100 000, 0.32 s at 25 000.
constexpr uint32_t kNodeBudget = GC_SEARCH_NODE_BUDGET;
constexpr int kInf = 1 << 20;

enum { ST_OK = 0, ST_AT_GOAL = 1, ST_UNREACHABLE = 2, ST_BUDGET = 3, ST_JOINT_UNSUPPORTED = 4 };

struct Ctx {
  unsigned long long floorp;   // walkable squares of the planning world (floor minus frozen agents)
  unsigned long long region;   // squares of floorp the agent can ever stand on (its connected component)
  unsigned long long putable;  // squares an object may be put on: every non-walkable, non-delivery square
  unsigned long long blocked;  // level-1 only: squares of the other agents (obstacles that hold nothing)
  unsigned long long cut, deliv;
  const uint8_t* dstat;        // static all-pairs floor distances of the level (shared memory)
  gc_subtask st;
  uint32_t nodes;
  bool over;
};

struct Plan {
  uint32_t cell;                    // the agent's square
  uint32_t slot[GC_MAX_OBJECTS];    // holder field: 1 = held by THE agent, 0 lying, 7 dead
};

__device__ __forceinline__ uint32_t sq_of(uint32_t slot) { return (slot >> 7) & 63u; }
__device__ __forceinline__ bool lying(uint32_t slot) { return (slot >> 13) == 0u; }
__device__ __forceinline__ bool held(uint32_t slot) { return (slot >> 13) == 1u; }

// single-source hop distances over ctx.floorp into dist[64] (255 = unreachable)
__device__ __forceinline__ void bfs(const Ctx& cx, uint32_t from, uint8_t* dist) {
#pragma unroll
  for (int c = 0; c < 64; c += 4) *reinterpret_cast<uint32_t*>(dist + c) = 0xFFFFFFFFu;
  unsigned long long visited = 1ull << from, frontier = visited;
  dist[from] = 0;
  for (uint32_t d = 1; frontier; d++) {
    frontier = gcnav::neighbours(frontier) & cx.floorp & ~visited;
    visited |= frontier;
    for (unsigned long long f = frontier; f; f &= f - 1) dist[__ffsll((long long)f) - 1] = (uint8_t)d;
  }
}

// the floor square from which action `dir` faces square q, or 64 if that is off the board
__device__ __forceinline__ uint32_t approach_cell(uint32_t q, int dir) {
  const bool off = (dir == 3 && (q & 7u) == 0u) || (dir == 2 && (q & 7u) == 7u) || (dir == 0 && q < 8u) ||
                   (dir == 1 && q >= 56u);
  return off ? 64u : ((q - (uint32_t)gc::action_delta((uint32_t)dir)) & 63u);
}

// fewest steps to stand next to q (walk only), from the current BFS
__device__ __forceinline__ int reach(const Ctx& cx, const uint8_t* dist, uint32_t q) {
  int best = kInf;
#pragma unroll
  for (int dir = 0; dir < 4; dir++) {
    const uint32_t f = approach_cell(q, dir);
    if (f < 64u && ((cx.floorp >> f) & 1ull) && dist[f] != 255) best = min(best, (int)dist[f]);
  }
  return best;
}

// lower bound on: walk to q1, interact (+1), walk on to q2, interact (+1).  First leg exact
// (current BFS), second leg from the level's static table (frozen agents ignored -> admissible).
__device__ __forceinline__ int two_legs(const Ctx& cx, const uint8_t* dist, uint32_t q1, uint32_t q2) {
  int best = kInf;
#pragma unroll
  for (int d1 = 0; d1 < 4; d1++) {
    const uint32_t f1 = approach_cell(q1, d1);
    if (f1 >= 64u || !((cx.floorp >> f1) & 1ull) || dist[f1] == 255) continue;
#pragma unroll
    for (int d2 = 0; d2 < 4; d2++) {
      const uint32_t f2 = approach_cell(q2, d2);
      if (f2 >= 64u || !((cx.region >> f2) & 1ull)) continue;  // outside the agent's component: never
      const uint32_t leg = cx.dstat[f1 * 64 + f2];
      if (leg == 255) continue;
      best = min(best, (int)dist[f1] + 1 + (int)leg + 1);
    }
  }
  return best;
}

// Admissible lower bound (in steps) on the plan that still has to be executed.
__device__ int heuristic(const Ctx& cx, const Plan& p, const uint8_t* dist) {
  int hand = -1;
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++)
    if (held(p.slot[k])) hand = k;
  const uint32_t hand_mask = hand >= 0 ? (p.slot[hand] & 0x7fu) : 0u;
  int best = kInf;
  if (cx.st.kind == GC_ST_CHOP || cx.st.kind == GC_ST_DELIVER) {
    const unsigned long long targets = cx.st.kind == GC_ST_CHOP ? cx.cut : cx.deliv;
    // bring the start object (mask a) to a target square
    for (int k = 0; k < GC_MAX_OBJECTS; k++) {
      if ((p.slot[k] & 0x7fu) != cx.st.a || (p.slot[k] >> 13) >= 3u) continue;  // 3 = inert, 7 = dead
      if (held(p.slot[k])) {
        for (unsigned long long t = targets; t; t &= t - 1) {
          const int r = reach(cx, dist, (uint32_t)__ffsll((long long)t) - 1u);
          if (r < kInf) best = min(best, r + 1);
        }
      } else if (!((cx.deliv >> sq_of(p.slot[k])) & 1ull)) {  // nothing leaves a Delivery square
        const int extra = hand >= 0 ? 1 : 0;                   // free the hand first
        for (unsigned long long t = targets; t; t &= t - 1) {
          const int r = two_legs(cx, dist, sq_of(p.slot[k]), (uint32_t)__ffsll((long long)t) - 1u);
          if (r < kInf) best = min(best, r + extra);
        }
      }
    }
  } else {  // Merge: one of the two parts in hand, facing the other
    for (int ka = 0; ka < GC_MAX_OBJECTS; ka++) {
      if ((p.slot[ka] & 0x7fu) != cx.st.a || (p.slot[ka] >> 13) >= 3u) continue;
      for (int kb = 0; kb < GC_MAX_OBJECTS; kb++) {
        if (kb == ka || (p.slot[kb] & 0x7fu) != cx.st.b || (p.slot[kb] >> 13) >= 3u) continue;
        const bool a_stuck = lying(p.slot[ka]) && ((cx.deliv >> sq_of(p.slot[ka])) & 1ull);
        const bool b_stuck = lying(p.slot[kb]) && ((cx.deliv >> sq_of(p.slot[kb])) & 1ull);
        if (a_stuck || b_stuck) continue;
        if (held(p.slot[ka])) {
          const int r = reach(cx, dist, sq_of(p.slot[kb]));
          if (r < kInf) best = min(best, r + 1);
        } else if (held(p.slot[kb])) {
          const int r = reach(cx, dist, sq_of(p.slot[ka]));
          if (r < kInf) best = min(best, r + 1);
        } else {
          const int extra = (hand >= 0 && hand_mask != cx.st.a && hand_mask != cx.st.b) ? 1 : 0;
          const int r = min(two_legs(cx, dist, sq_of(p.slot[ka]), sq_of(p.slot[kb])),
                            two_legs(cx, dist, sq_of(p.slot[kb]), sq_of(p.slot[ka])));
          if (r < kInf) best = min(best, r + extra);
        }
      }
    }
  }
  return best;
}

// utils/interact.py:33-89 for the planning agent facing non-floor square q.  Returns false when
// nothing changes; `made` receives the mask of an object created or delivered by the action.
__device__ __forceinline__ bool apply_interaction(const Ctx& cx, Plan& p, uint32_t q, uint32_t& made,
                                                  bool& delivered) {
  made = 0;
  delivered = false;
  int hand = -1, on = -1;
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++) {
    if (held(p.slot[k])) hand = k;
    if (lying(p.slot[k]) && sq_of(p.slot[k]) == q) on = k;
  }
  const bool is_del = (cx.deliv >> q) & 1ull, is_cut = (cx.cut >> q) & 1ull;
  if (hand >= 0) {
    const uint32_t mH = p.slot[hand] & 0x7fu;
    if (is_del) {
      if (!gc::deliverable(mH)) return false;
      p.slot[hand] = mH | (q << 7);
      made = mH;
      delivered = true;
      return true;
    }
    if (on >= 0) {
      const uint32_t mT = p.slot[on] & 0x7fu;
      if (!gc::mergeable(mH, mT)) return false;
      p.slot[hand] |= mT;
      p.slot[on] = GC_SLOT_DEAD;
      made = mH | mT;
      return true;
    }
    if (is_cut && gc::needs_chopped(mH)) {
      p.slot[hand] |= mH << 4;
      made = mH | (mH << 4);
      return true;
    }
    p.slot[hand] = mH | (q << 7);
    return true;
  }
  if (on >= 0 && !is_del) {
    p.slot[on] = (p.slot[on] & 0x7fu) | (1u << 13);
    return true;
  }
  return false;
}

__device__ __forceinline__ bool is_goal(const Ctx& cx, uint32_t made, bool delivered) {
  if (made != cx.st.goal) return false;
  return cx.st.kind == GC_ST_DELIVER ? delivered : !delivered;
}

// depth-first search over interaction sequences with f = g + h <= bound (all in steps)
template <int D>
__device__ bool dfs(Ctx& cx, const Plan& p, int g, int bound) {
  if (++cx.nodes > kNodeBudget) {
    cx.over = true;
    return false;
  }
  uint8_t dist[64];
  bfs(cx, p.cell, dist);
  const int h = heuristic(cx, p, dist);
  if (h >= kInf || g + h > bound) return false;
  uint32_t hand_mask = 0;
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++)
    if (held(p.slot[k])) hand_mask = p.slot[k] & 0x7fu;
  // squares worth facing: lying objects (pick / merge), and with something in hand also
  // Delivery squares and every empty put-able square
  unsigned long long cand = 0, occupied = 0;
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++)
    if (lying(p.slot[k])) occupied |= 1ull << sq_of(p.slot[k]);
  if (hand_mask == 0u) {
    cand = occupied & ~cx.deliv;
  } else {
    cand = (occupied & ~cx.deliv) | (cx.putable & ~occupied);
    if (gc::deliverable(hand_mask)) cand |= cx.deliv;
  }
  for (; cand; cand &= cand - 1) {
    const uint32_t q = (uint32_t)__ffsll((long long)cand) - 1u;
#pragma unroll 1
    for (int dir = 0; dir < 4; dir++) {
      const uint32_t f = approach_cell(q, dir);
      if (f >= 64u || !((cx.floorp >> f) & 1ull) || dist[f] == 255) continue;
      const int g2 = g + (int)dist[f] + 1;
      if (g2 > bound) continue;
      Plan n = p;
      uint32_t made;
      bool delivered;
      if (!apply_interaction(cx, n, q, made, delivered)) break;  // same outcome from every side
      if (is_goal(cx, made, delivered)) return true;
      n.cell = f;
      if constexpr (D + 1 < kMaxDepth) {
        if (dfs<D + 1>(cx, n, g2, bound)) return true;
        if (cx.over) return false;
      }
    }
  }
  return false;
}

// minimal number of steps from plan state p to the goal; kInf = unreachable / over budget.  `known_max`: an upper
// bound the caller can prove (kInf = none) - the deepening stops below it and returns it unsearched (the last
// iteration of an iterative deepening costs more than all the earlier ones together).
__device__ int solve(Ctx& cx, const Plan& p, int first_bound, int known_max = kInf) {
  uint8_t dist[64];
  bfs(cx, p.cell, dist);
  const int h0 = heuristic(cx, p, dist);
  if (h0 >= kInf) return kInf;
  for (int bound = max(h0, first_bound); bound <= kMaxSteps; bound++) {
    if (bound >= known_max) return known_max;
    if (dfs<0>(cx, p, 0, bound)) return bound;
    if (cx.over) return kInf;
  }
  cx.over = true;
  return kInf;
}

template <int NA, bool MULTI>
__global__ void __launch_bounds__(kThreads)
subtask_q_kernel(const __grid_constant__ GcNavLevels levels, const __grid_constant__ GcPairs pairs,
                 const uint8_t* __restrict__ level_id, const uint4* __restrict__ state, float* __restrict__ v_out,
                 float* __restrict__ q_out, uint8_t* __restrict__ status_out, int64_t n, int n_levels) {
  extern __shared__ uint8_t s_dist[];  // [n_levels][64][64] static floor distances
  for (int l = 0; l < n_levels; l++) gcnav::fill_floor_distances(levels.lv[l].floor_mask, s_dist + l * 4096);
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= n * pairs.n) return;
  const int64_t env = idx / pairs.n;
  const int pi = (int)(idx - env * pairs.n);
  const int lvl = MULTI ? level_id[env] : 0;
  const GcNavLevel& L = levels.lv[lvl];
  const int sub = pairs.p[pi][0], ai = pairs.p[pi][1], aj = pairs.p[pi][2];
  const bool level1 = pairs.p[pi][3] != 0;  // e2e_brtdp._configure_planner_level :379-381: nobody is removed
  float* q = q_out ? q_out + idx * 25 : nullptr;
  if (q)
    for (int a = 0; a < 25; a++) q[a] = NAN;  // NaN = action not offered; +inf = offered but the goal is out of reach
  float v = INFINITY;
  uint8_t status = ST_UNREACHABLE;
  if (aj != 0xFF) {
    status = ST_JOINT_UNSUPPORTED;
  } else if ((uint32_t)sub < L.n_subtasks) {
    const uint4 s = state[env];
    Ctx cx;
    cx.cut = L.cut_mask;
    cx.deliv = L.deliv_mask;
    cx.dstat = s_dist + lvl * 4096;
    cx.st = L.st[sub];
    cx.nodes = 0;
    cx.over = false;
    // planning world (e2e_brtdp.py:386-406): other agents frozen into Agent-Counters
    unsigned long long frozen = 0;
#pragma unroll
    for (int i = 0; i < NA; i++)
      if (i != ai) frozen |= 1ull << ((s.x >> (6 * i)) & 63u);
    cx.floorp = L.floor_mask & ~frozen;
    cx.blocked = level1 ? frozen : 0ull;
    cx.putable = ~cx.floorp & ~cx.deliv & ~cx.blocked;
    Plan p;
    p.cell = (s.x >> (6 * ai)) & 63u;
    // With a single mover and static obstacles the set of squares it can stand on never
    // changes, so anything not adjacent to this component is out of play for good: the bound
    // below turns infinite and the pair is reported unreachable without searching.
    cx.region = 1ull << p.cell;
    for (unsigned long long grow = cx.region; grow;) {
      grow = gcnav::neighbours(grow) & cx.floorp & ~cx.region;
      cx.region |= grow;
    }
    int goal_objects = 0;
#pragma unroll
    for (int k = 0; k < GC_MAX_OBJECTS; k++) {
      uint32_t sl = gcnav::slot_of(s, k);
      const uint32_t holder = sl >> 13;
      // level 0 deletes what the frozen agents hold (:397-399); level 1 keeps it, out of reach (holder code 3)
      if (holder >= 1u && holder <= 4u)
        sl = (holder == (uint32_t)(ai + 1)) ? ((sl & 0x7fu) | (1u << 13))
                                            : (level1 ? ((sl & 0x7fu) | (3u << 13)) : GC_SLOT_DEAD);
      p.slot[k] = sl;
      if ((sl >> 13) != 7u && (sl & 0x7fu) == cx.st.goal &&
          (cx.st.kind != GC_ST_DELIVER || (lying(sl) && ((cx.deliv >> sq_of(sl)) & 1ull))))
        goal_objects++;
    }
    // Q(start, a) for the actions get_single_actions offers (navigation_planner/utils.py:55-90)
    int best_steps = kInf;
    int q_steps[4] = {kInf, kInf, kInf, kInf};
    bool offered[4];
    // with one food of each kind a second goal object can never appear: no search then
    const int v_start = goal_objects == 0 ? solve(cx, p, 0) : kInf;
    for (int a = 0; a < 4; a++) {
      const uint32_t tgt = (p.cell + (uint32_t)gc::action_delta((uint32_t)a)) & 63u;
      offered[a] = false;
      if ((cx.blocked >> tgt) & 1ull) continue;  // another agent stands there (navigation_planner/utils.py:71)
      Plan nx = p;
      bool goal_now = false;
      // A step that one more step undoes (a walk; a pick-up or put-down that creates nothing: the object goes back
      // where it was) leaves V* within one of the start's: V*(nx) <= V*(start) + 1 needs no search to be known.
      bool undoable = true;
      if ((cx.floorp >> tgt) & 1ull) {
        nx.cell = tgt;
        offered[a] = true;
      } else {
        uint32_t made;
        bool delivered;
        const bool changed = apply_interaction(cx, nx, tgt, made, delivered);
        offered[a] = changed || ((cx.deliv >> tgt) & 1ull);  // a Delivery square may always be faced (:77-78)
        goal_now = changed && is_goal(cx, made, delivered);
        // chop / merge / deliver cannot be taken back; nor can a pick-up from a cutboard (facing the board again
        // with something choppable in hand chops it instead of putting it back)
        undoable = !changed || (made == 0u && !((cx.cut >> tgt) & 1ull));
      }
      if (!offered[a] || v_start >= kInf || cx.over) continue;
      int steps = 0;
      if (!goal_now) steps = solve(cx, nx, max(0, v_start - 1), undoable ? v_start + 1 : kInf);
      if (steps < kInf) {
        q_steps[a] = steps + 1;
        best_steps = min(best_steps, steps + 1);
      }
    }
    if (cx.over) status = ST_BUDGET;
    else if (best_steps < kInf) status = ST_OK;
    if (best_steps < kInf) v = 1.1f * (float)best_steps;
    if (q) {
      for (int a = 0; a < 4; a++)
        if (offered[a]) q[a] = q_steps[a] < kInf ? 1.1f * (float)q_steps[a] : INFINITY;
      q[4] = 1.0f + v;  // staying is always offered (:88); it costs time only and changes nothing
    }
  }
  v_out[idx] = v;
  if (status_out) status_out[idx] = status;
}

}  // namespace

extern "C" {

int gc_subtask_q(const gc_level* levels, int n_levels, const uint8_t* level_id, const uint32_t* state,
                 const uint8_t* pairs, int n_pairs, float* v, float* q, uint8_t* status, int64_t n, int n_agents,
                 void* stream) {
  GcNavLevels lv;
  GcPairs pr;
  if (n_agents < 1 || n_agents > GC_MAX_AGENTS) return gc_fail(GC_E_ARG, "n_agents must be 1..4");
  if (int rc = gc_nav_levels_to_dev(levels, n_levels, &lv)) return rc;
  if (int rc = gc_pairs_to_dev(pairs, n_pairs, n_agents, &pr)) return rc;
  if (!state || !v || n < 0) return gc_fail(GC_E_ARG, "gc_subtask_q: null state/v or n < 0");
  if (n_levels > 1 && !level_id) return gc_fail(GC_E_ARG, "gc_subtask_q: n_levels > 1 needs level_id");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const int64_t threads = n * n_pairs;
  const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
  const size_t smem = (size_t)n_levels * 4096;
  auto* s4 = reinterpret_cast<const uint4*>(state);
  cudaStream_t st = (cudaStream_t)stream;
#define GC_SQ_LAUNCH(NA_)                                                                                          \
  do {                                                                                                             \
    if (n_levels > 1) {                                                                                            \
      cudaFuncSetAttribute(subtask_q_kernel<NA_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
      subtask_q_kernel<NA_, true><<<grid, kThreads, smem, st>>>(lv, pr, level_id, s4, v, q, status, n, n_levels);  \
    } else {                                                                                                       \
      subtask_q_kernel<NA_, false><<<grid, kThreads, smem, st>>>(lv, pr, level_id, s4, v, q, status, n, n_levels); \
    }                                                                                                              \
  } while (0)
  switch (n_agents) {
    case 1: GC_SQ_LAUNCH(1); break;
    case 2: GC_SQ_LAUNCH(2); break;
    case 3: GC_SQ_LAUNCH(3); break;
    default: GC_SQ_LAUNCH(4); break;
  }
#undef GC_SQ_LAUNCH
  return gc_check_launch("gc_subtask_q");
}

}  // extern "C"
