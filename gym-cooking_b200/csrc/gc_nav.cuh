// gc_nav.cuh - device structures and helpers shared by the planner kernels (path B).
#pragma once
#include <string.h>

#include "gc_device.cuh"

struct GcNavLevel {
  unsigned long long floor_mask, cut_mask, deliv_mask;
  uint32_t perimeter;   // 2 * (width + height), env:198
  uint32_t n_subtasks;
  gc_subtask st[GC_MAX_SUBTASKS];
};
static_assert(sizeof(GcNavLevel) == 32 + 4 * GC_MAX_SUBTASKS, "GcNavLevel layout");

struct GcNavLevels {
  GcNavLevel lv[GC_MAX_LEVELS];
};

struct GcPairs {
  uint8_t p[GC_MAX_PAIRS][4];  // subtask index, agent i, agent j (0xFF = single), pad; i < j
  int n;
};

int gc_nav_levels_to_dev(const gc_level* levels, int n_levels, GcNavLevels* out);
int gc_pairs_to_dev(const uint8_t* pairs, int n_pairs, int n_agents, GcPairs* out);

namespace gcnav {

constexpr unsigned long long kCol0 = 0x0101010101010101ull;  // x == 0 of every row
constexpr unsigned long long kCol7 = 0x8080808080808080ull;  // x == 7
constexpr uint32_t kFar = 255;                               // "no path" in the distance table

// squares one step away from any square of `f` (4-neighbourhood on the 8x8 board)
__device__ __forceinline__ unsigned long long neighbours(unsigned long long f) {
  return ((f & ~kCol7) << 1) | ((f & ~kCol0) >> 1) | (f << 8) | (f >> 8);
}

// All-pairs hop distance over `walkable` squares into dist[64][64] (bytes, kFar = unreachable
// or not walkable).  Threads 0..63 of the CTA each own one source square; callers
// __syncthreads() afterwards.  This is World.reachability_graph restricted to its (square,
// (0,0)) nodes: a collidable node hangs off exactly one floor square (world.py:93-98), so
// every graph distance the reference asks for is a floor distance plus one per collidable end.
__device__ __forceinline__ void fill_floor_distances(unsigned long long walkable, uint8_t* dist) {
  for (int s = threadIdx.x; s < 64; s += blockDim.x) {
    uint8_t* row = dist + s * 64;
    for (int c = 0; c < 64; c += 4) *reinterpret_cast<uint32_t*>(row + c) = 0xFFFFFFFFu;
    const unsigned long long src = 1ull << s;
    if (!(walkable & src)) continue;
    unsigned long long visited = src, frontier = src;
    row[s] = 0;
    for (uint32_t d = 1; frontier; d++) {
      frontier = neighbours(frontier) & walkable & ~visited;
      visited |= frontier;
      for (unsigned long long f = frontier; f; f &= f - 1) row[__ffsll((long long)f) - 1] = (uint8_t)d;
    }
  }
}

// working slot (mask | cell << 7 | holder << 13) of object k: the byte planes are decoded in gc_device.cuh
__device__ __forceinline__ uint32_t slot_of(const uint4& s, int k) { return gc::slot_of(s, k); }

// distance from floor square `from` to graph node (square c, approach a); a == 4 means c is
// itself a floor square.  kFar where networkx would raise (node missing / no path).
__device__ __forceinline__ uint32_t node_dist(const uint8_t* dist, unsigned long long floor_mask, uint32_t from,
                                              uint32_t c, int a) {
  if (a == 4) return dist[from * 64 + c];
  const uint32_t f = (c + (uint32_t)gc::action_delta((uint32_t)a)) & 63u;
  // the approach square must be a floor neighbour inside the board
  const bool wraps = ((a == 2) && (c & 7u) == 0u) || ((a == 3) && (c & 7u) == 7u) ||
                     ((a == 0) && c >= 56u) || ((a == 1) && c < 8u);
  if (wraps || !((floor_mask >> f) & 1ull)) return kFar;
  const uint32_t d = dist[from * 64 + f];
  return d == kFar ? kFar : d + 1u;
}

// World.get_lower_bound_between_helper (utils/world.py:148-264) for one (A, B) location pair
template <int NAG>
__device__ __forceinline__ float lb_helper(const GcNavLevel& L, const uint8_t* dist, int kind, const uint32_t* ag,
                                           uint32_t A, uint32_t B) {
  const float perimeter = (float)L.perimeter;
  float lower = perimeter + 1.0f;
  const bool a_coll = !((L.floor_mask >> A) & 1ull), b_coll = !((L.floor_mask >> B) & 1ull);
  const int na0 = a_coll ? 0 : 4, na1 = a_coll ? 4 : 5, nb0 = b_coll ? 0 : 4, nb1 = b_coll ? 4 : 5;
  for (int na = na0; na < na1; na++) {
    // floor square the A-node hangs off (or A itself)
    const uint32_t fa = a_coll ? ((A + (uint32_t)gc::action_delta((uint32_t)na)) & 63u) : A;
    for (int nb = nb0; nb < nb1; nb++) {
      float bound;
      if (NAG == 1) {  // :178-189
        const uint32_t b1 = node_dist(dist, L.floor_mask, ag[0], A, na);
        if (b1 == kFar) continue;
        uint32_t b2;
        if (A == B && na == nb) {
          b2 = 0;
        } else {
          b2 = node_dist(dist, L.floor_mask, fa, B, nb);
          if (b2 == kFar) continue;
          b2 += a_coll ? 1u : 0u;
        }
        bound = (float)b1 + (float)b2 - 1.0f;
      } else {  // :193-258 - a missing node / path costs `perimeter`, the pair is still scored
        uint32_t d;
        const float b1A = (d = node_dist(dist, L.floor_mask, ag[0], A, na)) == kFar ? perimeter : (float)d;
        const float b2A = (d = node_dist(dist, L.floor_mask, ag[1], A, na)) == kFar ? perimeter : (float)d;
        const float b1B = (d = node_dist(dist, L.floor_mask, ag[0], B, nb)) == kFar ? perimeter : (float)d;
        const float b2B = (d = node_dist(dist, L.floor_mask, ag[1], B, nb)) == kFar ? perimeter : (float)d;
        float minA = fminf(b1A, b2A), minB = fminf(b1B, b2B);
        const int dx = (int)(A & 7u) - (int)(B & 7u), dy = (int)(A >> 3) - (int)(B >> 3);
        const float between = (float)(abs(dx) + abs(dy));  // manhattan_dist, navigation_planner/utils.py:95-98
        if (kind == GC_ST_MERGE) {
          if ((b1A == minA && b1B == minB) || (b2A == minA && b2B == minB)) {  // check_bound :266-283
            minA *= 2.0f;
            minB *= 2.0f;
          }
          bound = fmaxf(minA, minB) + (between - 1.0f) * 0.5f;
        } else {
          bound = minA + between - 1.0f;
        }
      }
      lower = fminf(lower, bound);
    }
  }
  return fmaxf(1.0f, lower);  // :264
}

// Locations of objects equal to `mask`: lying anywhere, or held by one of the subtask agents
// (env.get_AB_locs_given_objs :480-589).  Returns a count, fills cells[].
template <int NA>
__device__ __forceinline__ int object_locs(const uint4& s, uint32_t mask, int ai, int aj, uint32_t* cells) {
  int n = 0;
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++) {
    const uint32_t sl = slot_of(s, k);
    if ((sl >> 13) == 0u && (sl & 0x7fu) == mask) cells[n++] = (sl >> 7) & 63u;
  }
#pragma unroll
  for (int i = 0; i < NA; i++) {  // sim_agents order
    if (i != ai && i != aj) continue;
#pragma unroll
    for (int k = 0; k < GC_MAX_OBJECTS; k++) {
      const uint32_t sl = slot_of(s, k);
      if ((sl >> 13) == (uint32_t)(i + 1) && (sl & 0x7fu) == mask) cells[n++] = (s.x >> (6 * i)) & 63u;
    }
  }
  return n;
}

__device__ __forceinline__ int board_cells(unsigned long long m, uint32_t* cells) {
  int n = 0;
  for (; m && n < 8; m &= m - 1) cells[n++] = (uint32_t)__ffsll((long long)m) - 1u;
  return n;
}

// env.get_lower_bound_for_subtask_given_objs :594-664.  aj == 0xFF: single agent.
template <int NA>
__device__ __forceinline__ float lower_bound(const GcNavLevel& L, const uint8_t* dist, const uint4& s, int sub,
                                             int ai, int aj) {
  const float not_doable = (float)L.perimeter + 1.0f;
  if ((uint32_t)sub >= L.n_subtasks) return not_doable;
  const gc_subtask st = L.st[sub];
  const bool joint = aj != 0xFF;
  if (!joint) aj = -1;
  // holding penalty :612-638
  float penalty = 0.0f;
  if (st.kind != GC_ST_MERGE) {
#pragma unroll
    for (int k = 0; k < GC_MAX_OBJECTS; k++) {
      const uint32_t sl = slot_of(s, k);
      const int holder = (int)(sl >> 13) - 1;
      if (holder >= 0 && holder < NA && (holder == ai || holder == aj)) {
        const uint32_t m = sl & 0x7fu;
        if (m != st.a && m != st.goal) penalty += 1.0f;
      }
    }
  }
  penalty = fminf(penalty, 1.0f);
  uint32_t ag[2];
  ag[0] = (s.x >> (6 * ai)) & 63u;
  ag[1] = joint ? (s.x >> (6 * aj)) & 63u : 0u;
  uint32_t A[10], B[10];
  int nA = 0, nB = 0;
  if (st.kind == GC_ST_CHOP) {  // :512-527
    nA = object_locs<NA>(s, st.a, ai, aj, A);
    nB = board_cells(L.cut_mask, B);
  } else if (st.kind == GC_ST_DELIVER) {  // :535-548
    nB = board_cells(L.deliv_mask, B);
    uint32_t tmp[10];
    const int n0 = object_locs<NA>(s, st.a, ai, aj, tmp);
    for (int k = 0; k < n0; k++)
      if (!((L.deliv_mask >> tmp[k]) & 1ull)) A[nA++] = tmp[k];
  } else if (st.kind == GC_ST_MERGE) {  // :573-584
    nA = object_locs<NA>(s, st.a, ai, aj, A);
    nB = object_locs<NA>(s, st.b, ai, aj, B);
  }
  float lower = not_doable;  // world.py:133
  for (int a = 0; a < nA; a++)
    for (int b = 0; b < nB; b++)
      lower = fminf(lower, joint ? lb_helper<2>(L, dist, st.kind, ag, A[a], B[b])
                                 : lb_helper<1>(L, dist, st.kind, ag, A[a], B[b]));
  return lower + penalty;
}

}  // namespace gcnav
