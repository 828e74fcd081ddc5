// gc_joint.cu - path B, part 3: exact level-0 values of JOINT (two-agent) subtask pairs.
//
// Same MDP as gc_search.cu, but with two movers the interaction-sequence trick no longer applies
// (simultaneous moves, the pairwise collision rule of env.is_collision :671-718, hand-overs across
// counters), so this kernel runs the real thing: uniform-cost search over full planning states
// with the reference's own transition rules -
//   actions   get_single_actions per agent (navigation_planner/utils.py:55-90), joint product
//             filtered by is_collision (e2e_brtdp.get_actions :151-206),
//   T         interact(agent_1) then interact(agent_2) (e2e_brtdp.T :132-143),
//   cost      1 + 0.1 per moving agent (:816-826)  -> integer tenths 10 / 11 / 12,
//   goal      one more goal object than at planning start (:435-566),
//   level 0   other agents frozen into Agent-Counters, their held object deleted (:360-406).
// One CTA owns one search = one (env, pair, root joint action): Dial's algorithm over a ring of
// 16 cost buckets, visited set = open-addressing hash table in a caller-provided global scratch
// arena keyed by an injective 64-bit packing of the planning state.  Q(start, a) = cost(a) +
// V*(T(start, a)); V*(start) = min_a Q.  Searches are budgeted (kSlots states); a search that
// exceeds its budget is reported (status 3), never guessed.
//
// Two kernels share those rules.  joint_tree_kernel (first) runs ONE search per (env, pair): a
// forward Dial pass from the start state that records every generated edge in per-state
// predecessor lists, out to cost V* + kSlack (or until the reachable space is exhausted), then a
// backward Dial pass from the goal states over the recorded edges, which yields the exact
// cost-to-go of every state whose optimal plan stays inside the explored ball - in particular of
// the start's successors, i.e. Q(start, a) for (almost) all 25 joint actions at once.  An action
// whose Q is not proven by that pass (Q above the explored radius, e.g. after an irreversible
// move) is handed to joint_q_kernel, the per-action search described above.
#include <stdlib.h>

#include "gc_device.cuh"
#include "gc_host.h"
#include "gc_nav.cuh"

namespace {

constexpr int kThreads = 128;
constexpr uint32_t kSlots = 1u << 16;        // hash capacity per search
constexpr uint32_t kMaxStates = 48 * 1024;   // state budget per search (load factor 0.75)
constexpr uint32_t kBucketCap = 16 * 1024;   // entries per cost bucket
constexpr int kBuckets = 16;                 // ring (edge costs 10..12 < 16)
constexpr int kMaxCost = 600;                // 60.0
constexpr unsigned long long kEmpty = ~0ull;
constexpr uint32_t kInfCost = 0xFFFFFFFFu;

enum { ST_OK = 0, ST_UNREACHABLE = 2, ST_BUDGET = 3, ST_UNSUPPORTED = 4 };

// tree search (one per problem)
constexpr uint32_t kSlots2 = 1u << 17;
#ifndef GC_JOINT_MAX_STATES
#define GC_JOINT_MAX_STATES (96 * 1024)
#endif
#ifndef GC_JOINT_MAX_SLACK
#define GC_JOINT_MAX_SLACK 48
#endif
#ifndef GC_JOINT_CTAS_PER_SM_DEFAULT
#define GC_JOINT_CTAS_PER_SM_DEFAULT 8
#endif
constexpr uint32_t kMaxStates2 = GC_JOINT_MAX_STATES;   // budget of a per-action A* (and the arenas' capacity)
#ifndef GC_JOINT_TREE_STATES
#define GC_JOINT_TREE_STATES (32 * 1024)
#endif
#ifndef GC_JOINT_WIDEN_STATES
#define GC_JOINT_WIDEN_STATES (32 * 1024)
#endif
constexpr unsigned long long kSkipUnit = ~0ull - 1ull;     // a slot of the per-action work list nobody filled
constexpr uint32_t kTreeStates = GC_JOINT_TREE_STATES;    // the tree search's budget BEYOND its first goal: what it cannot prove inside it
constexpr uint32_t kWidenStates = GC_JOINT_WIDEN_STATES;
constexpr uint32_t kSeedStates = 16 * 1024;                // room the seeded per-action searches of a problem may add  // is cheaper to prove action by action; no widening beyond this
constexpr int64_t kWideProblems = INT64_MAX;  // every launch: see the note at the launch site
// one search per CTA: large batches run 4 CTAs of 128 threads per SM (throughput); small ones
// (the delegation loop solving the few states it has not seen yet) run 2 CTAs of 512 threads so
// that more hash / edge atomics are in flight per search (latency)
constexpr int kTreeThreadsWide = 512;
constexpr int kRing = 64;                    // A* keys may jump by an edge (<= 12) plus a heuristic change (< 52)
constexpr uint32_t kRingCap = 16 * 1024;     // entries per key bucket
constexpr uint32_t kPool = 1u << 20;         // predecessor-list nodes
constexpr uint32_t kGoalCap = 4096;
constexpr uint32_t kTouchedCap = kMaxStates2 + 8192;
constexpr uint32_t kNil = 0xFFFFFFFFu;
constexpr int kSlack = 12;                   // first pass explores keys up to V* + 1.2, then widens by
constexpr int kSlackStep = 12;               // 1.2 per round while offered actions remain unproven,
constexpr int kMaxSlack = GC_JOINT_MAX_SLACK;                // up to V* + 4.8; what is still open goes to the per-action search

struct Arena2 {
  unsigned long long keys[kSlots2];
  uint4 states[kSlots2];
  uint32_t gcost[kSlots2];  // forward: 2*cost (+1 once settled)
  uint32_t val[kSlots2];    // backward: cost-to-go in tenths
  uint32_t head[kSlots2];   // predecessor list
  uint32_t bucket[kRing][kRingCap];      // forward open list: slot | g << 17, by key f
  uint32_t bbucket[kBuckets][kRingCap];  // backward pass: slots by cost-to-go
  uint2 pool[kPool];        // .x = predecessor slot | (edge cost - 10) << 30, .y = next node
  uint32_t goals[kGoalCap];
  uint32_t touched[kTouchedCap];  // slots inserted by the current search: cleaning them beats clearing 2^17 slots
};

// per-CTA scratch layout
struct Arena {
  unsigned long long keys[kSlots];
  uint4 states[kSlots];
  uint32_t gcost[kSlots];  // 2*cost (+1 once settled)
  uint32_t bucket[kBuckets][kBucketCap];
};

struct World {
  unsigned long long floorp, nonfloor, cut, deliv;
  unsigned long long blocked;  // level-1 only: squares of the other agents (cannot be faced or entered)
  uint32_t goal_mask, goal_kind, a_mask, b_mask;
};

// planning state: two agents, object slots with holder 1 / 2 (the two subtask agents)
struct PState {
  uint32_t cell[2];
  uint32_t slot[4];
};

__device__ __forceinline__ uint4 pack_state(const PState& p) {
  return make_uint4(p.cell[0] | (p.cell[1] << 6), p.slot[0] | (p.slot[1] << 16), p.slot[2] | (p.slot[3] << 16), 0u);
}
__device__ __forceinline__ PState unpack_state(const uint4& w) {
  PState p;
  p.cell[0] = w.x & 63u;
  p.cell[1] = (w.x >> 6) & 63u;
  p.slot[0] = w.y & 0xffffu;
  p.slot[1] = w.y >> 16;
  p.slot[2] = w.z & 0xffffu;
  p.slot[3] = w.z >> 16;
  return p;
}

// injective 64-bit key: 2 x 6-bit cells + 4 x (7-bit mask + 6-bit place), place = rank of the
// square among non-walkable squares (<= 60), 60/61 = held by agent 1/2, 63 = dead
__device__ __forceinline__ unsigned long long compact_key(const World& w, const PState& p) {
  unsigned long long k = p.cell[0] | (p.cell[1] << 6);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t s = p.slot[i], holder = s >> 13;
    uint32_t place;
    if (holder == 0u) place = (uint32_t)__popcll(w.nonfloor & ((1ull << ((s >> 7) & 63u)) - 1ull));
    else place = holder == 7u ? 63u : 59u + holder;
    k |= (unsigned long long)((s & 0x7fu) | (place << 7)) << (12 + 13 * i);
  }
  return k;
}

__device__ __forceinline__ uint32_t hash_key(unsigned long long k) {
  k ^= k >> 31;
  k *= 0x9E3779B97F4A7C15ull;
  k ^= k >> 29;
  return (uint32_t)k & (kSlots - 1u);
}

__device__ __forceinline__ int lying_at(const PState& p, uint32_t q) {
  int on = -1;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if ((p.slot[k] >> 13) == 0u && ((p.slot[k] >> 7) & 63u) == q) on = k;
  return on;
}
__device__ __forceinline__ int held_by(const PState& p, int agent) {
  int h = -1;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if ((p.slot[k] >> 13) == (uint32_t)(agent + 1)) h = k;
  return h;
}

// nav_utils.get_single_actions (navigation_planner/utils.py:55-90): bit a = action a offered
__device__ __forceinline__ uint32_t single_actions(const World& w, const PState& p, int agent) {
  uint32_t valid = 1u << 4;
  const int hand = held_by(p, agent);
  for (uint32_t a = 0; a < 4; a++) {
    const uint32_t t = (p.cell[agent] + (uint32_t)gc::action_delta(a)) & 63u;
    if (t == p.cell[0] || t == p.cell[1] || ((w.blocked >> t) & 1ull)) continue;  // :71 an agent stands there
    if (((w.floorp >> t) & 1ull) || ((w.deliv >> t) & 1ull)) {  // :74-78
      valid |= 1u << a;
      continue;
    }
    const int on = lying_at(p, t);
    if (on < 0 && hand >= 0) valid |= 1u << a;                    // :80-81 put down / chop
    else if (on >= 0 && hand < 0) valid |= 1u << a;               // :82-83 pick up
    else if (on >= 0 && hand >= 0 && gc::mergeable(p.slot[hand] & 0x7fu, p.slot[on] & 0x7fu)) valid |= 1u << a;
  }
  return valid;
}

// env.is_collision :671-718: both agents may execute
__device__ __forceinline__ bool joint_ok(const World& w, const PState& p, uint32_t a1, uint32_t a2) {
  uint32_t n1 = (p.cell[0] + (uint32_t)gc::action_delta(a1)) & 63u, n2 = (p.cell[1] + (uint32_t)gc::action_delta(a2)) & 63u;
  if (!((w.floorp >> n1) & 1ull)) n1 = p.cell[0];
  if (!((w.floorp >> n2) & 1ull)) n2 = p.cell[1];
  if (n1 == n2) return false;  // every branch of :704-711 cancels at least one action
  if (p.cell[0] == n2 && p.cell[1] == n1) return false;
  return true;
}

// utils/interact.py:4-89 on the planning world
__device__ __forceinline__ void interact(const World& w, PState& p, int agent, uint32_t a) {
  if (a == 4u) return;
  const uint32_t t = (p.cell[agent] + (uint32_t)gc::action_delta(a)) & 63u;
  if ((w.floorp >> t) & 1ull) {
    p.cell[agent] = t;
    return;
  }
  const int hand = held_by(p, agent), on = lying_at(p, t);
  const bool is_del = (w.deliv >> t) & 1ull, is_cut = (w.cut >> t) & 1ull;
  if (hand >= 0) {
    const uint32_t mH = p.slot[hand] & 0x7fu;
    if (is_del) {
      if (gc::deliverable(mH)) p.slot[hand] = mH | (t << 7);
    } else if (on >= 0) {
      if (gc::mergeable(mH, p.slot[on] & 0x7fu)) {
        p.slot[hand] |= p.slot[on] & 0x7fu;
        p.slot[on] = GC_SLOT_DEAD;
      }
    } else if (is_cut && gc::needs_chopped(mH)) {
      p.slot[hand] |= mH << 4;
    } else {
      p.slot[hand] = mH | (t << 7);
    }
  } else if (on >= 0 && !is_del) {
    p.slot[on] = (p.slot[on] & 0x7fu) | ((uint32_t)(agent + 1) << 13);
  }
}

// e2e_brtdp._define_goal_state :435-566 with a base count of zero
__device__ __forceinline__ bool is_goal(const World& w, const PState& p) {
  bool g = false;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t s = p.slot[k];
    if ((s >> 13) == 7u || (s & 0x7fu) != w.goal_mask) continue;
    if (w.goal_kind == GC_ST_DELIVER) g |= (s >> 13) == 0u && ((w.deliv >> ((s >> 7) & 63u)) & 1ull);
    else g = true;
  }
  return g;
}

// Necessary condition for the goal to be reachable at all, from square adjacency alone.  comp_k
// = floor component of agent k (ignoring the partner), S_k = squares next to it.  An object gets
// into k's hands only if k holds it, or it lies on a non-delivery square of S_k, or the partner
// can get hold of it and a non-delivery square lies in both S_1 and S_2 (hand-over).  The final
// chop / deliver / merge needs the acting agent next to a cutboard / delivery / the other part.
// Occupancy is ignored, so the test only ever errs towards "maybe reachable"; pairs it rejects
// (e.g. both agents on the far side of a full divider) are reported unreachable without the
// exhaustive search that would otherwise burn the whole state budget 25 times.
__device__ bool maybe_reachable(const World& w, const PState& p, uint32_t a_mask, uint32_t b_mask) {
  unsigned long long comp[2], S[2];
  for (int k = 0; k < 2; k++) {
    comp[k] = 1ull << p.cell[k];
    for (unsigned long long grow = comp[k]; grow;) {
      grow = gcnav::neighbours(grow) & w.floorp & ~comp[k];
      comp[k] |= grow;
    }
    S[k] = gcnav::neighbours(comp[k]) & w.nonfloor & ~w.blocked;
  }
  const bool handover = (S[0] & S[1] & ~w.deliv) != 0ull;
  auto direct = [&](int k, uint32_t s) {  // slot s can get into agent k's hands without help
    const uint32_t holder = s >> 13;
    if (holder == (uint32_t)(k + 1)) return true;
    return holder == 0u && ((S[k] & ~w.deliv) >> ((s >> 7) & 63u)) & 1ull;
  };
  auto can_hold = [&](int k, uint32_t s) { return direct(k, s) || (handover && direct(1 - k, s)); };
  bool ok = false;
  for (int i = 0; i < 4; i++) {
    const uint32_t sa = p.slot[i];
    if ((sa >> 13) >= 3u || (sa & 0x7fu) != a_mask) continue;
    for (int k = 0; k < 2; k++) {
      if (w.goal_kind == GC_ST_CHOP) ok |= can_hold(k, sa) && (S[k] & w.cut) != 0ull;
      else if (w.goal_kind == GC_ST_DELIVER) ok |= can_hold(k, sa) && (S[k] & w.deliv) != 0ull;
      else {
        for (int j = 0; j < 4; j++) {
          const uint32_t sb = p.slot[j];
          if (j == i || (sb >> 13) >= 3u || (sb & 0x7fu) != b_mask) continue;
          // k ends up holding one part and facing the other, which therefore must be able to lie in S_k
          auto can_face = [&](uint32_t s) {
            return ((s >> 13) == 0u && ((S[k] & ~w.deliv) >> ((s >> 7) & 63u)) & 1ull) || can_hold(k, s) ||
                   (handover && direct(1 - k, s));
          };
          ok |= (can_hold(k, sa) && can_face(sb)) || (can_hold(k, sb) && can_face(sa));
        }
      }
    }
  }
  return ok;
}

// Planning world + start state of problem (env, pair pi).  Returns 0 = not a joint pair / bad
// subtask (nothing to do), 1 = outside the supported envelope, 2 = search, 3 = offered actions only
// (a goal object exists already, or square adjacency rules the goal out).
__device__ int setup_problem(const GcNavLevels& levels, const GcPairs& pairs, const uint8_t* level_id,
                             const uint4* state, int64_t env, int pi, int n_agents, World& w, PState& p,
                             gc_subtask& st) {
  const GcNavLevel& L = levels.lv[level_id ? level_id[env] : 0];
  const int sub = pairs.p[pi][0], ai = pairs.p[pi][1], aj = pairs.p[pi][2];
  const bool level1 = pairs.p[pi][3] != 0;
  if (aj == 0xFF || (uint32_t)sub >= L.n_subtasks) return 0;
  const uint4 s = state[env];
  st = L.st[sub];
  unsigned long long frozen = 0;
  for (int i = 0; i < n_agents; i++)
    if (i != ai && i != aj) frozen |= 1ull << ((s.x >> (6 * i)) & 63u);
  w.floorp = L.floor_mask & ~frozen;
  w.blocked = level1 ? frozen : 0ull;
  w.nonfloor = ~w.floorp;
  w.cut = L.cut_mask;
  w.deliv = L.deliv_mask;
  w.goal_mask = st.goal;
  w.goal_kind = st.kind;
  w.a_mask = st.a;
  w.b_mask = st.b;
  p.cell[0] = (s.x >> (6 * ai)) & 63u;
  p.cell[1] = (s.x >> (6 * aj)) & 63u;
  bool supported = gcnav::slot_of(s, 4) == GC_SLOT_DEAD && gcnav::slot_of(s, 5) == GC_SLOT_DEAD;
  for (int k = 0; k < 4; k++) {
    uint32_t sl = gcnav::slot_of(s, k);
    const uint32_t holder = sl >> 13;
    if (holder >= 1u && holder <= 4u)
      sl = (int)holder == ai + 1 ? ((sl & 0x7fu) | (1u << 13))
           : (int)holder == aj + 1 ? ((sl & 0x7fu) | (2u << 13))
           : level1 ? ((sl & 0x7fu) | (3u << 13)) : GC_SLOT_DEAD;  // level 1: kept but out of reach
    p.slot[k] = sl;
  }
  supported = supported && __popcll(w.nonfloor) <= 60;  // ranks 60..63 are reserved by compact_key
  if (!supported) return 1;
  // a goal object is already there: the count can never rise (one food of each kind); or the
  // squares the two agents can touch rule the goal out
  return (is_goal(w, p) || !maybe_reachable(w, p, st.a, st.b)) ? 3 : 2;
}

// insert / relax `p` with cost c (tenths); pushes it into its bucket when it improved
__device__ __forceinline__ void relax(const World& w, Arena* A, uint32_t* bcount, uint32_t* n_states, int* over,
                                      const PState& p, uint32_t c) {
  const unsigned long long k = compact_key(w, p);
  uint32_t h = hash_key(k);
  for (uint32_t probe = 0; probe < kSlots; probe++, h = (h + 1u) & (kSlots - 1u)) {
    const unsigned long long old = atomicCAS(&A->keys[h], kEmpty, k);
    if (old == kEmpty) {
      A->states[h] = pack_state(p);
      if (atomicAdd(n_states, 1u) >= kMaxStates) atomicExch(over, 1);
    } else if (old != k) {
      continue;
    }
    const uint32_t prev = atomicMin(&A->gcost[h], 2u * c);
    if (prev > 2u * c) {
      const uint32_t b = c & (kBuckets - 1);
      const uint32_t pos = atomicAdd(&bcount[b], 1u);
      if (pos < kBucketCap) A->bucket[b][pos] = h;
      else atomicExch(over, 1);
    }
    return;
  }
  atomicExch(over, 1);
}

// ---------------------------------------------------------------------------------------
// tree search: one forward A* pass + one backward Dial pass per (env, pair)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_key2(unsigned long long k) {
  k ^= k >> 31;
  k *= 0x9E3779B97F4A7C15ull;
  k ^= k >> 29;
  return (uint32_t)k & (kSlots2 - 1u);
}

__device__ __forceinline__ uint32_t find2(const World& w, const Arena2* A, const PState& p) {
  const unsigned long long k = compact_key(w, p);
  uint32_t h = hash_key2(k);
  for (uint32_t probe = 0; probe < kSlots2; probe++, h = (h + 1u) & (kSlots2 - 1u)) {
    const unsigned long long cur = A->keys[h];
    if (cur == k) return h;
    if (cur == kEmpty) return kNil;
  }
  return kNil;
}

// Distance tables of one planning world (shared memory, bytes, gcnav::kFar = no path):
//   walk[x][y]  steps an agent needs between floor squares (the partner ignored),
//   carry[x][y] steps an OBJECT needs between any two squares: it moves one square per step, either
//               carried over floor or put on / picked from a counter next to the carrier, so edges
//               join 4-neighbours of which at least one is floor (hand-overs across a divider included),
//   reach[c]    carry distance from square c to where the subtask's last action happens: a floor
//               square next to a cutboard (Chop), or a delivery square itself (Deliver).
struct Tables {
  uint8_t walk[64 * 64];
  uint8_t carry[64 * 64];
  uint8_t reach[64];
};

__device__ void fill_tables(const World& w, Tables* T) {
  gcnav::fill_floor_distances(w.floorp, T->walk);
  for (int s = threadIdx.x; s < 64; s += blockDim.x) {
    uint8_t* row = T->carry + s * 64;
    for (int c = 0; c < 64; c += 4) *reinterpret_cast<uint32_t*>(row + c) = 0xFFFFFFFFu;
    unsigned long long visited = 1ull << s, frontier = visited;
    row[s] = 0;
    for (uint32_t d = 1; frontier; d++) {
      frontier = (gcnav::neighbours(frontier & w.floorp) | (gcnav::neighbours(frontier & ~w.floorp) & w.floorp)) & ~visited;
      visited |= frontier;
      for (unsigned long long f = frontier; f; f &= f - 1) row[__ffsll((long long)f) - 1] = (uint8_t)d;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 64; c += blockDim.x) {
    unsigned long long tgt = w.goal_kind == GC_ST_DELIVER ? w.deliv
                             : w.goal_kind == GC_ST_CHOP  ? (gcnav::neighbours(w.cut) & w.floorp) : 0ull;
    uint32_t best = gcnav::kFar;
    for (; tgt; tgt &= tgt - 1) best = min(best, (uint32_t)T->carry[c * 64 + (__ffsll((long long)tgt) - 1)]);
    T->reach[c] = (uint8_t)best;
  }
  __syncthreads();
}

// Admissible estimate (tenths) of the cost still to pay from planning state p; kInfCost = dead end.
// Every step costs at least 11 (one mover), so it is 11 x a lower bound T on the number of steps:
//   nobody holds the object(s) yet -> the nearest agent first walks next to one of them (approach),
//   Chop:    the object then travels `reach` squares and is chopped in one more step,
//   Deliver: it travels `reach` squares (the last one is the put-down),
//   Merge:   the two parts close their carry distance D by at most 2 per step until they are
//            neighbours, and merge in one more step.
__device__ uint32_t heuristic(const World& w, const Tables* T, const PState& p) {
  uint32_t best = kInfCost;
  auto pos_of = [&](uint32_t sl) { return (sl >> 13) == 0u ? ((sl >> 7) & 63u) : p.cell[(sl >> 13) - 1u]; };
  // an object on a delivery square is never picked up again, and facing it does not merge (interact.py:41-46, 73)
  auto stuck = [&](uint32_t sl) { return (sl >> 13) == 0u && ((w.deliv >> ((sl >> 7) & 63u)) & 1ull); };
  auto approach = [&](uint32_t sl) -> uint32_t {  // steps until some agent stands next to a lying object
    if ((sl >> 13) != 0u) return 0u;
    const uint32_t q = (sl >> 7) & 63u;
    uint32_t ap = gcnav::kFar;
    unsigned long long adj = gcnav::neighbours(1ull << q) & w.floorp;
    for (; adj; adj &= adj - 1) {
      const uint32_t f = (uint32_t)__ffsll((long long)adj) - 1u;
      ap = min(ap, min((uint32_t)T->walk[p.cell[0] * 64 + f], (uint32_t)T->walk[p.cell[1] * 64 + f]));
    }
    return ap;
  };
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t sa = p.slot[i];
    if ((sa >> 13) >= 3u || (sa & 0x7fu) != w.a_mask || stuck(sa)) continue;
    if (w.goal_kind != GC_ST_MERGE) {
      const uint32_t ap = approach(sa), r = T->reach[pos_of(sa)];
      if (ap == gcnav::kFar || r == gcnav::kFar) continue;
      best = min(best, ap + r + (w.goal_kind == GC_ST_CHOP ? 1u : 0u));
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t sb = p.slot[j];
        if (j == i || (sb >> 13) >= 3u || (sb & 0x7fu) != w.b_mask || stuck(sb)) continue;
        const uint32_t D = T->carry[pos_of(sa) * 64 + pos_of(sb)];
        if (D == gcnav::kFar) continue;
        uint32_t ap = 0u;
        if ((sa >> 13) == 0u && (sb >> 13) == 0u) {
          ap = min(approach(sa), approach(sb));
          if (ap == gcnav::kFar) continue;
        }
        best = min(best, ap + (D > 1u ? D / 2u : 0u) + 1u);  // ceil((D - 1) / 2) + 1
      }
    }
  }
  return best == kInfCost ? kInfCost : 11u * best;
}

// insert / relax successor `p` of state slot `pred` (kNil for the start) reached with total cost g
// over an edge of cost 10 + code; records the edge in p's predecessor list.  fmin = the parent's
// f, so that keys never decrease along a path (pathmax).  Bucket entries carry (slot, g).
template <bool kRecord = true>
__device__ __forceinline__ void relax2(const World& w, const Tables* T, Arena2* A, uint32_t* bcount, uint32_t* n_states,
                                       uint32_t* n_pool, int* over, const PState& p, uint32_t g, uint32_t fmin,
                                       uint32_t pred, uint32_t code, uint32_t max_states, uint32_t gbase = 0u) {
  const unsigned long long k = compact_key(w, p);
  uint32_t h = hash_key2(k);
  for (uint32_t probe = 0; probe < kSlots2; probe++, h = (h + 1u) & (kSlots2 - 1u)) {
    const unsigned long long old = atomicCAS(&A->keys[h], kEmpty, k);
    if (old == kEmpty) {
      A->states[h] = pack_state(p);
      const uint32_t idx = atomicAdd(n_states, 1u);
      if (idx < kTouchedCap) A->touched[idx] = h;
      if (idx >= max_states) atomicExch(over, 1);
    } else if (old != k) {
      continue;
    }
    // (the cost min and the list-head exchange do not depend on each other: both are issued before either
    // result is used, one global round trip instead of two)
    // gbase: the tag of a seeded per-action search (0 in the forward pass), see seeded searches in the tree kernel
    const uint32_t prev = atomicMin(&A->gcost[h], gbase + 2u * g);
    if (kRecord && pred != kNil) {  // (the per-action A* needs no backward pass: no edges)
      const uint32_t node = atomicAdd(n_pool, 1u);
      if (node < kPool) {
        const uint32_t next = atomicExch(&A->head[h], node);
        A->pool[node] = make_uint2(pred | (code << 30), next);
      } else {
        atomicExch(over, 1);
      }
    }
    // strictly cheaper only: `prev == 2g + 1` is this state already SETTLED at the same cost - re-opening it would
    // expand it a second time and record every one of its edges twice.  (The min has cleared its settled bit by
    // then, which is harmless: the one open-list entry that carried this cost has been consumed.)
    if (prev > gbase + 2u * g + 1u) {
      const uint32_t est = is_goal(w, p) ? 0u : heuristic(w, T, p);
      if (est == kInfCost) return;  // dead end: recorded (it has a slot and its edge) but never expanded
      const uint32_t f = max(fmin, g + est);
      if (pred == kNil) *n_pool = f;  // the start: its key is where the sweep begins (n_pool is reset by the caller)
      else if (f >= fmin + (uint32_t)kRing || g >= (1u << 15)) {
        atomicExch(over, 1);
        return;
      }
      const uint32_t b = f & (kRing - 1);
      const uint32_t pos = atomicAdd(&bcount[b], 1u);
      if (pos < kRingCap) A->bucket[b][pos] = h | (g << 17);
      else atomicExch(over, 1);
    }
    return;
  }
  atomicExch(over, 1);
}

// Forward pass over the open-list entries (slot | g << 17) popped at key `cur`, in two phases so that the
// parallelism is (entry x joint action) and not just entries: a frontier of this search has a handful of states
// far more often than a CTA's worth, and a thread that relaxes all 24 successors of one state itself pays 24
// dependent chains of global atomics (hash CAS, edge node, cost min, bucket push) one after the other - the
// kernel sat at 8 of 32 lanes and 3.5 % issue utilisation (profiles/r02_planner_kernels_ncu.csv).
// settle_entry: claim the state (stale entries and states expanded already fail the CAS), absorb goal states;
// returns 0 when there is nothing to expand, else bit 10 | the two agents' offered single actions (5 bits each)
// and the state itself in `packed`.
__device__ __forceinline__ uint32_t settle_entry(const World& w, Arena2* A, uint32_t* n_goals, int* over, int* result,
                                                 uint32_t entry, uint4& packed, uint32_t gbase = 0u, bool seeded = false) {
  const uint32_t h = entry & (kSlots2 - 1u), g = entry >> 17;
  if (atomicCAS(&A->gcost[h], gbase + 2u * g, gbase + 2u * g + 1u) != gbase + 2u * g) return 0u;
  if (seeded) {  // a state whose exact cost-to-go the backward pass left behind ends the path right here
    const uint32_t v = A->val[h];
    if (v != kInfCost) {
      atomicMin(result, (int)(g + v));
      return 0u;
    }
  }
  packed = A->states[h];
  const PState p = unpack_state(packed);
  if (is_goal(w, p)) {
    atomicMin(result, (int)g);
    A->val[h] = 0u;
    if (!seeded) {
      const uint32_t gi = atomicAdd(n_goals, 1u);
      if (gi < kGoalCap) A->goals[gi] = h;
      else atomicExch(over, 1);
    }
    return 0u;
  }
  return (1u << 10) | single_actions(w, p, 0) | (single_actions(w, p, 1) << 5);
}

// expand_action: relax the successor of a settled entry under joint action `act` (0..23; stay-stay is a self loop)
template <bool kRecord>
__device__ __forceinline__ void expand_action(const World& w, const Tables* T, Arena2* A, uint32_t* bcount,
                                              uint32_t* n_states, uint32_t* n_pool, int* over, uint32_t entry,
                                              const uint4& packed, uint32_t masks, uint32_t act, uint32_t cur,
                                              uint32_t max_states, uint32_t gbase = 0u) {
  const uint32_t h = entry & (kSlots2 - 1u), g = entry >> 17;
  const uint32_t b1 = act / 5u, b2 = act % 5u;
  if (!((masks >> b1) & 1u) || !((masks >> (5u + b2)) & 1u)) return;
  const PState p = unpack_state(packed);
  if (!joint_ok(w, p, b1, b2)) return;
  PState nx = p;
  interact(w, nx, 0, b1);
  interact(w, nx, 1, b2);
  const uint32_t code = (b1 != 4u) + (b2 != 4u);
  relax2<kRecord>(w, T, A, bcount, n_states, n_pool, over, nx, g + 10u + code, cur, h, code, max_states, gbase);
}

// One key bucket of the forward pass, swept until it stops growing (children on the same plateau - pathmax - land in
// this very bucket), then emptied.  Shared by the tree search and the per-action A*; every thread of the CTA calls it.
template <int kTreeThreads, bool kRecord>
__device__ __forceinline__ void sweep_bucket(const World& w, const Tables& T, Arena2* A, uint32_t* bcount, uint32_t b,
                                             uint32_t cur, uint32_t* n_states, uint32_t* n_pool, uint32_t* n_goals, int* over,
                                             int* result, uint32_t* s_cnt, uint4* s_state, uint32_t* s_ent, uint32_t* s_msk,
                                             uint32_t max_states, uint32_t gbase = 0u, bool seeded = false) {
  for (uint32_t done = 0;;) {
    __syncthreads();
    if (threadIdx.x == 0) *s_cnt = min(bcount[b], kRingCap);
    __syncthreads();
    const uint32_t cnt = *s_cnt;
    if (cnt == done) break;
#ifdef GC_JOINT_PER_ENTRY  // the earlier shape, kept for A/B builds: one thread settles AND expands an entry
    for (uint32_t e = done + threadIdx.x; e < cnt; e += kTreeThreads) {
      uint4 packed;
      const uint32_t entry = A->bucket[b][e];
      const uint32_t masks = settle_entry(w, A, n_goals, over, result, entry, packed, gbase, seeded);
      if (masks)
        for (uint32_t act = 0; act < 24u; act++)
          expand_action<kRecord>(w, &T, A, bcount, n_states, n_pool, over, entry, packed, masks, act, cur, max_states, gbase);
    }
#else
    // chunks of one entry per thread: settle (phase A, state and action masks parked in shared memory), then
    // all (entry, action) pairs of the chunk (phase B)
    for (uint32_t base = done; base < cnt; base += kTreeThreads) {
      const uint32_t m = min(cnt - base, (uint32_t)kTreeThreads);
      if (threadIdx.x < m) {
        uint4 packed = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t entry = A->bucket[b][base + threadIdx.x];
        s_msk[threadIdx.x] = settle_entry(w, A, n_goals, over, result, entry, packed, gbase, seeded);
        s_ent[threadIdx.x] = entry;
        s_state[threadIdx.x] = packed;
      }
      __syncthreads();
      for (uint32_t item = threadIdx.x; item < m * 24u; item += kTreeThreads) {
        const uint32_t i = item / 24u, masks = s_msk[i];
        if (masks) expand_action<kRecord>(w, &T, A, bcount, n_states, n_pool, over, s_ent[i], s_state[i], masks, item % 24u, cur, max_states, gbase);
      }
      __syncthreads();
    }
#endif
    done = cnt;
    if (cnt == kRingCap) break;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (bcount[b] > kRingCap) *over = 1;
    bcount[b] = 0;
  }
  __syncthreads();
}

template <int kTreeThreads>
__global__ void __launch_bounds__(kTreeThreads)
joint_tree_kernel(const __grid_constant__ GcNavLevels levels, const __grid_constant__ GcPairs pairs,
                  const uint8_t* __restrict__ level_id, const uint4* __restrict__ state, Arena2* __restrict__ arenas,
                  float* __restrict__ q_out, int* __restrict__ flags, uint32_t* __restrict__ todo, int64_t n,
                  int n_agents, uint32_t max_states) {
  __shared__ uint32_t bcount[kRing], bbcount[kBuckets];
  __shared__ uint32_t n_states, n_pool, n_goals, s_todo, s_v1, s_v2, s_f0, s_cnt;
  __shared__ int over, result, s_kind;
  __shared__ PState start;
  __shared__ World w;
  __shared__ __align__(16) Tables T;
  __shared__ __align__(16) uint4 s_state[kTreeThreads];  // the chunk of open-list entries being expanded
  __shared__ uint32_t s_ent[kTreeThreads], s_msk[kTreeThreads];
  Arena2* A = arenas + blockIdx.x;
  const int64_t n_prob = n * pairs.n;
  auto clear_all = [&]() {
    for (uint32_t k = threadIdx.x; k < kSlots2; k += kTreeThreads) {
      A->keys[k] = kEmpty;
      A->gcost[k] = kInfCost;
      A->val[k] = kInfCost;
      A->head[k] = kNil;
    }
  };
  clear_all();  // once per CTA; every search cleans up the slots it inserted
  for (int64_t prob = blockIdx.x; prob < n_prob; prob += gridDim.x) {
    const int64_t env = prob / pairs.n;
    const int pi = (int)(prob - env * pairs.n);
    __syncthreads();
    if (threadIdx.x == 0) {
      PState p;
      gc_subtask st;
      s_kind = setup_problem(levels, pairs, level_id, state, env, pi, n_agents, w, p, st);
      start = p;
      if (s_kind >= 2) {
        s_v1 = single_actions(w, p, 0);
        s_v2 = single_actions(w, p, 1);
      }
      n_states = n_pool = n_goals = s_todo = 0;
      over = 0;
      result = 0x7fffffff;
    }
    if (threadIdx.x < kRing) bcount[threadIdx.x] = 0;
    __syncthreads();
    if (s_kind == 0) continue;  // a single-agent pair: not ours
    if (s_kind == 1) {
      if (threadIdx.x == 0) {
        atomicOr(&flags[prob], 4);
        todo[prob] = 0;
      }
      continue;
    }
    const uint32_t a1 = threadIdx.x / 5u, a2 = threadIdx.x % 5u;
    const bool offered = threadIdx.x < 25 && ((s_v1 >> a1) & 1u) && ((s_v2 >> a2) & 1u) && joint_ok(w, start, a1, a2);
    if (offered) q_out[prob * 25 + threadIdx.x] = INFINITY;  // e2e_brtdp.get_actions :151-206
    if (s_kind == 3) {  // nothing to search: a goal object exists, or adjacency rules the goal out
      if (threadIdx.x == 0) todo[prob] = 0;
      continue;
    }
    fill_tables(w, &T);  // (ends with a barrier; the table is clean: see the end of the search)
    if (threadIdx.x == 0) {
      relax2(w, &T, A, bcount, &n_states, &n_pool, &over, start, 0u, 0u, kNil, 0u, max_states);
      s_f0 = n_pool;
      n_pool = 0;
    }
    __syncthreads();
    // ---- forward A*: every state with key f <= V* + kSlack is expanded, every generated edge recorded ----
    int empty_run = 0, limit = kMaxCost, cur = (int)s_f0, slack = kSlack;
    bool complete = false;
    // budget: until a goal comes into view this is plain A* and every state is needed (max_states, the arena's
    // capacity); what follows - the ball of radius V* + slack that proves the other actions - gets kTreeStates more
    uint32_t budget = max_states;
    for (;;) {  // widen the explored region until every offered action is proven (or kMaxSlack / the budget is hit)
      for (; cur <= limit; cur++) {
        const uint32_t b = (uint32_t)cur & (kRing - 1);
        if (bcount[b] == 0) {  // uniform: bcount only changes between barriers
          if (++empty_run >= kRing) {
            complete = true;  // nothing left to expand: the whole space that can still reach the goal was covered
            break;
          }
          continue;
        }
        empty_run = 0;
        sweep_bucket<kTreeThreads, true>(w, T, A, bcount, b, (uint32_t)cur, &n_states, &n_pool, &n_goals, &over, &result,
                                         &s_cnt, s_state, s_ent, s_msk, budget);
        if (result != 0x7fffffff && budget == max_states) budget = min(max_states, n_states + kTreeStates);  // uniform
        if (over) break;
        if (result != 0x7fffffff) limit = min(kMaxCost, result + slack);
      }
      // every state with key <= radius has been expanded
      const int radius = complete ? 0x3fffffff : (over ? cur - 1 : (cur > limit ? limit : cur - 1));
      __syncthreads();
      if (threadIdx.x < kBuckets) bbcount[threadIdx.x] = 0;
      if (threadIdx.x == 0) s_todo = 0;
      __syncthreads();
      if (result == 0x7fffffff) {  // no goal inside the explored region
        if (threadIdx.x == 0) {
          // complete: every offered Q is +inf, exactly.  Otherwise plain A* from the start ran out of budget: unknown
          // (an A* from any successor would explore the same space)
          if (!complete) atomicOr(&flags[prob], 1);
          todo[prob] = 0;
        }
        break;
      }
      // ---- backward: exact cost-to-go inside the region, from the goal states over the recorded edges
      // (its own buckets: the forward pass may resume from its open list) ----
      for (uint32_t i = threadIdx.x; i < min(n_goals, kGoalCap); i += kTreeThreads) {
        const uint32_t pos = atomicAdd(&bbcount[0], 1u);
        if (pos < kRingCap) A->bbucket[0][pos] = A->goals[i];
        else atomicExch(&over, 2);
      }
      __syncthreads();
      const int bmax = complete ? kMaxCost : radius;
      int brun = 0;
      for (int bc = 0; bc <= bmax; bc++) {
        const uint32_t b = (uint32_t)bc & (kBuckets - 1);
        const uint32_t cnt = min(bbcount[b], kRingCap);
        if (cnt == 0) {
          if (++brun >= kBuckets) break;
          continue;
        }
        brun = 0;
        for (uint32_t e = threadIdx.x; e < cnt; e += kTreeThreads) {
          const uint32_t h = A->bbucket[b][e];
          if (A->val[h] != (uint32_t)bc) continue;  // improved since it was pushed
          // predecessor list: the next node is fetched while this node's atomic is in flight
          uint32_t node = A->head[h];
          uint2 pn = node != kNil ? A->pool[node] : make_uint2(0u, kNil);
          while (node != kNil) {
            const uint32_t x = pn.x & 0x3FFFFFFFu;
            const uint32_t nv = (uint32_t)bc + 10u + (pn.x >> 30);
            const uint32_t old = atomicMin(&A->val[x], nv);
            node = pn.y;
            if (node != kNil) pn = A->pool[node];
            if (old > nv) {
              const uint32_t pos = atomicAdd(&bbcount[nv & (kBuckets - 1)], 1u);
              if (pos < kRingCap) A->bbucket[nv & (kBuckets - 1)][pos] = x;
              else atomicExch(&over, 2);
            }
          }
        }
        __syncthreads();
        if (threadIdx.x == 0) bbcount[b] = 0;
        __syncthreads();
      }
      // ---- Q(start, a) = cost(a) + cost-to-go of T(start, a), proven when it fits inside the region ----
      if (offered && threadIdx.x != 24) {
        PState nx = start;
        interact(w, nx, 0, a1);
        interact(w, nx, 1, a2);
        const uint32_t code = (a1 != 4u) + (a2 != 4u);
        const float step_cost = 1.0f + 0.1f * (float)code;
        if (is_goal(w, nx)) {
          q_out[prob * 25 + threadIdx.x] = step_cost;
        } else {
          const uint32_t h = find2(w, A, nx);
          const uint32_t v = h == kNil ? kInfCost : A->val[h];
          const bool proven = over != 2 && v != kInfCost && (complete || (int)(v + 10u + code) <= radius);
          const bool dead = heuristic(w, &T, nx) == kInfCost;  // no way back: +inf is exact
          if (proven) q_out[prob * 25 + threadIdx.x] = step_cost + 0.1f * (float)v;
          else if (!dead && !(complete && over != 2)) atomicOr(&s_todo, 1u << threadIdx.x);  // complete: +inf is exact
        }
      }
      __syncthreads();
      const uint32_t open_actions = s_todo;
      // Widening pays while the region is small.  Once it is not (or the budget ran out) the actions still open go to
      // the per-action A* (joint_astar_kernel): from T(start, a) with NO slack it only visits states on optimal plans,
      // where this search - which needs the ball of radius V* + slack around the start to prove an action whose Q is
      // that far above V* - explodes with every step of slack (a partner's idle move costs 0.1).  over == 2 (the
      // backward pass overflowed a bucket) leaves nothing proven beyond doubt: the open actions go the same way.
      if (open_actions == 0u || complete || over || slack >= kMaxSlack || limit >= kMaxCost || n_states > kWidenStates) {
        uint32_t left = open_actions;
#ifdef GC_JOINT_SEEDED
        // ---- seeded per-action searches, in this arena (opt-in build: -DGC_JOINT_SEEDED) ----
        // Measured at the end of round 2: the full GPU suite passes with it (the two solvers still agree bit for bit),
        // but cfg-5's partial-divider batches run 2.66 / 7.27 s with it against 2.31 / 6.86 s without - the open actions
        // mostly come from problems that stopped on their state budget, where few ball states are provably exact and
        // the seeded search ends up handing over to joint_astar_kernel anyway.  Kept for the next round to instrument.
        // After the backward pass a state x of the explored ball with g(x) + val(x) <= radius carries its EXACT
        // cost-to-go (its optimal plan stays among the expanded states: the argument of `proven` above).  An A* from
        // T(start, a) may therefore stop at the first such state it settles - f = g + val(x) is a finished plan - and
        // the plan of an action that was not proven re-enters the ball within a step or two, where the per-action A*
        // of joint_astar_kernel walks all the way to a goal.  The forward costs are no longer needed: `gcost` is wiped
        // once and reused with a per-action tag in its high bits, descending, so that what an earlier action left
        // behind always compares as "not reached yet".
        if (left && over != 2 && n_states + kSeedStates <= max_states && n_states <= kTouchedCap) {
          for (uint32_t i = threadIdx.x; i < n_states; i += kTreeThreads) {
            const uint32_t h = A->touched[i], v = A->val[h];
            if (v != kInfCost && (A->gcost[h] >> 1) + v > (uint32_t)radius) A->val[h] = kInfCost;  // not proven exact
            A->gcost[h] = kInfCost;
          }
          const uint32_t seed_budget = n_states + kSeedStates;
          const int over_before = over;
          __syncthreads();
          if (threadIdx.x == 0) over = 0;
          uint32_t tag = 24u;
          for (uint32_t m = left; m; m &= m - 1u, tag--) {
            const uint32_t act = (uint32_t)__ffs((int)m) - 1u, b1 = act / 5u, b2 = act % 5u;
            const uint32_t gbase = tag << 20;
            __syncthreads();
            if (threadIdx.x < kRing) bcount[threadIdx.x] = 0;
            if (threadIdx.x == 0) result = 0x7fffffff;
            __syncthreads();
            if (threadIdx.x == 0) {
              PState nx = start;
              interact(w, nx, 0, b1);
              interact(w, nx, 1, b2);
              relax2<false>(w, &T, A, bcount, &n_states, &n_pool, &over, nx, 0u, 0u, kNil, 0u, seed_budget, gbase);
              s_f0 = n_pool;  // the root's key
            }
            __syncthreads();
            int run = 0;
            bool exhausted = false;
            for (int c = (int)s_f0; c <= kMaxCost; c++) {
              const uint32_t b = (uint32_t)c & (kRing - 1);
              if (bcount[b] == 0) {
                if (++run >= kRing) {
                  exhausted = true;
                  break;
                }
                continue;
              }
              run = 0;
              sweep_bucket<kTreeThreads, false>(w, T, A, bcount, b, (uint32_t)c, &n_states, &n_pool, &n_goals, &over, &result,
                                                &s_cnt, s_state, s_ent, s_msk, seed_budget, gbase, true);
              if (over || (result != 0x7fffffff && result <= c + 1)) break;  // every open key is >= c + 1 now
            }
            if (over) break;  // out of room: this action and the ones after it go to joint_astar_kernel
            if (result != 0x7fffffff) {
              if (threadIdx.x == 0) q_out[prob * 25 + act] = 1.0f + 0.1f * (float)((b1 != 4u) + (b2 != 4u)) + 0.1f * (float)result;
              left &= ~(1u << act);
            } else if (exhausted) {
              left &= ~(1u << act);  // no plan at all: the +inf written for an offered action is exact
            }
          }
          __syncthreads();
          if (threadIdx.x == 0) over = over_before;
        }
#endif
        if (threadIdx.x == 0) todo[prob] = left;
        break;
      }
      // widen: the forward pass resumes where it stopped; cost-to-go values are rebuilt from the goals
      slack += kSlackStep;
      limit = min(kMaxCost, result + slack);
      // only the slots this search inserted can hold a value (a 512 KB sweep of the whole table per widening
      // round was a good part of the kernel's 27 GB of DRAM traffic per 2^12-env launch)
      if (n_states <= kTouchedCap) {
        for (uint32_t i = threadIdx.x; i < n_states; i += kTreeThreads) A->val[A->touched[i]] = kInfCost;
      } else {
        for (uint32_t k = threadIdx.x; k < kSlots2; k += kTreeThreads) A->val[k] = kInfCost;
      }
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < min(n_goals, kGoalCap); i += kTreeThreads) A->val[A->goals[i]] = 0u;
      __syncthreads();
    }  // widening loop
    __syncthreads();
    if (n_states <= kTouchedCap) {
      for (uint32_t i = threadIdx.x; i < n_states; i += kTreeThreads) {
        const uint32_t h = A->touched[i];
        A->keys[h] = kEmpty;
        A->gcost[h] = kInfCost;
        A->val[h] = kInfCost;
        A->head[h] = kNil;
      }
    } else {
      clear_all();
    }
  }
}

// Per-action A*: Q(start, a) = cost(a) + V*(T(start, a)) for the actions the tree search left open (todo mask),
// each by its own forward A* from the successor with the same keys, heuristic and pathmax rule as the tree search
// but NO slack and NO edge recording: the first goal state settled ends the search (keys are monotone, a strictly
// cheaper path re-opens a state, so that cost is V*), and an exhausted open list proves +inf.
template <int kTreeThreads>
__global__ void __launch_bounds__(kTreeThreads)
joint_astar_kernel(const __grid_constant__ GcNavLevels levels, const __grid_constant__ GcPairs pairs,
                   const uint8_t* __restrict__ level_id, const uint4* __restrict__ state, Arena2* __restrict__ arenas,
                   float* __restrict__ q_out, int* __restrict__ flags, uint32_t* __restrict__ queue,
                   const unsigned long long* __restrict__ units, int64_t unit_cap, int64_t n, int n_agents,
                   uint32_t max_states) {
  __shared__ uint32_t bcount[kRing];
  __shared__ uint32_t n_states, n_pool, n_goals, s_f0, s_cnt;
  __shared__ int over, result;
  __shared__ PState start;
  __shared__ World w;
  __shared__ __align__(16) Tables T;
  __shared__ __align__(16) uint4 s_state[kTreeThreads];
  __shared__ uint32_t s_ent[kTreeThreads], s_msk[kTreeThreads];
  __shared__ unsigned long long s_unit;
  Arena2* A = arenas + blockIdx.x;  // left clean by the tree kernel; every search cleans up after itself
  // Work units are (problem, action) pairs listed by joint_units_kernel and handed out through one global counter:
  // the problems with open actions are few and uneven (one can bring 24 searches of 10^4 states), and with a static
  // stride over problems the launch lasted as long as its unluckiest CTA.
  const uint32_t n_units = (uint32_t)min((int64_t)queue[0], unit_cap);
  int64_t last_prob = -1;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long un = ~0ull;
      for (;;) {
        const uint32_t u = atomicAdd(&queue[1], 1u);
        if (u >= n_units) {
          un = ~0ull;
          break;
        }
        un = units[u];
        if (un == kSkipUnit) continue;
        // a pair one of whose searches ran out of budget is reported as status 3 whatever its other searches find:
        // they are skipped (read by one thread, so that the whole CTA sees one answer)
        if (!(*reinterpret_cast<volatile int*>(&flags[un >> 5]) & 1)) break;
      }
      s_unit = un;
    }
    __syncthreads();
    const unsigned long long unit = s_unit;
    if (unit == ~0ull) break;
    const int64_t prob = (int64_t)(unit >> 5);
    const uint32_t act = (uint32_t)(unit & 31ull);
    if (prob != last_prob) {
      const int64_t env = prob / pairs.n;
      const int pi = (int)(prob - env * pairs.n);
      if (threadIdx.x == 0) {
        PState p;
        gc_subtask st;
        setup_problem(levels, pairs, level_id, state, env, pi, n_agents, w, p, st);
        start = p;
      }
      __syncthreads();
      fill_tables(w, &T);
      last_prob = prob;
    }
    {
      const uint32_t a1 = act / 5u, a2 = act % 5u;
      const uint32_t code = (a1 != 4u) + (a2 != 4u);
      __syncthreads();
      if (threadIdx.x < kRing) bcount[threadIdx.x] = 0;
      if (threadIdx.x == 0) {
        n_states = n_pool = n_goals = 0;
        over = 0;
        result = 0x7fffffff;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        PState nx = start;
        interact(w, nx, 0, a1);
        interact(w, nx, 1, a2);
        relax2<false>(w, &T, A, bcount, &n_states, &n_pool, &over, nx, 0u, 0u, kNil, 0u, max_states);
        s_f0 = n_pool;  // the root's key (relax2 leaves it there for the start)
        n_pool = 0;
      }
      __syncthreads();
      int empty_run = 0;
      bool complete = false;
      for (int cur = (int)s_f0; cur <= kMaxCost; cur++) {
        const uint32_t b = (uint32_t)cur & (kRing - 1);
        if (bcount[b] == 0) {
          if (++empty_run >= kRing) {
            complete = true;
            break;
          }
          continue;
        }
        empty_run = 0;
        sweep_bucket<kTreeThreads, false>(w, T, A, bcount, b, (uint32_t)cur, &n_states, &n_pool, &n_goals, &over, &result,
                                          &s_cnt, s_state, s_ent, s_msk, max_states);
        if (over || result != 0x7fffffff) break;
      }
      if (threadIdx.x == 0) {
        if (result != 0x7fffffff) q_out[prob * 25 + act] = 1.0f + 0.1f * (float)code + 0.1f * (float)result;
        else if (!complete) atomicOr(&flags[prob], 1);  // budget (or the cost ceiling): this Q stays unknown
      }
      __syncthreads();
      if (n_states <= kTouchedCap) {
        for (uint32_t i = threadIdx.x; i < n_states; i += kTreeThreads) {
          const uint32_t h = A->touched[i];
          A->keys[h] = kEmpty;
          A->gcost[h] = kInfCost;
          A->val[h] = kInfCost;
          A->head[h] = kNil;
        }
      } else {
        for (uint32_t k = threadIdx.x; k < kSlots2; k += kTreeThreads) {
          A->keys[k] = kEmpty;
          A->gcost[k] = kInfCost;
          A->val[k] = kInfCost;
          A->head[k] = kNil;
        }
      }
    }  // one (problem, action) search
  }
}

// the open actions of all problems as a list of (problem << 5 | action) units.  The list holds every open action
// of a batch of up to 4 096 problems and at least one slot per problem beyond that (0.1-0.5 per problem are used);
// a problem whose actions do not fit any more is reported as over budget.
__global__ void joint_units_kernel(const uint32_t* __restrict__ todo, int* __restrict__ flags, uint32_t* __restrict__ queue,
                                   unsigned long long* __restrict__ units, int64_t n_prob, int64_t unit_cap) {
  const int64_t prob = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (prob >= n_prob) return;
  uint32_t m = todo[prob];
  if (m == 0u) return;
  const uint32_t k = (uint32_t)__popc(m);
  const uint32_t base = atomicAdd(&queue[0], k);
  if ((int64_t)base + k > unit_cap) {
    atomicOr(&flags[prob], 1);
    for (int64_t i = base; i < unit_cap; i++) units[i] = kSkipUnit;  // the counter has moved past these slots
    return;
  }
  for (uint32_t i = 0; m; i++, m &= m - 1u) units[base + i] = ((unsigned long long)prob << 5) | (uint32_t)(__ffs((int)m) - 1);
}

__global__ void __launch_bounds__(kThreads)
joint_q_kernel(const __grid_constant__ GcNavLevels levels, const __grid_constant__ GcPairs pairs,
               const uint8_t* __restrict__ level_id, const uint4* __restrict__ state, Arena* __restrict__ arenas,
               float* __restrict__ q_out, int* __restrict__ flags, const uint32_t* __restrict__ todo, int64_t n,
               int n_agents) {
  __shared__ uint32_t bcount[kBuckets];
  __shared__ uint32_t n_states;
  __shared__ int over, result;
  __shared__ PState root;
  __shared__ World w;
  __shared__ int root_state;  // 0 search, 1 invalid action, 2 goal right away
  Arena* A = arenas + blockIdx.x;
  // work units: with a todo mask (after the tree search) one problem, whose open actions are walked
  // here - nearly all masks are 0, and one uniform read per PROBLEM skips them; without it (the
  // first-generation path) one (problem, action)
  const int64_t n_prob = n * pairs.n;
  const int64_t n_units = todo ? n_prob : n_prob * 25;
  for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int64_t prob = todo ? unit : unit / 25;
    uint32_t open = todo ? todo[prob] : (1u << (int)(unit - prob * 25));
    while (open) {
      const int act = __ffs((int)open) - 1;
      open &= open - 1u;
      const int64_t env = prob / pairs.n;
      const int pi = (int)(prob - env * pairs.n);
      const uint32_t a1 = (uint32_t)(act / 5), a2 = (uint32_t)(act % 5);
      __syncthreads();
      if (threadIdx.x == 0) {
        root_state = 1;
        PState p;
        gc_subtask st;
        const bool wanted = true;
        const int kind = wanted ? setup_problem(levels, pairs, level_id, state, env, pi, n_agents, w, p, st) : 0;
        if (kind == 1) {
          atomicOr(&flags[prob], 4);
        } else if (kind >= 2 && wanted) {
          const bool goal_exists = kind == 3;
          const uint32_t v1 = single_actions(w, p, 0), v2 = single_actions(w, p, 1);
          if (((v1 >> a1) & 1u) && ((v2 >> a2) & 1u) && joint_ok(w, p, a1, a2)) {
            q_out[prob * 25 + act] = INFINITY;  // offered (e2e_brtdp.get_actions :151-206)
            // (stay, stay) leaves the state unchanged: Q = 1 + V*(start), filled in by the finalize kernel
            if (!goal_exists && act != 24) {
              interact(w, p, 0, a1);
              interact(w, p, 1, a2);
              root = p;
              root_state = is_goal(w, p) ? 2 : 0;
              // an irreversible move can put the goal out of reach: same adjacency test as at the start
              if (root_state == 0 && !maybe_reachable(w, p, st.a, st.b)) root_state = 1;
            }
          }
        }
        n_states = 0;
        over = 0;
        result = 0x7fffffff;
      }
      if (threadIdx.x < kBuckets) bcount[threadIdx.x] = 0;
      __syncthreads();
      const float step_cost = 1.0f + 0.1f * (float)((a1 != 4u) + (a2 != 4u));
      if (root_state == 1) continue;
      if (root_state == 2) {
        if (threadIdx.x == 0) q_out[prob * 25 + act] = step_cost;
        continue;
      }
      // ---- clear the table, seed the root ----
      for (uint32_t k = threadIdx.x; k < kSlots; k += kThreads) {
        A->keys[k] = kEmpty;
        A->gcost[k] = kInfCost;
      }
      __syncthreads();
      if (threadIdx.x == 0) relax(w, A, bcount, &n_states, &over, root, 0u);
      __syncthreads();
      // ---- Dial's algorithm ----
      int empty_run = 0;
      for (int cur = 0; cur <= kMaxCost && empty_run < kBuckets; cur++) {
        const uint32_t b = (uint32_t)cur & (kBuckets - 1);
        const uint32_t cnt = min(bcount[b], kBucketCap);
        if (cnt == 0) {
          empty_run++;
          continue;  // uniform: bcount is shared and only changes between barriers
        }
        empty_run = 0;
        for (uint32_t e = threadIdx.x; e < cnt; e += kThreads) {
          const uint32_t h = A->bucket[b][e];
          // settle exactly once, and only if this entry still carries the best cost
          if (atomicCAS(&A->gcost[h], 2u * (uint32_t)cur, 2u * (uint32_t)cur + 1u) != 2u * (uint32_t)cur) continue;
          const PState p = unpack_state(A->states[h]);
          if (is_goal(w, p)) {
            atomicMin(&result, cur);
            continue;
          }
          const uint32_t v1 = single_actions(w, p, 0), v2 = single_actions(w, p, 1);
          for (uint32_t b1 = 0; b1 < 5; b1++) {
            if (!((v1 >> b1) & 1u)) continue;
            for (uint32_t b2 = 0; b2 < 5; b2++) {
              if (!((v2 >> b2) & 1u) || (b1 == 4u && b2 == 4u) || !joint_ok(w, p, b1, b2)) continue;
              PState nx = p;
              interact(w, nx, 0, b1);
              interact(w, nx, 1, b2);
              relax(w, A, bcount, &n_states, &over, nx, (uint32_t)cur + 10u + (b1 != 4u) + (b2 != 4u));
            }
          }
        }
        __syncthreads();
        if (threadIdx.x == 0) bcount[b] = 0;
        __syncthreads();
        if (result != 0x7fffffff || over) break;
      }
      if (threadIdx.x == 0) {
        if (result != 0x7fffffff) q_out[prob * 25 + act] = step_cost + 0.1f * (float)result;
        else if (over) atomicOr(&flags[prob], 1);  // budget exceeded: this Q stays +inf and is flagged
      }
    }  // open actions
  }
}

// per problem: v = min_a q, q(stay, stay) = 1 + v, status
__global__ void joint_finalize_kernel(const __grid_constant__ GcPairs pairs, float* __restrict__ v_out,
                                      float* __restrict__ q_out, uint8_t* __restrict__ status_out,
                                      const int* __restrict__ flags, int64_t n) {
  const int64_t prob = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (prob >= n * pairs.n) return;
  if (pairs.p[prob % pairs.n][2] == 0xFF) return;  // single-agent pair: not ours
  float best = INFINITY;
  for (int a = 0; a < 24; a++) best = fminf(best, q_out[prob * 25 + a]);
  if (isfinite(best)) {
    // (stay, stay) changes nothing: Q = 1 + V*(start).  It is offered whenever both agents may stay,
    // which is always (navigation_planner/utils.py:88), and never collides.
    q_out[prob * 25 + 24] = 1.0f + best;
  }
  v_out[prob] = best;
  if (status_out)
    status_out[prob] = (flags[prob] & 4) ? ST_UNSUPPORTED : (flags[prob] & 1) ? ST_BUDGET : isfinite(best) ? ST_OK
                                                                                                           : ST_UNREACHABLE;
}

__global__ void joint_init_kernel(const __grid_constant__ GcPairs pairs, float* __restrict__ q_out,
                                  int* __restrict__ flags, uint32_t* __restrict__ todo, uint32_t* __restrict__ queue,
                                  int64_t n) {
  const int64_t prob = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (prob >= n * pairs.n) return;
  if (prob == 0) queue[0] = queue[1] = 0u;  // units listed / units handed out (joint_astar_kernel)
  flags[prob] = 0;
  todo[prob] = 0;
  if (pairs.p[prob % pairs.n][2] == 0xFF) return;
  for (int a = 0; a < 25; a++) q_out[prob * 25 + a] = NAN;  // NaN = not offered; +inf = offered, goal out of reach
}

}  // namespace

extern "C" {

// scratch layout: [arenas: max(per-action arenas, tree arenas)] [flags: int per problem] [todo: u32 per problem]
static void joint_ctas(int64_t n, int n_pairs, int* tree_ctas, int* act_ctas, bool* wide = nullptr) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  // searches (CTAs, each with its own 18 MB arena) per SM: 8, of which 2 of the default 512-thread shape are
  // resident at a time (21 GB of arenas, capped at a quarter of the device memory).  History: 4 x 128 threads in
  // round 1; 16 x 64 when a thread still expanded whole entries (2.0x on cfg-3, scripts/joint_ctas_probe.sh);
  // the wide shape since the (entry x action) expansion (note at the launch site).  GC_JOINT_CTAS_PER_SM=k overrides.
  static const int per_sm = [] {
    const char* e = getenv("GC_JOINT_CTAS_PER_SM");
    const int v = e ? atoi(e) : GC_JOINT_CTAS_PER_SM_DEFAULT;
    return v >= 1 && v <= 16 ? v : GC_JOINT_CTAS_PER_SM_DEFAULT;
  }();
  static const int64_t arena_cap = [] {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
      cudaGetLastError();
      total_b = (size_t)64 << 30;
    }
    const int64_t k = (int64_t)(total_b / 4 / sizeof(Arena2));
    return k < 64 ? (int64_t)64 : k;
  }();
  int64_t cap = (int64_t)sms * per_sm;
  if (cap > arena_cap) cap = arena_cap;
  const int64_t problems = n * n_pairs;
  *tree_ctas = (int)(problems < cap ? (problems > 0 ? problems : 1) : cap);
  *act_ctas = (int)(problems * 25 < cap ? (problems > 0 ? problems * 25 : 1) : cap);
  if (wide) *wide = false;
}

// slots of the per-action work list (joint_units_kernel)
static int64_t joint_unit_cap(int64_t problems) {
  const int64_t small = 4096;
  return problems <= small ? problems * 24 : (problems > small * 24 ? problems : small * 24);
}

static int64_t joint_arena_bytes(int tree_ctas, int act_ctas) {
  const int64_t a = (int64_t)tree_ctas * (int64_t)sizeof(Arena2), b = (int64_t)act_ctas * (int64_t)sizeof(Arena);
  return ((a > b ? a : b) + 255) & ~(int64_t)255;
}

int64_t gc_joint_q_scratch_bytes(int64_t n, int n_pairs, int* n_ctas_out) {
  int tree_ctas = 0, act_ctas = 0;
  joint_ctas(n, n_pairs, &tree_ctas, &act_ctas);
  if (n_ctas_out) *n_ctas_out = tree_ctas;
  // + per problem: flags, todo mask, one unit slot; + the two queue counters
  return joint_arena_bytes(tree_ctas, act_ctas) + n * n_pairs * (int64_t)(sizeof(int) + sizeof(uint32_t)) +
         joint_unit_cap(n * n_pairs) * (int64_t)sizeof(unsigned long long) + 16;
}

int gc_joint_q(const gc_level* levels, int n_levels, const uint8_t* level_id, const uint32_t* state,
               const uint8_t* pairs, int n_pairs, float* v, float* q, uint8_t* status, void* scratch,
               int64_t scratch_bytes, int64_t n, int n_agents, void* stream) {
  GcNavLevels lv;
  GcPairs pr;
  if (n_agents < 2 || n_agents > GC_MAX_AGENTS) return gc_fail(GC_E_ARG, "gc_joint_q: n_agents must be 2..4");
  if (int rc = gc_nav_levels_to_dev(levels, n_levels, &lv)) return rc;
  if (int rc = gc_pairs_to_dev(pairs, n_pairs, n_agents, &pr)) return rc;
  if (!state || !v || !q || n < 0) return gc_fail(GC_E_ARG, "gc_joint_q: null state/v/q or n < 0");
  if (n_levels > 1 && !level_id) return gc_fail(GC_E_ARG, "gc_joint_q: n_levels > 1 needs level_id");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  int tree_ctas = 0, act_ctas = 0;
  bool wide = false;
  joint_ctas(n, n_pairs, &tree_ctas, &act_ctas, &wide);
  const int64_t need = gc_joint_q_scratch_bytes(n, n_pairs, nullptr);
  if (!scratch || scratch_bytes < need)
    return gc_fail(GC_E_ARG, "gc_joint_q: scratch of %lld bytes needed, %lld given", (long long)need, (long long)scratch_bytes);
  char* base = reinterpret_cast<char*>(scratch);
  int* flags = reinterpret_cast<int*>(base + joint_arena_bytes(tree_ctas, act_ctas));
  uint32_t* todo = reinterpret_cast<uint32_t*>(flags + n * n_pairs);
  unsigned long long* units = reinterpret_cast<unsigned long long*>(todo + n * n_pairs);  // 8 B per problem before it: aligned
  const int64_t unit_cap = joint_unit_cap(n * n_pairs);
  uint32_t* queue = reinterpret_cast<uint32_t*>(units + unit_cap);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t probs = n * n_pairs;
  const unsigned pgrid = (unsigned)((probs + 255) / 256);
  const uint4* s4 = reinterpret_cast<const uint4*>(state);
  const uint8_t* lid = n_levels > 1 ? level_id : nullptr;
  static const bool per_action_only = getenv("GC_JOINT_PER_ACTION") != nullptr;  // the first-generation path, for A/B runs
  joint_init_kernel<<<pgrid, 256, 0, st>>>(pr, q, flags, todo, queue, n);
  if (per_action_only) {
    joint_q_kernel<<<(unsigned)act_ctas, kThreads, 0, st>>>(lv, pr, lid, s4, reinterpret_cast<Arena*>(base), q, flags,
                                                            nullptr, n, n_agents);
  } else {
    // CTA shape.  While a thread expanded a whole open-list entry the searches were latency-bound with a handful of
    // busy lanes, and many small CTAs (16 of 64 threads per SM) beat few large ones for big batches.  With the
    // (entry x action) expansion a search keeps 512 threads busy, and the wide shape - 2 resident CTAs of 512
    // threads per SM, a grid of 8 per SM - is as fast or faster everywhere: cfg-4 loop 3.2 -> 2.3 s, cfg-3 chunk
    // equal, 4 096 x 11 problems 0.09 -> 0.06 s, and it needs half the arenas (21 GB instead of 43 GB).
    // GC_JOINT_WIDE_PROBLEMS=<n> restores the narrow shape for launches of more than n problems (A/B runs).
    static const int64_t wide_limit = getenv("GC_JOINT_WIDE_PROBLEMS") ? atoll(getenv("GC_JOINT_WIDE_PROBLEMS")) : kWideProblems;
    static const bool narrow = !(getenv("GC_JOINT_THREADS") && atoi(getenv("GC_JOINT_THREADS")) == 128);  // 128: the round-1 shape
    // GC_JOINT_UCS_FALLBACK=1: the open actions go to the first-generation uniform-cost search instead (A/B runs)
    static const bool ucs_fallback = getenv("GC_JOINT_UCS_FALLBACK") != nullptr;
    // GC_JOINT_BUDGET=<states>: a smaller per-search state budget than the arenas' capacity (experiments)
    static const uint32_t budget = [] {
      const char* e = getenv("GC_JOINT_BUDGET");
      const long v = e ? atol(e) : 0;
      return v >= 1024 && v < (long)kMaxStates2 ? (uint32_t)v : kMaxStates2;
    }();
    Arena2* a2 = reinterpret_cast<Arena2*>(base);
    if (probs > wide_limit && narrow) {
      joint_tree_kernel<64><<<(unsigned)tree_ctas, 64, 0, st>>>(lv, pr, lid, s4, a2, q, flags, todo, n, n_agents, budget);
      if (!ucs_fallback) {
        joint_units_kernel<<<pgrid, 256, 0, st>>>(todo, flags, queue, units, probs, unit_cap);
        joint_astar_kernel<64><<<(unsigned)tree_ctas, 64, 0, st>>>(lv, pr, lid, s4, a2, q, flags, queue, units, unit_cap, n, n_agents, budget);
      }
    } else if (probs > wide_limit) {
      joint_tree_kernel<kThreads><<<(unsigned)tree_ctas, kThreads, 0, st>>>(lv, pr, lid, s4, a2, q, flags, todo, n, n_agents,
                                                                          budget);
      if (!ucs_fallback) {
        joint_units_kernel<<<pgrid, 256, 0, st>>>(todo, flags, queue, units, probs, unit_cap);
        joint_astar_kernel<kThreads><<<(unsigned)tree_ctas, kThreads, 0, st>>>(lv, pr, lid, s4, a2, q, flags, queue, units, unit_cap,
                                                                             n, n_agents, budget);
      }
    } else {
      const int wide_ctas = tree_ctas;
      joint_tree_kernel<kTreeThreadsWide><<<(unsigned)wide_ctas, kTreeThreadsWide, 0, st>>>(lv, pr, lid, s4, a2, q, flags, todo,
                                                                                         n, n_agents, budget);
      if (!ucs_fallback) {
        joint_units_kernel<<<pgrid, 256, 0, st>>>(todo, flags, queue, units, probs, unit_cap);
        joint_astar_kernel<kTreeThreadsWide><<<(unsigned)wide_ctas, kTreeThreadsWide, 0, st>>>(lv, pr, lid, s4, a2, q, flags,
                                                                                              queue, units, unit_cap, n, n_agents, budget);
      }
    }
    if (ucs_fallback)
      joint_q_kernel<<<(unsigned)act_ctas, kThreads, 0, st>>>(lv, pr, lid, s4, reinterpret_cast<Arena*>(base), q, flags, todo,
                                                              n, n_agents);
  }
  joint_finalize_kernel<<<pgrid, 256, 0, st>>>(pr, v, q, status, flags, n);
  return gc_check_launch("gc_joint_q");
}

}  // extern "C"
