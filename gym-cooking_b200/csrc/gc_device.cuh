// gc_device.cuh - device-side packed-state logic shared by the env-step, rollout and planner
// kernels.  Bit layout: include/gymcook.h.  Reference semantics cited per function
// (paths relative to /root/reference/gym_cooking/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gymcook.h"

// Per-level tables in the form the kernels consume: three 64-bit cell bitboards instead of a
// type byte per cell, and each Deliver goal pre-combined with the delivery cell into the
// 16-bit slot value a delivered goal object has (holder = 0).
struct GcLevelDev {
  unsigned long long floor_mask;  // bit c set: cell c is Floor (the only non-collidable square)
  unsigned long long cut_mask;    // Cutboard squares
  unsigned long long deliv_mask;  // Delivery squares (all of them: interact.py:35 tests the type)
  uint32_t goal_slot[GC_MAX_GOALS];  // goal_mask | first_delivery_cell << 7; unused entries repeat goal 0
  uint32_t n_goals;
  uint32_t max_t;  // 0 = no limit
  uint32_t init[4];
};
static_assert(sizeof(GcLevelDev) == 64, "GcLevelDev layout");

struct GcLevelsDev {
  GcLevelDev lv[GC_MAX_LEVELS];
};

namespace gc {

// World.NAV_ACTIONS + stay as cell offsets (cell = y*8+x): 0:(0,+1)=+8 1:(0,-1)=-8 2:(-1,0)=-1
// 3:(+1,0)=+1 4:stay=0 (utils/world.py:16).  Any value > 4 is treated as stay.
__device__ __forceinline__ int action_delta(uint32_t a) {
  uint32_t t = __funnelshift_rc(0x01FFF808u, 0u, a * 8u);  // byte a of the table, 0 for a >= 4
  return (int)(int8_t)(t & 0xffu);
}

// every Food in the object is in its last state (Food.done, utils/core.py:293-296)
__device__ __forceinline__ bool foods_done(uint32_t m) { return ((m & 7u) & ~(m >> 4)) == 0u; }
// Object.is_deliverable utils/core.py:214-219
__device__ __forceinline__ bool deliverable(uint32_t m) {
  return foods_done(m) && (__popc(m & 15u) > 1);
}
// mergeable utils/core.py:222-241
__device__ __forceinline__ bool mergeable(uint32_t a, uint32_t b) {
  return ((a & b & 8u) == 0u) && foods_done(a | b);
}
// Object.needs_chopped utils/core.py:176-178: a single fresh food, i.e. mask in {1,2,4}
__device__ __forceinline__ bool needs_chopped(uint32_t m) { return ((0x16u >> min(m, 31u)) & 1u) != 0u; }

// Unpacked working form: agent cells and object slots in registers.
template <int NOBJ>
struct Env {
  uint32_t cell[GC_MAX_AGENTS];
  uint32_t slot[NOBJ];
  uint32_t t;
};

// ---- packed state <-> working slots -----------------------------------------------------------
// In HBM an env keeps its objects as byte planes (include/gymcook.h: word 1 = place bytes of objects
// 0..3, word 2 = their masks, word 3 = place, place, mask, mask of objects 4 and 5), the form the step
// kernels compute on (gc_step2.cuh).  The planner / render kernels and the reference form of the step
// below work on one 16-bit *slot* per object: bits 0-6 mask, 7-12 cell (0 while held), 13-15 holder
// (0 lying, 1..4 agent, 7 dead; a dead slot is GC_SLOT_DEAD).  The two are converted here and only here.
__device__ __forceinline__ uint32_t slot_from_bytes(uint32_t place, uint32_t mask) {
  const uint32_t p9 = (place & GC_PLACE_HELD) ? ((place & 7u) << 6) : place;  // holder << 6 | cell
  return (p9 << 7) | mask;
}
__device__ __forceinline__ uint32_t place_of_slot(uint32_t slot) {
  const uint32_t holder = slot >> 13;
  return holder ? (GC_PLACE_HELD | holder) : ((slot >> 7) & 63u);
}
// slot of object k (0..5) of a packed state
__device__ __forceinline__ uint32_t slot_of(const uint4& s, int k) {
  const uint32_t place = k < 4 ? (s.y >> (8 * k)) : (s.w >> (8 * (k - 4)));
  const uint32_t mask = k < 4 ? (s.z >> (8 * k)) : (s.w >> (8 * (k - 2)));
  return slot_from_bytes(place & 0x7Fu, mask & 0x7Fu);
}
// the three object words from six slots
__device__ __forceinline__ void words_from_slots(const uint32_t (&sl)[GC_MAX_OBJECTS], uint32_t& y, uint32_t& z,
                                                 uint32_t& w) {
  uint32_t P[GC_MAX_OBJECTS], M[GC_MAX_OBJECTS];
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++) {
    P[k] = place_of_slot(sl[k]);
    M[k] = sl[k] & 0x7Fu;
  }
  y = P[0] | (P[1] << 8) | (P[2] << 16) | (P[3] << 24);
  z = M[0] | (M[1] << 8) | (M[2] << 16) | (M[3] << 24);
  w = P[4] | (P[5] << 8) | (M[4] << 16) | (M[5] << 24);
}

template <int NA, int NOBJ>
__device__ __forceinline__ void unpack(const uint4& s, Env<NOBJ>& e) {
#pragma unroll
  for (int i = 0; i < NA; i++) e.cell[i] = (s.x >> (6 * i)) & 63u;
  e.t = (s.x >> 24) & 127u;
#pragma unroll
  for (int k = 0; k < NOBJ; k++) e.slot[k] = slot_of(s, k);
}

template <int NA, int NOBJ>
__device__ __forceinline__ uint4 pack(const Env<NOBJ>& e, bool done) {
  uint32_t x = (e.t << 24) | (done ? 0x80000000u : 0u);
#pragma unroll
  for (int i = 0; i < NA; i++) x |= e.cell[i] << (6 * i);
  uint32_t sl[GC_MAX_OBJECTS];
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++) sl[k] = k < NOBJ ? e.slot[k] : GC_SLOT_DEAD;
  uint4 r;
  r.x = x;
  words_from_slots(sl, r.y, r.z, r.w);
  return r;
}

// (mask >> c) & 1 for a 64-bit bitboard: one funnel shift (SHF.R.U64) instead of building 1 << c
__device__ __forceinline__ bool board_bit(unsigned long long m, uint32_t c) {
  uint32_t lo;  // asm: keeps nvcc from rewriting the test as ((1 << c) & m) != 0 (4 extra instructions)
  asm("{\n\t.reg .u64 t;\n\tshr.u64 t, %1, %2;\n\tcvt.u32.u64 %0, t;\n\t}" : "=r"(lo) : "l"(m), "r"(c));
  return (lo & 1u) != 0u;
}

// utils/interact.py:33-89 for an agent whose target square is NOT floor (counter, cutboard or
// delivery); arglist.play == False.  `hp` = (agent index + 1) << 6 is the 9-bit "place" of the
// agent's hand (cell 0, holder i+1); a lying object's place is its cell (holder 0).  Written as
// a gather (hand slot / target slot), a handful of predicates, and a scatter with recomputed
// compares, so that nothing but slot values stays live across the case analysis.
// Returns true when an object was put on a Delivery square (the only way done() can flip).
template <int NOBJ>
__device__ __forceinline__ bool interact_square(Env<NOBJ>& e, uint32_t hp, uint32_t tgt, bool is_del,
                                                bool is_cut) {
  uint32_t sH = 0, sT = 0;  // slot in hand / slot lying on the target square (0 = none)
#pragma unroll
  for (int k = 0; k < NOBJ; k++) {
    const uint32_t place = e.slot[k] >> 7;
    sH = (place == hp) ? e.slot[k] : sH;
    sT = (place == tgt) ? e.slot[k] : sT;
  }
  const uint32_t mH = sH & 0x7fu, mT = sT & 0x7fu;
  const bool holding = sH != 0u, occupied = sT != 0u;
  // hand side
  const bool chop = is_cut & !occupied & needs_chopped(mH);                 // :63-65 (mH == 0 -> false)
  const bool drop = is_del ? deliverable(mH) : (!occupied & !chop);         // :35-40 deliver, :66-70 put down
  const bool merge = !is_del & occupied & holding & mergeable(mH, mT);      // :43-52
  uint32_t newH = sH;  // sH == 0 (not holding) matches no slot below
  newH = chop ? (sH | (mH << 4)) : newH;
  newH = merge ? (sH | mT) : newH;
  newH = drop ? (mH | (tgt << 7)) : newH;
  // square side: merged away (:48-49) or picked up (:77-84, never from a Delivery)
  const bool pick = !holding & !is_del;
  uint32_t newT = sT;  // sT == 0 (empty square) matches no slot below
  newT = merge ? GC_SLOT_DEAD : newT;
  newT = pick ? (mT | (hp << 7)) : newT;
  // scatter by VALUE (the gathered slot values are unique among live slots whenever they
  // change), which also keeps the compiler from carrying eight compare predicates across
#pragma unroll
  for (int k = 0; k < NOBJ; k++) {
    uint32_t v = e.slot[k];
    v = (v == sT) ? newT : v;
    v = (e.slot[k] == sH) ? newH : v;
    e.slot[k] = v;
  }
  return holding && is_del && drop;
}

// One joint transition: env.step (envs/overcooked_environment.py:255-306) minus the copies.
// Returns the number of CollisionRepr (env:747-752); act[] is overwritten with the executed
// (post-collision) actions; `done`/`success` receive env.done() / env.successful.
// Precondition: the env is not done (so not every Deliver goal is met yet).
template <int NA, int NOBJ, typename LV>
__device__ __forceinline__ uint32_t step(Env<NOBJ>& e, uint32_t (&act)[NA], const LV& L, bool& done,
                                         bool& success) {
  e.t = min(e.t + 1u, 127u);  // env:257
  // where each agent would stand after its own action (is_collision :692-700)
  uint32_t tgt[NA], nxt[NA];
  bool floor_t[NA];
#pragma unroll
  for (int i = 0; i < NA; i++) {
    act[i] = min(act[i], 4u);
    tgt[i] = (e.cell[i] + (uint32_t)action_delta(act[i])) & 63u;
    floor_t[i] = board_bit(L.floor_mask, tgt[i]);
    nxt[i] = floor_t[i] ? tgt[i] : e.cell[i];
  }
  // check_collisions :724-762 - all pairs on the ORIGINAL actions, cancellations applied after
  uint32_t ncoll = 0;
  bool cancel[NA];
#pragma unroll
  for (int i = 0; i < NA; i++) cancel[i] = false;
#pragma unroll
  for (int i = 0; i < NA; i++) {
#pragma unroll
    for (int j = i + 1; j < NA; j++) {
      const bool same = nxt[i] == nxt[j];
      const bool swap = (e.cell[i] == nxt[j]) & (e.cell[j] == nxt[i]);
      const bool bi = (nxt[i] == e.cell[i]) & (act[i] != 4u);  // i faces a square: keeps its action
      const bool bj = (nxt[j] == e.cell[j]) & (act[j] != 4u);
      cancel[i] |= (same & !bi) | (!same & swap);  // :704-717
      cancel[j] |= (same & (bi | !bj)) | (!same & swap);
      ncoll += (same | swap) ? 1u : 0u;
    }
  }
  // execute_navigation :767-770 - sequential in agent order
  bool delivered = false;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    act[i] = cancel[i] ? 4u : act[i];  // :757-761
    e.cell[i] = cancel[i] ? e.cell[i] : nxt[i];  // interact.py:29-30 (nxt == cell unless floor ahead)
    if (act[i] != 4u && !floor_t[i])
      delivered |= interact_square<NOBJ>(e, (uint32_t)(i + 1) << 6, tgt[i], board_bit(L.deliv_mask, tgt[i]),
                                         board_bit(L.cut_mask, tgt[i]));
  }
  // env.done :316-363 - timeout first, then every Deliver goal lying on the delivery square.
  // Goals can only become complete on a step that put something on a Delivery square.
  bool all_goals = false;
  if (delivered) {
    all_goals = true;
#pragma unroll
    for (int g = 0; g < GC_MAX_GOALS; g++) {
      bool found = false;  // unused goal entries repeat goal 0 (gc_levels_to_dev)
#pragma unroll
      for (int k = 0; k < NOBJ; k++) found |= e.slot[k] == L.goal_slot[g];
      all_goals &= found;
    }
  }
  const bool timeout = L.max_t != 0u && e.t >= L.max_t;
  done = timeout || all_goals;
  success = all_goals && !timeout;
  return ncoll;
}

// splitmix64 finaliser
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z ^= z >> 30;
  z *= 0xbf58476d1ce4e5b9ull;
  z ^= z >> 27;
  z *= 0x94d049bb133111ebull;
  z ^= z >> 31;
  return z;
}

__device__ __forceinline__ void cswap(uint32_t& a, uint32_t& b) {
  uint32_t lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}

// Canonical 64-bit state hash (include/gymcook.h "canonical item key").
// Always works on all six slots so that it is independent of the NOBJ specialisation.
template <int NA>
__device__ __forceinline__ unsigned long long state_hash(const uint4& s) {
  uint32_t key[GC_MAX_OBJECTS];
  unsigned long long W0 = (unsigned long long)((s.x >> 24) & 127u) << 52;
  uint32_t hm[GC_MAX_AGENTS] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++) {
    uint32_t sl = slot_of(s, k);
    uint32_t holder = sl >> 13;
    uint32_t cell = (sl >> 7) & 63u, held = 0;
#pragma unroll
    for (int i = 0; i < NA; i++)
      if (holder == (uint32_t)(i + 1)) {
        cell = (s.x >> (6 * i)) & 63u;
        held = 1;
        hm[i] = sl & 0x7fu;
      }
    key[k] = (holder == 7u) ? 0x3FFFu : (((sl & 0x7fu) << 7) | (cell << 1) | held);
  }
#pragma unroll
  for (int i = 0; i < NA; i++)
    W0 |= (unsigned long long)(((s.x >> (6 * i)) & 63u) | (hm[i] << 6)) << (13 * i);
  // 12-comparator sorting network for 6 keys
  cswap(key[0], key[5]); cswap(key[1], key[3]); cswap(key[2], key[4]);
  cswap(key[1], key[2]); cswap(key[3], key[4]);
  cswap(key[0], key[3]); cswap(key[2], key[5]);
  cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[4], key[5]);
  cswap(key[1], key[2]); cswap(key[3], key[4]);
  unsigned long long W1 = key[0] | ((unsigned long long)key[1] << 14) | ((unsigned long long)key[2] << 28);
  unsigned long long W2 = key[3] | ((unsigned long long)key[4] << 14) | ((unsigned long long)key[5] << 28);
  unsigned long long h = mix64(W0 + 0x9E3779B97F4A7C15ull);
  h = mix64(h ^ W1);
  h = mix64(h ^ W2);
  return h;
}

// philox4x32-10 (Salmon et al., SC'11); counter (t, env_lo, env_hi, 0), key = seed.
// action[agent] = mulhi(word[agent], 5).
__device__ __forceinline__ void philox_actions(unsigned long long seed, uint32_t t, unsigned long long env,
                                               uint32_t (&out)[4]) {
  uint32_t c0 = t, c1 = (uint32_t)env, c2 = (uint32_t)(env >> 32), c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = __umulhi(c0, 5u);
  out[1] = __umulhi(c1, 5u);
  out[2] = __umulhi(c2, 5u);
  out[3] = __umulhi(c3, 5u);
}

// streaming 128-bit accesses: every env word is touched exactly once per step, so keep it out
// of L1 (ld.global.nc would be wrong for the in-place state: it is written by this kernel)
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

}  // namespace gc
