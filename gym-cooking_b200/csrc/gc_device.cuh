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
  uint32_t goal_slot[GC_MAX_GOALS];  // goal_mask | first_delivery_cell << 7; unused = 0xFFFFFFFF
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

__device__ __forceinline__ bool bit64(unsigned long long m, uint32_t c) { return (m >> c) & 1ull; }

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
__device__ __forceinline__ bool needs_chopped(uint32_t m) { return m < 8u && ((0x16u >> m) & 1u); }

// Unpacked working form: agent cells and object slots in registers.
template <int NOBJ>
struct Env {
  uint32_t cell[GC_MAX_AGENTS];
  uint32_t slot[NOBJ];
  uint32_t t;
};

template <int NA, int NOBJ>
__device__ __forceinline__ void unpack(const uint4& s, Env<NOBJ>& e) {
#pragma unroll
  for (int i = 0; i < NA; i++) e.cell[i] = (s.x >> (6 * i)) & 63u;
  e.t = (s.x >> 24) & 127u;
  const uint32_t w[3] = {s.y, s.z, s.w};
#pragma unroll
  for (int k = 0; k < NOBJ; k++) e.slot[k] = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu);
}

template <int NA, int NOBJ>
__device__ __forceinline__ uint4 pack(const Env<NOBJ>& e, bool done) {
  uint32_t x = (e.t << 24) | (done ? 0x80000000u : 0u);
#pragma unroll
  for (int i = 0; i < NA; i++) x |= e.cell[i] << (6 * i);
  uint32_t w[3] = {0xE000E000u, 0xE000E000u, 0xE000E000u};
#pragma unroll
  for (int k = 0; k < NOBJ; k += 2) {
    uint32_t hi = (k + 1 < NOBJ) ? e.slot[k + 1] : 0xE000u;
    w[k >> 1] = e.slot[k] | (hi << 16);
  }
  return make_uint4(x, w[0], w[1], w[2]);
}

// utils/interact.py:33-89 for an agent whose target square is NOT floor (counter, cutboard or
// delivery); arglist.play == False.  `hp` = (agent index + 1) << 6 is the 9-bit "place" of the
// agent's hand (cell 0, holder i+1); a lying object's place is its cell (holder 0).
template <int NOBJ>
__device__ __forceinline__ void interact_square(Env<NOBJ>& e, uint32_t hp, uint32_t tgt, bool is_del,
                                                bool is_cut) {
  uint32_t mH = 0, mT = 0;  // mask in hand / mask lying on the target square
  bool isH[NOBJ], isT[NOBJ];
#pragma unroll
  for (int k = 0; k < NOBJ; k++) {
    uint32_t place = e.slot[k] >> 7;
    isH[k] = place == hp;
    isT[k] = place == tgt;
    if (isH[k]) mH = e.slot[k] & 0x7fu;
    if (isT[k]) mT = e.slot[k] & 0x7fu;
  }
  const bool holding = mH != 0u, occupied = mT != 0u;
  uint32_t newH = mH | (hp << 7), newT = 0;
  bool chgT = false;
  if (holding) {
    if (is_del) {  // :35-40 deliver if deliverable, else nothing
      if (deliverable(mH)) newH = mH | (tgt << 7);
    } else if (occupied) {  // :43-52 merge into the hand; the counter object dies
      if (mergeable(mH, mT)) {
        newH |= mT;
        newT = GC_SLOT_DEAD;
        chgT = true;
      }
    } else if (is_cut && needs_chopped(mH)) {  // :63-65 chop in hand
      newH |= mH << 4;
    } else {  // :66-70 put down
      newH = mH | (tgt << 7);
    }
  } else if (occupied && !is_del) {  // :77-84 pick up (never from a Delivery)
    newT = mT | (hp << 7);
    chgT = true;
  }
#pragma unroll
  for (int k = 0; k < NOBJ; k++) {
    if (isH[k]) e.slot[k] = newH;
    else if (isT[k] && chgT) e.slot[k] = newT;
  }
}

// One joint transition: env.step (envs/overcooked_environment.py:255-306) minus the copies.
// Returns the number of CollisionRepr (env:747-752); act[] is overwritten with the executed
// (post-collision) actions; `done`/`success` receive env.done() / env.successful.
template <int NA, int NOBJ, typename LV>
__device__ __forceinline__ uint32_t step(Env<NOBJ>& e, uint32_t (&act)[NA], const LV& L, bool& done,
                                         bool& success) {
  if (e.t < 127u) e.t += 1u;  // env:257
  // where each agent would stand after its own action (is_collision :692-700)
  uint32_t tgt[NA], nxt[NA];
  bool floor_t[NA];
#pragma unroll
  for (int i = 0; i < NA; i++) {
    if (act[i] > 4u) act[i] = 4u;
    tgt[i] = (e.cell[i] + (uint32_t)action_delta(act[i])) & 63u;
    floor_t[i] = bit64(L.floor_mask, tgt[i]);
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
      const bool swap = (e.cell[i] == nxt[j]) && (e.cell[j] == nxt[i]);
      const bool bi = (nxt[i] == e.cell[i]) && (act[i] != 4u);  // i faces a square: keeps its action
      const bool bj = (nxt[j] == e.cell[j]) && (act[j] != 4u);
      const bool ci = same ? !bi : swap;               // :704-717
      const bool cj = same ? (bi || !bj) : swap;
      cancel[i] |= ci;
      cancel[j] |= cj;
      ncoll += (same || swap) ? 1u : 0u;
    }
  }
  // execute_navigation :767-770 - sequential in agent order
#pragma unroll
  for (int i = 0; i < NA; i++) {
    if (cancel[i]) act[i] = 4u;  // :757-761
    if (act[i] != 4u) {
      if (floor_t[i]) {
        e.cell[i] = tgt[i];  // interact.py:29-30
      } else {
        interact_square<NOBJ>(e, (uint32_t)(i + 1) << 6, tgt[i], bit64(L.deliv_mask, tgt[i]),
                              bit64(L.cut_mask, tgt[i]));
      }
    }
  }
  // env.done :316-363 - timeout first, then every Deliver goal lying on the delivery square
  bool all_goals = true;
  for (uint32_t g = 0; g < L.n_goals; g++) {
    bool found = false;
#pragma unroll
    for (int k = 0; k < NOBJ; k++) found |= e.slot[k] == L.goal_slot[g];
    all_goals &= found;
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
  const uint32_t w[3] = {s.y, s.z, s.w};
  uint32_t key[GC_MAX_OBJECTS];
  unsigned long long W0 = (unsigned long long)((s.x >> 24) & 127u) << 52;
  uint32_t hm[GC_MAX_AGENTS] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < GC_MAX_OBJECTS; k++) {
    uint32_t sl = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu);
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
