// gc_step2.cuh - the env transition on byte planes (state layout: include/gymcook.h).
//
// The object part of a state is two byte planes: P = the place byte of objects 0..3 and M = their
// content masks (objects 4, 5: the halves of the fourth word).  Every byte has bit 7 clear, so
// "which object is in this hand / lies on this square" is one SIMD byte compare for all objects
//     u = ~(((P ^ pattern) + 0x7F7F7F7F) >> 7) & 0x01010101        (0x01 in every matching byte)
// the gather of its mask is one dot product (dp4a(M, u)), and the scatter of an interaction's
// outcome is two multiply-adds and an XOR per plane (u * delta moves delta into the matching
// byte).  The case split of interact() itself stays the five 1 KB outcome tables of the first
// table-driven version (profiles/r01_*: that kernel needed 58 instructions per agent interaction,
// two thirds of them finding and re-writing 16-bit slots one by one, and ran them under a divergent
// branch at 27 % lane occupancy; this form needs about 35 and is branch-free).
//
// Semantics: env.step (envs/overcooked_environment.py:255-306) = t += 1, check_collisions
// (:724-762, is_collision :671-718), execute_navigation -> interact (utils/interact.py:4-89) in
// agent order, done()/reward() (:316-376).  gc::step (gc_device.cuh) is the plain-ALU restatement
// of the same rules on unpacked slots; the GPU tests compare the two bit for bit.
#pragma once
#include <stdint.h>

#include "../../include/gymcook.h"

#if defined(__CUDACC__)
#define GC_HD __host__ __device__ __forceinline__
#else
#define GC_HD inline
#endif

namespace gcs2 {

constexpr uint32_t kOnes = 0x01010101u;
constexpr uint32_t kLow7 = 0x7F7F7F7Fu;

// level-independent tables, built at compile time (identical to the first table-driven version)
struct StaticTables {
  uint8_t hprops[128];  // b0 holding, b1 foods done, b2 deliverable, b3 needs chopping, b4 plate
  uint8_t tprops[128];  // b5 occupied, b6 foods done, b7 plate
  uint8_t chop[1024], merge[1024], drop[1024], pick[1024], delivered[1024];
};

constexpr bool c_foods_done(uint32_t m) { return ((m & 7u) & ~(m >> 4)) == 0u; }
constexpr int c_popc4(uint32_t m) { return (int)((m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u)); }

constexpr StaticTables make_static_tables() {
  StaticTables t{};
  for (uint32_t m = 0; m < 128; m++) {
    const bool fd = c_foods_done(m), plate = (m & 8u) != 0u;
    const bool deliverable = fd && c_popc4(m) > 1;          // utils/core.py:214-219
    const bool needs_chop = m == 1u || m == 2u || m == 4u;  // utils/core.py:176-178
    t.hprops[m] = (uint8_t)((m != 0u ? 1u : 0u) | (fd ? 2u : 0u) | (deliverable ? 4u : 0u) | (needs_chop ? 8u : 0u) |
                            (plate ? 16u : 0u));
    t.tprops[m] = (uint8_t)((m != 0u ? 32u : 0u) | (fd ? 64u : 0u) | (plate ? 128u : 0u));
  }
  for (uint32_t idx = 0; idx < 1024; idx++) {
    const uint32_t kind = idx >> 8;  // 1 counter, 2 cutboard, 3 delivery; 0 = no interaction: all flags 0
    const bool holding = idx & 1u, fdH = idx & 2u, delivH = idx & 4u, chopH = idx & 8u, plateH = idx & 16u;
    const bool occupied = idx & 32u, fdT = idx & 64u, plateT = idx & 128u;
    uint8_t c = 0, m = 0, d = 0, p = 0, g = 0;
    if (kind != 0u) {
      if (holding) {  // utils/interact.py:33-70
        if (kind == 3u) {
          if (delivH) d = 1, g = 1;
        } else if (occupied) {
          if (!(plateH && plateT) && fdH && fdT) m = 1;  // mergeable, utils/core.py:222-241
        } else if (kind == 2u && chopH) {
          c = 1;
        } else {
          d = 1;
        }
      } else if (occupied && kind != 3u) {  // :73-84
        p = 1;
      }
    }
    t.chop[idx] = c;
    t.merge[idx] = m;
    t.drop[idx] = d;
    t.pick[idx] = p;
    t.delivered[idx] = g;
  }
  return t;
}

// per-level geometry, indexed by cell * 8 + action (actions 5..7 = stay)
struct LevelTables {
  uint8_t tg[512];      // the square the agent faces
  uint8_t nxt[512];     // where it stands afterwards if nothing vetoes: tg when tg is floor, else cell
  uint16_t kind8[512];  // kind(tg) << 8: 0 floor, 0x100 counter, 0x200 cutboard, 0x300 delivery
  uint32_t deliv_b;     // first Delivery cell in every byte (env.done :349)
  uint32_t goal_b[GC_MAX_GOALS];  // goal mask in every byte; unused entries repeat goal 0
  uint32_t max_t24;     // max_num_timesteps << 24 (0 = no limit)
  uint32_t init[4];     // the level's reset state
  uint32_t pad[6];
};
static_assert(sizeof(LevelTables) == 2048 + 64, "LevelTables layout");

struct Tables {
  StaticTables st;
  LevelTables lv;
};

GC_HD uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __dp4a(a, b, c);
#else
  uint32_t r = c;
  for (int k = 0; k < 4; k++) r += ((a >> (8 * k)) & 0xFFu) * ((b >> (8 * k)) & 0xFFu);
  return r;
#endif
}

// a * b + c that stays an IMAD (FMA pipe): the compiler otherwise rewrites flag * delta into SEL + IADD on
// the ALU pipe, which is the busier one in this kernel
GC_HD uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
#else
  return a * b + c;
#endif
}

// 0x01 in every byte of `plane` that equals the matching byte of `pattern` (all bytes < 0x80)
GC_HD uint32_t eq_units(uint32_t plane, uint32_t pattern, uint32_t lanes = kOnes) {
  return ~(((plane ^ pattern) + kLow7) >> 7) & lanes;
}

// Working form of one env: x = word 0 (agent cells, t, done), P/M = planes of objects 0..3,
// P2/M2 = planes of objects 4, 5 in the two low bytes (NOBJ == 6 only).
template <int NOBJ>
struct Env {
  uint32_t x, P, M, P2, M2;
};

template <int NOBJ>
GC_HD void unpack(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, Env<NOBJ>& e) {
  e.x = w0;
  e.P = w1;
  e.M = w2;
  if (NOBJ > 4) {
    e.P2 = w3 & 0xFFFFu;
    e.M2 = w3 >> 16;
  } else {
    e.P2 = w3;  // carried through to pack() untouched (GC_W3_EMPTY in a well-formed state).  Keeping the loaded
    e.M2 = 0;   // word live matters: a dead fourth register of the prefetching 128-bit load gets reused as a
  }             // scratch register while the load is in flight, and that write waits for the load (WAW)
}

// one agent's interaction with the square it faces, branch-free: k8 == 0 (floor ahead, stay, or a vetoed
// action) selects the all-zero rows of the outcome tables
template <int NOBJ>
GC_HD uint32_t interact(Env<NOBJ>& e, int agent, uint32_t tg, uint32_t k8, const StaticTables& S) {
  const uint32_t hpc = GC_PLACE_HELD + 1u + (uint32_t)agent;  // place byte of "in this agent's hand"
  const uint32_t tgb = tg * kOnes;
  const uint32_t uH = eq_units(e.P, hpc * kOnes), uT = eq_units(e.P, tgb);
  uint32_t mH = dp4a_u(e.M, uH, 0u), mT = dp4a_u(e.M, uT, 0u);
  uint32_t uH2 = 0, uT2 = 0;
  if (NOBJ > 4) {
    uH2 = eq_units(e.P2, hpc * kOnes, 0x0101u);
    uT2 = eq_units(e.P2, tgb, 0x0101u);
    mH = dp4a_u(e.M2, uH2, mH);
    mT = dp4a_u(e.M2, uT2, mT);
  }
  // at most one object is in a hand or lies on a counter / cutboard; a Delivery square may hold several
  // (interact.py:38), whose masks add up - the low 7 bits keep the index in range, and the outcome on a
  // Delivery square does not depend on what lies there (merge = pick = 0 below)
  mT &= 0x7Fu;
  const uint32_t idx = (uint32_t)S.hprops[mH] + (uint32_t)S.tprops[mT] + k8;
  const uint32_t c = S.chop[idx], m = S.merge[idx], d = S.drop[idx], p = S.pick[idx];
  // XOR deltas.  hand object: mask gains its chopped bit or the merged contents (disjoint bits), place
  // hand -> square when put down / delivered.  square object: picked up (square -> hand) or merged away
  // (mask -> 0, place -> dead)
  const uint32_t xa = tg ^ hpc, xd = tg ^ GC_PLACE_DEAD;
  const uint32_t dMT = m * mT;
  const uint32_t dMH = imad(c * mH, 16u, dMT);
  const uint32_t dPH = d * xa;
  const uint32_t dPT = imad(p, xa, m * xd);
  e.M ^= imad(uH, dMH, uT * dMT);  // distinct bytes (or both zero): the sum is the XOR
  e.P ^= imad(uH, dPH, uT * dPT);
  if (NOBJ > 4) {
    e.M2 ^= imad(uH2, dMH, uT2 * dMT);
    e.P2 ^= imad(uH2, dPH, uT2 * dPT);
  }
  return S.delivered[idx];
}

// One joint transition of an env that is not done.  `aw` = the raw action word (agent i's action in
// byte i; values > 4 are "stay").  Returns the number of CollisionRepr (env:747-752); `exec` receives the
// executed (post-collision) actions in the same byte form.
template <int NA, int NOBJ, bool WANT_EXEC>
GC_HD uint32_t step(Env<NOBJ>& e, uint32_t aw, const StaticTables& S, const LevelTables& L, bool& done, bool& success,
                    uint32_t& exec) {
  uint32_t cell[NA], act[NA], tg[NA], nxt[NA], k8[NA];
#pragma unroll
  for (int i = 0; i < NA; i++) {
    cell[i] = (e.x >> (6 * i)) & 63u;
    const uint32_t a = (aw >> (8 * i)) & 0xFFu;
    act[i] = a < 4u ? a : 4u;
    const uint32_t idx = cell[i] * 8u + act[i];
    tg[i] = L.tg[idx];
    nxt[i] = L.nxt[idx];
    k8[i] = L.kind8[idx];
  }
  // check_collisions :724-762 - all pairs on the ORIGINAL actions.  "agent i faces a square and keeps its
  // action" (:705-708) is exactly k8[i] != 0: a non-stay action that leaves the agent in place.  With three
  // or more agents the pairwise rule lets two agents end up on one square, so `same` and `swap` can both
  // hold: the reference tests `same` first (:704) and `swap` only otherwise (:714).
  uint32_t ncoll = 0;
  bool cancel[NA];
#pragma unroll
  for (int i = 0; i < NA; i++) cancel[i] = false;
#pragma unroll
  for (int i = 0; i < NA; i++) {
#pragma unroll
    for (int j = i + 1; j < NA; j++) {
      const bool same = nxt[i] == nxt[j];
      const bool swap = (cell[i] == nxt[j]) & (cell[j] == nxt[i]);
      const bool bi = k8[i] != 0u, bj = k8[j] != 0u;
      cancel[i] |= same ? !bi : swap;
      cancel[j] |= same ? (bi | !bj) : swap;
      if (WANT_EXEC) ncoll += (same | swap) ? 1u : 0u;
    }
  }
  // execute_navigation :767-770 - sequential in agent order
  uint32_t delivered = 0, lo = 0;
  exec = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    const uint32_t c = cancel[i] ? cell[i] : nxt[i];  // interact.py:29-30 (nxt == cell unless floor ahead)
    lo += c << (6 * i);
    if (WANT_EXEC) exec |= (cancel[i] ? 4u : act[i]) << (8 * i);  // :757-761
    delivered += interact<NOBJ>(e, i, tg[i], cancel[i] ? 0u : k8[i], S);
  }
  // env.done :316-363 - timeout first, then every Deliver goal lying on the first Delivery square.
  // Goals can only become complete on a step that put something there.
  bool all_goals = false;
  if (delivered) {
    const uint32_t on = eq_units(e.P, L.deliv_b), on2 = NOBJ > 4 ? eq_units(e.P2, L.deliv_b, 0x0101u) : 0u;
    all_goals = true;
#pragma unroll
    for (int g = 0; g < GC_MAX_GOALS; g++) {
      uint32_t hit = on & eq_units(e.M, L.goal_b[g]);
      if (NOBJ > 4) hit |= on2 & eq_units(e.M2, L.goal_b[g], 0x0101u);
      all_goals &= hit != 0u;
    }
  }
  // t += 1 (env:257), saturating at 127 (the done bit is clear on entry, so the add cannot carry out)
  uint32_t hi = e.x + 0x01000000u;
  hi = (hi < 0x7FFFFFFFu ? hi : 0x7FFFFFFFu) & 0x7F000000u;
  const bool timeout = L.max_t24 != 0u && hi >= L.max_t24;
  done = timeout || all_goals;
  success = all_goals && !timeout;
  e.x = hi | lo | (done ? 0x80000000u : 0u);
  return ncoll;
}

template <int NOBJ>
GC_HD void pack(const Env<NOBJ>& e, uint32_t& w0, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  w0 = e.x;
  w1 = e.P;
  w2 = e.M;
  w3 = NOBJ > 4 ? (e.P2 | (e.M2 << 16)) : e.P2;
}

// the level's reset state in packed form (env.reset :201-250): agents on their start cells, objects lying
// where the level file puts them (gc_level.object_init = mask | cell << 7), t = 0
inline void initial_state(const gc_level& s, int n_agents, uint32_t w[4]) {
  uint32_t P[GC_MAX_OBJECTS], M[GC_MAX_OBJECTS];
  for (int k = 0; k < GC_MAX_OBJECTS; k++) {
    const bool live = k < s.n_objects;
    P[k] = live ? (((uint32_t)s.object_init[k] >> 7) & 63u) : GC_PLACE_DEAD;
    M[k] = live ? ((uint32_t)s.object_init[k] & 0x7Fu) : 0u;
  }
  w[0] = 0;
  for (int i = 0; i < n_agents; i++) w[0] |= (uint32_t)(s.agent_cell[i] & 63) << (6 * i);
  w[1] = P[0] | P[1] << 8 | P[2] << 16 | P[3] << 24;
  w[2] = M[0] | M[1] << 8 | M[2] << 16 | M[3] << 24;
  w[3] = P[4] | P[5] << 8 | M[4] << 16 | M[5] << 24;
}

// host: the per-level tables from a gc_level
inline void fill_level_tables(const gc_level& s, int n_agents, LevelTables* out) {
  static const int delta[5] = {8, -8, -1, 1, 0};
  for (int c = 0; c < 64; c++)
    for (int a = 0; a < 8; a++) {
      const int t = (c + delta[a < 5 ? a : 4]) & 63;
      const int kind = s.cell_type[t];
      out->tg[c * 8 + a] = (uint8_t)t;
      out->nxt[c * 8 + a] = (uint8_t)(kind == GC_CELL_FLOOR ? t : c);
      out->kind8[c * 8 + a] = (uint16_t)(kind << 8);
    }
  out->deliv_b = (uint32_t)s.delivery_cell * kOnes;
  for (int g = 0; g < GC_MAX_GOALS; g++) out->goal_b[g] = (uint32_t)s.goal_mask[g < s.n_goals ? g : 0] * kOnes;
  out->max_t24 = (uint32_t)s.max_timesteps << 24;
  initial_state(s, n_agents, out->init);
  for (int k = 0; k < 6; k++) out->pad[k] = 0;
}

}  // namespace gcs2
