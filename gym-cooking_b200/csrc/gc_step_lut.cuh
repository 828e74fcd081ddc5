// gc_step_lut.cuh - table-driven form of the env transition for single-level batches.
//
// ncu on the first (pure ALU) version of step_kernel showed the integer ALU pipe at 83 % of its
// half-rate peak with the FMA and LSU pipes idle (profiles/r01_step_kernel_v1_ncu.csv): the
// kernel was bound by ISETP/SEL/LOP3/SHF count, not by HBM.  This version moves work off the
// ALU pipe:
//   * square geometry        -> one shared-memory look-up  move[cell*8 + action] = target | kind<<6
//   * object predicates      -> two look-ups  hprops[mask in hand], tprops[mask on the square]
//   * interact() case split  -> five 1 KB look-ups of 0/1 flags (chop, merge, drop, pick, delivered)
//                               indexed by hprops | tprops | kind<<8
//   * slot gather / scatter  -> 0/1 "is the hand slot" / "is the square slot" flags per slot, then
//                               flag * value + acc  (IMAD, FMA pipe; inline PTX keeps it an IMAD)
// Semantics are those of gc::step (gc_device.cuh), which stays the reference form in the
// multi-level kernels and is compared bit for bit against this one in the GPU tests.
#pragma once
#include "gc_device.cuh"

namespace gclut {

constexpr int kMoveBytes = 512;                          // per level: 64 cells x 8 (5 used)
constexpr int kPropBytes = 128;                          // per table
constexpr int kOutcomeBytes = 1024;                      // per table
constexpr int kStaticBytes = 2 * kPropBytes + 5 * kOutcomeBytes;  // 5376, level independent
constexpr uint32_t kDeadPlace = 0x1C0u;                  // holder 7, cell 0 (slot 0xE000)

// level-independent tables, built at compile time
struct StaticTables {
  uint8_t hprops[kPropBytes];  // b0 holding, b1 foods done, b2 deliverable, b3 needs chopping, b4 plate
  uint8_t tprops[kPropBytes];  // b5 occupied, b6 foods done, b7 plate
  uint8_t chop[kOutcomeBytes], merge[kOutcomeBytes], drop[kOutcomeBytes], pick[kOutcomeBytes],
      delivered[kOutcomeBytes];
};

constexpr bool c_foods_done(uint32_t m) { return ((m & 7u) & ~(m >> 4)) == 0u; }
constexpr int c_popc4(uint32_t m) { return (int)((m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u)); }

constexpr StaticTables make_static_tables() {
  StaticTables t{};
  for (uint32_t m = 0; m < 128; m++) {
    const bool fd = c_foods_done(m), plate = (m & 8u) != 0u;
    const bool deliverable = fd && c_popc4(m) > 1;             // utils/core.py:214-219
    const bool needs_chop = m == 1u || m == 2u || m == 4u;     // utils/core.py:176-178
    t.hprops[m] = (uint8_t)((m != 0u ? 1u : 0u) | (fd ? 2u : 0u) | (deliverable ? 4u : 0u) | (needs_chop ? 8u : 0u) |
                            (plate ? 16u : 0u));
    t.tprops[m] = (uint8_t)((m != 0u ? 32u : 0u) | (fd ? 64u : 0u) | (plate ? 128u : 0u));
  }
  for (uint32_t idx = 0; idx < 1024; idx++) {
    const uint32_t kind = idx >> 8;  // 1 counter, 2 cutboard, 3 delivery (0 = floor: never looked up)
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

// per-level geometry: move[cell*8 + action] = target cell | kind(target) << 6
struct MoveTable {
  uint8_t v[kMoveBytes];
};

// shared-memory image used by the kernels
struct Tables {
  StaticTables st;
  MoveTable mv;
};

// Working form of one env: the object slots stay in their packed 16-bit form, slot = place * 128 + mask
// with place = slot >> 7: the cell of a lying object, (holder << 6) while held, kDeadPlace when dead.
// One IMAD then moves mask and place of a slot together, and pack / the goal test need no re-assembly.
template <int NOBJ>
struct Env {
  uint32_t cell[GC_MAX_AGENTS];
  uint32_t sl[NOBJ];
  uint32_t t;
};

template <int NA, int NOBJ>
__device__ __forceinline__ void unpack(const uint4& s, Env<NOBJ>& e) {
#pragma unroll
  for (int i = 0; i < NA; i++) e.cell[i] = (s.x >> (6 * i)) & 63u;
  e.t = (s.x >> 24) & 127u;
  const uint32_t w[3] = {s.y, s.z, s.w};
#pragma unroll
  for (int k = 0; k < NOBJ; k++) e.sl[k] = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu);
}

template <int NA, int NOBJ>
__device__ __forceinline__ uint4 pack(const Env<NOBJ>& e, bool done) {
  uint32_t x = e.t * 0x1000000u + (done ? 0x80000000u : 0u);
#pragma unroll
  for (int i = 0; i < NA; i++) x += e.cell[i] << (6 * i);
  uint32_t w[3] = {0xE000E000u, 0xE000E000u, 0xE000E000u};
#pragma unroll
  for (int k = 0; k < NOBJ; k += 2) {
    const uint32_t hi = (k + 1 < NOBJ) ? e.sl[k + 1] : 0xE000u;
    w[k >> 1] = hi * 65536u + e.sl[k];
  }
  return make_uint4(x, w[0], w[1], w[2]);
}

// a * b + c that stays an IMAD (inline PTX: the compiler otherwise rewrites flag * delta into SEL + IADD)
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// env.step (envs/overcooked_environment.py:255-306) - same contract as gc::step.
template <int NA, int NOBJ>
__device__ __forceinline__ uint32_t step(Env<NOBJ>& e, uint32_t (&act)[NA], const StaticTables& S,
                                         const uint8_t* mv, const GcLevelDev& L, bool& done, bool& success) {
  e.t = min(e.t + 1u, 127u);  // env:257
  uint32_t tgt[NA], nxt[NA], kind8[NA];  // kind8 = kind(target) << 8, 0 for floor
#pragma unroll
  for (int i = 0; i < NA; i++) {
    act[i] = min(act[i], 4u);
    const uint32_t m = mv[e.cell[i] * 8u + act[i]];
    tgt[i] = m & 63u;
    kind8[i] = (m & 0xC0u) << 2;
    nxt[i] = kind8[i] ? e.cell[i] : tgt[i];  // is_collision :692-700
  }
  // check_collisions :724-762.  "agent i faces a square and keeps its action" (:705-708) is
  // exactly kind8[i] != 0: a non-stay action that leaves the agent in place.
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
      const bool bi = kind8[i] != 0u, bj = kind8[j] != 0u;
      cancel[i] |= (same & !bi) | (!same & swap);
      cancel[j] |= (same & (bi | !bj)) | (!same & swap);
      ncoll += (same | swap) ? 1u : 0u;
    }
  }
  // execute_navigation :767-770 - sequential in agent order
  uint32_t delivered = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    act[i] = cancel[i] ? 4u : act[i];            // :757-761
    e.cell[i] = cancel[i] ? e.cell[i] : nxt[i];  // interact.py:29-30
    if (!cancel[i] && kind8[i] != 0u) {          // utils/interact.py:33-89
      const uint32_t hp = (uint32_t)(i + 1) << 6, tg = tgt[i];
      // 0/1 flags per slot once, then every gather / scatter is an IMAD on the (idle) FMA pipe
      uint32_t fH[NOBJ], fT[NOBJ], gH = 0, gT = 0;
#pragma unroll
      for (int k = 0; k < NOBJ; k++) {
        const uint32_t place = e.sl[k] >> 7;
        fH[k] = place == hp ? 1u : 0u;
        fT[k] = place == tg ? 1u : 0u;
        gH = imad(fH[k], e.sl[k], gH);
        gT = imad(fT[k], e.sl[k], gT);
      }
      // at most one object is in a hand or lies on a counter / cutboard; a Delivery square may hold
      // several (interact.py:38), whose slots add up here - the low 7 bits keep the table index in range
      // (the outcome on a Delivery square does not depend on what lies there, and m = p = 0 below)
      const uint32_t mH = gH & 0x7Fu, mT = gT & 0x7Fu;
      const uint32_t idx = S.hprops[mH] + S.tprops[mT] + kind8[i];
      const uint32_t c = S.chop[idx], m = S.merge[idx], d = S.drop[idx], p = S.pick[idx];
      delivered += S.delivered[idx];
      // hand slot:   mask += chop bits | merged contents;   place: hand -> square when dropped
      // square slot: picked up (square -> hand) or merged away (mask 0, dead place)
      const uint32_t dmaskH = imad(c, mH * 16u, m * mT), dplaceH = d * (tg - hp);
      const uint32_t dmaskT = m * (0u - mT), dplaceT = imad(p, hp - tg, m * (kDeadPlace - tg));
      const uint32_t dslotH = imad(dplaceH, 128u, dmaskH), dslotT = imad(dplaceT, 128u, dmaskT);
#pragma unroll
      for (int k = 0; k < NOBJ; k++) {
        e.sl[k] = imad(fH[k], dslotH, e.sl[k]);
        e.sl[k] = imad(fT[k], dslotT, e.sl[k]);
      }
    }
  }
  // env.done :316-363 (goals can only complete on a step that delivered something)
  bool all_goals = false;
  if (delivered) {
    all_goals = true;
#pragma unroll
    for (int g = 0; g < GC_MAX_GOALS; g++) {
      bool found = false;
#pragma unroll
      for (int k = 0; k < NOBJ; k++) found |= e.sl[k] == L.goal_slot[g];
      all_goals &= found;
    }
  }
  const bool timeout = L.max_t != 0u && e.t >= L.max_t;
  done = timeout || all_goals;
  success = all_goals && !timeout;
  return ncoll;
}

// move[cell*8 + action] = target | kind(target) << 6 from a level's bitboards, filled by the CTA
// (multi-level batches: one table per level is built in the kernel prologue)
// NT = threads of the CTA, a compile-time constant: with blockDim.x the loop bounds cost an integer
// division per loop and ~150 prologue instructions per thread
template <int NT>
__device__ __forceinline__ void fill_move_table_dev(const GcLevelDev& L, uint8_t* mv) {
#pragma unroll
  for (int k0 = 0; k0 < kMoveBytes; k0 += NT) {
    const int k = k0 + (int)threadIdx.x;
    if (kMoveBytes % NT != 0 && k >= kMoveBytes) break;
    const uint32_t c = (uint32_t)k >> 3, a = (uint32_t)k & 7u;
    const uint32_t t = (c + (uint32_t)gc::action_delta(a < 5u ? a : 4u)) & 63u;
    const uint32_t kind = ((L.floor_mask >> t) & 1ull) ? 0u : ((L.cut_mask >> t) & 1ull) ? 2u : ((L.deliv_mask >> t) & 1ull) ? 3u : 1u;
    mv[k] = (uint8_t)(t | (kind << 6));
  }
}

// cooperative load of the tables into shared memory (call from every thread, then __syncthreads)
template <int NT>
__device__ __forceinline__ void load_static_tables(StaticTables* dst, const StaticTables* g_static) {
  static_assert(sizeof(StaticTables) % 16 == 0, "vector copy");
  constexpr int n4 = (int)(sizeof(StaticTables) / 16);
  const uint4* src = reinterpret_cast<const uint4*>(g_static);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int k0 = 0; k0 < n4; k0 += NT) {
    const int k = k0 + (int)threadIdx.x;
    if (k < n4) d4[k] = src[k];
  }
}

template <int NT>
__device__ __forceinline__ void load_tables(Tables* dst, const StaticTables* g_static, const MoveTable& mv_param) {
  load_static_tables<NT>(&dst->st, g_static);
  const uint32_t* ms = reinterpret_cast<const uint32_t*>(&mv_param);
  uint32_t* md = reinterpret_cast<uint32_t*>(&dst->mv);
#pragma unroll
  for (int k0 = 0; k0 < kMoveBytes / 4; k0 += NT) {
    const int k = k0 + (int)threadIdx.x;
    if (k < kMoveBytes / 4) md[k] = ms[k];
  }
}

}  // namespace gclut
