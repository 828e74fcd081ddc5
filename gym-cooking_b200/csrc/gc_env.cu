// gc_env.cu - path A kernels: reset, step, fused rollout, action stream, hash, statistics.
// One thread per env; a warp is a tile of 32 envs whose 128-bit states are read and written
// with one fully coalesced 512-byte transaction each way.  Static level tables arrive as
// kernel parameters (constant bank, warp-uniform) when the batch has one level, and are
// staged through shared memory when envs carry a per-env level id.
#include "gc_device.cuh"
#include "gc_host.h"
#include "gc_step_lut.cuh"

#include <stdlib.h>

namespace {

constexpr int kThreads = 256;

template <int NA>
__device__ __forceinline__ void load_actions(const uint8_t* __restrict__ actions, int64_t i, uint32_t (&a)[NA]) {
  if constexpr (NA == 1) {
    a[0] = actions[i];
  } else if constexpr (NA == 2) {
    uint32_t v = reinterpret_cast<const uint16_t*>(actions)[i];
    a[0] = v & 0xffu;
    a[1] = v >> 8;
  } else if constexpr (NA == 4) {
    uint32_t v = reinterpret_cast<const uint32_t*>(actions)[i];
    a[0] = v & 0xffu;
    a[1] = (v >> 8) & 0xffu;
    a[2] = (v >> 16) & 0xffu;
    a[3] = v >> 24;
  } else {
#pragma unroll
    for (int k = 0; k < NA; k++) a[k] = actions[i * NA + k];
  }
}

// the same in two halves: the global load (issued one iteration ahead by the persistent kernels, kept
// as the raw word so that nothing waits on it early) and the byte extraction at the point of use
template <int NA>
__device__ __forceinline__ uint32_t load_actions_raw(const uint8_t* __restrict__ actions, int64_t i) {
  if constexpr (NA == 1) {
    return actions[i];
  } else if constexpr (NA == 2) {
    return reinterpret_cast<const uint16_t*>(actions)[i];
  } else if constexpr (NA == 4) {
    return reinterpret_cast<const uint32_t*>(actions)[i];
  } else {
    return (uint32_t)actions[i * 3] | ((uint32_t)actions[i * 3 + 1] << 8) | ((uint32_t)actions[i * 3 + 2] << 16);
  }
}
template <int NA>
__device__ __forceinline__ void unpack_actions(uint32_t v, uint32_t (&a)[NA]) {
#pragma unroll
  for (int k = 0; k < NA; k++) a[k] = (v >> (8 * k)) & 0xffu;
}
template <int NA>
__device__ __forceinline__ void store_actions(uint8_t* __restrict__ out, int64_t i, const uint32_t (&a)[NA]) {
#pragma unroll
  for (int k = 0; k < NA; k++) out[i * NA + k] = (uint8_t)a[k];
}

// ---------------------------------------------------------------------------------------
// reset: every env := level.init                                          (env.reset :201-250)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
reset_kernel(const __grid_constant__ GcLevelsDev levels, const uint8_t* __restrict__ level_id,
             uint4* __restrict__ state, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const GcLevelDev& L = levels.lv[level_id ? level_id[i] : 0];
  state[i] = make_uint4(L.init[0], L.init[1], L.init[2], L.init[3]);
}

// ---------------------------------------------------------------------------------------
// step                                                                       (env.step :255-306)
// ---------------------------------------------------------------------------------------
// Persistent, software-pipelined: a thread walks envs i, i + stride, ... and issues the loads of
// its NEXT env before computing the current one, so DRAM latency hides behind ~300 ALU
// instructions of the same warp instead of needing a fresh wave of CTAs (at 2^20 envs a
// one-env-per-thread grid is only 3.5 waves deep and spends half its time in ramp and tail).
template <int NA, int NOBJ, bool MULTI>
__global__ void __launch_bounds__(kThreads)
step_kernel(const __grid_constant__ GcLevelsDev levels, const uint8_t* __restrict__ level_id,
            uint4* __restrict__ state, const uint8_t* __restrict__ actions,
            uint8_t* __restrict__ reward_done, unsigned long long* __restrict__ hash,
            uint32_t* __restrict__ collisions, uint8_t* __restrict__ executed, int64_t n) {
  __shared__ GcLevelDev s_levels[MULTI ? GC_MAX_LEVELS : 1];
  if constexpr (MULTI) {
    // stage the (<= 1 KB) level tables once per CTA: per-lane level ids would serialise
    // constant-bank reads, shared memory serves divergent indices at full rate
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&levels);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_levels);
    for (int k = threadIdx.x; k < (int)(sizeof(GcLevelsDev) / 4); k += kThreads) dst[k] = src[k];
    __syncthreads();
  }
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  uint4 s_next = gc::ld_stream(state + i);
  uint32_t a_next[NA];
  load_actions<NA>(actions, i, a_next);
  uint32_t lvl_next = 0;
  if constexpr (MULTI) lvl_next = level_id[i];

  for (; i < n; i += stride) {
    uint4 s = s_next;
    uint32_t act[NA];
#pragma unroll
    for (int k = 0; k < NA; k++) act[k] = a_next[k];
    const uint32_t lvl = lvl_next;
    const int64_t inext = i + stride;
    if (inext < n) {  // prefetch
      s_next = gc::ld_stream(state + inext);
      load_actions<NA>(actions, inext, a_next);
      if constexpr (MULTI) lvl_next = level_id[inext];
    }
    const GcLevelDev& L = MULTI ? s_levels[lvl] : levels.lv[0];

    bool done, success;
    if (s.x >> 31) {
      // sticky done: the episode is over, nothing mutates; re-report the stored outcome
      const uint32_t t = (s.x >> 24) & 127u;
      done = true;
      success = !(L.max_t != 0u && t >= L.max_t);
#pragma unroll
      for (int k = 0; k < NA; k++) act[k] = 4u;
    } else {
      gc::Env<NOBJ> e;
      gc::unpack<NA, NOBJ>(s, e);
      const uint32_t ncoll = gc::step<NA, NOBJ>(e, act, L, done, success);
      s = gc::pack<NA, NOBJ>(e, done);
      gc::st_stream(state + i, s);
      if (collisions && ncoll) collisions[i] += ncoll;
    }
    if (reward_done) reward_done[i] = (uint8_t)((done ? GC_RD_DONE : 0) | (success ? GC_RD_REWARD : 0));
    if (hash) hash[i] = gc::state_hash<NA>(s);
    if (executed) store_actions<NA>(executed, i, act);
  }
}

// ---------------------------------------------------------------------------------------
// rollout: n_steps fused transitions, state in registers, philox actions in-kernel
// ---------------------------------------------------------------------------------------
template <int NA, int NOBJ, bool MULTI>
__global__ void __launch_bounds__(kThreads)
rollout_kernel(const __grid_constant__ GcLevelsDev levels, const uint8_t* __restrict__ level_id,
               uint4* __restrict__ state, uint8_t* __restrict__ reward_done,
               unsigned long long* __restrict__ hash_trace, uint32_t* __restrict__ collisions,
               int64_t n, int n_steps, uint32_t t0, int64_t env0, unsigned long long seed) {
  __shared__ GcLevelDev s_levels[MULTI ? GC_MAX_LEVELS : 1];
  if constexpr (MULTI) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&levels);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_levels);
    for (int k = threadIdx.x; k < (int)(sizeof(GcLevelsDev) / 4); k += kThreads) dst[k] = src[k];
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const GcLevelDev& L = MULTI ? s_levels[level_id[i]] : levels.lv[0];

  uint4 s = gc::ld_stream(state + i);
  gc::Env<NOBJ> e;
  gc::unpack<NA, NOBJ>(s, e);
  bool done = s.x >> 31;
  bool success = done && !(L.max_t != 0u && e.t >= L.max_t);
  uint32_t ncoll = 0;
  for (int k = 0; k < n_steps; k++) {
    if (!done) {
      uint32_t r[4], act[NA];
      gc::philox_actions(seed, t0 + (uint32_t)k, (unsigned long long)(env0 + i), r);
#pragma unroll
      for (int a = 0; a < NA; a++) act[a] = r[a];
      ncoll += gc::step<NA, NOBJ>(e, act, L, done, success);
    }
    if (hash_trace) hash_trace[(int64_t)k * n + i] = gc::state_hash<NA>(gc::pack<NA, NOBJ>(e, done));
  }
  gc::st_stream(state + i, gc::pack<NA, NOBJ>(e, done));
  if (reward_done) reward_done[i] = (uint8_t)((done ? GC_RD_DONE : 0) | (success ? GC_RD_REWARD : 0));
  if (collisions && ncoll) collisions[i] += ncoll;
}

__global__ void __launch_bounds__(kThreads)
fill_actions_kernel(uint8_t* __restrict__ actions, int64_t n, int n_agents, int n_steps, uint32_t t0,
                    int64_t env0, unsigned long long seed) {
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= n * n_steps) return;
  const int64_t k = idx / n, i = idx % n;
  uint32_t r[4];
  gc::philox_actions(seed, t0 + (uint32_t)k, (unsigned long long)(env0 + i), r);
  for (int a = 0; a < n_agents; a++) actions[idx * n_agents + a] = (uint8_t)r[a];
}

template <int NA>
__global__ void __launch_bounds__(kThreads)
hash_kernel(const uint4* __restrict__ state, unsigned long long* __restrict__ hash, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  hash[i] = gc::state_hash<NA>(state[i]);
}

// ---------------------------------------------------------------------------------------
// episode statistics: warp-shuffle reduce -> shared-memory histogram -> one atomic per bin
// per CTA.  stats layout: include/gymcook.h (GC_STATS_LEN).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
stats_kernel(const __grid_constant__ GcLevelsDev levels, const uint8_t* __restrict__ level_id,
             const uint4* __restrict__ state, const uint32_t* __restrict__ collisions,
             unsigned long long* __restrict__ stats, int64_t n) {
  __shared__ unsigned int s_hist[128];
  __shared__ unsigned long long s_acc[5];
  for (int k = threadIdx.x; k < 128; k += kThreads) s_hist[k] = 0;
  if (threadIdx.x < 5) s_acc[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long v[5] = {0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const uint32_t w0 = state[i].x;
    const GcLevelDev& L = levels.lv[level_id ? level_id[i] : 0];
    const uint32_t t = (w0 >> 24) & 127u;
    const bool done = w0 >> 31;
    v[0] += 1;
    if (done) {
      v[1] += !(L.max_t != 0u && t >= L.max_t);
      v[2] += t;
      atomicAdd(&s_hist[t], 1u);
    } else {
      v[4] += 1;
    }
    if (collisions) v[3] += collisions[i];
  }
#pragma unroll
  for (int k = 0; k < 5; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(&s_acc[k], v[k]);
  }
  __syncthreads();
  if (threadIdx.x < 5 && s_acc[threadIdx.x]) atomicAdd(&stats[threadIdx.x], s_acc[threadIdx.x]);
  for (int k = threadIdx.x; k < 128; k += kThreads)
    if (s_hist[k]) atomicAdd(&stats[5 + k], (unsigned long long)s_hist[k]);
}

// reward/done bytes -> bit planes: bits[2*w] = done of envs 32w..32w+31, bits[2*w+1] = reward
__global__ void __launch_bounds__(kThreads)
pack_rd_kernel(const uint8_t* __restrict__ rd, uint32_t* __restrict__ bits, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const uint32_t v = i < n ? rd[i] : 0u;  // every lane votes: no early exit before the ballots
  const uint32_t d = __ballot_sync(0xffffffffu, v & GC_RD_DONE), r = __ballot_sync(0xffffffffu, v & GC_RD_REWARD);
  if ((threadIdx.x & 31) == 0 && i < n) {
    bits[2 * (i >> 5)] = d;
    bits[2 * (i >> 5) + 1] = r;
  }
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

// bit planes from the reward/done bytes, for the step kernels that do not write them themselves
inline void pack_rd(const uint8_t* rd, uint32_t* bits, int64_t n, cudaStream_t st) {
  const int64_t words = (n + 31) / 32;
  pack_rd_kernel<<<(unsigned)((words * 32 + kThreads - 1) / kThreads), kThreads, 0, st>>>(rd, bits, n);
}

// Grid of the step kernel.  Measured on B200 (profiles/r01_step_kernel.md): the kernel is bound
// by the integer ALU pipe, not by DRAM latency, so one env per thread (a full grid) beats a
// persistent grid-stride loop with register prefetch by ~10 %.  GC_STEP_CTAS_PER_SM=k (1..8)
// switches to the persistent form with k CTAs per SM for experiments.
inline unsigned step_grid(int64_t n) {
  static int sms = 0, per_sm = -1;
  if (per_sm < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const char* e = getenv("GC_STEP_CTAS_PER_SM");
    per_sm = e ? atoi(e) : 0;
    if (per_sm < 0 || per_sm > 8) per_sm = 0;
  }
  const unsigned full = grid_for(n);
  if (per_sm == 0) return full;
  const unsigned cap = (unsigned)(sms * per_sm);
  return full < cap ? full : cap;
}

// ---------------------------------------------------------------------------------------
// step, table-driven (single-level batches): see gc_step_lut.cuh
// ---------------------------------------------------------------------------------------
__device__ const gclut::StaticTables g_static_tables = gclut::make_static_tables();

struct StepLutParams {
  GcLevelDev lv;
  gclut::MoveTable mv;
};

#ifndef GC_LUT_MIN_CTAS
#define GC_LUT_MIN_CTAS 4
#endif
// The plain step (no optional outputs) fits 40 registers without spilling for <= 2 agents and <= 4
// objects (6 CTAs of 256 threads per SM) and 48 registers otherwise (5 CTAs; <4,6> spills 12 bytes).
#ifdef GC_LUT_MIN_CTAS_PLAIN_ALL  // experiments: one bound for every instantiation
#define GC_LUT_MIN_CTAS_PLAIN(NA, NOBJ) (GC_LUT_MIN_CTAS_PLAIN_ALL)
#else
#define GC_LUT_MIN_CTAS_PLAIN(NA, NOBJ) (((NA) <= 2 && (NOBJ) <= 4) ? 6 : 5)
#endif
#ifndef GC_LUT_THREADS
#define GC_LUT_THREADS 256
#endif
#ifndef GC_LUT_L2_PREFETCH
#define GC_LUT_L2_PREFETCH 1  // tiles per thread prefetched into L2 ahead of griddepcontrol.wait
#endif
constexpr int kLutThreads = GC_LUT_THREADS;  // block size of the single-level table-driven step kernel
// EXTRAS = false is the plain gym step (state in place + reward/done byte): the optional outputs
// (hash, collision counters, executed actions) and everything computed only for them drop out of
// the loop at compile time instead of costing uniform branches and registers.
// BITS (plain step only) also writes the results as two bit planes per 32 envs - rd_bits[2*w] = done,
// rd_bits[2*w+1] = reward of envs 32w..32w+31, the format gc_env_step_host sends over PCIe - with two
// ballots per warp; the warp then walks its tiles in lockstep (lanes past n idle but vote).
template <int NA, int NOBJ, bool EXTRAS, bool BITS>
__global__ void __launch_bounds__(kLutThreads, EXTRAS ? GC_LUT_MIN_CTAS : GC_LUT_MIN_CTAS_PLAIN(NA, NOBJ))
step_lut_kernel(const __grid_constant__ StepLutParams P, uint4* __restrict__ state,
                const uint8_t* __restrict__ actions, uint8_t* __restrict__ reward_done,
                unsigned long long* __restrict__ hash, uint32_t* __restrict__ collisions,
                uint8_t* __restrict__ executed, uint32_t* __restrict__ rd_bits,
                uint32_t n) {  // n < 2^31: the host splits larger batches
  static_assert(!(EXTRAS && BITS), "bit planes come with the plain step only");
  __shared__ __align__(16) gclut::Tables T;
  __shared__ __align__(16) uint4 s_stage[kLutThreads];  // each thread's NEXT state, filled by cp.async
  // Persistent CTAs: the 5.9 KB of tables are loaded once per CTA and reused for every env the CTA
  // walks (a one-env-per-thread grid spent ~20 % of its instructions refilling them).  DRAM latency
  // is hidden by software pipelining WITHOUT registers: while a thread computes env i, the 16-byte
  // state of its next env streams into its private shared-memory slot with cp.async (LDGSTS), and
  // the next action word waits in one register.  (A register-prefetch variant pushed the kernel
  // over 32 registers / 100 % occupancy and lost more than it won.)
  const uint32_t stride = gridDim.x * kLutThreads;  // 32-bit indices: one IMAD.WIDE per address
  uint32_t i = blockIdx.x * kLutThreads + threadIdx.x;
  const uint32_t slot = (uint32_t)__cvta_generic_to_shared(&s_stage[threadIdx.x]);
  uint32_t a_next = 0x04040404u;  // raw action word of the next env (all "stay")
  // Programmatic dependent launch: the tables do not depend on earlier kernels, so this grid may
  // start (and fill them) while the previous kernel of the stream drains; everything that can
  // have been written by it (state, actions) is read only after griddepcontrol.wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  gclut::load_tables<kLutThreads>(&T, &g_static_tables, P.mv);
#if GC_LUT_L2_PREFETCH > 0
  // While the previous kernel drains, pull this CTA's first tiles from DRAM into L2.  L2 is the
  // coherence point of the device, so a line prefetched there can never be stale: whatever the
  // previous kernel still writes lands in the same L2 line.  Nothing is READ before the wait.
#pragma unroll
  for (int d = 0; d < GC_LUT_L2_PREFETCH; d++) {
    const uint32_t ip = i + (uint32_t)d * stride;
    if (ip < n && ip >= i) {
      if ((threadIdx.x & 7u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(state + ip));
      if ((threadIdx.x & 31u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(actions + (size_t)ip * NA));
    }
  }
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (i < n) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(state + i) : "memory");
    a_next = load_actions_raw<NA>(actions, i);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  __syncthreads();
  const GcLevelDev& L = P.lv;
  const uint32_t lane = threadIdx.x & 31u;
  for (; BITS ? (i - lane < n) : (i < n); i += stride) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    uint4 s = s_stage[threadIdx.x];  // written by this thread's own cp.async: no CTA barrier needed
    const bool valid = !BITS || i < n;
    if (BITS && !valid) s.x = 0x80000000u;  // a lane past the end: nothing to compute, nothing stored
    uint32_t act[NA];
    unpack_actions<NA>(a_next, act);
    const uint32_t inext = i + stride;
    if (inext < n) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(state + inext) : "memory");
      a_next = load_actions_raw<NA>(actions, inext);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    bool done, success;
    if (s.x >> 31) {
      const uint32_t t = (s.x >> 24) & 127u;
      done = true;
      success = !(L.max_t != 0u && t >= L.max_t);
#pragma unroll
      for (int k = 0; k < NA; k++) act[k] = 4u;
    } else {
      gclut::Env<NOBJ> e;
      gclut::unpack<NA, NOBJ>(s, e);
      const uint32_t ncoll = gclut::step<NA, NOBJ>(e, act, T.st, T.mv.v, L, done, success);
      s = gclut::pack<NA, NOBJ>(e, done);
      gc::st_stream(state + i, s);
      if constexpr (EXTRAS) {
        if (collisions && ncoll) collisions[i] += ncoll;
      }
    }
    const uint8_t rd = (uint8_t)((done ? GC_RD_DONE : 0) | (success ? GC_RD_REWARD : 0));
    if constexpr (EXTRAS) {
      if (reward_done) reward_done[i] = rd;
      if (hash) hash[i] = gc::state_hash<NA>(s);
      if (executed) store_actions<NA>(executed, i, act);
    } else {
      if (valid) reward_done[i] = rd;
      if constexpr (BITS) {
        const uint32_t d = __ballot_sync(0xffffffffu, done && valid), r = __ballot_sync(0xffffffffu, success && valid);
        if (lane == 0u) *reinterpret_cast<uint2*>(rd_bits + 2u * (i >> 5)) = make_uint2(d, r);
      }
    }
  }
}


// Table-driven step for multi-level batches (per-env level_id).  Same pipeline as step_lut_kernel;
// the level tables (64 B of bitboards / goals each) arrive as a kernel parameter, every CTA stages
// them in shared memory and derives one 512 B move table per level in its prologue - nothing is
// uploaded per launch.  Dynamic shared memory: n_levels x 512 B.
struct MultiShared {
  gclut::StaticTables st;
  GcLevelDev lv[GC_MAX_LEVELS];
};

template <int NA, int NOBJ, bool EXTRAS>
__global__ void __launch_bounds__(kThreads, EXTRAS ? 4 : 5)
step_lut_multi_kernel(const __grid_constant__ GcLevelsDev P, int n_levels, const uint8_t* __restrict__ level_id,
                      uint4* __restrict__ state, const uint8_t* __restrict__ actions,
                      uint8_t* __restrict__ reward_done, unsigned long long* __restrict__ hash,
                      uint32_t* __restrict__ collisions, uint8_t* __restrict__ executed, uint32_t n) {
  __shared__ __align__(16) MultiShared S;
  __shared__ __align__(16) uint4 s_stage[kThreads];
  extern __shared__ __align__(16) uint8_t s_mv[];  // [n_levels][kMoveBytes]
  const uint32_t stride = gridDim.x * kThreads;
  uint32_t i = blockIdx.x * kThreads + threadIdx.x;
  const uint32_t slot = (uint32_t)__cvta_generic_to_shared(&s_stage[threadIdx.x]);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    gclut::load_static_tables<kThreads>(&S.st, &g_static_tables);
    const uint32_t* ls = reinterpret_cast<const uint32_t*>(&P);
    uint32_t* ld = reinterpret_cast<uint32_t*>(S.lv);
#pragma unroll 1
    for (int k = threadIdx.x; k < n_levels * (int)(sizeof(GcLevelDev) / 4); k += kThreads) ld[k] = ls[k];
    __syncthreads();
    for (int l = 0; l < n_levels; l++) gclut::fill_move_table_dev<kThreads>(S.lv[l], s_mv + l * gclut::kMoveBytes);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint32_t a_next = 0x04040404u, l_next = 0;
  if (i < n) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(state + i) : "memory");
    a_next = load_actions_raw<NA>(actions, i);
    l_next = level_id[i];
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  __syncthreads();
  for (; i < n; i += stride) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    uint4 s = s_stage[threadIdx.x];
    uint32_t act[NA];
    unpack_actions<NA>(a_next, act);
    const uint32_t lvl = min(l_next, (uint32_t)(n_levels - 1));
    const uint32_t inext = i + stride;
    if (inext < n) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(state + inext) : "memory");
      a_next = load_actions_raw<NA>(actions, inext);
      l_next = level_id[inext];
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const GcLevelDev& L = S.lv[lvl];
    bool done, success;
    if (s.x >> 31) {
      const uint32_t t = (s.x >> 24) & 127u;
      done = true;
      success = !(L.max_t != 0u && t >= L.max_t);
#pragma unroll
      for (int k = 0; k < NA; k++) act[k] = 4u;
    } else {
      gclut::Env<NOBJ> e;
      gclut::unpack<NA, NOBJ>(s, e);
      const uint32_t ncoll = gclut::step<NA, NOBJ>(e, act, S.st, s_mv + lvl * gclut::kMoveBytes, L, done, success);
      s = gclut::pack<NA, NOBJ>(e, done);
      gc::st_stream(state + i, s);
      if constexpr (EXTRAS) {
        if (collisions && ncoll) collisions[i] += ncoll;
      }
    }
    const uint8_t rd = (uint8_t)((done ? GC_RD_DONE : 0) | (success ? GC_RD_REWARD : 0));
    if constexpr (EXTRAS) {
      if (reward_done) reward_done[i] = rd;
      if (hash) hash[i] = gc::state_hash<NA>(s);
      if (executed) store_actions<NA>(executed, i, act);
    } else {
      reward_done[i] = rd;
    }
  }
}

// Fused rollout, table-driven form (single-level batches): same philox stream and the same
// transitions as rollout_kernel, with gclut::step instead of gc::step.
template <int NA, int NOBJ>
__global__ void __launch_bounds__(kThreads)
rollout_lut_kernel(const __grid_constant__ StepLutParams P, uint4* __restrict__ state,
                   uint8_t* __restrict__ reward_done, unsigned long long* __restrict__ hash_trace,
                   uint32_t* __restrict__ collisions, int64_t n, int n_steps, uint32_t t0, int64_t env0,
                   unsigned long long seed) {
  __shared__ __align__(16) gclut::Tables T;
  gclut::load_tables<kThreads>(&T, &g_static_tables, P.mv);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const GcLevelDev& L = P.lv;
  uint4 s = gc::ld_stream(state + i);
  gclut::Env<NOBJ> e;
  gclut::unpack<NA, NOBJ>(s, e);
  bool done = s.x >> 31;
  bool success = done && !(L.max_t != 0u && e.t >= L.max_t);
  uint32_t ncoll = 0;
  for (int k = 0; k < n_steps; k++) {
    if (!done) {
      uint32_t r[4], act[NA];
      gc::philox_actions(seed, t0 + (uint32_t)k, (unsigned long long)(env0 + i), r);
#pragma unroll
      for (int a = 0; a < NA; a++) act[a] = r[a];
      ncoll += gclut::step<NA, NOBJ>(e, act, T.st, T.mv.v, L, done, success);
    }
    if (hash_trace) hash_trace[(int64_t)k * n + i] = gc::state_hash<NA>(gclut::pack<NA, NOBJ>(e, done));
  }
  gc::st_stream(state + i, gclut::pack<NA, NOBJ>(e, done));
  if (reward_done) reward_done[i] = (uint8_t)((done ? GC_RD_DONE : 0) | (success ? GC_RD_REWARD : 0));
  if (collisions && ncoll) collisions[i] += ncoll;
}

// host: move[cell*8 + action] = target | kind(target) << 6 from the level's bitboards
void fill_move_table(const GcLevelDev& L, gclut::MoveTable* mv) {
  static const int delta[5] = {8, -8, -1, 1, 0};
  for (int c = 0; c < 64; c++)
    for (int a = 0; a < 8; a++) {
      const int t = (c + delta[a < 5 ? a : 4]) & 63;
      const unsigned long long b = 1ull << t;
      const int kind = (L.floor_mask & b) ? 0 : (L.cut_mask & b) ? 2 : (L.deliv_mask & b) ? 3 : 1;
      mv->v[c * 8 + a] = (uint8_t)(t | (kind << 6));
    }
}

// persistent grid of step_lut_kernel: as many CTAs as are resident at once (occupancy of the
// instantiation: 5 per SM for the plain step at 48 registers, 4 with the optional outputs), unless
// GC_LUT_CTAS_PER_SM overrides it
template <int NA, int NOBJ, bool EXTRAS, bool BITS>
unsigned lut_step_grid(int64_t n) {
  static int resident = 0;  // CTAs per device
  if (!resident) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const char* e = getenv("GC_LUT_CTAS_PER_SM");
    if (e) per_sm = atoi(e);
    if (per_sm < 1 || per_sm > 8) {
      per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_lut_kernel<NA, NOBJ, EXTRAS, BITS>, kLutThreads, 0) !=
              cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 4;
      }
    }
    resident = sms * per_sm;
  }
  const unsigned full = (unsigned)((n + kLutThreads - 1) / kLutThreads);
  return full < (unsigned)resident ? full : (unsigned)resident;
}

template <int NA, int NOBJ, bool EXTRAS>
unsigned lut_multi_grid(int64_t n, size_t dyn_smem) {
  static int resident = 0;
  static size_t resident_smem = 0;
  if (!resident || resident_smem != dyn_smem) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const char* e = getenv("GC_LUT_CTAS_PER_SM");
    if (e) per_sm = atoi(e);
    if (per_sm < 1 || per_sm > 8) {
      per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_lut_multi_kernel<NA, NOBJ, EXTRAS>, kThreads,
                                                        dyn_smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 4;
      }
    }
    resident = sms * per_sm;
    resident_smem = dyn_smem;
  }
  const unsigned full = (unsigned)((n + kThreads - 1) / kThreads);
  return full < (unsigned)resident ? full : (unsigned)resident;
}

inline bool use_generic_step() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GC_STEP_GENERIC");  // 1 = the table-free reference form of the kernel
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

template <int NA, int NOBJ>
int launch_step(bool multi, int n_levels, const GcLevelsDev& lv, const uint8_t* level_id, uint32_t* state,
                const uint8_t* actions, uint8_t* rd, uint64_t* hash, uint32_t* coll, uint8_t* executed,
                uint32_t* rd_bits, int64_t n, cudaStream_t st) {
  auto* s4 = reinterpret_cast<uint4*>(state);
  auto* h = reinterpret_cast<unsigned long long*>(hash);
  if (!multi && !use_generic_step()) {
    StepLutParams P;
    P.lv = lv.lv[0];
    fill_move_table(P.lv, &P.mv);
    static const bool pdl = getenv("GC_STEP_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kLutThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    // the kernel indexes with 32 bits: batches beyond 2^30 envs go in slices
    const int64_t slice = (int64_t)1 << 30;
    for (int64_t lo = 0; lo < n; lo += slice) {
      const int64_t m = n - lo < slice ? n - lo : slice;
      const bool extras = !rd || h || coll || executed;
      const bool bits = rd_bits && !extras;
      cfg.gridDim = dim3(extras ? lut_step_grid<NA, NOBJ, true, false>(m)
                                : bits ? lut_step_grid<NA, NOBJ, false, true>(m) : lut_step_grid<NA, NOBJ, false, false>(m));
      const cudaError_t err = cudaLaunchKernelEx(
          &cfg,
          extras ? step_lut_kernel<NA, NOBJ, true, false>
                 : bits ? step_lut_kernel<NA, NOBJ, false, true> : step_lut_kernel<NA, NOBJ, false, false>,
          P, s4 + lo, actions + lo * NA, rd ? rd + lo : nullptr, h ? h + lo : nullptr, coll ? coll + lo : nullptr,
          executed ? executed + lo * NA : nullptr, bits ? rd_bits + lo / 16 : nullptr, (uint32_t)m);
      if (err != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_env_step: launch failed: %s", cudaGetErrorString(err));
      if (rd_bits && !bits) pack_rd(rd + lo, rd_bits + lo / 16, m, st);
    }
    return gc_check_launch("gc_env_step");
  }
  if (multi && !use_generic_step()) {
    static const bool pdl = getenv("GC_STEP_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)n_levels * gclut::kMoveBytes;
    const bool extras = !rd || h || coll || executed;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const int64_t slice = (int64_t)1 << 30;
    for (int64_t lo = 0; lo < n; lo += slice) {
      const int64_t m = n - lo < slice ? n - lo : slice;
      cfg.gridDim = dim3(extras ? lut_multi_grid<NA, NOBJ, true>(m, cfg.dynamicSmemBytes)
                                : lut_multi_grid<NA, NOBJ, false>(m, cfg.dynamicSmemBytes));
      const cudaError_t err = cudaLaunchKernelEx(
          &cfg, extras ? step_lut_multi_kernel<NA, NOBJ, true> : step_lut_multi_kernel<NA, NOBJ, false>, lv, n_levels, level_id + lo, s4 + lo, actions + lo * NA,
          rd ? rd + lo : nullptr, h ? h + lo : nullptr, coll ? coll + lo : nullptr,
          executed ? executed + lo * NA : nullptr, (uint32_t)m);
      if (err != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_env_step: launch failed: %s", cudaGetErrorString(err));
    }
    if (rd_bits) pack_rd(rd, rd_bits, n, st);
    return gc_check_launch("gc_env_step");
  }
  if (multi)
    step_kernel<NA, NOBJ, true><<<step_grid(n), kThreads, 0, st>>>(lv, level_id, s4, actions, rd, h, coll, executed, n);
  else
    step_kernel<NA, NOBJ, false><<<step_grid(n), kThreads, 0, st>>>(lv, level_id, s4, actions, rd, h, coll, executed, n);
  if (rd_bits) pack_rd(rd, rd_bits, n, st);
  return gc_check_launch("gc_env_step");
}

template <int NA, int NOBJ>
int launch_rollout(bool multi, const GcLevelsDev& lv, const uint8_t* level_id, uint32_t* state, uint8_t* rd,
                   uint64_t* hash_trace, uint32_t* coll, int64_t n, int n_steps, int t0, int64_t env0,
                   uint64_t seed, cudaStream_t st) {
  auto* s4 = reinterpret_cast<uint4*>(state);
  auto* h = reinterpret_cast<unsigned long long*>(hash_trace);
  if (!multi && !use_generic_step()) {
    StepLutParams P;
    P.lv = lv.lv[0];
    fill_move_table(P.lv, &P.mv);
    rollout_lut_kernel<NA, NOBJ><<<grid_for(n), kThreads, 0, st>>>(P, s4, rd, h, coll, n, n_steps, (uint32_t)t0, env0,
                                                                   seed);
    return gc_check_launch("gc_env_rollout");
  }
  if (multi)
    rollout_kernel<NA, NOBJ, true><<<grid_for(n), kThreads, 0, st>>>(lv, level_id, s4, rd, h, coll, n, n_steps,
                                                                     (uint32_t)t0, env0, seed);
  else
    rollout_kernel<NA, NOBJ, false><<<grid_for(n), kThreads, 0, st>>>(lv, level_id, s4, rd, h, coll, n, n_steps,
                                                                      (uint32_t)t0, env0, seed);
  return gc_check_launch("gc_env_rollout");
}

#define GC_DISPATCH_NA_NOBJ(FN, ...)                                   \
  do {                                                                 \
    const bool six = max_objs > 4;                                     \
    switch (n_agents * 2 + (six ? 1 : 0)) {                            \
      case 2: return FN<1, 4>(__VA_ARGS__);                            \
      case 3: return FN<1, 6>(__VA_ARGS__);                            \
      case 4: return FN<2, 4>(__VA_ARGS__);                            \
      case 5: return FN<2, 6>(__VA_ARGS__);                            \
      case 6: return FN<3, 4>(__VA_ARGS__);                            \
      case 7: return FN<3, 6>(__VA_ARGS__);                            \
      case 8: return FN<4, 4>(__VA_ARGS__);                            \
      case 9: return FN<4, 6>(__VA_ARGS__);                            \
    }                                                                  \
  } while (0)

}  // namespace

extern "C" {

int gc_env_reset(const gc_level* levels, int n_levels, const uint8_t* level_id, uint32_t* state, int64_t n,
                 int n_agents, void* stream) {
  GcLevelsDev lv;
  int max_objs = 0;
  if (int rc = gc_levels_to_dev(levels, n_levels, n_agents, &lv, &max_objs)) return rc;
  if (!state || n < 0) return gc_fail(GC_E_ARG, "gc_env_reset: bad state/n");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  reset_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(lv, n_levels > 1 ? level_id : nullptr,
                                                                  reinterpret_cast<uint4*>(state), n);
  return gc_check_launch("gc_env_reset");
}

// gc_env_step plus the optional bit planes of the results (internal: gc_env_step_host)
static int env_step_impl(const gc_level* levels, int n_levels, const uint8_t* level_id, uint32_t* state,
                         const uint8_t* actions, uint8_t* reward_done, uint64_t* hash, uint32_t* collisions,
                         uint8_t* executed, uint32_t* rd_bits, int64_t n, int n_agents, void* stream) {
  GcLevelsDev lv;
  int max_objs = 0;
  if (int rc = gc_levels_to_dev(levels, n_levels, n_agents, &lv, &max_objs)) return rc;
  if (!state || !actions || n < 0) return gc_fail(GC_E_ARG, "gc_env_step: null state/actions or n < 0");
  if (n_levels > 1 && !level_id) return gc_fail(GC_E_ARG, "gc_env_step: n_levels > 1 needs level_id");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const bool multi = n_levels > 1;
  if (rd_bits && !reward_done) return gc_fail(GC_E_ARG, "gc_env_step: bit planes need the reward_done buffer");
  GC_DISPATCH_NA_NOBJ(launch_step, multi, n_levels, lv, level_id, state, actions, reward_done, hash, collisions, executed,
                      rd_bits, n, (cudaStream_t)stream);
  return gc_fail(GC_E_ARG, "gc_env_step: n_agents must be 1..4");
}

int gc_env_step(const gc_level* levels, int n_levels, const uint8_t* level_id, uint32_t* state,
                const uint8_t* actions, uint8_t* reward_done, uint64_t* hash, uint32_t* collisions,
                uint8_t* executed, int64_t n, int n_agents, void* stream) {
  return env_step_impl(levels, n_levels, level_id, state, actions, reward_done, hash, collisions, executed, nullptr, n,
                       n_agents, stream);
}

int gc_env_step_host(const gc_level* levels, int n_levels, const uint8_t* level_id, uint32_t* state,
                     const uint8_t* actions_host, uint8_t* actions_dev, uint8_t* reward_done_dev,
                     uint8_t* reward_done_host, uint32_t* rd_bits_dev, uint32_t* rd_bits_host, uint32_t* collisions,
                     int64_t n, int n_agents, void* stream) {
  if (!actions_host || !actions_dev || !reward_done_dev)
    return gc_fail(GC_E_ARG, "gc_env_step_host: null host/device action or reward_done buffer");
  if (!reward_done_host && !(rd_bits_dev && rd_bits_host))
    return gc_fail(GC_E_ARG, "gc_env_step_host: need reward_done_host or the rd_bits pair for the results");
  if (n < 0 || n_agents < 1 || n_agents > GC_MAX_AGENTS) return gc_fail(GC_E_ARG, "gc_env_step_host: bad n / n_agents");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(actions_dev, actions_host, (size_t)n * n_agents, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_env_step_host: copy in failed: %s", cudaGetErrorString(e));
  const bool want_bits = rd_bits_dev && rd_bits_host;  // two bit planes per 32 envs: a quarter of the bytes over PCIe
  if (int rc = env_step_impl(levels, n_levels, level_id, state, actions_dev, reward_done_dev, nullptr, collisions, nullptr,
                             want_bits ? rd_bits_dev : nullptr, n, n_agents, stream))
    return rc;
  if (want_bits) {
    const int64_t words = (n + 31) / 32;
    e = cudaMemcpyAsync(rd_bits_host, rd_bits_dev, (size_t)words * 8, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess && reward_done_host)
    e = cudaMemcpyAsync(reward_done_host, reward_done_dev, (size_t)n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_env_step_host: copy out failed: %s", cudaGetErrorString(e));
  return GC_OK;
}

int gc_env_rollout(const gc_level* levels, int n_levels, const uint8_t* level_id, uint32_t* state,
                   uint8_t* reward_done, uint64_t* hash_trace, uint32_t* collisions, int64_t n, int n_agents,
                   int n_steps, int t0, int64_t env0, uint64_t seed, void* stream) {
  GcLevelsDev lv;
  int max_objs = 0;
  if (int rc = gc_levels_to_dev(levels, n_levels, n_agents, &lv, &max_objs)) return rc;
  if (!state || n < 0 || n_steps < 0) return gc_fail(GC_E_ARG, "gc_env_rollout: bad state/n/n_steps");
  if (n_levels > 1 && !level_id) return gc_fail(GC_E_ARG, "gc_env_rollout: n_levels > 1 needs level_id");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const bool multi = n_levels > 1;
  GC_DISPATCH_NA_NOBJ(launch_rollout, multi, lv, level_id, state, reward_done, hash_trace, collisions, n, n_steps,
                      t0, env0, seed, (cudaStream_t)stream);
  return gc_fail(GC_E_ARG, "gc_env_rollout: n_agents must be 1..4");
}

int gc_fill_random_actions(uint8_t* actions, int64_t n, int n_agents, int n_steps, int t0, int64_t env0,
                           uint64_t seed, void* stream) {
  if (!actions || n < 0 || n_steps < 0 || n_agents < 1 || n_agents > 4)
    return gc_fail(GC_E_ARG, "gc_fill_random_actions: bad arguments");
  if (n * n_steps == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  fill_actions_kernel<<<grid_for(n * n_steps), kThreads, 0, (cudaStream_t)stream>>>(actions, n, n_agents, n_steps,
                                                                                    (uint32_t)t0, env0, seed);
  return gc_check_launch("gc_fill_random_actions");
}

int gc_state_hash(const uint32_t* state, uint64_t* hash, int64_t n, int n_agents, void* stream) {
  if (!state || !hash || n < 0) return gc_fail(GC_E_ARG, "gc_state_hash: bad arguments");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  auto* s4 = reinterpret_cast<const uint4*>(state);
  auto* h = reinterpret_cast<unsigned long long*>(hash);
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_agents) {
    case 1: hash_kernel<1><<<grid_for(n), kThreads, 0, st>>>(s4, h, n); break;
    case 2: hash_kernel<2><<<grid_for(n), kThreads, 0, st>>>(s4, h, n); break;
    case 3: hash_kernel<3><<<grid_for(n), kThreads, 0, st>>>(s4, h, n); break;
    case 4: hash_kernel<4><<<grid_for(n), kThreads, 0, st>>>(s4, h, n); break;
    default: return gc_fail(GC_E_ARG, "gc_state_hash: n_agents must be 1..4");
  }
  return gc_check_launch("gc_state_hash");
}

int gc_stats_reduce(const uint32_t* state, const uint32_t* collisions, const gc_level* levels, int n_levels,
                    const uint8_t* level_id, uint64_t* stats, int64_t n, void* stream) {
  GcLevelsDev lv;
  int max_objs = 0;
  if (int rc = gc_levels_to_dev(levels, n_levels, 1, &lv, &max_objs)) return rc;
  if (!state || !stats || n < 0) return gc_fail(GC_E_ARG, "gc_stats_reduce: bad arguments");
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  unsigned grid = grid_for(n);
  if (grid > 148u * 8u) grid = 148u * 8u;  // grid-stride: one resident wave
  stats_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(lv, n_levels > 1 ? level_id : nullptr,
                                                            reinterpret_cast<const uint4*>(state), collisions,
                                                            reinterpret_cast<unsigned long long*>(stats), n);
  return gc_check_launch("gc_stats_reduce");
}

}  // extern "C"
