// gc_env.cu - path A kernels: reset, step, fused rollout, action stream, hash, statistics.
// One thread per env; a warp is a tile of 32 envs whose 128-bit states are read and written
// with one fully coalesced 512-byte transaction each way.  The default step (step2_kernel, on the byte planes
// of gc_step2.cuh) stages cached device tables in shared memory; the generic kernels (step_kernel /
// rollout_kernel on gc::step, the plain-ALU reference form) take the level bitboards as kernel parameters.
#include "gc_device.cuh"
#include "gc_host.h"
#include "gc_step2.cuh"

#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

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
  const GcLevelDev& L = levels.lv[level_id ? (level_id[i] & (GC_MAX_LEVELS - 1)) : 0];  // in bounds whatever the byte says
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
  if constexpr (MULTI) lvl_next = level_id[i] & (GC_MAX_LEVELS - 1);

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
      if constexpr (MULTI) lvl_next = level_id[inext] & (GC_MAX_LEVELS - 1);
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
  const GcLevelDev& L = MULTI ? s_levels[level_id[i] & (GC_MAX_LEVELS - 1)] : levels.lv[0];

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
// the levels' subtasks in mask form (gc_level.subtask), for the completed-subtask count of the statistics
struct StatsSubtasks {
  uint8_t n[GC_MAX_LEVELS];
  gc_subtask st[GC_MAX_LEVELS][GC_MAX_SUBTASKS];
};

// Subtasks a state has completed, read off the objects (the batched stand-in for the Bag's
// num_completed_subtasks, metrics_bag.py:55-61, which intersects the agents' own incomplete lists):
// Chop(X) once some live object carries X chopped, Merge(a, b) once some live object contains a | b,
// Deliver(m) once an object with mask m lies on a Delivery square.  Monotone along an episode, like the
// agents' bookkeeping (RealAgent.refresh_subtasks never re-adds a subtask, utils/agent.py:151-171).
__device__ __forceinline__ uint32_t completed_subtasks(const uint4& s, const GcLevelDev& L, const gc_subtask* st, int n_st) {
  uint32_t done = 0;
  for (int q = 0; q < n_st; q++) {
    bool hit = false;
#pragma unroll
    for (int k = 0; k < GC_MAX_OBJECTS; k++) {
      const uint32_t sl = gc::slot_of(s, k), m = sl & 0x7fu, holder = sl >> 13;
      if (holder == 7u) continue;
      if (st[q].kind == GC_ST_DELIVER)
        hit |= holder == 0u && m == st[q].goal && ((L.deliv_mask >> ((sl >> 7) & 63u)) & 1ull);
      else
        hit |= (m & st[q].goal) == st[q].goal;
    }
    done += hit ? 1u : 0u;
  }
  return done;
}

__global__ void __launch_bounds__(kThreads)
stats_kernel(const __grid_constant__ GcLevelsDev levels, const __grid_constant__ StatsSubtasks subtasks,
             const uint8_t* __restrict__ level_id, const uint4* __restrict__ state,
             const uint32_t* __restrict__ collisions, unsigned long long* __restrict__ stats, int64_t n) {
  constexpr int kAcc = 6;
  __shared__ unsigned int s_hist[128];
  __shared__ unsigned long long s_acc[kAcc];
  for (int k = threadIdx.x; k < 128; k += kThreads) s_hist[k] = 0;
  if (threadIdx.x < kAcc) s_acc[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long v[kAcc] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const uint4 s = state[i];
    const uint32_t w0 = s.x;
    const uint32_t lvl = level_id ? (level_id[i] & (GC_MAX_LEVELS - 1)) : 0;  // in bounds whatever the byte says
    const GcLevelDev& L = levels.lv[lvl];
    v[5] += completed_subtasks(s, L, subtasks.st[lvl], subtasks.n[lvl]);
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
  for (int k = 0; k < kAcc; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(&s_acc[k], v[k]);
  }
  __syncthreads();
  if (threadIdx.x < 5 && s_acc[threadIdx.x]) atomicAdd(&stats[threadIdx.x], s_acc[threadIdx.x]);
  if (threadIdx.x == 5 && s_acc[5]) atomicAdd(&stats[133], s_acc[5]);  // after the histogram: the layout only grows
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

// Grid of the generic (table-free) step kernel, the reference form kept for GC_STEP_GENERIC=1 and for the
// A/B tests: one env per thread.  GC_STEP_CTAS_PER_SM=k (1..8) switches it to a persistent grid-stride form
// with k CTAs per SM for experiments (profiles/r01_step_kernel_v1_ncu.csv is this kernel).
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
// step on byte planes (gc_step2.cuh): the default form
// ---------------------------------------------------------------------------------------
// Tables in global memory, built once per (device, level set, n_agents) and cached: a step call passes one
// pointer instead of rebuilding and shipping 2.6 KB of tables per launch (round 1: 640-byte kernel
// parameters plus a 512-entry table fill on the host per call made a plain launch loop host-bound).
struct DeviceTables {
  gcs2::StaticTables st;
  gcs2::LevelTables lv[GC_MAX_LEVELS];
};
constexpr size_t kTablesHead = sizeof(gcs2::StaticTables);
__host__ __device__ constexpr size_t tables_bytes(int n_levels) { return kTablesHead + (size_t)n_levels * sizeof(gcs2::LevelTables); }
static_assert(kTablesHead % 16 == 0 && sizeof(gcs2::LevelTables) % 16 == 0, "tables are copied as uint4");

// the tables depend on the map, the objects, the goals and the horizon - not on the subtask list that
// gc_level_set_subtasks rewrites: the key is the part of gc_level in front of it
constexpr size_t kLevelKeyBytes = offsetof(gc_level, subtask);
struct TableCacheEntry {
  int device, n_levels, n_agents;
  uint64_t stamp;
  uint8_t key[GC_MAX_LEVELS][kLevelKeyBytes];
  DeviceTables* dev_ptr;
};
inline bool same_levels(const TableCacheEntry& e, const gc_level* levels, int n_levels) {
  for (int l = 0; l < n_levels; l++)
    if (memcmp(e.key[l], &levels[l], kLevelKeyBytes) != 0) return false;
  return true;
}
// Every CTA of a launch copies the tables into its shared memory at the same moment: kTableCopies copies at
// different addresses spread those reads over the L2 slices instead of queueing ~900 CTAs on the same 59
// lines (GC_STEP_TABLE_COPIES=1 restores the single copy for A/B timing).
constexpr int kTableCopies = 32;
constexpr int kTableCache = 64;
TableCacheEntry g_table_cache[kTableCache];
int g_table_cache_used = 0;
uint64_t g_table_stamp = 0;
std::mutex g_table_mutex;
const gcs2::StaticTables g_static_tables_host = gcs2::make_static_tables();

// device tables for this level set (nullptr + gc_last_error on failure).  The first call for a level set
// allocates and copies (synchronously: do it outside a stream capture - gc_env_prepare, which KitchenBatch's
// constructor calls); later calls are a memcmp.  The cache holds kTableCache level sets per process and evicts
// the least recently used one; a CUDA graph captured with a level set's tables stays valid while the set is
// cached.
const DeviceTables* tables_for(const gc_level* levels, int n_levels, int n_agents) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    gc_fail(GC_E_CUDA, "cudaGetDevice failed");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(g_table_mutex);
  for (int k = 0; k < g_table_cache_used; k++) {
    TableCacheEntry& e = g_table_cache[k];
    if (e.device == dev && e.n_levels == n_levels && e.n_agents == n_agents && same_levels(e, levels, n_levels)) {
      e.stamp = ++g_table_stamp;
      return e.dev_ptr;
    }
  }
  int slot = g_table_cache_used;
  if (slot == kTableCache) {  // evict the least recently used entry (cudaFree waits for kernels using it)
    slot = 0;
    for (int k = 1; k < kTableCache; k++)
      if (g_table_cache[k].stamp < g_table_cache[slot].stamp) slot = k;
    cudaSetDevice(g_table_cache[slot].device);
    cudaFree(g_table_cache[slot].dev_ptr);
    cudaSetDevice(dev);
  }
  static DeviceTables host;  // under the mutex
  host.st = g_static_tables_host;
  for (int l = 0; l < n_levels; l++) gcs2::fill_level_tables(levels[l], n_agents, &host.lv[l]);
  DeviceTables* d = nullptr;
  const size_t stride = tables_bytes(n_levels);
  cudaError_t err = cudaMalloc(&d, stride * kTableCopies);
  for (int c = 0; c < kTableCopies && err == cudaSuccess; c++)
    err = cudaMemcpy(reinterpret_cast<uint8_t*>(d) + c * stride, &host, stride, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    gc_fail(GC_E_CUDA, "level tables: %s (the first step of a level set cannot run inside a stream capture)",
            cudaGetErrorString(err));
    cudaGetLastError();
    if (d) cudaFree(d);
    if (slot < g_table_cache_used) {  // the evicted entry is gone: compact
      g_table_cache[slot] = g_table_cache[g_table_cache_used - 1];
      g_table_cache_used--;
    }
    return nullptr;
  }
  TableCacheEntry& e = g_table_cache[slot];
  e.device = dev;
  e.n_levels = n_levels;
  e.n_agents = n_agents;
  e.stamp = ++g_table_stamp;
  for (int l = 0; l < n_levels; l++) memcpy(e.key[l], &levels[l], kLevelKeyBytes);
  e.dev_ptr = d;
  if (slot == g_table_cache_used) g_table_cache_used++;
  return d;
}

#ifndef GC_STEP2_THREADS
#define GC_STEP2_THREADS 256
#endif
// resident CTAs per SM the register allocation aims at: 6 (40 registers) for the plain step of <= 2 agents,
// 5 (48) for 3-4 agents, 4 (64) with the optional outputs (the hash alone needs ~30 live registers)
#ifdef GC_STEP2_MIN_CTAS_ALL  // experiments: one bound for every instantiation
#define GC_STEP2_MIN_CTAS(NA, EXTRAS) (GC_STEP2_MIN_CTAS_ALL)
#else
#define GC_STEP2_MIN_CTAS(NA, EXTRAS) ((EXTRAS) ? 4 : ((NA) <= 2 ? 6 : 5))
#endif
constexpr int kS2Threads = GC_STEP2_THREADS;

struct Step2Args {
  const DeviceTables* tables;
  uint4* state;
  const uint8_t* actions;
  uint8_t* reward_done;      // plain step: required; with EXTRAS: nullable
  unsigned long long* hash;  // EXTRAS, nullable
  uint32_t* collisions;      // EXTRAS, nullable
  uint8_t* executed;         // EXTRAS, nullable
  uint32_t* rd_bits;         // BITS
  const uint8_t* level_id;   // MULTI
  uint32_t n;                // < 2^31: the host slices larger batches
  int n_levels;
  int table_copies;          // copies of the tables behind `tables` (tables_bytes(n_levels) apart)
};

template <int NA>
__device__ __forceinline__ void store_action_word(uint8_t* __restrict__ out, uint32_t i, uint32_t w) {
  if constexpr (NA == 1) {
    out[i] = (uint8_t)w;
  } else if constexpr (NA == 2) {
    reinterpret_cast<uint16_t*>(out)[i] = (uint16_t)w;
  } else if constexpr (NA == 4) {
    reinterpret_cast<uint32_t*>(out)[i] = w;
  } else {
    uint8_t* p = out + (size_t)i * 3;
    p[0] = (uint8_t)w;
    p[1] = (uint8_t)(w >> 8);
    p[2] = (uint8_t)(w >> 16);
  }
}

// one env of the persistent loop: transition (or the frozen outcome of a finished episode) and the outputs.
// EXTRAS = false is the plain gym step (state in place + the reward/done byte): the optional outputs (hash,
// collision counters, executed actions) and everything computed only for them drop out at compile time.
// BITS (plain step only) also writes the results as two bit planes per 32 envs - rd_bits[2*w] = done,
// rd_bits[2*w+1] = reward of envs 32w..32w+31, the format gc_env_step_host sends over PCIe - with two
// ballots per warp; lanes past n (valid == false) idle but vote.
template <int NA, int NOBJ, bool EXTRAS, bool BITS>
__device__ __forceinline__ void step2_one(const gcs2::StaticTables& S, const gcs2::LevelTables& L, const Step2Args& A,
                                          uint4 s, uint32_t aw, uint32_t i, bool valid) {
  bool done, success;
  uint32_t exec = 0x04040404u;
  if (s.x >> 31) {
    // sticky done: the episode is over, nothing mutates; re-report the stored outcome
    done = true;
    success = !(L.max_t24 != 0u && (s.x & 0x7F000000u) >= L.max_t24);
  } else {
    gcs2::Env<NOBJ> e;
    gcs2::unpack<NOBJ>(s.x, s.y, s.z, s.w, e);
    const uint32_t ncoll = gcs2::step<NA, NOBJ, EXTRAS>(e, aw, S, L, done, success, exec);
    gcs2::pack<NOBJ>(e, s.x, s.y, s.z, s.w);
    gc::st_stream(A.state + i, s);
    if constexpr (EXTRAS) {
      if (A.collisions && ncoll) A.collisions[i] += ncoll;
    }
  }
  const uint8_t rd = (uint8_t)((done ? GC_RD_DONE : 0) | (success ? GC_RD_REWARD : 0));
  if constexpr (EXTRAS) {
    if (A.reward_done) A.reward_done[i] = rd;
    if (A.hash) A.hash[i] = gc::state_hash<NA>(s);
    if (A.executed) store_action_word<NA>(A.executed, i, exec);
  } else {
    if (valid) A.reward_done[i] = rd;
    if constexpr (BITS) {
      const uint32_t d = __ballot_sync(0xffffffffu, done && valid), r = __ballot_sync(0xffffffffu, success && valid);
      if ((threadIdx.x & 31u) == 0u) *reinterpret_cast<uint2*>(A.rd_bits + 2u * (i >> 5)) = make_uint2(d, r);
    }
  }
}

// Two envs of one thread, branch-free (ILP variant of the plain step): both transitions are computed in one
// basic block so that the compiler can interleave their instruction streams; a finished (or out-of-range:
// frozen) env computes a transition that is thrown away - its stores are predicated off.
template <int NA, int NOBJ>
__device__ __forceinline__ void step2_pair(const gcs2::StaticTables& S, const gcs2::LevelTables& L, const Step2Args& A,
                                           uint4 (&s)[2], const uint32_t (&aw)[2], const uint32_t (&idx)[2],
                                           const bool (&valid)[2]) {
  bool done[2], success[2], live[2];
  uint32_t exec;
#pragma unroll
  for (int u = 0; u < 2; u++) {
    live[u] = !(s[u].x >> 31);
    const bool old_success = !(L.max_t24 != 0u && (s[u].x & 0x7F000000u) >= L.max_t24);
    gcs2::Env<NOBJ> e;
    gcs2::unpack<NOBJ>(s[u].x, s[u].y, s[u].z, s[u].w, e);
    e.x &= 0x7FFFFFFFu;  // the transition assumes the done bit clear
    gcs2::step<NA, NOBJ, false>(e, aw[u], S, L, done[u], success[u], exec);
    gcs2::pack<NOBJ>(e, s[u].x, s[u].y, s[u].z, s[u].w);
    done[u] = live[u] ? done[u] : true;
    success[u] = live[u] ? success[u] : old_success;
  }
#pragma unroll
  for (int u = 0; u < 2; u++) {
    if (live[u]) gc::st_stream(A.state + idx[u], s[u]);
    if (valid[u]) A.reward_done[idx[u]] = (uint8_t)((done[u] ? GC_RD_DONE : 0) | (success[u] ? GC_RD_REWARD : 0));
  }
}

#ifndef GC_STEP2_ILP
#define GC_STEP2_ILP 1  // 2: two envs per thread and iteration (plain single-level step only)
#endif
#ifndef GC_STEP2_L2_AHEAD
#define GC_STEP2_L2_AHEAD 1  // pull the tile after next into L2 while the next one loads into registers (-0.3 us per 2^20-env launch)
#endif
template <int NA>
__device__ __forceinline__ void l2_prefetch(const Step2Args& A, uint32_t i, uint32_t n) {
  if (i < n) {
    if ((threadIdx.x & 7u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.state + i));
    if ((threadIdx.x & 31u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.actions + (size_t)i * NA));
  }
}

// Persistent CTAs (as many as are resident at once) walk tiles of kS2Threads envs with a grid stride; a warp's
// 32 states are one coalesced 512-byte transaction each way.  The loop is software-pipelined in registers
// and unrolled by two: the loads of a thread's NEXT env are issued before the current one is computed, so
// DRAM latency hides behind ~160 instructions of every resident warp.  (Round 1 staged the next state
// through shared memory with cp.async to save registers; on sm_100a every LDGSTS drags three predicated-off
// LDS fillers along, and the byte-plane step needs so few registers that five more fit under the
// 6-CTAs-per-SM bound.)  Programmatic dependent launch: the tables do not depend on earlier kernels, so
// this grid may start, stage them and prefetch its first tile into L2 while the previous kernel of the
// stream drains; everything an earlier kernel can have written is read only after griddepcontrol.wait
// (L2 is the coherence point: a prefetched line cannot be stale).
// joint-action form (GC_PLAN_JOINT_ACTIONS): one index per env, j = sum_i action_i * 5^(NA-1-i) (for two
// agents 5 * a_1 + a_2, the planners' joint-action index) in a uint8 (NA <= 3) or uint16 (NA = 4) - half /
// a third / half of the bytes a caller with host-resident actions sends over PCIe
template <int NA>
__device__ __forceinline__ uint32_t load_joint_raw(const uint8_t* __restrict__ actions, uint32_t i) {
  if constexpr (NA == 4) return reinterpret_cast<const uint16_t*>(actions)[i];
  return actions[i];
}
template <int NA>
__device__ __forceinline__ uint32_t joint_to_bytes(uint32_t j) {
  uint32_t aw = 0;
#pragma unroll
  for (int i = NA - 1; i >= 0; i--) {
    const uint32_t q = (j * 0xCCCDu) >> 18;  // j / 5 for j < 2^16
    aw |= (j - 5u * q) << (8 * i);
    j = q;
  }
  return aw | (j ? 0x04040404u : 0u);  // an index >= 5^NA: everybody stays
}

template <int NA, int NOBJ, bool EXTRAS, bool BITS, bool MULTI, bool JOINT = false>
__global__ void __launch_bounds__(kS2Threads, GC_STEP2_MIN_CTAS(NA, EXTRAS))
step2_kernel(const __grid_constant__ Step2Args A) {
  static_assert(!(EXTRAS && BITS) && !(MULTI && BITS), "bit planes come with the plain single-level step only");
  static_assert(!JOINT || (!EXTRAS && !MULTI), "joint-action indices come with the plain single-level step only");
  extern __shared__ __align__(16) uint8_t s_tables[];  // StaticTables, then n_levels x LevelTables
  const gcs2::StaticTables& S = *reinterpret_cast<const gcs2::StaticTables*>(s_tables);
  const gcs2::LevelTables* LV = reinterpret_cast<const gcs2::LevelTables*>(s_tables + kTablesHead);
  const uint32_t stride = gridDim.x * kS2Threads;  // 32-bit indices: one IMAD.WIDE per address
  const uint32_t n = A.n;
  uint32_t i = blockIdx.x * kS2Threads + threadIdx.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    const int n4 = (int)(tables_bytes(MULTI ? A.n_levels : 1) / 16);
    const uint4* src = reinterpret_cast<const uint4*>(A.tables) + (size_t)(blockIdx.x % (unsigned)A.table_copies) * n4;
    uint4* dst = reinterpret_cast<uint4*>(s_tables);
    for (int k = (int)threadIdx.x; k < n4; k += kS2Threads) dst[k] = __ldg(src + k);
  }
  if (i < n) {
    if ((threadIdx.x & 7u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.state + i));
    if ((threadIdx.x & 31u) == 0u)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(A.actions + (size_t)i * (JOINT ? (NA == 4 ? 2 : 1) : NA)));
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint4 frozen = make_uint4(0x80000000u, 0u, 0u, 0u);  // a lane past the end: nothing to compute or store
  uint4 s_next = frozen;
  uint32_t a_next = 0, l_next = 0;
  if (i < n) {
    s_next = gc::ld_stream(A.state + i);
    a_next = JOINT ? load_joint_raw<NA>(A.actions, i) : load_actions_raw<NA>(A.actions, i);
    if constexpr (MULTI) l_next = A.level_id[i];
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lmax = MULTI ? (uint32_t)(A.n_levels - 1) : 0u;
  // BITS: the warp walks its tiles in lockstep (every lane votes), so the bound is the warp's first env.
  // The copies s = s_next, aw = a_next at the top are the point: they read the registers the previous
  // iteration's loads wrote, so the wait for that data sits BEFORE this iteration's loads are issued.  ptxas
  // tracks all of a warp's global loads with one counting scoreboard, and waiting on it means "every load
  // issued so far has landed": with the loads issued first (an unrolled ping-pong of two register sets, the
  // first version of this loop) the first use of the current state also waited for the loads just issued -
  // a prefetch distance of zero, 7.3 warps per issue slot parked on the long scoreboard
  // (profiles/r02_step2_v1_ncu.csv).
#if GC_STEP2_ILP == 2
  if constexpr (!EXTRAS && !BITS && !MULTI) {
    // two tiles per iteration: envs i and i + stride; the pair after that prefetched into registers
    uint4 t_next = frozen;
    uint32_t b_next = 0;
    if (i + stride < n) {
      t_next = gc::ld_stream(A.state + i + stride);
      b_next = JOINT ? load_joint_raw<NA>(A.actions, i + stride) : load_actions_raw<NA>(A.actions, i + stride);
    }
#pragma unroll 1
    for (; i < n; i += 2u * stride) {
      uint4 s[2] = {s_next, t_next};
      const uint32_t aw[2] = {JOINT ? joint_to_bytes<NA>(a_next) : a_next, JOINT ? joint_to_bytes<NA>(b_next) : b_next};
      const uint32_t idx[2] = {i, i + stride};
      const bool valid[2] = {true, i + stride < n};
      const uint32_t n0 = i + 2u * stride, n1 = i + 3u * stride;
      s_next = frozen;
      t_next = frozen;
      if (n0 < n) {
        s_next = gc::ld_stream(A.state + n0);
        a_next = JOINT ? load_joint_raw<NA>(A.actions, n0) : load_actions_raw<NA>(A.actions, n0);
      }
      if (n1 < n) {
        t_next = gc::ld_stream(A.state + n1);
        b_next = JOINT ? load_joint_raw<NA>(A.actions, n1) : load_actions_raw<NA>(A.actions, n1);
      }
#if GC_STEP2_L2_AHEAD
      l2_prefetch<JOINT ? (NA == 4 ? 2 : 1) : NA>(A, n0 + 2u * stride, n);
      l2_prefetch<JOINT ? (NA == 4 ? 2 : 1) : NA>(A, n1 + 2u * stride, n);
#endif
      step2_pair<NA, NOBJ>(S, LV[0], A, s, aw, idx, valid);
    }
    return;
  }
#endif
#pragma unroll 1
  for (; BITS ? (i - lane < n) : (i < n); i += stride) {
    const uint4 s = s_next;
    const uint32_t aw = JOINT ? joint_to_bytes<NA>(a_next) : a_next, lvl = l_next;
    const uint32_t inext = i + stride;
    if constexpr (BITS) s_next = frozen;
    if (inext < n) {
      s_next = gc::ld_stream(A.state + inext);
      a_next = JOINT ? load_joint_raw<NA>(A.actions, inext) : load_actions_raw<NA>(A.actions, inext);
      if constexpr (MULTI) l_next = A.level_id[inext];
    }
#if GC_STEP2_L2_AHEAD
    l2_prefetch<JOINT ? (NA == 4 ? 2 : 1) : NA>(A, inext + stride, n);
#endif
    step2_one<NA, NOBJ, EXTRAS, BITS>(S, LV[MULTI ? min(lvl, lmax) : 0u], A, s, aw, i, !BITS || i < n);
  }
}

// Fused rollout on byte planes (single-level batches): same philox stream and the same transitions as
// rollout_kernel, state in registers between steps.
template <int NA, int NOBJ>
__global__ void __launch_bounds__(kThreads)
rollout2_kernel(const DeviceTables* __restrict__ tables, uint4* __restrict__ state, uint8_t* __restrict__ reward_done,
                unsigned long long* __restrict__ hash_trace, uint32_t* __restrict__ collisions, int64_t n,
                int n_steps, uint32_t t0, int64_t env0, unsigned long long seed) {
  __shared__ __align__(16) uint8_t s_tables[tables_bytes(1)];
  {
    const uint4* src = reinterpret_cast<const uint4*>(tables);
    uint4* dst = reinterpret_cast<uint4*>(s_tables);
    for (int k = (int)threadIdx.x; k < (int)(tables_bytes(1) / 16); k += kThreads) dst[k] = __ldg(src + k);
  }
  __syncthreads();
  const gcs2::StaticTables& S = *reinterpret_cast<const gcs2::StaticTables*>(s_tables);
  const gcs2::LevelTables& L = *reinterpret_cast<const gcs2::LevelTables*>(s_tables + kTablesHead);
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  uint4 s = gc::ld_stream(state + i);
  gcs2::Env<NOBJ> e;
  gcs2::unpack<NOBJ>(s.x, s.y, s.z, s.w, e);
  bool done = s.x >> 31;
  bool success = done && !(L.max_t24 != 0u && (s.x & 0x7F000000u) >= L.max_t24);
  uint32_t ncoll = 0;
  for (int k = 0; k < n_steps; k++) {
    if (!done) {
      uint32_t r[4], aw = 0, exec;
      gc::philox_actions(seed, t0 + (uint32_t)k, (unsigned long long)(env0 + i), r);
#pragma unroll
      for (int a = 0; a < NA; a++) aw |= r[a] << (8 * a);
      ncoll += gcs2::step<NA, NOBJ, true>(e, aw, S, L, done, success, exec);
    }
    if (hash_trace) {
      gcs2::pack<NOBJ>(e, s.x, s.y, s.z, s.w);
      hash_trace[(int64_t)k * n + i] = gc::state_hash<NA>(s);
    }
  }
  gcs2::pack<NOBJ>(e, s.x, s.y, s.z, s.w);
  gc::st_stream(state + i, s);
  if (reward_done) reward_done[i] = (uint8_t)((done ? GC_RD_DONE : 0) | (success ? GC_RD_REWARD : 0));
  if (collisions && ncoll) collisions[i] += ncoll;
}

// persistent grid of step2_kernel: as many CTAs as are resident at once (from the occupancy of the
// instantiation), unless GC_LUT_CTAS_PER_SM overrides it
template <int NA, int NOBJ, bool EXTRAS, bool BITS, bool MULTI, bool JOINT = false>
unsigned step2_grid(int64_t n, size_t dyn_smem) {
  static int resident = 0;  // CTAs per device
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
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step2_kernel<NA, NOBJ, EXTRAS, BITS, MULTI, JOINT>, kS2Threads,
                                                        dyn_smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 4;
      }
    }
    resident = sms * per_sm;
    resident_smem = dyn_smem;
  }
  const unsigned full = (unsigned)((n + kS2Threads - 1) / kS2Threads);
  if (full <= (unsigned)resident) return full;
  static const int balanced = getenv("GC_STEP_BALANCED") ? atoi(getenv("GC_STEP_BALANCED")) : 0;
  if (balanced) {  // experiment: the same number of tiles for every CTA (k = ceil(tiles / resident))
    const unsigned k = (full + (unsigned)resident - 1) / (unsigned)resident;
    return (full + k - 1) / k;
  }
  return (unsigned)resident;
}

template <int NA, int NOBJ, bool EXTRAS, bool BITS, bool MULTI, bool JOINT = false>
cudaError_t launch_step2(const Step2Args& A, cudaStream_t st) {
  static const bool pdl = getenv("GC_STEP_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kS2Threads);
  cfg.dynamicSmemBytes = tables_bytes(MULTI ? A.n_levels : 1);
  cfg.gridDim = dim3(step2_grid<NA, NOBJ, EXTRAS, BITS, MULTI, JOINT>(A.n, cfg.dynamicSmemBytes));
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, step2_kernel<NA, NOBJ, EXTRAS, BITS, MULTI, JOINT>, A);
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
int launch_step(const gc_level* levels, int n_levels, const GcLevelsDev& lv, const uint8_t* level_id, uint32_t* state,
                const uint8_t* actions, uint8_t* rd, uint64_t* hash, uint32_t* coll, uint8_t* executed,
                uint32_t* rd_bits, int64_t n, int n_agents, cudaStream_t st) {
  auto* s4 = reinterpret_cast<uint4*>(state);
  auto* h = reinterpret_cast<unsigned long long*>(hash);
  const bool multi = n_levels > 1;
  if (!use_generic_step()) {
    const DeviceTables* tables = tables_for(levels, n_levels, n_agents);
    if (!tables) return GC_E_CUDA;
    const bool extras = !rd || h || coll || executed;
    const bool bits = rd_bits && !extras && !multi;  // the plain single-level step writes the bit planes itself
    // the kernel indexes with 32 bits: batches beyond 2^30 envs go in slices
    const int64_t slice = (int64_t)1 << 30;
    for (int64_t lo = 0; lo < n; lo += slice) {
      const int64_t m = n - lo < slice ? n - lo : slice;
      Step2Args A;
      A.tables = tables;
      A.state = s4 + lo;
      A.actions = actions + lo * NA;
      A.reward_done = rd ? rd + lo : nullptr;
      A.hash = h ? h + lo : nullptr;
      A.collisions = coll ? coll + lo : nullptr;
      A.executed = executed ? executed + lo * NA : nullptr;
      A.rd_bits = bits ? rd_bits + lo / 16 : nullptr;
      A.level_id = multi ? level_id + lo : nullptr;
      A.n = (uint32_t)m;
      A.n_levels = n_levels;
      static const int copies = getenv("GC_STEP_TABLE_COPIES") ? atoi(getenv("GC_STEP_TABLE_COPIES")) : kTableCopies;
      A.table_copies = copies >= 1 && copies <= kTableCopies ? copies : kTableCopies;
      cudaError_t err;
      if (multi)
        err = extras ? launch_step2<NA, NOBJ, true, false, true>(A, st) : launch_step2<NA, NOBJ, false, false, true>(A, st);
      else if (extras)
        err = launch_step2<NA, NOBJ, true, false, false>(A, st);
      else
        err = bits ? launch_step2<NA, NOBJ, false, true, false>(A, st) : launch_step2<NA, NOBJ, false, false, false>(A, st);
      if (err != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_env_step: launch failed: %s", cudaGetErrorString(err));
      if (rd_bits && !bits) pack_rd(rd + lo, rd_bits + lo / 16, m, st);
    }
    return GC_OK;
  }
  if (multi)
    step_kernel<NA, NOBJ, true><<<step_grid(n), kThreads, 0, st>>>(lv, level_id, s4, actions, rd, h, coll, executed, n);
  else
    step_kernel<NA, NOBJ, false><<<step_grid(n), kThreads, 0, st>>>(lv, level_id, s4, actions, rd, h, coll, executed, n);
  if (rd_bits) pack_rd(rd, rd_bits, n, st);
  return gc_check_launch("gc_env_step");
}

template <int NA, int NOBJ>
int launch_rollout(const gc_level* levels, int n_levels, const GcLevelsDev& lv, const uint8_t* level_id, uint32_t* state,
                   uint8_t* rd, uint64_t* hash_trace, uint32_t* coll, int64_t n, int n_agents, int n_steps, int t0,
                   int64_t env0, uint64_t seed, cudaStream_t st) {
  auto* s4 = reinterpret_cast<uint4*>(state);
  auto* h = reinterpret_cast<unsigned long long*>(hash_trace);
  const bool multi = n_levels > 1;
  if (!multi && !use_generic_step()) {
    const DeviceTables* tables = tables_for(levels, 1, n_agents);
    if (!tables) return GC_E_CUDA;
    rollout2_kernel<NA, NOBJ><<<grid_for(n), kThreads, 0, st>>>(tables, s4, rd, h, coll, n, n_steps, (uint32_t)t0, env0,
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

// ---- prepared steps -------------------------------------------------------------------------------
// Everything gc_env_step derives per call (level validation, table lookup, argument block, launch
// configuration) fixed once; a step is then one cudaLaunchKernelEx.
struct gc_step_plan {
  gc_level level;  // looked up in the table cache at every run: an evicted level set is rebuilt, never dangling
  int device, n_agents, flags;
  int64_t n;
  Step2Args args;
  cudaError_t (*launch)(const Step2Args&, cudaStream_t);       // plain step
  cudaError_t (*launch_bits)(const Step2Args&, cudaStream_t);  // plain step + result bit planes
  uint8_t* actions_dev;  // staging of gc_step_plan_run_host (owned)
  uint32_t* bits_dev;
  size_t action_bytes;
  const void* zc_host;   // last rd_bits_host seen and its device alias (pinned, mapped memory), or null
  uint32_t* zc_dev;
  cudaStream_t side[4];  // chunked gc_step_plan_run_host: one stream per chunk (created on first use)
  cudaEvent_t ev_start, ev_done[4];
  int n_side;
};

namespace {
template <int NA, int NOBJ>
int plan_bind(gc_step_plan* p) {
  if (p->flags & GC_PLAN_JOINT_ACTIONS) {
    p->launch = launch_step2<NA, NOBJ, false, false, false, true>;
    p->launch_bits = launch_step2<NA, NOBJ, false, true, false, true>;
  } else {
    p->launch = launch_step2<NA, NOBJ, false, false, false, false>;
    p->launch_bits = launch_step2<NA, NOBJ, false, true, false, false>;
  }
  return GC_OK;
}
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

int gc_env_prepare(const gc_level* levels, int n_levels, int n_agents) {
  GcLevelsDev lv;
  int max_objs = 0;
  if (int rc = gc_levels_to_dev(levels, n_levels, n_agents, &lv, &max_objs)) return rc;
  if (int rc = gc_require_device()) return rc;
  return tables_for(levels, n_levels, n_agents) ? GC_OK : GC_E_CUDA;
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
  if (rd_bits && !reward_done) return gc_fail(GC_E_ARG, "gc_env_step: bit planes need the reward_done buffer");
  if (rd_bits && ((uintptr_t)rd_bits & 7u)) return gc_fail(GC_E_ARG, "gc_env_step_host: rd_bits_dev must be 8-byte aligned");
  GC_DISPATCH_NA_NOBJ(launch_step, levels, n_levels, lv, level_id, state, actions, reward_done, hash, collisions, executed,
                      rd_bits, n, n_agents, (cudaStream_t)stream);
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

int gc_step_plan_create(const gc_level* level, uint32_t* state, uint8_t* reward_done, int64_t n, int n_agents,
                        int flags, gc_step_plan** out) {
  if (!out) return gc_fail(GC_E_ARG, "gc_step_plan_create: null out");
  *out = nullptr;
  GcLevelsDev lv;
  int max_objs = 0;
  if (int rc = gc_levels_to_dev(level, 1, n_agents, &lv, &max_objs)) return rc;
  if (!state || !reward_done || n < 1 || n > ((int64_t)1 << 30))
    return gc_fail(GC_E_ARG, "gc_step_plan_create: need state, reward_done and 1 <= n <= 2^30");
  if (int rc = gc_require_device()) return rc;
  const DeviceTables* tables = tables_for(level, 1, n_agents);
  if (!tables) return GC_E_CUDA;
  gc_step_plan* p = new gc_step_plan();
  p->level = *level;
  cudaGetDevice(&p->device);
  p->n_agents = n_agents;
  p->flags = flags;
  p->n = n;
  p->args = Step2Args();
  p->args.tables = tables;
  p->args.state = reinterpret_cast<uint4*>(state);
  p->args.reward_done = reward_done;
  p->args.n = (uint32_t)n;
  p->args.n_levels = 1;
  p->args.table_copies = kTableCopies;
  p->action_bytes = (flags & GC_PLAN_JOINT_ACTIONS) ? (size_t)n * (n_agents == 4 ? 2 : 1) : (size_t)n * n_agents;
  auto bind = [&]() -> int {
    GC_DISPATCH_NA_NOBJ(plan_bind, p);
    return gc_fail(GC_E_ARG, "gc_step_plan_create: n_agents must be 1..4");
  };
  if (int rc = bind()) {
    delete p;
    return rc;
  }
  *out = p;
  return GC_OK;
}

int gc_step_plan_run(const gc_step_plan* p, const uint8_t* actions, void* stream) {
  if (!p || !actions) return gc_fail(GC_E_ARG, "gc_step_plan_run: null plan / actions");
  Step2Args A = p->args;
  A.tables = tables_for(&p->level, 1, p->n_agents);  // a memcmp against the cached level sets
  if (!A.tables) return GC_E_CUDA;
  A.actions = actions;
  const cudaError_t err = p->launch(A, (cudaStream_t)stream);
  if (err != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_step_plan_run: launch failed: %s", cudaGetErrorString(err));
  return GC_OK;
}

static int plan_run_host(gc_step_plan* p, const uint8_t* actions_host, uint32_t* rd_bits_host, void* stream, bool wait);

int gc_step_plan_run_host(gc_step_plan* p, const uint8_t* actions_host, uint32_t* rd_bits_host, void* stream) {
  return plan_run_host(p, actions_host, rd_bits_host, stream, true);
}

int gc_step_plan_enqueue_host(gc_step_plan* p, const uint8_t* actions_host, uint32_t* rd_bits_host, void* stream) {
  return plan_run_host(p, actions_host, rd_bits_host, stream, false);
}

static int plan_run_host(gc_step_plan* p, const uint8_t* actions_host, uint32_t* rd_bits_host, void* stream, bool wait) {
  if (!p || !actions_host || !rd_bits_host) return gc_fail(GC_E_ARG, "gc_step_plan_run_host: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bits_bytes = (size_t)((p->n + 31) / 32) * 8;
  cudaError_t e = cudaSuccess;
  if (!p->actions_dev) {  // staging buffers on first use
    e = cudaMalloc(&p->actions_dev, p->action_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->bits_dev, bits_bytes);
    if (e != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_step_plan_run_host: staging: %s", cudaGetErrorString(e));
  }
  // Results: when `rd_bits_host` is pinned, mapped host memory (any cudaHostAlloc / torch pin_memory buffer
  // under unified addressing) the kernel stores the bit planes straight into it - posted writes over PCIe,
  // visible to the host once the stream has drained - and the device-to-host copy with its ~10 us of
  // launch and completion latency disappears from the step.  Opt-in (GC_E2E_ZEROCOPY=1): on the measured box
  // the 32 K eight-byte PCIe writes cost more than the copy they replace (84 vs 65 us per 2^20-env step,
  // scripts/e2e_ab.sh), so the default keeps the device-to-host copy.
  if (p->zc_host != rd_bits_host) {
    static const bool no_zc = getenv("GC_E2E_ZEROCOPY") == nullptr;  // opt-in: measured slower (below)
    p->zc_host = rd_bits_host;
    p->zc_dev = nullptr;
    cudaPointerAttributes at;
    if (!no_zc && cudaPointerGetAttributes(&at, rd_bits_host) == cudaSuccess && at.type == cudaMemoryTypeHost &&
        at.devicePointer && ((uintptr_t)at.devicePointer & 7u) == 0)
      p->zc_dev = static_cast<uint32_t*>(at.devicePointer);
    cudaGetLastError();
  }
  Step2Args A = p->args;
  A.tables = tables_for(&p->level, 1, p->n_agents);
  if (!A.tables) return GC_E_CUDA;
  // Synchronous call on a large batch: the batch goes in `chunks` pieces, each on its own stream (copy in, step,
  // copy out), so that piece k+1's host-to-device copy overlaps piece k's kernel and device-to-host copy - the
  // two copy directions have their own engines.  (The first copy of this call, issued above for the whole batch
  // in the single-chunk form, is issued per piece here.)  GC_E2E_CHUNKS=1..4, default 2 from 2^18 envs.
  static const int want_chunks = [] {
    const char* c = getenv("GC_E2E_CHUNKS");
    const int v = c ? atoi(c) : 2;
    return v >= 1 && v <= 4 ? v : 2;
  }();
  const int chunks = (wait && !p->zc_dev && p->n >= ((int64_t)1 << 18)) ? want_chunks : 1;
  if (chunks > 1) {
    if (p->n_side < chunks) {
      if (!p->ev_start && cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming) != cudaSuccess)
        return gc_fail(GC_E_CUDA, "gc_step_plan_run_host: event");
      for (int c = p->n_side; c < chunks; c++) {
        if (cudaStreamCreateWithFlags(&p->side[c], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_done[c], cudaEventDisableTiming) != cudaSuccess)
          return gc_fail(GC_E_CUDA, "gc_step_plan_run_host: side stream");
        p->n_side = c + 1;
      }
    }
    const size_t ab = p->action_bytes / (size_t)p->n;  // action bytes per env
    e = cudaEventRecord(p->ev_start, st);
    for (int c = 0; c < chunks && e == cudaSuccess; c++) {
      const int64_t lo = (p->n * c / chunks) & ~(int64_t)1023, hi = c + 1 == chunks ? p->n : ((p->n * (c + 1) / chunks) & ~(int64_t)1023);
      cudaStream_t s = p->side[c];
      e = cudaStreamWaitEvent(s, p->ev_start, 0);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(p->actions_dev + lo * ab, actions_host + lo * ab, (size_t)(hi - lo) * ab, cudaMemcpyHostToDevice, s);
      Step2Args B = A;
      B.state = A.state + lo;
      B.actions = p->actions_dev + lo * ab;
      B.reward_done = A.reward_done + lo;
      B.rd_bits = p->bits_dev + lo / 16;
      B.n = (uint32_t)(hi - lo);
      if (e == cudaSuccess) e = p->launch_bits(B, s);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(rd_bits_host + lo / 16, p->bits_dev + lo / 16, (size_t)((hi - lo + 31) / 32) * 8, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaEventRecord(p->ev_done[c], s);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, p->ev_done[c], 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_step_plan_run_host: %s", cudaGetErrorString(e));
    return GC_OK;
  }
  e = cudaMemcpyAsync(p->actions_dev, actions_host, p->action_bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_step_plan_run_host: copy in failed: %s", cudaGetErrorString(e));
  A.actions = p->actions_dev;
  A.rd_bits = p->zc_dev ? p->zc_dev : p->bits_dev;
  e = p->launch_bits(A, st);
  if (e == cudaSuccess && !p->zc_dev)
    e = cudaMemcpyAsync(rd_bits_host, p->bits_dev, bits_bytes, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && wait) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return gc_fail(GC_E_CUDA, "gc_step_plan_run_host: %s", cudaGetErrorString(e));
  return GC_OK;
}

void gc_step_plan_destroy(gc_step_plan* p) {
  if (!p) return;
  if (p->actions_dev) cudaFree(p->actions_dev);
  if (p->bits_dev) cudaFree(p->bits_dev);
  for (int c = 0; c < p->n_side; c++) {
    cudaStreamDestroy(p->side[c]);
    cudaEventDestroy(p->ev_done[c]);
  }
  if (p->ev_start) cudaEventDestroy(p->ev_start);
  delete p;
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
  GC_DISPATCH_NA_NOBJ(launch_rollout, levels, n_levels, lv, level_id, state, reward_done, hash_trace, collisions, n,
                      n_agents, n_steps, t0, env0, seed, (cudaStream_t)stream);
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
  static StatsSubtasks sub;  // 2 KB: filled per call, passed by value
  memset(&sub, 0, sizeof(sub));
  for (int l = 0; l < n_levels; l++) {
    const int ns = levels[l].n_subtasks < 0 ? 0 : (levels[l].n_subtasks > GC_MAX_SUBTASKS ? GC_MAX_SUBTASKS : levels[l].n_subtasks);
    sub.n[l] = (uint8_t)ns;
    for (int q = 0; q < ns; q++) sub.st[l][q] = levels[l].subtask[q];
  }
  stats_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(lv, sub, n_levels > 1 ? level_id : nullptr,
                                                            reinterpret_cast<const uint4*>(state), collisions,
                                                            reinterpret_cast<unsigned long long*>(stats), n);
  return gc_check_launch("gc_stats_reduce");
}

}  // extern "C"
