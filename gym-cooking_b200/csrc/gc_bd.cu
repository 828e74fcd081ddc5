// gc_bd.cu - path C: Bayesian-Delegation posterior update, one lane group per (env, observer).
//
// Restates BayesianDelegator.bayes_update (delegation_planner/bayesian_delegator.py:1045-1072)
// with prob_nav_actions (:461-689) on dumped inputs: per likelihood row p a max-subtracted
// softmax over the valid actions (scipy.special.softmax, bd:627, 686) evaluated at the taken
// action; per hypothesis h the weighted sum of its rows' likelihoods (bd:1046-1066); multiply
// into the prior (dutils.update :177-178) and normalise (dutils.normalize :186-193, total == 0
// -> uniform over the surviving hypotheses).
//
// Mapping: a group of G = 8/16/32 lanes owns one row, so a warp carries 32/G rows and the
// probs / hyp_pair / qdiff blocks of neighbouring rows are contiguous in memory.  Lane p of
// the group computes likelihood L[p] (log-sum-exp over <= 25 actions), lane h then gathers
// the L values of its hypothesis with warp shuffles and the group reduces the normaliser with
// xor-shuffles - no shared memory, no atomics.
#include <stdlib.h>
#include <string.h>

#include "gc_device.cuh"
#include "gc_host.h"

namespace {

constexpr int kThreads = 256;

template <typename T>
__device__ __forceinline__ T exp_t(T x);
template <>
__device__ __forceinline__ float exp_t<float>(float x) { return __expf(x); }
template <>
__device__ __forceinline__ double exp_t<double>(double x) { return exp(x); }

template <typename T, int G>
__global__ void __launch_bounds__(kThreads)
bd_posterior_kernel(T* __restrict__ probs, const uint8_t* __restrict__ alive,
                    const uint8_t* __restrict__ hyp_pair, const uint8_t* __restrict__ pair_w,
                    const T* __restrict__ qdiff, const uint8_t* __restrict__ n_valid,
                    const uint8_t* __restrict__ act_idx, T beta, int64_t n, int H, int P, int A,
                    int n_entries) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (G - 1);          // lane within the group
  const int gbase = lane & ~(G - 1);       // first lane of the group (shuffle source base)
  const int64_t row = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
  const bool row_ok = row < n;
  const int64_t r = row_ok ? row : 0;

  // ---- likelihood rows: P can exceed G (4 agents: up to ~24 rows), so loop in chunks of G
  // and keep chunk c's value in Lreg[c] (P <= 4*G is enforced by the host) ----
  T Lreg[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const int p = c * G + sub;
    if (c * G < P && p < P && row_ok) {
      const int nv = n_valid[r * P + p];
      if (nv > 0) {
        const T* qd = qdiff + (r * P + p) * A;
        T mx = beta * qd[0];
        for (int a = 1; a < nv; a++) mx = max(mx, beta * qd[a]);
        T sum = T(0);
        for (int a = 0; a < nv; a++) sum += exp_t<T>(beta * qd[a] - mx);
        Lreg[c] = exp_t<T>(beta * qd[act_idx[r * P + p]] - mx) / sum;
      }
    }
  }

  // ---- hypotheses: lane sub handles h = sub, sub+G, ... ----
  T total = T(0);
  int n_alive = 0;
  T mine[GC_MAX_HYPOTHESES / 8];  // H <= 12*G is enforced by the host; G=8 -> 12 slots
  const int n_chunks = (H + G - 1) / G;
#pragma unroll
  for (int c = 0; c < GC_MAX_HYPOTHESES / 8; c++) {
    mine[c] = T(0);
    if (c < n_chunks) {  // warp-uniform: shuffles below need all lanes of the group
      const int h = c * G + sub;
      const bool ok = row_ok && h < H && (!alive || alive[r * H + h]);
      T update = T(0);
      for (int e = 0; e < n_entries; e++) {
        int p = ok ? hyp_pair[(r * H + h) * n_entries + e] : 0xFF;
        const bool used = p != 0xFF;
        p = used ? p : 0;
        // gather L[p] from lane (p % G) of this group, register chunk p / G
        T v = T(0);
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          T got = __shfl_sync(0xffffffffu, Lreg[cc], gbase + (p & (G - 1)));
          if ((p / G) == cc) v = got;
        }
        if (used) update += T(pair_w[r * P + p]) * v;
      }
      if (ok) {
        mine[c] = probs[r * H + h] * update;
        total += mine[c];
        n_alive += 1;
      }
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, o);
    n_alive += __shfl_xor_sync(0xffffffffu, n_alive, o);
  }
#pragma unroll
  for (int c = 0; c < GC_MAX_HYPOTHESES / 8; c++) {
    if (c < n_chunks) {
      const int h = c * G + sub;
      if (row_ok && h < H) {
        const bool ok = !alive || alive[r * H + h];
        T out = T(0);
        if (ok) out = (total == T(0)) ? T(1) / T(n_alive) : mine[c] * (T(1) / total);
        probs[r * H + h] = out;
      }
    }
  }
}


// ---- staged form -------------------------------------------------------------------------
// A CTA owns R = 16 or 32 consecutive rows.  Every input of those rows is one contiguous byte
// range per array, so the CTA copies them into shared memory with 16-byte cp.async (fully
// coalesced, ~8-100 KB in flight per CTA), computes out of shared memory with the same lane
// grouping as above (likelihood values are exchanged through a shared row instead of
// shuffles) and writes the posteriors back with 16-byte stores.
__device__ __forceinline__ void stage_bytes(unsigned char* dst, const unsigned char* src, int nbytes) {
  const int full = nbytes & ~15;
  for (int i = threadIdx.x * 16; i < full; i += blockDim.x * 16) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + i);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
  }
  for (int i = full + threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = src[i];  // ragged last CTA
}

// bulk-copy (TMA, non-tensor form) helpers: one elected thread moves a contiguous byte range
// global -> shared and signals an mbarrier with the byte count; the reverse direction is a
// bulk-group store.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t nbytes, uint32_t bar) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(src), "r"(nbytes), "r"(bar)
               : "memory");
}

// shared-memory image of one tile of R rows (offsets in bytes from the stage base)
struct TileLayout {
  uint32_t qd, pr, hp, pw, nv, ai, al, bytes;
};

template <typename T>
__host__ __device__ inline TileLayout tile_layout(int R, int H, int P, int A, int E) {
  TileLayout t;
  t.qd = 0;
  t.pr = t.qd + (uint32_t)(R * P * A) * (uint32_t)sizeof(T);
  t.hp = t.pr + (uint32_t)(R * H) * (uint32_t)sizeof(T);
  t.pw = t.hp + (uint32_t)(R * H * E);
  t.nv = t.pw + (uint32_t)(R * P);
  t.ai = t.nv + (uint32_t)(R * P);
  t.al = t.ai + (uint32_t)(R * P);
  t.bytes = t.al + (uint32_t)(R * H);  // every section is a multiple of 16 bytes because R is
  return t;
}

// one row group's share of a staged tile.  ONE: P <= G and H <= G, so lane `sub` owns at most
// one likelihood row and one hypothesis and the loops disappear.
template <typename T, int G, int AT, bool ONE>
__device__ __forceinline__ void tile_compute(unsigned char* st, const TileLayout& lay, T* s_L, bool has_alive, T beta,
                                             int rows, int H, int P, int A_rt, int E) {
  const int A = AT ? AT : A_rt;
  const T* s_qd = reinterpret_cast<const T*>(st + lay.qd);
  T* s_pr = reinterpret_cast<T*>(st + lay.pr);
  const uint8_t* s_hp = st + lay.hp;
  const uint8_t* s_pw = st + lay.pw;
  const uint8_t* s_nv = st + lay.nv;
  const uint8_t* s_ai = st + lay.ai;
  const uint8_t* s_al = st + lay.al;
  const int rl = threadIdx.x / G, sub = threadIdx.x & (G - 1);
  const bool row_ok = rl < rows;
  // likelihood rows (bd:626-641, 682-689)
  for (int p = sub; p < P; p += G) {
    const int rp = rl * P + p;
    const int nv = row_ok ? s_nv[rp] : 0;
    T L = T(0);
    if (nv > 0) {
      const T* qd = s_qd + rp * A;
      const T qa = beta * qd[s_ai[rp]];
      T mx = beta * qd[0], sum = T(0);
      if (AT) {
        T x[AT ? AT : 1];
#pragma unroll
        for (int a = 0; a < AT; a++) x[a] = beta * qd[a];
#pragma unroll
        for (int a = 1; a < AT; a++) mx = a < nv ? max(mx, x[a]) : mx;
#pragma unroll
        for (int a = 0; a < AT; a++) sum += a < nv ? exp_t<T>(x[a] - mx) : T(0);
      } else {
        for (int a = 1; a < nv; a++) mx = max(mx, beta * qd[a]);
        for (int a = 0; a < nv; a++) sum += exp_t<T>(beta * qd[a] - mx);
      }
      L = exp_t<T>(qa - mx) / sum;
    }
    if (row_ok) s_L[rp] = L;
    if (ONE) break;
  }
  __syncwarp();  // a row's G <= 32 lanes live in one warp

  T total = T(0);
  int n_alive = 0;
  T mine1 = T(0);
  bool ok1 = false;
  for (int h = sub; h < H; h += G) {
    const int rh = rl * H + h;
    const bool ok = row_ok && (!has_alive || s_al[rh]);
    T update = T(0);
    if (ok) {
#pragma unroll
      for (int e = 0; e < GC_MAX_AGENTS; e++) {
        if (e < E) {
          const int p0 = s_hp[rh * E + e];
          const int p = p0 == 0xFF ? 0 : p0;
          const T term = T(s_pw[rl * P + p]) * s_L[rl * P + p];
          update += p0 == 0xFF ? T(0) : term;
        }
      }
    }
    const T mine = ok ? s_pr[rh] * update : T(0);
    total += mine;
    n_alive += ok ? 1 : 0;
    if (ONE) {
      mine1 = mine;
      ok1 = ok;
      break;
    }
    if (row_ok) s_pr[rh] = mine;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, o);
    n_alive += __shfl_xor_sync(0xffffffffu, n_alive, o);
  }
  const bool zero = total == T(0);
  const T scale = zero ? T(0) : T(1) / total;
  const T uni = zero ? T(1) / T(n_alive > 0 ? n_alive : 1) : T(0);
  if (ONE) {
    if (row_ok && sub < H) s_pr[rl * H + sub] = ok1 ? mine1 * scale + uni : T(0);
  } else if (row_ok) {
    for (int h = sub; h < H; h += G) {
      const int rh = rl * H + h;
      const bool ok = !has_alive || s_al[rh];
      s_pr[rh] = ok ? s_pr[rh] * scale + uni : T(0);
    }
  }
}

// ---- persistent form: full tiles ----------------------------------------------------------
// Each CTA walks tiles blockIdx.x, +gridDim.x, ... of R consecutive rows with a two-stage
// pipeline: thread 0 issues the seven bulk copies of tile k+1 (one contiguous byte range per
// input array) while all threads compute tile k out of shared memory; posteriors go back with
// one bulk store per tile.
template <typename T, int G, int AT, bool ONE>
__global__ void bd_posterior_tiles_kernel(T* __restrict__ probs, const uint8_t* __restrict__ alive,
                                          const uint8_t* __restrict__ hyp_pair, const uint8_t* __restrict__ pair_w,
                                          const T* __restrict__ qdiff, const uint8_t* __restrict__ n_valid,
                                          const uint8_t* __restrict__ act_idx, T beta, int n_tiles, int H, int P,
                                          int A_rt, int E, int R) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int A = AT ? AT : A_rt;
  const TileLayout lay = tile_layout<T>(R, H, P, A, E);
  // [0,16): two mbarriers; then the likelihood exchange row; then two stages
  T* s_L = reinterpret_cast<T*>(smem + 16);
  unsigned char* stage0 = smem + 16 + (uint32_t)(R * P) * (uint32_t)sizeof(T);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(smem);
  const bool has_alive = alive != nullptr;
  const uint32_t tx = lay.bytes - (has_alive ? 0u : (uint32_t)(R * H));

  auto issue = [&](int tile, int s) {
    unsigned char* st = stage0 + (size_t)s * lay.bytes;
    const uint32_t bar = bar0 + 8u * (uint32_t)s;
    const int64_t row0 = (int64_t)tile * R;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
    bulk_load(st + lay.qd, qdiff + row0 * P * A, lay.pr - lay.qd, bar);
    bulk_load(st + lay.pr, probs + row0 * H, lay.hp - lay.pr, bar);
    bulk_load(st + lay.hp, hyp_pair + row0 * H * E, lay.pw - lay.hp, bar);
    bulk_load(st + lay.pw, pair_w + row0 * P, lay.nv - lay.pw, bar);
    bulk_load(st + lay.nv, n_valid + row0 * P, lay.ai - lay.nv, bar);
    bulk_load(st + lay.ai, act_idx + row0 * P, lay.al - lay.ai, bar);
    if (has_alive) bulk_load(st + lay.al, alive + row0 * H, lay.bytes - lay.al, bar);
  };

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  }
  __syncthreads();  // barriers initialised before anyone polls them

  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const int s = it & 1;
    unsigned char* st = stage0 + (size_t)s * lay.bytes;
    if (threadIdx.x == 0 && tile + (int)gridDim.x < n_tiles) {
      // the other stage was last read by tile it-1's compute (all threads passed the barrier at
      // the end of that iteration) and by its bulk store: wait for the store to have read it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue(tile + gridDim.x, s ^ 1);
    }
    const uint32_t bar = bar0 + 8u * (uint32_t)s, parity = (uint32_t)(it >> 1) & 1u;
    uint32_t ready = 0;
    while (!ready) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
          : "=r"(ready)
          : "r"(bar), "r"(parity)
          : "memory");
    }
    tile_compute<T, G, AT, ONE>(st, lay, s_L, has_alive, beta, R, H, P, A_rt, E);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(st + lay.pr);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(probs + (int64_t)tile * R * H),
                   "r"(src), "r"(lay.hp - lay.pr)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- ragged tail (< R rows) and shapes the bulk path does not take: plain staged CTA -------
template <typename T, int G>
__global__ void bd_posterior_tail_kernel(T* __restrict__ probs, const uint8_t* __restrict__ alive,
                                         const uint8_t* __restrict__ hyp_pair, const uint8_t* __restrict__ pair_w,
                                         const T* __restrict__ qdiff, const uint8_t* __restrict__ n_valid,
                                         const uint8_t* __restrict__ act_idx, T beta, int64_t row0, int rows, int H,
                                         int P, int A, int E, int R) {
  extern __shared__ __align__(16) unsigned char smem[];
  const TileLayout lay = tile_layout<T>(R, H, P, A, E);
  T* s_L = reinterpret_cast<T*>(smem + 16);
  unsigned char* st = smem + 16 + (uint32_t)(R * P) * (uint32_t)sizeof(T);
  stage_bytes(st + lay.qd, reinterpret_cast<const unsigned char*>(qdiff + row0 * P * A), rows * P * A * (int)sizeof(T));
  stage_bytes(st + lay.pr, reinterpret_cast<const unsigned char*>(probs + row0 * H), rows * H * (int)sizeof(T));
  stage_bytes(st + lay.hp, hyp_pair + row0 * H * E, rows * H * E);
  stage_bytes(st + lay.pw, pair_w + row0 * P, rows * P);
  stage_bytes(st + lay.nv, n_valid + row0 * P, rows * P);
  stage_bytes(st + lay.ai, act_idx + row0 * P, rows * P);
  if (alive) stage_bytes(st + lay.al, alive + row0 * H, rows * H);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  tile_compute<T, G, 0, false>(st, lay, s_L, alive != nullptr, beta, rows, H, P, A, E);
  __syncthreads();
  const T* s_pr = reinterpret_cast<const T*>(st + lay.pr);
  for (int i = threadIdx.x; i < rows * H; i += blockDim.x) probs[row0 * H + i] = s_pr[i];
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

template <typename T, int G, int AT, bool ONE>
bool launch_tiles(T* probs, const uint8_t* alive, const uint8_t* hyp_pair, const uint8_t* pair_w, const T* qdiff,
                  const uint8_t* n_valid, const uint8_t* act_idx, T beta, int n_tiles, int H, int P, int A, int E,
                  int R, size_t smem_bytes, cudaStream_t st) {
  auto kern = bd_posterior_tiles_kernel<T, G, AT, ONE>;
  static int per_sm = -1;
  static size_t configured = 0;
  if (smem_bytes > configured) {
    if (smem_bytes > 48u * 1024u &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = smem_bytes;
    per_sm = -1;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, R * G, smem_bytes) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    return false;
  }
  per_sm = occ;
  const int cap = sm_count() * per_sm;
  kern<<<n_tiles < cap ? n_tiles : cap, R * G, smem_bytes, st>>>(probs, alive, hyp_pair, pair_w, qdiff, n_valid,
                                                                 act_idx, beta, n_tiles, H, P, A, E, R);
  return true;
}

template <typename T, int G>
bool launch_staged(T* probs, const uint8_t* alive, const uint8_t* hyp_pair, const uint8_t* pair_w, const T* qdiff,
                   const uint8_t* n_valid, const uint8_t* act_idx, T beta, int64_t n, int H, int P, int A, int E,
                   cudaStream_t st) {
  static const bool force_v1 = getenv("GC_BD_UNSTAGED") != nullptr;
  if (force_v1) return false;
  if (!aligned16(probs) || !aligned16(hyp_pair) || !aligned16(pair_w) || !aligned16(qdiff) || !aligned16(n_valid) ||
      !aligned16(act_idx) || (alive && !aligned16(alive)))
    return false;
  int R = 32;
  if (R * G > 1024 || tile_layout<T>(R, H, P, A, E).bytes > 24u * 1024u) R = 16;
  const TileLayout lay = tile_layout<T>(R, H, P, A, E);
  const size_t head = 16u + (size_t)R * P * sizeof(T);
  if (head + 2u * lay.bytes > 200u * 1024u) return false;
  if (n / R > 0x7fffffff) return false;
  const int n_tiles = (int)(n / R);
  const bool one = P <= G && H <= G;
  const int tail = (int)(n - (int64_t)n_tiles * R);
  if (tail > 0 && head + lay.bytes > 48u * 1024u &&
      cudaFuncSetAttribute(bd_posterior_tail_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(head + lay.bytes)) != cudaSuccess) {
    cudaGetLastError();
    return false;  // nothing launched yet: the caller falls back to the unstaged kernel
  }
  if (n_tiles > 0) {
    const size_t smem_bytes = head + 2u * lay.bytes;
    bool ok;
#define GC_BD_TILES(AT)                                                                                              \
  (one ? launch_tiles<T, G, AT, true>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n_tiles, H, P, \
                                      A, E, R, smem_bytes, st)                                                      \
       : launch_tiles<T, G, AT, false>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n_tiles, H,   \
                                       P, A, E, R, smem_bytes, st))
    if (A == 5) ok = GC_BD_TILES(5);
    else if (A == 25) ok = GC_BD_TILES(25);
    else ok = GC_BD_TILES(0);
#undef GC_BD_TILES
    if (!ok) return false;
  }
  if (tail > 0)
    bd_posterior_tail_kernel<T, G><<<1, R * G, head + lay.bytes, st>>>(
        probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, (int64_t)n_tiles * R, tail, H, P, A, E, R);
  return true;
}

template <typename T>
int launch(T* probs, const uint8_t* alive, const uint8_t* hyp_pair, const uint8_t* pair_w, const T* qdiff,
           const uint8_t* n_valid, const uint8_t* act_idx, T beta, int64_t n, int H, int P, int A, int n_entries,
           void* stream) {
  if (!probs || !hyp_pair || !pair_w || !qdiff || !n_valid || !act_idx)
    return gc_fail(GC_E_ARG, "gc_bd_posterior: null array");
  if (n < 0 || H < 1 || P < 1 || A < 1 || n_entries < 1 || n_entries > GC_MAX_AGENTS)
    return gc_fail(GC_E_ARG, "gc_bd_posterior: bad sizes (n=%lld H=%d P=%d A=%d entries=%d)", (long long)n, H, P, A,
                   n_entries);
  if (H > GC_MAX_HYPOTHESES || P > 128 || A > 32)
    return gc_fail(GC_E_LIMIT, "gc_bd_posterior: H <= %d, P <= 128, A <= 32", GC_MAX_HYPOTHESES);
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const int need = H > P ? H : P;
  cudaStream_t st = (cudaStream_t)stream;
  // staged kernels: the group is sized for the likelihood rows (the expensive phase: up to 25 exps
  // per row); hypotheses loop in chunks of G.  GC_BD_GROUP_BY_MAX=1 restores "G covers max(H, P)".
  static const bool by_max = getenv("GC_BD_GROUP_BY_MAX") != nullptr;
  const int lanes = by_max ? need : P;
  if (lanes <= 8) {
    if (launch_staged<T, 8>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries, st))
      return gc_check_launch("gc_bd_posterior");
  } else if (lanes <= 16) {
    if (launch_staged<T, 16>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries, st))
      return gc_check_launch("gc_bd_posterior");
  } else {
    if (launch_staged<T, 32>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries, st))
      return gc_check_launch("gc_bd_posterior");
  }
  if (need <= 8 && P <= 32) {
    const int64_t threads = n * 8;
    bd_posterior_kernel<T, 8><<<(unsigned)((threads + kThreads - 1) / kThreads), kThreads, 0, st>>>(
        probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries);
  } else if (need <= 16 && P <= 64) {
    const int64_t threads = n * 16;
    bd_posterior_kernel<T, 16><<<(unsigned)((threads + kThreads - 1) / kThreads), kThreads, 0, st>>>(
        probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries);
  } else {
    const int64_t threads = n * 32;
    bd_posterior_kernel<T, 32><<<(unsigned)((threads + kThreads - 1) / kThreads), kThreads, 0, st>>>(
        probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries);
  }
  return gc_check_launch("gc_bd_posterior");
}


// ---- likelihood rows (prob_nav_actions :461-689): softmax inputs from the planner's Q rows ------
struct RowTable {
  int32_t pair[128];
  uint8_t kind[128], agent[128], agent2[128];
};

template <typename T, int A>
__global__ void __launch_bounds__(kThreads)
bd_rows_kernel(const float* __restrict__ q_table, const int64_t* __restrict__ q_row, int n_pairs,
               const __grid_constant__ RowTable rows, const uint8_t* __restrict__ executed,
               const uint8_t* __restrict__ n_moves, int observer, T none_p, T q_cap, T* __restrict__ qdiff,
               uint8_t* __restrict__ n_valid, uint8_t* __restrict__ act_idx, int64_t n, int P, int n_agents) {
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= n * P) return;
  const int64_t env = idx / P;
  const int p = (int)(idx - env * P);
  const int kind = rows.kind[p], ag = rows.agent[p];
  T* out = qdiff + idx * A;
  int nv = 0, taken_rank = 0;
  if (kind == 0) {  // doing nothing (bd:618-641)
    const int k = min((int)n_moves[env], 4);
    nv = k + 1;
    out[0] = none_p;
    for (int a = 1; a <= k; a++) out[a] = (T(1) - none_p) / T(k);
    for (int a = k + 1; a < A; a++) out[a] = T(0);
    taken_rank = executed[env * n_agents + ag] == 4 ? 0 : min(1, nv - 1);
  } else if (kind == 3) {
    // a joint row the observer is not part of (three or more agents): every offered joint action is valid
    // (no partner filter, bd:677 is false), taken = 5 * a_i + a_j
    if constexpr (A >= 25) {
      const float* base = q_table + (q_row[env] * n_pairs + rows.pair[p]) * 25;
      const int taken = 5 * min((int)executed[env * n_agents + ag], 4) + min((int)executed[env * n_agents + rows.agent2[p]], 4);
      const float qt = base[taken];
      const T old = (isnan(qt) || isinf(qt)) ? q_cap : min((T)qt, q_cap);
      for (int a = 0; a < 25; a++) {
        const float q = base[a];
        if (!isnan(q) || a == taken) {
          if (a == taken) taken_rank = nv;
          out[nv++] = old - ((isnan(q) || isinf(q)) ? q_cap : min((T)q, q_cap));
        }
      }
      for (int a = nv; a < A; a++) out[a] = T(0);
    }
  } else {
    const float* base = q_table + (q_row[env] * n_pairs + rows.pair[p]) * 25;
    int taken = executed[env * n_agents + ag];
    int stride = 1, offset = 0;  // entry a of the row is base[offset + a * stride]
    if (kind == 2) {             // only joint actions matching the partner's executed move (bd:677-679)
      const int ag2 = rows.agent2[p];
      if (observer == ag) {
        stride = 5;
        offset = min((int)executed[env * n_agents + ag2], 4);
      } else {
        offset = 5 * min(taken, 4);
        taken = executed[env * n_agents + ag2];
      }
    }
    taken = min(taken, 4);
    T qc[5], res[5] = {T(0), T(0), T(0), T(0), T(0)};
    bool valid[5];
#pragma unroll
    for (int a = 0; a < 5; a++) {
      const float q = base[offset + a * stride];
      valid[a] = !isnan(q) || a == taken;
      qc[a] = (isnan(q) || isinf(q)) ? q_cap : min((T)q, q_cap);
    }
    T old = qc[0];
#pragma unroll
    for (int a = 1; a < 5; a++) old = a == taken ? qc[a] : old;
#pragma unroll
    for (int a = 0; a < 5; a++) {
      if (valid[a]) {
        if (a == taken) taken_rank = nv;
#pragma unroll
        for (int o = 0; o < 5; o++)
          if (o == nv) res[o] = old - qc[a];
        nv++;
      }
    }
#pragma unroll
    for (int a = 0; a < 5; a++) out[a] = res[a];
    for (int a = 5; a < A; a++) out[a] = T(0);
  }
  n_valid[idx] = (uint8_t)nv;
  act_idx[idx] = (uint8_t)taken_rank;
}

template <typename T>
int launch_rows(const float* q_table, const int64_t* q_row, int n_pairs, const int32_t* row_pair,
                const uint8_t* row_kind, const uint8_t* row_agent, const uint8_t* row_agent2, const uint8_t* executed,
                const uint8_t* n_moves, int observer, T none_p, T q_cap, T* qdiff, uint8_t* n_valid, uint8_t* act_idx,
                int64_t n, int P, int n_agents, int A, void* stream) {
  if (!q_table || !q_row || !row_pair || !row_kind || !row_agent || !row_agent2 || !executed || !n_moves || !qdiff ||
      !n_valid || !act_idx)
    return gc_fail(GC_E_ARG, "gc_bd_likelihood_rows: null array");
  if (n < 0 || P < 1 || P > 128 || n_pairs < 1 || n_agents < 1 || n_agents > GC_MAX_AGENTS || observer < 0 ||
      observer >= n_agents)
    return gc_fail(GC_E_ARG, "gc_bd_likelihood_rows: bad sizes (n=%lld P=%d pairs=%d agents=%d observer=%d)",
                   (long long)n, P, n_pairs, n_agents, observer);
  if (A != 5 && A != 25) return gc_fail(GC_E_ARG, "gc_bd_likelihood_rows: A must be 5 or 25");
  RowTable rt;
  memset(&rt, 0, sizeof(rt));
  for (int p = 0; p < P; p++) {
    const int kind = row_kind[p];
    if (kind > 3 || row_agent[p] >= n_agents || (kind >= 2 && row_agent2[p] >= n_agents) ||
        (kind != 0 && (row_pair[p] < 0 || row_pair[p] >= n_pairs)))
      return gc_fail(GC_E_ARG, "gc_bd_likelihood_rows: row %d is malformed", p);
    if (kind == 2 && observer != row_agent[p] && observer != row_agent2[p])
      return gc_fail(GC_E_ARG, "gc_bd_likelihood_rows: joint row %d does not contain the observer (that is kind 3)", p);
    if (kind == 3 && (A != 25 || observer == row_agent[p] || observer == row_agent2[p]))
      return gc_fail(GC_E_ARG, "gc_bd_likelihood_rows: row %d: kind 3 is a joint row WITHOUT the observer and needs A = 25", p);
    rt.pair[p] = row_pair[p];
    rt.kind[p] = (uint8_t)kind;
    rt.agent[p] = row_agent[p];
    rt.agent2[p] = row_agent2[p];
  }
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  const int64_t total = n * P;
  const unsigned grid = (unsigned)((total + kThreads - 1) / kThreads);
  if (A == 25)
    bd_rows_kernel<T, 25><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        q_table, q_row, n_pairs, rt, executed, n_moves, observer, none_p, q_cap, qdiff, n_valid, act_idx, n, P, n_agents);
  else
    bd_rows_kernel<T, 5><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
        q_table, q_row, n_pairs, rt, executed, n_moves, observer, none_p, q_cap, qdiff, n_valid, act_idx, n, P, n_agents);
  return gc_check_launch("gc_bd_likelihood_rows");
}


// ---- fused update over per-env hypothesis lists (three / four agents) ----------------------------------
// With three or four agents the hypothesis table of a level has 10^2..4*10^4 rows (add_subtasks, bd:792-886)
// of which an env keeps the few that survive pruning (bd:200-256), so each env carries a LIST of table rows
// (`rid`) next to its probabilities instead of a dense [H] vector, and the table (`hyp_pair [H][E]`) is shared
// by all envs.  One warp per env: lanes build the env's likelihood values straight from the planner's Q rows
// (the arithmetic of bd_rows_kernel followed by the softmax of bd_posterior_kernel, same order of operations),
// park them in shared memory, then walk the list.
struct RowTableW {
  RowTable t;
  uint8_t w[128];
};

constexpr int kListWarps = 8;

template <typename T>
__device__ T row_likelihood(const RowTable& rows, int p, const float* __restrict__ q_table, int64_t qrow, int n_pairs,
                            const uint8_t* __restrict__ ex, int n_moves, int observer, T none_p, T q_cap, T beta) {
  const int kind = rows.kind[p], ag = rows.agent[p];
  if (kind == 0) {  // doing nothing (bd:618-641)
    const int k = min(n_moves, 4);
    const int nv = k + 1;
    const int taken_rank = ex[ag] == 4 ? 0 : min(1, nv - 1);
    const T other = (T(1) - none_p) / T(k);
    T mx = beta * none_p;
    if (k > 0) mx = max(mx, beta * other);
    T sum = exp_t<T>(beta * none_p - mx);
    for (int a = 1; a <= k; a++) sum += exp_t<T>(beta * other - mx);
    return exp_t<T>(beta * (taken_rank == 0 ? none_p : other) - mx) / sum;
  }
  const float* base = q_table + (qrow * n_pairs + rows.pair[p]) * 25;
  int taken = min((int)ex[ag], 4), stride = 1, offset = 0, count = 5;
  if (kind == 3) {  // joint row without the observer: all 25 joint actions, taken = 5 a_i + a_j
    taken = 5 * taken + min((int)ex[rows.agent2[p]], 4);
    count = 25;
  } else if (kind == 2) {  // only joint actions matching the partner's executed move (bd:677-679)
    const int ag2 = rows.agent2[p];
    if (observer == ag) {
      stride = 5;
      offset = min((int)ex[ag2], 4);
    } else {
      offset = 5 * taken;
      taken = min((int)ex[ag2], 4);
    }
  }
  const float qt = base[offset + taken * stride];
  const T old = (isnan(qt) || isinf(qt)) ? q_cap : min((T)qt, q_cap);
  T mx = T(0);  // the taken action is always valid and its difference is 0
  bool first = true;
  for (int a = 0; a < count; a++) {
    const float q = base[offset + a * stride];
    if (!isnan(q) || a == taken) {
      const T d = beta * (old - ((isnan(q) || isinf(q)) ? q_cap : min((T)q, q_cap)));
      mx = first ? d : max(mx, d);
      first = false;
    }
  }
  T sum = T(0);
  for (int a = 0; a < count; a++) {
    const float q = base[offset + a * stride];
    if (!isnan(q) || a == taken) sum += exp_t<T>(beta * (old - ((isnan(q) || isinf(q)) ? q_cap : min((T)q, q_cap))) - mx);
  }
  return exp_t<T>(beta * (old - old) - mx) / sum;
}

template <typename T>
__global__ void __launch_bounds__(kListWarps * 32)
bd_update_lists_kernel(T* __restrict__ probs, const uint8_t* __restrict__ alive, const int64_t* __restrict__ rid, int W,
                       const uint8_t* __restrict__ hyp_pair, int H, int E, const float* __restrict__ q_table,
                       const int64_t* __restrict__ q_row, int n_pairs, const __grid_constant__ RowTableW rows,
                       const uint8_t* __restrict__ executed, const uint8_t* __restrict__ n_moves, int observer, T none_p,
                       T q_cap, T beta, int64_t n, int P, int n_agents) {
  __shared__ T s_l[kListWarps][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * kListWarps + warp;
  if (env >= n) return;
  const uint8_t* ex = executed + env * n_agents;
  const int64_t qrow = q_row[env];
  const int nm = n_moves[env];
  for (int p = lane; p < P; p += 32)
    s_l[warp][p] = T(rows.w[p]) * row_likelihood<T>(rows.t, p, q_table, qrow, n_pairs, ex, nm, observer, none_p, q_cap, beta);
  __syncwarp();
  T* pr = probs + env * W;
  const uint8_t* al = alive + env * W;
  const int64_t* rd = rid + env * W;
  T total = T(0);
  int n_alive = 0;
  for (int k = lane; k < W; k += 32) {
    const int64_t r = rd[k];
    T v = T(0);
    if (al[k] && r >= 0 && r < H) {
      T update = T(0);
      for (int e = 0; e < E; e++) {
        const int p = hyp_pair[r * E + e];
        if (p != 0xFF) update += s_l[warp][p];
      }
      v = pr[k] * update;
      n_alive += 1;
      total += v;
    }
    pr[k] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, o);
    n_alive += __shfl_xor_sync(0xffffffffu, n_alive, o);
  }
  __syncwarp();
  for (int k = lane; k < W; k += 32) {
    const int64_t r = rd[k];
    if (al[k] && r >= 0 && r < H) pr[k] = (total == T(0)) ? T(1) / T(n_alive) : pr[k] * (T(1) / total);
  }
}

template <typename T>
int launch_lists(T* probs, const uint8_t* alive, const int64_t* rid, int W, const uint8_t* hyp_pair, int H, int E,
                 const uint8_t* pair_w, const float* q_table, const int64_t* q_row, int n_pairs, const int32_t* row_pair,
                 const uint8_t* row_kind, const uint8_t* row_agent, const uint8_t* row_agent2, const uint8_t* executed,
                 const uint8_t* n_moves, int observer, T none_p, T q_cap, T beta, int64_t n, int P, int n_agents,
                 void* stream) {
  if (!probs || !alive || !rid || !hyp_pair || !pair_w || !q_table || !q_row || !row_pair || !row_kind || !row_agent ||
      !row_agent2 || !executed || !n_moves)
    return gc_fail(GC_E_ARG, "gc_bd_update_lists: null array");
  if (n < 0 || W < 1 || H < 1 || E < 1 || E > GC_MAX_AGENTS || P < 1 || P > 128 || n_pairs < 1 || n_agents < 1 ||
      n_agents > GC_MAX_AGENTS || observer < 0 || observer >= n_agents)
    return gc_fail(GC_E_ARG, "gc_bd_update_lists: bad sizes (n=%lld W=%d H=%d E=%d P=%d pairs=%d agents=%d observer=%d)",
                   (long long)n, W, H, E, P, n_pairs, n_agents, observer);
  RowTableW rt;
  memset(&rt, 0, sizeof(rt));
  for (int p = 0; p < P; p++) {
    const int kind = row_kind[p];
    if (kind > 3 || row_agent[p] >= n_agents || (kind >= 2 && row_agent2[p] >= n_agents) ||
        (kind != 0 && (row_pair[p] < 0 || row_pair[p] >= n_pairs)))
      return gc_fail(GC_E_ARG, "gc_bd_update_lists: row %d is malformed", p);
    const bool inside = observer == row_agent[p] || observer == row_agent2[p];
    if ((kind == 2 && !inside) || (kind == 3 && inside))
      return gc_fail(GC_E_ARG, "gc_bd_update_lists: row %d: kind 2 = joint row with the observer, kind 3 = without", p);
    rt.t.pair[p] = row_pair[p];
    rt.t.kind[p] = (uint8_t)kind;
    rt.t.agent[p] = row_agent[p];
    rt.t.agent2[p] = row_agent2[p];
    rt.w[p] = pair_w[p];
  }
  if (n == 0) return GC_OK;
  if (int rc = gc_require_device()) return rc;
  bd_update_lists_kernel<T><<<(unsigned)((n + kListWarps - 1) / kListWarps), kListWarps * 32, 0, (cudaStream_t)stream>>>(
      probs, alive, rid, W, hyp_pair, H, E, q_table, q_row, n_pairs, rt, executed, n_moves, observer, none_p, q_cap, beta,
      n, P, n_agents);
  return gc_check_launch("gc_bd_update_lists");
}

}  // namespace

extern "C" {

int gc_bd_update_lists_f32(float* probs, const uint8_t* alive, const int64_t* rid, int W, const uint8_t* hyp_pair, int H,
                           int n_entries, const uint8_t* pair_w, const float* q_table, const int64_t* q_row, int n_pairs,
                           const int32_t* row_pair, const uint8_t* row_kind, const uint8_t* row_agent,
                           const uint8_t* row_agent2, const uint8_t* executed, const uint8_t* n_moves, int observer,
                           float none_action_prob, float q_cap, float beta, int64_t n, int P, int n_agents, void* stream) {
  return launch_lists<float>(probs, alive, rid, W, hyp_pair, H, n_entries, pair_w, q_table, q_row, n_pairs, row_pair,
                             row_kind, row_agent, row_agent2, executed, n_moves, observer, none_action_prob, q_cap, beta,
                             n, P, n_agents, stream);
}

int gc_bd_update_lists_f64(double* probs, const uint8_t* alive, const int64_t* rid, int W, const uint8_t* hyp_pair, int H,
                           int n_entries, const uint8_t* pair_w, const float* q_table, const int64_t* q_row, int n_pairs,
                           const int32_t* row_pair, const uint8_t* row_kind, const uint8_t* row_agent,
                           const uint8_t* row_agent2, const uint8_t* executed, const uint8_t* n_moves, int observer,
                           double none_action_prob, double q_cap, double beta, int64_t n, int P, int n_agents,
                           void* stream) {
  return launch_lists<double>(probs, alive, rid, W, hyp_pair, H, n_entries, pair_w, q_table, q_row, n_pairs, row_pair,
                              row_kind, row_agent, row_agent2, executed, n_moves, observer, none_action_prob, q_cap, beta,
                              n, P, n_agents, stream);
}

int gc_bd_posterior_f32(float* probs, const uint8_t* alive, const uint8_t* hyp_pair, const uint8_t* pair_w,
                        const float* qdiff, const uint8_t* n_valid, const uint8_t* act_idx, float beta, int64_t n,
                        int H, int P, int A, int n_entries, void* stream) {
  return launch<float>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries, stream);
}

int gc_bd_posterior_f64(double* probs, const uint8_t* alive, const uint8_t* hyp_pair, const uint8_t* pair_w,
                        const double* qdiff, const uint8_t* n_valid, const uint8_t* act_idx, double beta, int64_t n,
                        int H, int P, int A, int n_entries, void* stream) {
  return launch<double>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries, stream);
}

int gc_bd_likelihood_rows_f32(const float* q_table, const int64_t* q_row, int n_pairs, const int32_t* row_pair,
                              const uint8_t* row_kind, const uint8_t* row_agent, const uint8_t* row_agent2,
                              const uint8_t* executed, const uint8_t* n_moves, int observer, float none_action_prob,
                              float q_cap, float* qdiff, uint8_t* n_valid, uint8_t* act_idx, int64_t n, int P,
                              int n_agents, int A, void* stream) {
  return launch_rows<float>(q_table, q_row, n_pairs, row_pair, row_kind, row_agent, row_agent2, executed, n_moves,
                         observer, none_action_prob, q_cap, qdiff, n_valid, act_idx, n, P, n_agents, A, stream);
}

int gc_bd_likelihood_rows_f64(const float* q_table, const int64_t* q_row, int n_pairs, const int32_t* row_pair,
                              const uint8_t* row_kind, const uint8_t* row_agent, const uint8_t* row_agent2,
                              const uint8_t* executed, const uint8_t* n_moves, int observer, double none_action_prob,
                              double q_cap, double* qdiff, uint8_t* n_valid, uint8_t* act_idx, int64_t n, int P,
                              int n_agents, int A, void* stream) {
  return launch_rows<double>(q_table, q_row, n_pairs, row_pair, row_kind, row_agent, row_agent2, executed, n_moves,
                         observer, none_action_prob, q_cap, qdiff, n_valid, act_idx, n, P, n_agents, A, stream);
}

}  // extern "C"
