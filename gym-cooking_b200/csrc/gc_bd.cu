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

template <typename T, int G, int AT>
__global__ void bd_posterior_staged_kernel(T* __restrict__ probs, const uint8_t* __restrict__ alive,
                                           const uint8_t* __restrict__ hyp_pair, const uint8_t* __restrict__ pair_w,
                                           const T* __restrict__ qdiff, const uint8_t* __restrict__ n_valid,
                                           const uint8_t* __restrict__ act_idx, T beta, int64_t n, int H, int P,
                                           int A_rt, int E, int R) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int A = AT ? AT : A_rt;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int rows = (int)min((int64_t)R, n - row0);
  // [0,16): mbarrier; every section after it is a multiple of 16 bytes because R is
  T* s_qd = reinterpret_cast<T*>(smem + 16);
  T* s_pr = s_qd + R * P * A;
  T* s_L = s_pr + R * H;
  uint8_t* s_hp = reinterpret_cast<uint8_t*>(s_L + R * P);
  uint8_t* s_pw = s_hp + R * H * E;
  uint8_t* s_nv = s_pw + R * P;
  uint8_t* s_ai = s_nv + R * P;
  uint8_t* s_al = s_ai + R * P;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem);

  if (rows == R) {  // full CTA: seven bulk copies issued by one thread
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t b_qd = (uint32_t)(R * P * A) * (uint32_t)sizeof(T), b_pr = (uint32_t)(R * H) * (uint32_t)sizeof(T);
      const uint32_t b_hp = (uint32_t)(R * H * E), b_p = (uint32_t)(R * P), b_al = alive ? (uint32_t)(R * H) : 0u;
      const uint32_t total = b_qd + b_pr + b_hp + 3u * b_p + b_al;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
      bulk_load(s_qd, qdiff + row0 * P * A, b_qd, bar);
      bulk_load(s_pr, probs + row0 * H, b_pr, bar);
      bulk_load(s_hp, hyp_pair + row0 * H * E, b_hp, bar);
      bulk_load(s_pw, pair_w + row0 * P, b_p, bar);
      bulk_load(s_nv, n_valid + row0 * P, b_p, bar);
      bulk_load(s_ai, act_idx + row0 * P, b_p, bar);
      if (alive) bulk_load(s_al, alive + row0 * H, b_al, bar);
    }
    __syncthreads();  // barrier initialised before anyone polls it
    uint32_t ready = 0;
    while (!ready) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
          : "=r"(ready)
          : "r"(bar)
          : "memory");
    }
  } else {  // ragged last CTA
    stage_bytes(reinterpret_cast<unsigned char*>(s_qd), reinterpret_cast<const unsigned char*>(qdiff + row0 * P * A),
                rows * P * A * (int)sizeof(T));
    stage_bytes(reinterpret_cast<unsigned char*>(s_pr), reinterpret_cast<const unsigned char*>(probs + row0 * H),
                rows * H * (int)sizeof(T));
    stage_bytes(s_hp, hyp_pair + row0 * H * E, rows * H * E);
    stage_bytes(s_pw, pair_w + row0 * P, rows * P);
    stage_bytes(s_nv, n_valid + row0 * P, rows * P);
    stage_bytes(s_ai, act_idx + row0 * P, rows * P);
    if (alive) stage_bytes(s_al, alive + row0 * H, rows * H);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }

  const int rl = threadIdx.x / G, sub = threadIdx.x & (G - 1);
  const bool row_ok = rl < rows;
  // likelihood rows (bd:626-641, 682-689): lane sub takes p = sub, sub+G, ...
  if (row_ok) {
    for (int p = sub; p < P; p += G) {
      const int rp = rl * P + p;
      const int nv = s_nv[rp];
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
      s_L[rp] = L;
    }
  }
  __syncwarp();  // a row's G <= 32 lanes live in one warp

  T total = T(0);
  int n_alive = 0;
  if (row_ok) {
    for (int h = sub; h < H; h += G) {
      const int rh = rl * H + h;
      const bool ok = !alive || s_al[rh];
      T mine = T(0);
      if (ok) {
        T update = T(0);
#pragma unroll
        for (int e = 0; e < GC_MAX_AGENTS; e++) {
          if (e < E) {
            const int p = s_hp[rh * E + e];
            if (p != 0xFF) update += T(s_pw[rl * P + p]) * s_L[rl * P + p];
          }
        }
        mine = s_pr[rh] * update;
        total += mine;
        n_alive += 1;
      }
      s_pr[rh] = mine;
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, o);
    n_alive += __shfl_xor_sync(0xffffffffu, n_alive, o);
  }
  if (row_ok) {
    const bool zero = total == T(0);
    const T scale = zero ? T(0) : T(1) / total;
    const T uni = zero ? T(1) / T(n_alive > 0 ? n_alive : 1) : T(0);
    for (int h = sub; h < H; h += G) {
      const int rh = rl * H + h;
      const bool ok = !alive || s_al[rh];
      s_pr[rh] = ok ? s_pr[rh] * scale + uni : T(0);
    }
  }
  if (rows == R) {  // posteriors back with one bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(s_pr);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(probs + row0 * H), "r"(src),
                   "r"((uint32_t)(R * H) * (uint32_t)sizeof(T))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    __syncthreads();
    for (int i = threadIdx.x; i < rows * H; i += blockDim.x) probs[row0 * H + i] = s_pr[i];
  }
}

template <typename T>
size_t staged_bytes(int R, int H, int P, int A, int E) {
  return 16u + (size_t)R * ((size_t)P * A * sizeof(T) + (size_t)H * sizeof(T) + (size_t)P * sizeof(T) + (size_t)H * E +
                      3u * P + (size_t)H);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
  if (R * G > 1024 || staged_bytes<T>(R, H, P, A, E) > 48u * 1024u) R = 16;
  const size_t bytes = staged_bytes<T>(R, H, P, A, E);
  if (bytes > 100u * 1024u) return false;
  auto kern = A == 5 ? bd_posterior_staged_kernel<T, G, 5>
                     : (A == 25 ? bd_posterior_staged_kernel<T, G, 25> : bd_posterior_staged_kernel<T, G, 0>);
  if (bytes > 48u * 1024u &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  kern<<<(unsigned)((n + R - 1) / R), R * G, bytes, st>>>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx,
                                                          beta, n, H, P, A, E, R);
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
  // smallest group that covers max(H, P) in one chunk; larger tables loop in chunks of 32
  if (need <= 8) {
    if (launch_staged<T, 8>(probs, alive, hyp_pair, pair_w, qdiff, n_valid, act_idx, beta, n, H, P, A, n_entries, st))
      return gc_check_launch("gc_bd_posterior");
  } else if (need <= 16) {
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

}  // namespace

extern "C" {

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

}  // extern "C"
