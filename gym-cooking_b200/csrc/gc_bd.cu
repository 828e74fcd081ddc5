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
