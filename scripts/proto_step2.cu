// proto_step2.cu - timing prototype of the byte-plane step kernel (plain step, single level).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I. scripts/proto_step2.cu -o scripts/proto_step2
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../gym-cooking_b200/csrc/gc_step2.cuh"

#ifndef NT
#define NT 256
#endif
#ifndef MINB
#define MINB 6
#endif

using namespace gcs2;

template <int NA, int NOBJ>
__global__ void __launch_bounds__(NT, MINB)
step2_kernel(const Tables* __restrict__ gT, uint4* __restrict__ state, const uint8_t* __restrict__ actions,
             uint8_t* __restrict__ reward_done, uint32_t n) {
  __shared__ __align__(16) Tables T;
  __shared__ __align__(16) uint4 s_stage[NT];
  const uint32_t stride = gridDim.x * NT;
  uint32_t i = blockIdx.x * NT + threadIdx.x;
  const uint32_t slot = (uint32_t)__cvta_generic_to_shared(&s_stage[threadIdx.x]);
  uint32_t a_next = 0x04040404u;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    const uint4* s = reinterpret_cast<const uint4*>(gT);
    uint4* d = reinterpret_cast<uint4*>(&T);
#pragma unroll
    for (int k0 = 0; k0 < (int)(sizeof(Tables) / 16); k0 += NT) {
      const int k = k0 + (int)threadIdx.x;
      if (k < (int)(sizeof(Tables) / 16)) d[k] = __ldg(s + k);
    }
  }
  if (i < n) {
    if ((threadIdx.x & 7u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(state + i));
    if ((threadIdx.x & 31u) == 0u) asm volatile("prefetch.global.L2 [%0];" ::"l"(actions + (size_t)i * NA));
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (i < n) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(state + i) : "memory");
    a_next = reinterpret_cast<const uint16_t*>(actions)[i];
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  __syncthreads();
  const uint32_t max_t24 = T.lv.max_t24;
  for (; i < n; i += stride) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    uint4 s = s_stage[threadIdx.x];
    const uint32_t aw = a_next;
    const uint32_t inext = i + stride;
    if (inext < n) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(state + inext) : "memory");
      a_next = reinterpret_cast<const uint16_t*>(actions)[inext];
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    bool done, success;
    if (s.x >> 31) {
      done = true;
      success = !(max_t24 != 0u && (s.x & 0x7F000000u) >= max_t24);
    } else {
      Env<NOBJ> e;
      unpack<NOBJ>(s.x, s.y, s.z, s.w, e);
      uint32_t ex;
      step<NA, NOBJ, false>(e, aw, T.st, T.lv, done, success, ex);
      pack<NOBJ>(e, s.x, s.y, s.z, s.w);
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(state + i), "r"(s.x), "r"(s.y),
                   "r"(s.z), "r"(s.w)
                   : "memory");
    }
    reward_done[i] = (uint8_t)((done ? 1 : 0) | (success ? 2 : 0));
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

int main(int argc, char** argv) {
  const uint32_t n = 1u << 20;
  const int RING = 16, HORIZON = 100;
  // partial-divider_tl
  const char* rows[7] = {"-----t-", "/  -  l", "/  -  -", "*  -  -", "-  -  -", "-     p", "-----p-"};
  gc_level g;
  memset(&g, 0, sizeof g);
  memset(g.cell_type, GC_CELL_COUNTER, 64);
  for (int k = 0; k < 6; k++) g.object_init[k] = GC_SLOT_DEAD;
  for (int y = 0; y < 7; y++)
    for (int x = 0; x < 7; x++) {
      char ch = rows[y][x];
      int c = y * 8 + x, m = 0;
      if (ch == ' ') g.cell_type[c] = GC_CELL_FLOOR;
      else if (ch == '/') g.cell_type[c] = GC_CELL_CUTBOARD;
      else if (ch == '*') { g.cell_type[c] = GC_CELL_DELIVERY; g.delivery_cell = c; }
      else if (ch == 't') m = 1; else if (ch == 'l') m = 2; else if (ch == 'p') m = 8;
      if (m) g.object_init[g.n_objects++] = (uint16_t)(m | (c << 7));
    }
  g.n_goals = 2; g.goal_mask[0] = 8 | 1 | 16; g.goal_mask[1] = 8 | 2 | 32; g.max_timesteps = HORIZON;
  g.agent_cell[0] = 1 * 8 + 2; g.agent_cell[1] = 1 * 8 + 4; g.n_agent_starts = 2;
  static Tables T;
  T.st = make_static_tables();
  fill_level_tables(g, 2, &T.lv);
  Tables* dT;
  CK(cudaMalloc(&dT, sizeof(Tables)));
  CK(cudaMemcpy(dT, &T, sizeof(Tables), cudaMemcpyHostToDevice));
  std::vector<uint4*> st(RING);
  std::vector<uint8_t*> act(RING), rd(RING);
  std::vector<uint4> init(n, make_uint4(T.lv.init[0], T.lv.init[1], T.lv.init[2], T.lv.init[3]));
  std::vector<uint8_t> hact((size_t)n * 2 * HORIZON);
  srand(1);
  for (auto& a : hact) a = rand() % 5;
  for (int r = 0; r < RING; r++) {
    CK(cudaMalloc(&st[r], (size_t)n * 16));
    CK(cudaMalloc(&act[r], (size_t)n * 2 * HORIZON));
    CK(cudaMalloc(&rd[r], n));
    CK(cudaMemcpy(st[r], init.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(act[r], hact.data(), hact.size(), cudaMemcpyHostToDevice));
  }
  int per_sm = 0, sms = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step2_kernel<2, 4>, NT, 0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  if (argc > 1) per_sm = atoi(argv[1]);
  const unsigned grid = sms * per_sm;
  printf("threads %d, CTAs/SM %d, grid %u\n", NT, per_sm, grid);
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NT);
  cfg.gridDim = dim3(grid);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  auto launch = [&](int r, int t) {
    CK(cudaLaunchKernelEx(&cfg, step2_kernel<2, 4>, (const Tables*)dT, st[r], (const uint8_t*)(act[r] + (size_t)t * n * 2), rd[r], n));
  };
  // graph of 25 t-values x RING
  cudaGraph_t graph;
  cudaGraphExec_t gexec;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int t = 0; t < 25; t++)
    for (int r = 0; r < RING; r++) launch(r, t);
  CK(cudaStreamEndCapture(s, &graph));
  CK(cudaGraphInstantiate(&gexec, graph, 0));
  CK(cudaGraphLaunch(gexec, s));
  CK(cudaStreamSynchronize(s));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; rep++) {
    // reset so that the envs are mid-episode (not done)
    for (int r = 0; r < RING; r++) CK(cudaMemcpyAsync(st[r], init.data(), (size_t)n * 16, cudaMemcpyHostToDevice, s));
    CK(cudaGraphLaunch(gexec, s));  // t 0..24
    CK(cudaEventRecord(e0, s));
    CK(cudaGraphLaunch(gexec, s));  // 400 more steps (t 25..49 in episode time)
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("graph: %.3f us per 2^20-env step, %.1f GB/s algorithmic (35 B/env)\n", ms * 1e3 / 400, 35.0 * n / (ms * 1e-3 / 400) / 1e9);
  }
  // plain loop
  for (int r = 0; r < RING; r++) CK(cudaMemcpyAsync(st[r], init.data(), (size_t)n * 16, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(e0, s));
  for (int t = 0; t < 25; t++)
    for (int r = 0; r < RING; r++) launch(r, t);
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("loop:  %.3f us per step\n", ms * 1e3 / 400);
  // checksum
  std::vector<uint4> out(n);
  CK(cudaMemcpy(out.data(), st[0], (size_t)n * 16, cudaMemcpyDeviceToHost));
  unsigned long long h = 0;
  for (auto& v : out) h = h * 1000003ull + v.x + 3ull * v.y + 7ull * v.z + 11ull * v.w;
  printf("checksum %llx\n", h);
  return 0;
}
