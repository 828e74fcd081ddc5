// step2_host_shim.cpp - TEST INFRASTRUCTURE: the product's transition function (gym-cooking_b200/csrc/gc_step2.cuh,
// the code the CUDA step kernels inline) compiled as HOST code behind a two-function C interface, so that the
// CPU test suite can hold it against the reference fixtures and the oracle without a GPU
// (tests/test_step2_host.py builds this file with g++).  Nothing in the product uses it.
#include <string.h>

#include "../gym-cooking_b200/csrc/gc_step2.cuh"

namespace {
const gcs2::StaticTables kStatic = gcs2::make_static_tables();

template <int NA, int NOBJ>
void run(const gcs2::LevelTables& L, uint32_t* state, const uint8_t* actions, uint8_t* rd, uint8_t* executed, uint32_t* ncoll,
         long long n) {
  for (long long i = 0; i < n; i++) {
    uint32_t* w = state + 4 * i;
    uint32_t aw = 0, exec = 0x04040404u, nc = 0;
    for (int a = 0; a < NA; a++) aw |= (uint32_t)actions[i * NA + a] << (8 * a);
    bool done = true, success = !(L.max_t24 != 0u && (w[0] & 0x7F000000u) >= L.max_t24);
    if (!(w[0] >> 31)) {  // sticky done, as in step2_one (gc_env.cu)
      gcs2::Env<NOBJ> e;
      gcs2::unpack<NOBJ>(w[0], w[1], w[2], w[3], e);
      nc = gcs2::step<NA, NOBJ, true>(e, aw, kStatic, L, done, success, exec);
      gcs2::pack<NOBJ>(e, w[0], w[1], w[2], w[3]);
    }
    rd[i] = (uint8_t)((done ? GC_RD_DONE : 0) | (success ? GC_RD_REWARD : 0));
    if (ncoll) ncoll[i] = nc;
    if (executed)
      for (int a = 0; a < NA; a++) executed[i * NA + a] = (uint8_t)(exec >> (8 * a));
  }
}
}  // namespace

extern "C" {

// the level's reset state (what gc_env_reset writes)
void s2h_initial_state(const gc_level* level, int n_agents, uint32_t w[4]) { gcs2::initial_state(*level, n_agents, w); }

// one step of n envs in place (state uint32[n][4], actions uint8[n][n_agents]); 0 on success
int s2h_step(const gc_level* level, int n_agents, uint32_t* state, const uint8_t* actions, uint8_t* rd, uint8_t* executed,
             uint32_t* ncoll, long long n) {
  gcs2::LevelTables L;
  gcs2::fill_level_tables(*level, n_agents, &L);
  const bool six = level->n_objects > 4;
  switch (n_agents * 2 + (six ? 1 : 0)) {
    case 2: run<1, 4>(L, state, actions, rd, executed, ncoll, n); return 0;
    case 3: run<1, 6>(L, state, actions, rd, executed, ncoll, n); return 0;
    case 4: run<2, 4>(L, state, actions, rd, executed, ncoll, n); return 0;
    case 5: run<2, 6>(L, state, actions, rd, executed, ncoll, n); return 0;
    case 6: run<3, 4>(L, state, actions, rd, executed, ncoll, n); return 0;
    case 7: run<3, 6>(L, state, actions, rd, executed, ncoll, n); return 0;
    case 8: run<4, 4>(L, state, actions, rd, executed, ncoll, n); return 0;
    case 9: run<4, 6>(L, state, actions, rd, executed, ncoll, n); return 0;
  }
  return -1;
}

}  // extern "C"
