// step2_host_check.cpp - runs gcs2::step (gc_step2.cuh, compiled as host code) against the C oracle's
// gco_step on random walks: every level file given on the command line x 1..4 agents x many envs.
//   g++ -O2 -I. scripts/step2_host_check.cpp oracle/gc_oracle.c -o /tmp/step2_check -lm -lpthread
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../gym-cooking_b200/csrc/gc_step2.cuh"
extern "C" {
#include "../oracle/gc_oracle.h"
}

static const gcs2::StaticTables kStatic = gcs2::make_static_tables();

static void to_gc_level(const gco_level& o, gc_level* g) {
  memset(g, 0, sizeof(*g));
  memset(g->cell_type, GC_CELL_COUNTER, sizeof(g->cell_type));
  g->width = o.width; g->height = o.height; g->n_agent_starts = o.n_agent_starts; g->n_objects = o.n_objs;
  g->n_goals = o.n_goals; g->delivery_cell = o.delivery_y * 8 + o.delivery_x; g->max_timesteps = o.max_timesteps;
  for (int y = 0; y < o.height; y++) for (int x = 0; x < o.width; x++) g->cell_type[y * 8 + x] = (uint8_t)o.type[y][x];
  for (int i = 0; i < o.n_agent_starts; i++) g->agent_cell[i] = (uint8_t)(o.agent_y[i] * 8 + o.agent_x[i]);
  for (int k = 0; k < GC_MAX_OBJECTS; k++)
    g->object_init[k] = k < o.n_objs ? (uint16_t)(o.obj_mask[k] | ((o.obj_y[k] * 8 + o.obj_x[k]) << 7)) : GC_SLOT_DEAD;
  for (int k = 0; k < o.n_goals; k++) g->goal_mask[k] = (uint8_t)o.goal_mask[k];
}

static void env_to_words(const gco_env& e, uint32_t w[4]) {
  w[0] = 0;
  for (int i = 0; i < e.n_agents; i++) w[0] |= (uint32_t)(e.ag[i].y * 8 + e.ag[i].x) << (6 * i);
  w[0] |= (uint32_t)(e.t & 127) << 24 | (e.done ? 0x80000000u : 0u);
  uint32_t P[6], M[6];
  for (int k = 0; k < 6; k++) {
    P[k] = GC_PLACE_DEAD; M[k] = 0;
    if (k < e.n_objs && e.ob[k].alive) {
      M[k] = (uint32_t)e.ob[k].mask;
      P[k] = e.ob[k].held_by >= 0 ? GC_PLACE_HELD + 1 + e.ob[k].held_by : (uint32_t)(e.ob[k].y * 8 + e.ob[k].x);
    }
  }
  w[1] = P[0] | P[1] << 8 | P[2] << 16 | P[3] << 24;
  w[2] = M[0] | M[1] << 8 | M[2] << 16 | M[3] << 24;
  w[3] = P[4] | P[5] << 8 | M[4] << 16 | M[5] << 24;
}

template <int NA, int NOBJ>
static long run(const gco_level& lv, const gcs2::Tables& T, int n_envs, int n_steps, unsigned seed) {
  long bad = 0, steps = 0;
  srand(seed);
  for (int n = 0; n < n_envs; n++) {
    gco_env e;
    gco_reset(&lv, NA, &e);
    uint32_t w[4];
    env_to_words(e, w);
    if (memcmp(w, T.lv.init, 16) != 0) { printf("init differs\n"); return 1; }
    for (int s = 0; s < n_steps; s++) {
      uint8_t act[4], ex[4];
      uint32_t aw = 0;
      for (int i = 0; i < NA; i++) { act[i] = (uint8_t)(rand() % 6 == 5 ? 5 + rand() % 250 : rand() % 5); aw |= (uint32_t)act[i] << (8 * i); }
      const bool was_done = e.done;
      const int ncoll = gco_step(&lv, &e, act, ex);
      uint32_t want[4];
      env_to_words(e, want);
      bool done = true, success = false;
      uint32_t nc = 0, exw = 0x04040404u;
      if (!(w[0] >> 31)) {
        gcs2::Env<NOBJ> E;
        gcs2::unpack<NOBJ>(w[0], w[1], w[2], w[3], E);
        nc = gcs2::step<NA, NOBJ, true>(E, aw, T.st, T.lv, done, success, exw);
        gcs2::pack<NOBJ>(E, w[0], w[1], w[2], w[3]);
      }
      steps++;
      bool ok = memcmp(w, want, 16) == 0 && (int)nc == ncoll;
      if (!was_done) {
        ok = ok && done == (bool)e.done && success == (bool)e.successful;
        for (int i = 0; i < NA; i++) ok = ok && ((exw >> (8 * i)) & 0xFF) == ex[i];
      }
      if (!ok) {
        if (bad < 5)
          printf("mismatch env %d step %d: got %08x %08x %08x %08x want %08x %08x %08x %08x ncoll %u/%d done %d/%d succ %d/%d exec %08x want %d %d %d %d act %d %d %d %d\n", n, s,
                 w[0], w[1], w[2], w[3], want[0], want[1], want[2], want[3], nc, ncoll, done, e.done, success, e.successful, exw, ex[0], ex[1], ex[2], ex[3], act[0], act[1], act[2], act[3]);
        bad++;
        memcpy(w, want, 16);
      }
    }
  }
  printf("  NA=%d NOBJ=%d: %ld steps, %ld mismatches\n", NA, NOBJ, steps, bad);
  return bad;
}

int main(int argc, char** argv) {
  long bad = 0;
  const int n_envs = 3000;
  for (int a = 1; a < argc; a++) {
    FILE* f = fopen(argv[a], "rb");
    if (!f) { perror(argv[a]); return 2; }
    std::string txt; char buf[4096]; size_t r;
    while ((r = fread(buf, 1, sizeof buf, f)) > 0) txt.append(buf, r);
    fclose(f);
    for (int max_t : {100, 30, 0}) {
      gco_level lv;
      if (gco_level_parse(txt.c_str(), max_t, &lv)) { printf("parse failed %s\n", argv[a]); return 2; }
      gc_level g;
      to_gc_level(lv, &g);
      printf("%s max_t=%d objs=%d\n", argv[a], max_t, lv.n_objs);
      for (int na = 1; na <= lv.n_agent_starts && na <= 4; na++) {
        gcs2::Tables T;
        T.st = kStatic;
        gcs2::fill_level_tables(g, na, &T.lv);
        const bool six = lv.n_objs > 4;
        const int steps = max_t ? 110 : 140;
        switch (na * 2 + six) {
          case 2: bad += run<1, 4>(lv, T, n_envs, steps, 1); break;
          case 3: bad += run<1, 6>(lv, T, n_envs, steps, 1); break;
          case 4: bad += run<2, 4>(lv, T, n_envs, steps, 2); break;
          case 5: bad += run<2, 6>(lv, T, n_envs, steps, 2); break;
          case 6: bad += run<3, 4>(lv, T, n_envs, steps, 3); break;
          case 7: bad += run<3, 6>(lv, T, n_envs, steps, 3); break;
          case 8: bad += run<4, 4>(lv, T, n_envs, steps, 4); break;
          case 9: bad += run<4, 6>(lv, T, n_envs, steps, 4); break;
        }
        if (!six && na == 2) bad += run<2, 6>(lv, T, 500, steps, 9);  // the 6-object form on a 4-object level
      }
    }
  }
  printf(bad ? "FAILED: %ld mismatches\n" : "all equal\n", bad);
  return bad ? 1 : 0;
}
