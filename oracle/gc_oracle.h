/*
 * gc_oracle.h - CPU restatement of the reference algorithms on the three hot paths.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or call this.  The product library
 * (gym-cooking_b200/csrc -> libgymcook.so) never includes or links anything under oracle/.
 *
 * The oracle deliberately uses a different formulation from the CUDA kernels: an unpacked
 * array-of-structs world, explicit loops, doubles - a plain reading of the reference's
 * Python (file:line cited at every function, relative to /root/reference/gym_cooking/).
 * Parity status: PINNED for path A, the distance heuristic of path B and path C by the
 * committed fixtures under tests/golden/ that oracle/gen_golden.py produced by running the
 * unmodified reference in the build container (the reference itself has no tests or golden
 * vectors).
 */
#ifndef GC_ORACLE_H
#define GC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCO_MAX_AGENTS 4
#define GCO_MAX_OBJS 6
#define GCO_MAX_GOALS 4

enum { GCO_FLOOR = 0, GCO_COUNTER = 1, GCO_CUTBOARD = 2, GCO_DELIVERY = 3 };

typedef struct {
  int width, height;
  int type[8][8]; /* [y][x] */
  int n_agent_starts;
  int agent_x[GCO_MAX_AGENTS], agent_y[GCO_MAX_AGENTS];
  int n_objs;
  int obj_mask[GCO_MAX_OBJS], obj_x[GCO_MAX_OBJS], obj_y[GCO_MAX_OBJS];
  int n_goals;
  int goal_mask[GCO_MAX_GOALS];
  int delivery_x, delivery_y; /* first Delivery square, -1 if none */
  int max_timesteps;
} gco_level;

typedef struct {
  int x, y;
  int hold; /* object index or -1 */
} gco_agent;

typedef struct {
  int alive;
  int mask;
  int x, y;    /* square it lies on; undefined while held */
  int held_by; /* agent index or -1 */
} gco_obj;

typedef struct {
  int t, done, successful;
  int n_agents, n_objs;
  gco_agent ag[GCO_MAX_AGENTS];
  gco_obj ob[GCO_MAX_OBJS];
} gco_env;

/* level .txt -> tables (env.load_level :130-198).  0 on success. */
int gco_level_parse(const char* txt, int max_timesteps, gco_level* out);
/* env.reset :201-250 */
void gco_reset(const gco_level* lv, int n_agents, gco_env* e);
/* env.step :255-306.  actions[i] in 0..4.  Returns the number of CollisionRepr appended
 * (env:747-752); executed[] (nullable) receives the post-collision actions. */
int gco_step(const gco_level* lv, gco_env* e, const uint8_t* actions, uint8_t* executed);

/* packed 128-bit form shared with the product (include/gymcook.h) */
void gco_pack(const gco_env* e, uint32_t w[4]);
void gco_unpack(const uint32_t w[4], int n_agents, gco_env* e);
/* canonical hash of a packed state */
uint64_t gco_hash_packed(const uint32_t w[4], int n_agents);
/* canonical sorted item keys (mask<<7|cell<<1|held), 0x3FFF padded; returns live count */
int gco_canonical_keys(const uint32_t w[4], uint16_t keys[GCO_MAX_OBJS]);

/* philox4x32-10 action stream shared with gc_env_rollout / gc_fill_random_actions */
void gco_philox_actions(uint64_t seed, uint32_t t, uint64_t env, uint8_t out[4]);

/* the same stream materialised: actions[n_steps][n][n_agents] */
void gco_fill_actions(uint8_t* actions, int64_t n, int n_agents, int n_steps, int t0, int64_t env0, uint64_t seed);

/* whole-episode replay from reset (tests) */
void gco_replay(const gco_level* lv, int n_agents, const uint8_t* actions, int n_steps,
                uint32_t* states, uint8_t* reward_done, uint8_t* ncoll, uint8_t* executed);

/* Batched helpers (also the CPU baseline legs of bench.py). */
/* n envs, one step, packed in place; OpenMP over envs when n_threads > 1. */
void gco_step_batch(const gco_level* lv, uint32_t* state, const uint8_t* actions,
                    uint8_t* reward_done, uint32_t* collisions, int64_t n, int n_agents,
                    int n_threads);
/* rollout with the philox stream; hash_trace nullable [n_steps][n] */
void gco_rollout_batch(const gco_level* lv, uint32_t* state, uint8_t* reward_done,
                       uint64_t* hash_trace, uint32_t* collisions, int64_t n, int n_agents,
                       int n_steps, int t0, int64_t env0, uint64_t seed, int n_threads);

/* ---- path C: posterior (bayesian_delegator.py:1045-1072, 461-689; dutils:177-193) ---- */
void gco_bd_posterior(double* probs, const uint8_t* alive, const uint8_t* hyp_pair,
                      const uint8_t* pair_w, const double* qdiff, const uint8_t* n_valid,
                      const uint8_t* act_idx, double beta, int64_t n, int H, int P, int A,
                      int n_entries);

/* ---- path B ---------------------------------------------------------------------------- */
typedef struct {
  int kind; /* 1 chop 2 merge 3 deliver */
  int a, b, goal;
} gco_subtask;
/* env.get_lower_bound_for_subtask_given_objs :594-664 (+ world.py:115-283).  agent j = -1
 * for a single-agent pair. */
double gco_lower_bound(const gco_level* lv, const gco_env* e, const gco_subtask* st, int ai, int aj);
/* exact level-0 V* and Q(start, .) by forward uniform-cost search over full env states with
 * the real transition function.  q has 25 entries (5 used when aj < 0), +inf = invalid.
 * Returns status: 0 ok, 1 goal satisfied at start, 2 unreachable, 3 budget exceeded. */
void gco_set_planner_level(int level); /* 0 (default) or 1: planning world of the next calls */
int gco_subtask_q(const gco_level* lv, const gco_env* e, const gco_subtask* st, int ai, int aj,
                  double* v, double* q, int max_states);

#ifdef __cplusplus
}
#endif
#endif
