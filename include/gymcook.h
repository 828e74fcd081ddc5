/*
 * gymcook.h - C-ABI of libgymcook.so, the B200-native batched drop-in for the three
 * data-parallel hot paths of deletfsi/gym-cooking:
 *
 *   (A) OvercookedEnvironment.step      gym_cooking/envs/overcooked_environment.py:255-306
 *   (B) E2E_BRTDP subtask value / Q      gym_cooking/navigation_planner/planners/e2e_brtdp.py:987-1076
 *   (C) BayesianDelegator posterior      gym_cooking/delegation_planner/bayesian_delegator.py:1026-1072
 *
 * The reference is pure Python and has no FFI of its own; the binding a maintainer adds is
 * the ctypes stub shown in INTEGRATION.md.  Conventions (all entry points):
 *   - plain C, extern "C"; no torch / C++ types in any signature;
 *   - every array pointer marked "device" is a CUDA device pointer owned by the caller
 *     (e.g. torch.Tensor.data_ptr()); the library never frees or retains it;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all
 *     kernels are asynchronous on it, the caller synchronises;
 *   - return value 0 = success, negative = error (GC_E_*), message via gc_last_error()
 *     (thread-local);
 *   - there is NO CPU fallback: if no CUDA device is usable the compute entry points fail
 *     with GC_E_CUDA.
 *
 * Action encoding (World.NAV_ACTIONS order + stay; utils/world.py:16, navigation_planner/utils.py:65,88):
 *     0=(0,+1)  1=(0,-1)  2=(-1,0)  3=(+1,0)  4=(0,0)
 *
 * Packed env state: one 128-bit word per env (uint32[4]), grid <= 8x8, <= 4 agents,
 * <= 6 movable objects.  cell = y*8 + x.
 *   w[0]  bits  0-5   agent-1 cell      bits  6-11 agent-2 cell
 *         bits 12-17  agent-3 cell      bits 18-23 agent-4 cell   (unused agents: 0)
 *         bits 24-30  t (env.t, saturates at 127)                  bit 31 done (sticky)
 *   w[1]  place bytes of objects 0..3 (object k in byte k), w[2] their content-mask bytes,
 *   w[3]  objects 4 and 5: byte 0 = place 4, byte 1 = place 5, byte 2 = mask 4, byte 3 = mask 5.
 *         mask byte:  bit0 Tomato, bit1 Lettuce, bit2 Onion, bit3 Plate present;
 *                     bit4/5/6 Tomato/Lettuce/Onion chopped; bit 7 clear
 *         place byte: 0..63 = cell of an object lying on a counter/cutboard/delivery square,
 *                     GC_PLACE_HELD + h (0x41..0x44) = held by agent-<h>,
 *                     GC_PLACE_DEAD (0x47) = dead object (its contents were merged into another
 *                     object; mask 0); bit 7 clear
 *         A state without objects 4 and 5 has w[3] = GC_W3_EMPTY.
 *   Byte planes are what the step kernel computes on: "which object is in this hand / on this
 *   square" is one SIMD byte compare over a plane, gathering its mask one dot product (dp4a).
 *   Objects are created in level-file scan order (row by row; env.load_level :149-174).  On a
 *   merge the HELD object absorbs the counter object (SimAgent.acquire, utils/agent.py:408-414)
 *   so the held object survives and the counter object dies.
 *   (ABI version 1 packed each object as a 16-bit slot mask | cell << 7 | holder << 13; that form
 *   survives as gc_level.object_init and as the planners' internal working form.)
 *
 * Canonical item key (used by the hash and by parity tests, independent of slot order):
 *   key = mask<<7 | cell<<1 | held, with cell = holder's cell for held objects; live items
 *   sorted ascending.
 */
#ifndef GYMCOOK_H
#define GYMCOOK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GC_ABI_VERSION 2

#define GC_MAX_AGENTS 4
#define GC_MAX_OBJECTS 6
#define GC_MAX_GOALS 4
#define GC_MAX_CELLS 64
#define GC_GRID_STRIDE 8
#define GC_MAX_SUBTASKS 32
#define GC_MAX_PAIRS 128
#define GC_MAX_JOINT_ACTIONS 25
#define GC_MAX_HYPOTHESES 128
#define GC_MAX_LEVELS 16

/* error codes */
#define GC_OK 0
#define GC_E_ARG (-1)    /* bad argument */
#define GC_E_PARSE (-2)  /* level text not understood / outside the supported envelope */
#define GC_E_CUDA (-3)   /* CUDA runtime error, or no device */
#define GC_E_LIMIT (-4)  /* a compile-time limit above was exceeded */

/* cell types (utils/core.py:59-120) */
#define GC_CELL_FLOOR 0
#define GC_CELL_COUNTER 1
#define GC_CELL_CUTBOARD 2
#define GC_CELL_DELIVERY 3

/* content-mask bits */
#define GC_M_TOMATO 0x01
#define GC_M_LETTUCE 0x02
#define GC_M_ONION 0x04
#define GC_M_PLATE 0x08
#define GC_M_CHOP_T 0x10
#define GC_M_CHOP_L 0x20
#define GC_M_CHOP_O 0x40
#define GC_SLOT_DEAD 0xE000u
/* place byte of an object (packed state words 1..3): 0..63 = the cell it lies on,
 * GC_PLACE_HELD + h = held by agent-<h> (h = 1..4), GC_PLACE_DEAD = merged into another object */
#define GC_PLACE_HELD 0x40u
#define GC_PLACE_DEAD 0x47u
#define GC_W3_EMPTY 0x00004747u /* word 3 of a state without objects 4 and 5 */

/* subtask kinds (recipe_planner/utils.py:114-162).  GC_ST_NONE is the `None` subtask. */
#define GC_ST_NONE 0
#define GC_ST_CHOP 1
#define GC_ST_MERGE 2
#define GC_ST_DELIVER 3

/* reward_done byte written by gc_env_step: bit0 = done(), bit1 = reward() (== successful) */
#define GC_RD_DONE 0x01
#define GC_RD_REWARD 0x02

/* One subtask in mask form.  For CHOP: a = fresh mask of the food, goal = chopped mask.
 * MERGE: a, b = masks of the two start objects (foods in their last state, env:573-584 via
 * nav_utils.get_subtask_obj :206-229), goal = a|b.  DELIVER: a = goal = mask to deliver. */
typedef struct gc_subtask {
  uint8_t kind;
  uint8_t a;
  uint8_t b;
  uint8_t goal;
} gc_subtask;

/* Static per-level tables.  Produced on the host by gc_level_parse from the reference's
 * level .txt format (utils/levels/<level>.txt; env.load_level :130-198), POD, 256 bytes. */
typedef struct gc_level {
  int32_t width;                         /* world.width  (env:196) */
  int32_t height;                        /* world.height (env:197) */
  int32_t n_agent_starts;                /* agent "x y" lines present in the file (<= 4) */
  int32_t n_objects;                     /* movable objects at reset (<= 6) */
  int32_t n_goals;                       /* Deliver goals, one per recipe (env.done :344-359) */
  int32_t delivery_cell;                 /* FIRST Delivery square (env:349), -1 if none */
  int32_t max_timesteps;                 /* arglist.max_num_timesteps (main.py:24); 0 = no limit */
  int32_t n_subtasks;                    /* filled by gc_level_set_subtasks (host recipe planner) */
  uint8_t cell_type[GC_MAX_CELLS];       /* GC_CELL_*, index y*8+x; cells outside the map = COUNTER */
  uint8_t agent_cell[GC_MAX_AGENTS];     /* start cells (file order) */
  uint16_t object_init[GC_MAX_OBJECTS];  /* objects at reset: mask | cell << 7; unused = GC_SLOT_DEAD */
  uint8_t goal_mask[GC_MAX_GOALS];       /* content mask that must lie on delivery_cell */
  gc_subtask subtask[GC_MAX_SUBTASKS];   /* recipe subtasks in the host's order */
  uint8_t recipe_code[GC_MAX_GOALS];     /* 1 SimpleTomato 2 SimpleLettuce 3 Salad 4 OnionSalad */
  uint8_t reserved[8];
} gc_level;

/* ---- library ------------------------------------------------------------------------ */
int gc_version(void);
const char* gc_last_error(void);
/* number of usable CUDA devices (0 without a GPU; never an error) */
int gc_device_count(void);

/* ---- level loader (host) -------------------------------------------------------------
 * Replaces OvercookedEnvironment.load_level (env:130-198) + the Deliver-goal part of done()
 * (env:344-359).  `txt` is the content of a level file: map rows, blank line, recipe class
 * names, blank line, agent "x y" lines.  Supported envelope: grid <= 8x8 whose outer ring
 * has no floor, <= 6 objects with at most one Tomato, one Lettuce and one Onion, recipes
 * SimpleTomato/SimpleLettuce/Salad/OnionSalad. */
int gc_level_parse(const char* txt, int len, int max_timesteps, gc_level* out);
int gc_level_set_subtasks(gc_level* lvl, const gc_subtask* subtasks, int n);

/* ---- (A) env transition --------------------------------------------------------------- */
/* reset(): every env := the level's initial state (env.reset :201-250).  `level_id` (device,
 * nullable) selects levels[level_id[i]] per env; NULL = levels[0] for all.  `levels` is a
 * HOST array; the tables are copied into kernel parameters. */
int gc_env_reset(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
                 uint32_t* state /*device uint32[n][4]*/, int64_t n, int n_agents, void* stream);

/* Builds and caches the device tables of a level set on the current device (what the first gc_env_step /
 * gc_env_rollout of that level set does implicitly: an allocation and a synchronous copy).  Call it before a
 * stream capture whose first captured operation would be that first step. */
int gc_env_prepare(const gc_level* levels, int n_levels, int n_agents);

/* step(): one joint transition for n envs, in place (env.step :255-306: t+=1,
 * check_collisions :724-762, execute_navigation :767-770 -> interact, done/reward :316-376).
 * Envs whose done bit is set are left untouched (sticky done) and report their old outcome.
 *   actions      device uint8[n][n_agents], values 0..4 (>4 is treated as stay)
 *   reward_done  device uint8[n]        nullable   GC_RD_* bits
 *   hash         device uint64[n]       nullable   canonical state hash after the step
 *   collisions   device uint32[n]       nullable   += number of CollisionRepr this step (env:747-752)
 *   executed     device uint8[n][n_agents] nullable  post-collision actions (env.agent_actions) */
int gc_env_step(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
                uint32_t* state /*device*/, const uint8_t* actions /*device*/,
                uint8_t* reward_done, uint64_t* hash, uint32_t* collisions, uint8_t* executed,
                int64_t n, int n_agents, void* stream);

/* step() for a caller whose actions and results live in HOST memory (the gym-style call:
 * `obs, reward, done, info = env.step(actions)`, env:255-306): copies `actions_host`
 * (uint8[n][n_agents], pinned for an asynchronous copy) into `actions_dev`, runs gc_env_step,
 * copies the reward/done bytes back into `reward_done_host` and waits for the stream - one call
 * instead of three stream operations plus a synchronisation on the caller's side.  The state
 * stays on the device.  `actions_dev` / `reward_done_dev` are caller-owned device buffers of
 * n * n_agents / n bytes.  Results come back as bytes (`reward_done_host`, n bytes, nullable) and/or
 * as two bit planes per 32 envs (`rd_bits_dev` / `rd_bits_host`, uint32[(n+31)/32][2], nullable
 * pair: word 0 = done bits, word 1 = reward bits of envs 32w .. 32w+31) - a quarter of the bytes
 * over PCIe.  `rd_bits_dev` must be 8-byte aligned (one 8-byte store per 32 envs). */
int gc_env_step_host(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
                     uint32_t* state /*device*/, const uint8_t* actions_host, uint8_t* actions_dev,
                     uint8_t* reward_done_dev, uint8_t* reward_done_host, uint32_t* rd_bits_dev,
                     uint32_t* rd_bits_host, uint32_t* collisions /*device*/, int64_t n, int n_agents,
                     void* stream);

/* Prepared steps: everything gc_env_step derives per call (validation, device tables, launch
 * configuration) fixed once for one single-level batch; gc_step_plan_run is then one kernel launch
 * (the plain step: state in place + reward_done bytes).  At 2^20 envs the kernel takes ~9 us, less
 * than the argument marshalling of a 12-argument call from Python.
 *   flags  GC_PLAN_JOINT_ACTIONS: `actions` holds ONE joint index per env instead of n_agents bytes,
 *          j = sum_i action_i * 5^(n_agents-1-i) (two agents: 5*a_1 + a_2), uint8 for <= 3 agents,
 *          uint16 for 4; an index >= 5^n_agents means "everybody stays".
 * gc_step_plan_run_host: the gym-style call for a caller whose actions live in (pinned) host memory -
 * copies them in (plan-owned staging), steps, copies the done / reward bit planes (uint32[(n+31)/32][2],
 * as gc_env_step_host) into `rd_bits_host` and waits for the stream.
 * A plan holds raw pointers to `state` / `reward_done`: it must not outlive them.  One plan per
 * (batch, device); the first plan of a level cannot be created inside a stream capture. */
#define GC_PLAN_JOINT_ACTIONS 1
typedef struct gc_step_plan gc_step_plan;
int gc_step_plan_create(const gc_level* level, uint32_t* state /*device*/, uint8_t* reward_done /*device*/,
                        int64_t n, int n_agents, int flags, gc_step_plan** out);
int gc_step_plan_run(const gc_step_plan* plan, const uint8_t* actions /*device*/, void* stream);
int gc_step_plan_run_host(gc_step_plan* plan, const uint8_t* actions_host, uint32_t* rd_bits_host, void* stream);
/* The same three stream operations without the final wait (the asynchronous half of a vector-env style
 * step_async / step_wait pair): the caller synchronises `stream` before reading `rd_bits_host`, and must not
 * touch `actions_host` until then.  Two batches on two streams overlap one batch's copies with the other's
 * kernel (PCIe is full duplex). */
int gc_step_plan_enqueue_host(gc_step_plan* plan, const uint8_t* actions_host, uint32_t* rd_bits_host, void* stream);
void gc_step_plan_destroy(gc_step_plan* plan);

/* rollout(): `n_steps` fused transitions with uniform-random actions generated in-kernel:
 * action[t][env][agent] = philox4x32-10(key=(seed_lo,seed_hi), ctr=(t0+t, env0+env, agent, 0)).x % 5
 * (SURVEY.md section 8d cfg-2).  State stays in registers between steps.
 *   hash_trace   device uint64[n_steps][n]  nullable  hash after every step
 *   stats        device uint64[GC_STATS_LEN] nullable  += episode statistics (see gc_stats_*) */
int gc_env_rollout(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
                   uint32_t* state /*device*/, uint8_t* reward_done, uint64_t* hash_trace,
                   uint32_t* collisions, int64_t n, int n_agents, int n_steps, int t0,
                   int64_t env0, uint64_t seed, void* stream);

/* the same philox stream, materialised: actions[n_steps][n][n_agents] (device uint8) */
int gc_fill_random_actions(uint8_t* actions /*device*/, int64_t n, int n_agents, int n_steps,
                           int t0, int64_t env0, uint64_t seed, void* stream);

/* canonical 64-bit state hash, device uint64[n] */
int gc_state_hash(const uint32_t* state /*device*/, uint64_t* hash /*device*/, int64_t n,
                  int n_agents, void* stream);

/* ---- episode statistics (the Bag fields that survive batching; misc/metrics/metrics_bag.py:5-72)
 * stats[0] episodes  [1] successes  [2] sum of t over done envs  [3] sum collisions
 * [4] envs still running  [5..5+128) histogram of t over done envs
 * [133] sum over envs of completed subtasks, read off the state with the levels' subtask tables
 *       (gc_level_set_subtasks): Chop(X) once X is chopped in some live object, Merge(a, b) once an
 *       object contains a | b, Deliver(m) once m lies on a Delivery square - the batched form of the
 *       Bag's num_completed_subtasks (metrics_bag.py:55-61) */
#define GC_STATS_LEN 134
int gc_stats_reduce(const uint32_t* state /*device*/, const uint32_t* collisions /*device, nullable*/,
                    const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
                    uint64_t* stats /*device uint64[GC_STATS_LEN], accumulated into*/,
                    int64_t n, void* stream);

/* ---- (C) Bayesian-Delegation posterior -------------------------------------------------
 * One posterior update per (env, observer) row: BayesianDelegator.bayes_update :1045-1072 with
 * prob_nav_actions :461-689 restated on dumped inputs.
 *   probs     device float32|float64 [n][H]   in place; entries with alive==0 are ignored and set to 0
 *   alive     device uint8 [n][H]             nullable (all alive)
 *   hyp_pair  device uint8 [n][H][n_agents]   for hypothesis h, the pair index (into the P
 *                                            likelihood rows) of each of its <= n_agents entries;
 *                                            0xFF = unused entry
 *   pair_w    device uint8 [n][P]             weight len(subtask_agent_names) (bd:1066), or 1 for greedy
 *   qdiff     device float [n][P][A]          Q(s,a_taken) - Q(s,a) for each valid action (bd:682);
 *                                            for a None pair the reference's pseudo-diffs (bd:626)
 *   n_valid   device uint8 [n][P]             number of valid actions in that row (<= A)
 *   act_idx   device uint8 [n][P]             index of the taken action among the valid ones (bd:689)
 * L[p] = softmax(beta*qdiff[p][:n_valid])[act_idx]; probs[h] *= sum_e pair_w*L; normalise
 * (delegation_planner/utils.py:177-193; total==0 -> uniform). */
int gc_bd_posterior_f32(float* probs, const uint8_t* alive, const uint8_t* hyp_pair,
                        const uint8_t* pair_w, const float* qdiff, const uint8_t* n_valid,
                        const uint8_t* act_idx, float beta, int64_t n, int H, int P, int A,
                        int n_entries, void* stream);
int gc_bd_posterior_f64(double* probs, const uint8_t* alive, const uint8_t* hyp_pair,
                        const uint8_t* pair_w, const double* qdiff, const uint8_t* n_valid,
                        const uint8_t* act_idx, double beta, int64_t n, int H, int P, int A,
                        int n_entries, void* stream);

/* The softmax inputs of prob_nav_actions (bd:461-689) for every (env, likelihood row), built on the
 * device from the planner's Q rows of obs_tm1 and the executed joint action (actions_tm1):
 *   None row   (kind 0, bd:618-641)  [p_none, (1-p_none)/k, ...] with k = n_moves[i], the OBSERVER's
 *                                    offered moves; taken = 0 when row_agent stayed, else 1
 *   single row (kind 1)              Q of row_agent's five actions
 *   joint row  (kind 2, bd:677-679)  the observer is row_agent or row_agent2: only joint actions
 *                                    that match the partner's executed move (five entries)
 *   joint row  (kind 3)              the observer is neither (three or more agents): all 25 joint actions
 *                                    a = 5 * a_i + a_j, no partner filter; needs A = 25
 * Valid = offered (Q not NaN) or taken; Q is capped at q_cap (+inf = goal out of reach); the valid
 * entries are compacted to the front in action order as  Q(taken) - Q(a).
 *   q_table  device float[..][n_pairs][25]  gc_subtask_q / gc_joint_q rows (any number of states)
 *   q_row    device int64[n]                row of q_table that holds env i's obs_tm1
 *   row_pair/row_kind/row_agent/row_agent2  HOST arrays [P] (pair index into n_pairs; agents 0-based)
 *   executed device uint8[n][n_agents]      env.agent_actions of the step that left obs_tm1
 *   n_moves  device uint8[n]
 * Outputs: qdiff [n][P][A] (A = 5, or 25 when the table has kind-3 rows), n_valid [n][P], act_idx [n][P] as
 * gc_bd_posterior_* reads them. */
int gc_bd_likelihood_rows_f32(const float* q_table, const int64_t* q_row, int n_pairs, const int32_t* row_pair,
                              const uint8_t* row_kind, const uint8_t* row_agent, const uint8_t* row_agent2,
                              const uint8_t* executed, const uint8_t* n_moves, int observer, float none_action_prob,
                              float q_cap, float* qdiff, uint8_t* n_valid, uint8_t* act_idx, int64_t n, int P,
                              int n_agents, int A, void* stream);
int gc_bd_likelihood_rows_f64(const float* q_table, const int64_t* q_row, int n_pairs, const int32_t* row_pair,
                              const uint8_t* row_kind, const uint8_t* row_agent, const uint8_t* row_agent2,
                              const uint8_t* executed, const uint8_t* n_moves, int observer, double none_action_prob,
                              double q_cap, double* qdiff, uint8_t* n_valid, uint8_t* act_idx, int64_t n, int P,
                              int n_agents, int A, void* stream);

/* bayes_update (bd:1026-1072) for envs that keep their hypotheses as a LIST of rows of a shared table - the form
 * three and four agents need, where the table of a level has up to 4*10^4 rows (add_subtasks, bd:792-886) and an env
 * only the few that survive pruning (bd:200-256).  Fuses gc_bd_likelihood_rows and gc_bd_posterior: the likelihood
 * values are built from the planner's Q rows inside the kernel (one warp per env), no [n][P][A] array exists.
 *   probs    device [n][W]            in place; entries that are not alive are set to 0
 *   alive    device uint8 [n][W]
 *   rid      device int64 [n][W]      row of the hypothesis table each list entry stands for (outside 0..H-1 = dead)
 *   hyp_pair device uint8 [H][n_entries]  likelihood-row index of each entry of each table row, 0xFF = unused
 *   pair_w   HOST uint8 [P]           weight of a likelihood row (bd:1066)
 * Row tables, executed, n_moves, q_table, q_row: as gc_bd_likelihood_rows_*.  total == 0 -> uniform over the alive. */
int gc_bd_update_lists_f32(float* probs, const uint8_t* alive, const int64_t* rid, int W, const uint8_t* hyp_pair, int H,
                           int n_entries, const uint8_t* pair_w, const float* q_table, const int64_t* q_row, int n_pairs,
                           const int32_t* row_pair, const uint8_t* row_kind, const uint8_t* row_agent,
                           const uint8_t* row_agent2, const uint8_t* executed, const uint8_t* n_moves, int observer,
                           float none_action_prob, float q_cap, float beta, int64_t n, int P, int n_agents, void* stream);
int gc_bd_update_lists_f64(double* probs, const uint8_t* alive, const int64_t* rid, int W, const uint8_t* hyp_pair, int H,
                           int n_entries, const uint8_t* pair_w, const float* q_table, const int64_t* q_row, int n_pairs,
                           const int32_t* row_pair, const uint8_t* row_kind, const uint8_t* row_agent,
                           const uint8_t* row_agent2, const uint8_t* executed, const uint8_t* n_moves, int observer,
                           double none_action_prob, double q_cap, double beta, int64_t n, int P, int n_agents,
                           void* stream);

/* What a RealAgent reads off the real env each step (utils/agent.py), per env and agent:
 *   gc_offered_actions     nav_utils.get_single_actions (navigation_planner/utils.py:55-90): bit a of
 *                          offered[env][agent] = move a (0..3) is offered (staying always is): the square faced holds
 *                          no agent and is floor / a delivery square, or a counter the agent can put its object on,
 *                          pick an object from, or merge with (core.mergeable)
 *   gc_subtasks_completed  RealAgent.def_subtask_completion (utils/agent.py:286-368): completed[env][agent] = the env
 *                          holds more objects equal to the goal of subtask[env][agent] after the step than before
 *                          (Deliver: lying on a delivery square); subtask >= the level's subtask count = none -> 0
 *   state/before/after device uint32[n][4]; offered, subtask, completed device uint8[n][n_agents] */
int gc_offered_actions(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/, const uint32_t* state,
                       uint8_t* offered, int64_t n, int n_agents, void* stream);
int gc_subtasks_completed(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
                          const uint32_t* before, const uint32_t* after, const uint8_t* subtask, uint8_t* completed,
                          int64_t n, int n_agents, void* stream);

/* ---- (B) navigation planner ------------------------------------------------------------
 * Distance lower bound of env.get_lower_bound_for_subtask_given_objs (env:594-664) =
 * World.get_lower_bound_between (utils/world.py:115-264) + holding penalty, for every
 * (env, pair): pair = (subtask index, agent i, agent j or 0xFF).
 *   pairs  HOST uint8[n_pairs][3]   lb  device float[n][n_pairs]  (29.0 = perimeter+1 = not doable)
 * For gc_subtask_q / gc_joint_q bit 7 of the subtask index selects the planning world: clear =
 * level 0 (other agents become Agent-Counters, their held object is deleted, e2e:386-406), set =
 * level 1 (everybody stays: the other agents are obstacles that can be neither entered nor used
 * as counters, e2e:379-381 with navigation_planner/utils.py:62-71). */
int gc_lower_bound(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
                   const uint32_t* state /*device*/, const uint8_t* pairs /*host*/, int n_pairs,
                   float* lb /*device*/, int64_t n, int n_agents, void* stream);

/* Exact level-0 subtask values: V*(s) of the deterministic shortest-path MDP the reference's
 * BRTDP brackets (e2e_brtdp.py:216-352), cost 1 + 0.1*#moving agents (:816-826), transitions =
 * interact, action set = get_single_actions (+ is_collision filter when joint, :151-206),
 * goal = "count of goal objects increased" (:435-566), other agents frozen (:360-406).
 *   v      device float[n][n_pairs]       V*(start); +inf when the goal is unreachable
 *   q      device float[n][n_pairs][25]   nullable; Q(start, a) = cost + V*(T(start,a)) for the
 *          joint action a = 5*a_i + a_j (single agent: a in 0..4); NaN for actions that are not
 *          offered, +inf for offered actions from which the goal is out of reach
 *   status device uint8[n][n_pairs]       nullable; 0 ok, 1 goal already satisfied at start,
 *          2 unreachable, 3 search budget exceeded */
int gc_subtask_q(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
                 const uint32_t* state /*device*/, const uint8_t* pairs /*host*/, int n_pairs,
                 float* v, float* q, uint8_t* status, int64_t n, int n_agents, void* stream);

/* Joint (two-agent) pairs of the same MDP: budgeted exact search over full planning states.  One
 * CTA per (env, pair) runs a forward A* pass that records every generated edge and a backward
 * pass over those edges, which proves Q(start, a) for all offered joint actions whose value fits
 * inside the explored region (widened up to V* + 4.8); what stays open is searched per action.
 * The visited sets live in a caller-provided scratch arena of gc_joint_q_scratch_bytes(n, n_pairs,
 * NULL) bytes.  Only pairs with agent j != 0xFF are written: v[n][n_pairs], q[n][n_pairs][25]
 * (joint action 5*a_i + a_j; NaN = not offered, +inf = offered but the goal is out of reach or the
 * value unproven), status (0 ok, 2 unreachable, 3 a search outgrew its budget of 96K states - the
 * proven Q values are kept, the others stay +inf, 4 unsupported: more than four objects).
 * Single-agent entries of v/q/status are left untouched, so both solvers fill the same arrays. */
int64_t gc_joint_q_scratch_bytes(int64_t n, int n_pairs, int* n_ctas_out);
int gc_joint_q(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
               const uint32_t* state /*device*/, const uint8_t* pairs /*host*/, int n_pairs, float* v, float* q,
               uint8_t* status, void* scratch /*device*/, int64_t scratch_bytes, int64_t n, int n_agents,
               void* stream);

/* ---- (A') optional image_obs renderer -------------------------------------------------
 * misc/game/game.py:56-185 geometry (80 px tiles) -> uint8[m][H*80][W*80][3], RGB.
 * `sprites`: device uint8[71][4][80][80][4] RGBA atlas - sprite 0 delivery, 1 cutboard, 2 plate, 3-6 agents
 * (blue, magenta, yellow, green), 7 + code food sprites with code = (mask & 7) | ((mask >> 4) & 7) << 3;
 * every sprite pre-scaled to the four sizes the reference draws at (frame 0: 80 px, 1: 56, 2: 40, 3: 28, in
 * the top-left corner of an 80 x 80 frame). */
int gc_render(const gc_level* levels, int n_levels, const uint8_t* level_id /*device*/,
              const uint32_t* state /*device*/, const uint8_t* sprites /*device, nullable*/,
              uint8_t* img /*device*/, int64_t m, int n_agents, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GYMCOOK_H */
