/*
 * gc_oracle.c - CPU restatement of path A (env transition) and path C (BD posterior).
 * TEST INFRASTRUCTURE, NOT PRODUCT - see gc_oracle.h.  Citations are file:line under
 * /root/reference/gym_cooking/.
 */
#include "gc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define M_T 0x01
#define M_L 0x02
#define M_O 0x04
#define M_P 0x08
#define M_FOODS 0x07

/* World.NAV_ACTIONS + stay (utils/world.py:16; navigation_planner/utils.py:65,88) */
static const int DX[5] = {0, 0, -1, 1, 0};
static const int DY[5] = {1, -1, 0, 0, 0};

/* ------------------------------------------------------------------------------------ */
/* object predicates (utils/core.py)                                                      */

static int n_contents(int mask) { return __builtin_popcount(mask & 0x0f); }

/* every Food in `mask` is in its last state (Food.done, core.py:293-296) */
static int foods_done(int mask) { return ((mask & M_FOODS) & ~(mask >> 4)) == 0; }

/* Object.needs_chopped core.py:176-178 + Food.needs_chopped :285-291 (Plate: False :365) */
static int needs_chopped(int mask) {
  if (n_contents(mask) != 1) return 0;
  if (mask & M_P) return 0;
  return !foods_done(mask);
}

/* Object.is_deliverable core.py:214-219: merged, and every content a Plate or a done Food */
static int is_deliverable(int mask) { return n_contents(mask) > 1 && foods_done(mask); }

/* mergeable core.py:222-241: at most one plate in total, every food done */
static int mergeable(int m1, int m2) {
  if ((m1 & M_P) && (m2 & M_P)) return 0;
  return foods_done(m1) && foods_done(m2);
}

/* Object.chop core.py:187-192 */
static int chop(int mask) { return mask | ((mask & M_FOODS) << 4); }

/* ------------------------------------------------------------------------------------ */
/* level loader: env.load_level :130-198                                                  */

static int recipe_goal(const char* name) {
  /* recipe_planner/recipe.py:199-228 + Recipe.add_goal :26-46: Delivered(full_plate_name),
   * goal object = every ingredient chopped, on a plate (nav_utils.get_subtask_obj :231-238) */
  if (!strcmp(name, "SimpleTomato")) return M_P | M_T | (M_T << 4);
  if (!strcmp(name, "SimpleLettuce")) return M_P | M_L | (M_L << 4);
  if (!strcmp(name, "Salad")) return M_P | M_T | M_L | ((M_T | M_L) << 4);
  if (!strcmp(name, "OnionSalad")) return M_P | M_FOODS | (M_FOODS << 4);
  return -1;
}

int gco_level_parse(const char* txt, int max_timesteps, gco_level* lv) {
  memset(lv, 0, sizeof(*lv));
  lv->delivery_x = lv->delivery_y = -1;
  lv->max_timesteps = max_timesteps;
  for (int y = 0; y < 8; y++)
    for (int x = 0; x < 8; x++) lv->type[y][x] = GCO_COUNTER;
  int phase = 1, y = 0, last_w = 0;
  const char* p = txt;
  while (*p) {
    const char* e = strchr(p, '\n');
    int len = e ? (int)(e - p) : (int)strlen(p);
    char line[64];
    if (len >= (int)sizeof(line)) return -2;
    memcpy(line, p, len);
    line[len] = 0;
    if (len == 0) {
      phase++; /* env:151-152 */
    } else if (phase == 1) {
      if (y >= 8 || len > 8) return -2;
      for (int x = 0; x < len; x++) { /* env:156-173 */
        char c = line[x];
        int m = c == 't' ? M_T : c == 'l' ? M_L : c == 'o' ? M_O : c == 'p' ? M_P : 0;
        if (m) {
          if (lv->n_objs >= GCO_MAX_OBJS) return -4;
          lv->type[y][x] = GCO_COUNTER;
          lv->obj_mask[lv->n_objs] = m;
          lv->obj_x[lv->n_objs] = x;
          lv->obj_y[lv->n_objs] = y;
          lv->n_objs++;
        } else if (c == '-') {
          lv->type[y][x] = GCO_COUNTER;
        } else if (c == '/') {
          lv->type[y][x] = GCO_CUTBOARD;
        } else if (c == '*') {
          lv->type[y][x] = GCO_DELIVERY;
          if (lv->delivery_x < 0) { /* done() uses the first Delivery, env:349 */
            lv->delivery_x = x;
            lv->delivery_y = y;
          }
        } else {
          lv->type[y][x] = GCO_FLOOR; /* ' ' and anything unknown, env:170-173 */
        }
      }
      last_w = len;
      y++;
    } else if (phase == 2) { /* env:178-182 */
      int g = recipe_goal(line);
      if (g < 0) return -2;
      if (lv->n_goals >= GCO_MAX_GOALS) return -4;
      lv->goal_mask[lv->n_goals++] = g;
    } else if (phase == 3) { /* env:186-193 */
      if (lv->n_agent_starts < GCO_MAX_AGENTS) {
        int ax, ay;
        char* sp = strchr(line, ' ');
        if (!sp) return -2;
        ax = atoi(line);
        ay = atoi(sp + 1);
        lv->agent_x[lv->n_agent_starts] = ax;
        lv->agent_y[lv->n_agent_starts] = ay;
        lv->n_agent_starts++;
      }
    }
    if (!e) break;
    p = e + 1;
  }
  lv->width = last_w; /* env:196 (x of the last map row + 1) */
  lv->height = y;     /* env:197 */
  if (lv->width < 1 || lv->height < 1 || lv->n_goals < 1 || lv->delivery_x < 0) return -2;
  return 0;
}

/* env.reset :201-250 */
void gco_reset(const gco_level* lv, int n_agents, gco_env* e) {
  memset(e, 0, sizeof(*e));
  e->n_agents = n_agents;
  e->n_objs = lv->n_objs;
  for (int i = 0; i < n_agents; i++) {
    e->ag[i].x = lv->agent_x[i];
    e->ag[i].y = lv->agent_y[i];
    e->ag[i].hold = -1;
  }
  for (int k = 0; k < lv->n_objs; k++) {
    e->ob[k].alive = 1;
    e->ob[k].mask = lv->obj_mask[k];
    e->ob[k].x = lv->obj_x[k];
    e->ob[k].y = lv->obj_y[k];
    e->ob[k].held_by = -1;
  }
}

/* ------------------------------------------------------------------------------------ */
/* world queries (utils/world.py)                                                         */

static int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* World.is_occupied :285-290 / get_object_at(find_held_objects=False) :389-419:
 * index of an un-held object lying on (x,y), or -1 */
static int object_on(const gco_env* e, int x, int y) {
  for (int k = 0; k < e->n_objs; k++)
    if (e->ob[k].alive && e->ob[k].held_by < 0 && e->ob[k].x == x && e->ob[k].y == y) return k;
  return -1;
}

/* utils/interact.py:4-89 (arglist.play == False) */
static void interact(const gco_level* lv, gco_env* e, int i, int action) {
  gco_agent* a = &e->ag[i];
  if (action == 4) return; /* :19-20 */
  int nx = clampi(a->x + DX[action], 0, lv->width - 1); /* world.inbounds :432-436 */
  int ny = clampi(a->y + DY[action], 0, lv->height - 1);
  int ty = lv->type[ny][nx];
  if (ty == GCO_FLOOR) { /* :29-30, SimAgent.move_to agent.py:420-423 */
    a->x = nx;
    a->y = ny;
    return;
  }
  if (a->hold >= 0) { /* :33 */
    gco_obj* h = &e->ob[a->hold];
    if (ty == GCO_DELIVERY) { /* :35-40 */
      if (is_deliverable(h->mask)) {
        h->x = nx;
        h->y = ny;
        h->held_by = -1;
        a->hold = -1;
      }
      return;
    }
    int o = object_on(e, nx, ny);
    if (o >= 0) { /* :43-52: the held object absorbs the counter object */
      if (mergeable(h->mask, e->ob[o].mask)) {
        h->mask |= e->ob[o].mask;
        e->ob[o].alive = 0;
      }
      return;
    }
    if (ty == GCO_CUTBOARD && needs_chopped(h->mask)) { /* :63-65 chop in hand */
      h->mask = chop(h->mask);
    } else { /* :66-70 put down */
      h->x = nx;
      h->y = ny;
      h->held_by = -1;
      a->hold = -1;
    }
    return;
  }
  /* :73-89 empty-handed */
  int o = object_on(e, nx, ny);
  if (o >= 0 && ty != GCO_DELIVERY) { /* :77-84 pick up */
    e->ob[o].held_by = i;
    a->hold = o;
  }
}

/* env.is_collision :671-718 */
static void is_collision(const gco_level* lv, int x1, int y1, int x2, int y2, int a1, int a2,
                         int* ex1, int* ex2) {
  *ex1 = *ex2 = 1;
  int nx1 = x1 + DX[a1], ny1 = y1 + DY[a1];
  /* get_gridsquare_at(next).collidable: everything but Floor (core.py:34,64).  Squares
   * outside the map make the reference assert (world.py:429); the supported envelope has a
   * non-floor outer ring so that never happens - treat as collidable. */
  if (nx1 < 0 || ny1 < 0 || nx1 >= lv->width || ny1 >= lv->height || lv->type[ny1][nx1] != GCO_FLOOR) {
    nx1 = x1;
    ny1 = y1;
  }
  int nx2 = x2 + DX[a2], ny2 = y2 + DY[a2];
  if (nx2 < 0 || ny2 < 0 || nx2 >= lv->width || ny2 >= lv->height || lv->type[ny2][nx2] != GCO_FLOOR) {
    nx2 = x2;
    ny2 = y2;
  }
  if (nx1 == nx2 && ny1 == ny2) { /* :704-711 */
    if (nx1 == x1 && ny1 == y1 && a1 != 4)
      *ex2 = 0;
    else if (nx2 == x2 && ny2 == y2 && a2 != 4)
      *ex1 = 0;
    else
      *ex1 = *ex2 = 0;
  } else if (x1 == nx2 && y1 == ny2 && x2 == nx1 && y2 == ny1) { /* :714-717 swap */
    *ex1 = *ex2 = 0;
  }
}

/* env.done :316-363 */
static void check_done(const gco_level* lv, gco_env* e) {
  if (lv->max_timesteps && e->t >= lv->max_timesteps) { /* :328-332 timeout first */
    e->done = 1;
    e->successful = 0;
    return;
  }
  for (int g = 0; g < lv->n_goals; g++) { /* :344-359 */
    int found = 0;
    for (int k = 0; k < e->n_objs; k++) {
      const gco_obj* o = &e->ob[k];
      if (!o->alive || o->mask != lv->goal_mask[g]) continue;
      int ox = o->held_by >= 0 ? e->ag[o->held_by].x : o->x;
      int oy = o->held_by >= 0 ? e->ag[o->held_by].y : o->y;
      if (ox == lv->delivery_x && oy == lv->delivery_y) found = 1;
    }
    if (!found) {
      e->done = 0;
      e->successful = 0;
      return;
    }
  }
  e->done = 1;
  e->successful = 1;
}

/* env.step :255-306 */
int gco_step(const gco_level* lv, gco_env* e, const uint8_t* actions, uint8_t* executed) {
  int act[GCO_MAX_AGENTS], exec_[GCO_MAX_AGENTS];
  if (e->done) { /* batched convention: finished episodes are frozen (main.py:97 stops there) */
    if (executed)
      for (int i = 0; i < e->n_agents; i++) executed[i] = 4;
    return 0;
  }
  if (e->t < 127) e->t += 1; /* :257; 7-bit field of the packed form */
  for (int i = 0; i < e->n_agents; i++) {
    act[i] = actions[i] > 4 ? 4 : actions[i];
    exec_[i] = 1;
  }
  /* check_collisions :724-762 - every pair judged on the ORIGINAL actions */
  int ncoll = 0;
  for (int i = 0; i < e->n_agents; i++)
    for (int j = i + 1; j < e->n_agents; j++) {
      int e1, e2;
      is_collision(lv, e->ag[i].x, e->ag[i].y, e->ag[j].x, e->ag[j].y, act[i], act[j], &e1, &e2);
      if (!e1) exec_[i] = 0;
      if (!e2) exec_[j] = 0;
      if (!(e1 && e2)) ncoll++;
    }
  for (int i = 0; i < e->n_agents; i++)
    if (!exec_[i]) act[i] = 4; /* :757-761 */
  /* execute_navigation :767-770 - sequential, agent order */
  for (int i = 0; i < e->n_agents; i++) {
    interact(lv, e, i, act[i]);
    if (executed) executed[i] = (uint8_t)act[i];
  }
  check_done(lv, e); /* :295-298 */
  return ncoll;
}

/* ------------------------------------------------------------------------------------ */
/* packed form (include/gymcook.h)                                                        */

/* place byte / mask byte of object k (include/gymcook.h: word 1 = places of objects 0..3, word 2 =
 * their masks, word 3 = place 4, place 5, mask 4, mask 5) */
static int place_byte(const uint32_t w[4], int k) {
  return (int)((k < 4 ? w[1] >> (8 * k) : w[3] >> (8 * (k - 4))) & 0xff);
}
static int mask_byte(const uint32_t w[4], int k) {
  return (int)((k < 4 ? w[2] >> (8 * k) : w[3] >> (8 * (k - 2))) & 0xff);
}
enum { PLACE_HELD = 0x40, PLACE_DEAD = 0x47 };

void gco_pack(const gco_env* e, uint32_t w[4]) {
  w[0] = w[1] = w[2] = w[3] = 0;
  for (int i = 0; i < e->n_agents; i++) w[0] |= (uint32_t)(e->ag[i].y * 8 + e->ag[i].x) << (6 * i);
  w[0] |= (uint32_t)(e->t & 127) << 24;
  w[0] |= (uint32_t)(e->done ? 1u : 0u) << 31;
  for (int k = 0; k < GCO_MAX_OBJS; k++) {
    uint32_t place = PLACE_DEAD, mask = 0;
    if (k < e->n_objs && e->ob[k].alive) {
      const gco_obj* o = &e->ob[k];
      mask = (uint32_t)o->mask;
      place = o->held_by >= 0 ? (uint32_t)(PLACE_HELD + o->held_by + 1) : (uint32_t)(o->y * 8 + o->x);
    }
    if (k < 4) {
      w[1] |= place << (8 * k);
      w[2] |= mask << (8 * k);
    } else {
      w[3] |= (place << (8 * (k - 4))) | (mask << (8 * (k - 2)));
    }
  }
}

void gco_unpack(const uint32_t w[4], int n_agents, gco_env* e) {
  memset(e, 0, sizeof(*e));
  e->n_agents = n_agents;
  e->n_objs = GCO_MAX_OBJS;
  e->t = (w[0] >> 24) & 127;
  e->done = (w[0] >> 31) & 1;
  for (int i = 0; i < n_agents; i++) {
    int c = (w[0] >> (6 * i)) & 63;
    e->ag[i].x = c & 7;
    e->ag[i].y = c >> 3;
    e->ag[i].hold = -1;
  }
  for (int k = 0; k < GCO_MAX_OBJS; k++) {
    const int place = place_byte(w, k);
    gco_obj* o = &e->ob[k];
    if (place == PLACE_DEAD) {
      o->alive = 0;
      o->held_by = -1;
      continue;
    }
    const int holder = place >= PLACE_HELD ? place - PLACE_HELD : 0;
    o->alive = 1;
    o->mask = mask_byte(w, k);
    o->x = holder ? 0 : (place & 7);
    o->y = holder ? 0 : (place >> 3);
    o->held_by = holder - 1;
    if (holder >= 1 && holder <= n_agents) e->ag[holder - 1].hold = k;
  }
}

int gco_canonical_keys(const uint32_t w[4], uint16_t keys[GCO_MAX_OBJS]) {
  int n = 0;
  for (int k = 0; k < GCO_MAX_OBJS; k++) {
    const int place = place_byte(w, k);
    if (place == PLACE_DEAD) continue;
    int cell = place, held = 0;
    if (place >= PLACE_HELD) {
      cell = (w[0] >> (6 * (place - PLACE_HELD - 1))) & 63;
      held = 1;
    }
    keys[n++] = (uint16_t)((mask_byte(w, k) << 7) | (cell << 1) | held);
  }
  for (int i = 1; i < n; i++) { /* insertion sort */
    uint16_t k = keys[i];
    int j = i - 1;
    while (j >= 0 && keys[j] > k) {
      keys[j + 1] = keys[j];
      j--;
    }
    keys[j + 1] = k;
  }
  for (int i = n; i < GCO_MAX_OBJS; i++) keys[i] = 0x3FFF;
  return n;
}

static uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
  z ^= z >> 30;
  z *= 0xbf58476d1ce4e5b9ull;
  z ^= z >> 27;
  z *= 0x94d049bb133111ebull;
  z ^= z >> 31;
  return z;
}

uint64_t gco_hash_packed(const uint32_t w[4], int n_agents) {
  uint16_t keys[GCO_MAX_OBJS];
  gco_canonical_keys(w, keys);
  /* W0: per agent 13 bits (cell | hold_mask<<6) at 13*i, t at bit 52 */
  uint64_t W0 = (uint64_t)((w[0] >> 24) & 127) << 52;
  for (int i = 0; i < n_agents; i++) {
    uint64_t cell = (w[0] >> (6 * i)) & 63, hm = 0;
    for (int k = 0; k < GCO_MAX_OBJS; k++)
      if (place_byte(w, k) == PLACE_HELD + i + 1) hm = (uint64_t)mask_byte(w, k);
    W0 |= (cell | (hm << 6)) << (13 * i);
  }
  uint64_t W1 = (uint64_t)keys[0] | ((uint64_t)keys[1] << 14) | ((uint64_t)keys[2] << 28);
  uint64_t W2 = (uint64_t)keys[3] | ((uint64_t)keys[4] << 14) | ((uint64_t)keys[5] << 28);
  uint64_t h = mix64(W0 + 0x9E3779B97F4A7C15ull);
  h = mix64(h ^ W1);
  h = mix64(h ^ W2);
  return h;
}

/* ------------------------------------------------------------------------------------ */
/* philox4x32-10 (Salmon et al. 2011), counter (t, env_lo, env_hi, 0), key = seed          */

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

void gco_philox_actions(uint64_t seed, uint32_t t, uint64_t env, uint8_t out[4]) {
  uint32_t c[4] = {t, (uint32_t)env, (uint32_t)(env >> 32), 0};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  for (int i = 0; i < 4; i++) out[i] = (uint8_t)(((uint64_t)c[i] * 5u) >> 32);
}

/* the same stream materialised: actions[n_steps][n][n_agents] (the CPU twin of gc_fill_random_actions) */
void gco_fill_actions(uint8_t* actions, int64_t n, int n_agents, int n_steps, int t0, int64_t env0, uint64_t seed) {
  for (int k = 0; k < n_steps; k++)
    for (int64_t i = 0; i < n; i++) {
      uint8_t a[4];
      gco_philox_actions(seed, (uint32_t)(t0 + k), (uint64_t)(env0 + i), a);
      for (int j = 0; j < n_agents; j++) actions[((int64_t)k * n + i) * n_agents + j] = a[j];
    }
}

/* ------------------------------------------------------------------------------------ */
/* batched drivers (pthread parallel-for over contiguous env ranges)                     */

typedef struct {
  const gco_level* lv;
  uint32_t* state;
  const uint8_t* actions; /* NULL -> philox rollout */
  uint8_t* reward_done;
  uint64_t* hash_trace;
  uint32_t* collisions;
  int64_t n, lo, hi;
  int n_agents, n_steps, t0;
  int64_t env0;
  uint64_t seed;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  const gco_level* lv = j->lv;
  for (int64_t i = j->lo; i < j->hi; i++) {
    gco_env e;
    gco_unpack(j->state + 4 * i, j->n_agents, &e);
    uint32_t nc = 0;
    for (int s = 0; s < j->n_steps; s++) {
      uint8_t a[4];
      const uint8_t* act = a;
      if (j->actions)
        act = j->actions + i * j->n_agents;
      else
        gco_philox_actions(j->seed, (uint32_t)(j->t0 + s), (uint64_t)(j->env0 + i), a);
      nc += (uint32_t)gco_step(lv, &e, act, 0);
      if (j->hash_trace) {
        uint32_t w[4];
        gco_pack(&e, w);
        j->hash_trace[(int64_t)s * j->n + i] = gco_hash_packed(w, j->n_agents);
      }
    }
    /* outcome of a finished episode, recomputed from the stored state (frozen envs too) */
    e.successful = e.done && !(lv->max_timesteps && e.t >= lv->max_timesteps);
    gco_pack(&e, j->state + 4 * i);
    if (j->reward_done) j->reward_done[i] = (uint8_t)((e.done ? 1 : 0) | (e.successful ? 2 : 0));
    if (j->collisions) j->collisions[i] += nc;
  }
  return 0;
}

static void run_batch(batch_job proto, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  if (n_threads == 1 || proto.n < 2 * n_threads) {
    proto.lo = 0;
    proto.hi = proto.n;
    batch_worker(&proto);
    return;
  }
  pthread_t th[256];
  batch_job jobs[256];
  for (int k = 0; k < n_threads; k++) {
    jobs[k] = proto;
    jobs[k].lo = proto.n * k / n_threads;
    jobs[k].hi = proto.n * (k + 1) / n_threads;
    pthread_create(&th[k], 0, batch_worker, &jobs[k]);
  }
  for (int k = 0; k < n_threads; k++) pthread_join(th[k], 0);
}

void gco_step_batch(const gco_level* lv, uint32_t* state, const uint8_t* actions,
                    uint8_t* reward_done, uint32_t* collisions, int64_t n, int n_agents,
                    int n_threads) {
  batch_job j = {lv, state, actions, reward_done, 0, collisions, n, 0, 0, n_agents, 1, 0, 0, 0};
  run_batch(j, n_threads);
}

void gco_rollout_batch(const gco_level* lv, uint32_t* state, uint8_t* reward_done,
                       uint64_t* hash_trace, uint32_t* collisions, int64_t n, int n_agents,
                       int n_steps, int t0, int64_t env0, uint64_t seed, int n_threads) {
  batch_job j = {lv, state, 0, reward_done, hash_trace, collisions, n, 0, 0, n_agents, n_steps, t0, env0, seed};
  run_batch(j, n_threads);
}

/* Replay one episode: states[0] = reset, states[s+1] = after actions[s] (stride 4 per step).
 * Used by the golden-trace tests so that the whole comparison runs in C + numpy. */
void gco_replay(const gco_level* lv, int n_agents, const uint8_t* actions, int n_steps,
                uint32_t* states, uint8_t* reward_done, uint8_t* ncoll, uint8_t* executed) {
  gco_env e;
  gco_reset(lv, n_agents, &e);
  gco_pack(&e, states);
  reward_done[0] = 0;
  ncoll[0] = 0;
  for (int i = 0; i < 4; i++) executed[i] = 4;
  for (int s = 0; s < n_steps; s++) {
    uint8_t ex[4] = {4, 4, 4, 4};
    int nc = gco_step(lv, &e, actions + 4 * s, ex);
    gco_pack(&e, states + 4 * (s + 1));
    reward_done[s + 1] = (uint8_t)((e.done ? 1 : 0) | (e.successful ? 2 : 0));
    ncoll[s + 1] = (uint8_t)nc;
    for (int i = 0; i < 4; i++) executed[4 * (s + 1) + i] = ex[i];
  }
}

/* ------------------------------------------------------------------------------------ */
/* path C: BayesianDelegator.bayes_update :1045-1072                                      */

void gco_bd_posterior(double* probs, const uint8_t* alive, const uint8_t* hyp_pair,
                      const uint8_t* pair_w, const double* qdiff, const uint8_t* n_valid,
                      const uint8_t* act_idx, double beta, int64_t n, int H, int P, int A,
                      int n_entries) {
  for (int64_t r = 0; r < n; r++) {
    double L[256];
    for (int p = 0; p < P; p++) {
      /* prob_nav_actions :682-689 (and :626-641 for None): scipy.special.softmax is
       * exp(x - max) / sum (scipy/special/_logsumexp.py) */
      const double* qd = qdiff + ((int64_t)r * P + p) * A;
      int nv = n_valid[r * P + p];
      if (nv == 0) {
        L[p] = 0.0;
        continue;
      }
      double mx = -INFINITY, sum = 0.0;
      for (int a = 0; a < nv; a++)
        if (beta * qd[a] > mx) mx = beta * qd[a];
      for (int a = 0; a < nv; a++) sum += exp(beta * qd[a] - mx);
      L[p] = exp(beta * qd[act_idx[r * P + p]] - mx) / sum;
    }
    double total = 0.0;
    int n_alive = 0;
    for (int h = 0; h < H; h++) {
      if (alive && !alive[r * H + h]) {
        probs[r * H + h] = 0.0;
        continue;
      }
      double update = 0.0; /* :1046-1066 */
      for (int e = 0; e < n_entries; e++) {
        int p = hyp_pair[((int64_t)r * H + h) * n_entries + e];
        if (p == 0xFF) continue;
        update += pair_w[r * P + p] * L[p];
      }
      probs[r * H + h] *= update; /* dutils.update :177-178 */
      total += probs[r * H + h];
      n_alive++;
    }
    for (int h = 0; h < H; h++) { /* dutils.normalize :186-193 */
      if (alive && !alive[r * H + h]) continue;
      if (total == 0.0)
        probs[r * H + h] = 1.0 / n_alive;
      else
        probs[r * H + h] *= 1.0 / total;
    }
  }
}
