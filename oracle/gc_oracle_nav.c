/*
 * gc_oracle_nav.c - CPU restatement of path B: the distance heuristic of the navigation
 * planner and the exact level-0 value of a (subtask, agent set) MDP.
 * TEST INFRASTRUCTURE, NOT PRODUCT - see gc_oracle.h.  Citations: file:line under
 * /root/reference/gym_cooking/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gc_oracle.h"

static const int DX[5] = {0, 0, -1, 1, 0};
static const int DY[5] = {1, -1, 0, 0, 0};
#define NOPATH 1000000
#define GCO_BLOCKED 4 /* level-1 planning: square of another agent */

/* ------------------------------------------------------------------------------------ */
/* World.reachability_graph (utils/world.py:67-107), queried like nx.shortest_path_length  */
/* Node = (square, approach): approach 4 = (0,0) for floor squares, 0..3 = NAV_ACTIONS for  */
/* collidable squares that have a floor square in that direction.                          */

typedef struct {
  int x, y, na;
} node_t;

static int is_floor(const gco_level* lv, int x, int y) {
  return x >= 0 && y >= 0 && x < lv->width && y < lv->height && lv->type[y][x] == GCO_FLOOR;
}

static int node_exists(const gco_level* lv, node_t n) {
  if (n.x < 0 || n.y < 0 || n.x >= lv->width || n.y >= lv->height) return 0;
  if (n.na == 4) return lv->type[n.y][n.x] == GCO_FLOOR;
  return lv->type[n.y][n.x] != GCO_FLOOR && is_floor(lv, n.x + DX[n.na], n.y + DY[n.na]);
}

/* BFS hop count over floor squares; NOPATH if disconnected */
static int floor_dist(const gco_level* lv, int sx, int sy, int tx, int ty) {
  int dist[8][8], qx[64], qy[64], qh = 0, qt = 0;
  memset(dist, -1, sizeof(dist));
  dist[sy][sx] = 0;
  qx[qt] = sx;
  qy[qt++] = sy;
  while (qh < qt) {
    int x = qx[qh], y = qy[qh++];
    if (x == tx && y == ty) return dist[y][x];
    for (int a = 0; a < 4; a++) {
      int nx = x + DX[a], ny = y + DY[a];
      if (is_floor(lv, nx, ny) && dist[ny][nx] < 0) {
        dist[ny][nx] = dist[y][x] + 1;
        qx[qt] = nx;
        qy[qt++] = ny;
      }
    }
  }
  return NOPATH;
}

/* nx.shortest_path_length(graph, s, t); NOPATH where networkx raises (missing node / no path) */
static int graph_dist(const gco_level* lv, node_t s, node_t t) {
  if (!node_exists(lv, s) || !node_exists(lv, t)) return NOPATH;
  if (s.x == t.x && s.y == t.y && s.na == t.na) return 0;
  int extra = 0, sx = s.x, sy = s.y, tx = t.x, ty = t.y;
  if (s.na != 4) { /* a collidable node hangs off exactly one floor square */
    sx += DX[s.na];
    sy += DY[s.na];
    extra++;
  }
  if (t.na != 4) {
    tx += DX[t.na];
    ty += DY[t.na];
    extra++;
  }
  int d = floor_dist(lv, sx, sy, tx, ty);
  return d >= NOPATH ? NOPATH : d + extra;
}

/* World.get_lower_bound_between_helper :148-264 */
static double lb_helper(const gco_level* lv, int kind, int n_ag, const int* agx, const int* agy, int ax, int ay,
                        int bx, int by) {
  const int perimeter = 2 * (lv->width + lv->height);
  double lower = perimeter + 1;
  int a_coll = lv->type[ay][ax] != GCO_FLOOR, b_coll = lv->type[by][bx] != GCO_FLOOR;
  for (int ia = 0; ia < (a_coll ? 4 : 1); ia++)
    for (int ib = 0; ib < (b_coll ? 4 : 1); ib++) {
      node_t A = {ax, ay, a_coll ? ia : 4}, B = {bx, by, b_coll ? ib : 4};
      double bound;
      if (n_ag == 1) { /* :178-189 */
        node_t s = {agx[0], agy[0], 4};
        int b1 = graph_dist(lv, s, A), b2 = graph_dist(lv, A, B);
        if (b1 >= NOPATH || b2 >= NOPATH) continue;
        bound = b1 + b2 - 1;
      } else { /* :193-258 */
        node_t s1 = {agx[0], agy[0], 4}, s2 = {agx[1], agy[1], 4};
        int d;
        double b1A = (d = graph_dist(lv, s1, A)) >= NOPATH ? perimeter : d;
        double b2A = (d = graph_dist(lv, s2, A)) >= NOPATH ? perimeter : d;
        double b1B = (d = graph_dist(lv, s1, B)) >= NOPATH ? perimeter : d;
        double b2B = (d = graph_dist(lv, s2, B)) >= NOPATH ? perimeter : d;
        double minA = b1A < b2A ? b1A : b2A, minB = b1B < b2B ? b1B : b2B;
        double between = fabs((double)(ax - bx)) + fabs((double)(ay - by)); /* manhattan_dist nutils:95-98 */
        if (kind == 1 || kind == 3) {
          bound = minA + between - 1;
        } else { /* Merge: check_bound :266-283 */
          if ((b1A == minA && b1B == minB) || (b2A == minA && b2B == minB)) {
            minA *= 2;
            minB *= 2;
          }
          bound = (minA > minB ? minA : minB) + (between - 1) / 2;
        }
      }
      if (bound < lower) lower = bound;
    }
  return lower < 1 ? 1 : lower; /* :264 */
}

/* locations of objects equal to `mask` that lie un-held, plus the subtask agents holding one
 * (env.get_AB_locs_given_objs :480-589) */
static int obj_locs(const gco_env* e, int mask, const int* ags, int n_ag, int* xs, int* ys) {
  int n = 0;
  for (int k = 0; k < e->n_objs; k++)
    if (e->ob[k].alive && e->ob[k].held_by < 0 && e->ob[k].mask == mask) {
      xs[n] = e->ob[k].x;
      ys[n++] = e->ob[k].y;
    }
  for (int i = 0; i < e->n_agents; i++) { /* sim_agents order */
    int in_set = 0;
    for (int q = 0; q < n_ag; q++) in_set |= ags[q] == i;
    if (in_set && e->ag[i].hold >= 0 && e->ob[e->ag[i].hold].mask == mask) {
      xs[n] = e->ag[i].x;
      ys[n++] = e->ag[i].y;
    }
  }
  return n;
}

static int square_locs(const gco_level* lv, int type, int* xs, int* ys) {
  int n = 0;
  for (int y = 0; y < lv->height; y++)
    for (int x = 0; x < lv->width; x++)
      if (lv->type[y][x] == type) {
        xs[n] = x;
        ys[n++] = y;
      }
  return n;
}

/* env.get_lower_bound_for_subtask_given_objs :594-664 */
double gco_lower_bound(const gco_level* lv, const gco_env* e, const gco_subtask* st, int ai, int aj) {
  int ags[2] = {ai, aj}, n_ag = aj >= 0 ? 2 : 1;
  if (n_ag == 2 && aj < ai) { /* agent_locs follow sim_agents order (:641) */
    ags[0] = aj;
    ags[1] = ai;
  }
  double penalty = 0; /* :612-638 */
  for (int q = 0; q < n_ag; q++) {
    const gco_agent* a = &e->ag[ags[q]];
    if (a->hold >= 0 && st->kind != 2) {
      int hm = e->ob[a->hold].mask;
      int start = st->a, goal = st->goal;
      if (hm != start && hm != goal) penalty += 1.0;
    }
  }
  if (penalty > 1) penalty = 1;
  int agx[2], agy[2];
  for (int q = 0; q < n_ag; q++) {
    agx[q] = e->ag[ags[q]].x;
    agy[q] = e->ag[ags[q]].y;
  }
  int axs[16], ays[16], bxs[16], bys[16], na = 0, nb = 0;
  if (st->kind == 1) { /* Chop :512-527 */
    na = obj_locs(e, st->a, ags, n_ag, axs, ays);
    nb = square_locs(lv, GCO_CUTBOARD, bxs, bys);
  } else if (st->kind == 3) { /* Deliver :535-548 */
    nb = square_locs(lv, GCO_DELIVERY, bxs, bys);
    int n0 = obj_locs(e, st->a, ags, n_ag, axs, ays);
    for (int k = 0; k < n0; k++) {
      int on_b = 0;
      for (int b = 0; b < nb; b++) on_b |= axs[k] == bxs[b] && ays[k] == bys[b];
      if (!on_b) {
        axs[na] = axs[k];
        ays[na++] = ays[k];
      }
    }
  } else if (st->kind == 2) { /* Merge :573-584 */
    na = obj_locs(e, st->a, ags, n_ag, axs, ays);
    nb = obj_locs(e, st->b, ags, n_ag, bxs, bys);
  }
  /* World.get_lower_bound_between :115-144 */
  double lower = 2 * (lv->width + lv->height) + 1;
  for (int a = 0; a < na; a++)
    for (int b = 0; b < nb; b++) {
      double bound = lb_helper(lv, st->kind, n_ag, agx, agy, axs[a], ays[a], bxs[b], bys[b]);
      if (bound < lower) lower = bound;
    }
  return lower + penalty;
}

/* ------------------------------------------------------------------------------------ */
/* exact level-0 value: uniform-cost search over the planner's own state space            */

/* Planning world of E2E_BRTDP._configure_planner_level :360-406 (level 0): agents outside
 * the subtask disappear, the object they hold is deleted, and their floor square becomes an
 * Agent-Counter (collidable; things can be put on it, utils/core.py:79-93). */
typedef struct {
  gco_level lv;
  int goal_kind, goal_mask, base_count;
} plan_t;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static int foods_done(int m) { return ((m & 7) & ~(m >> 4)) == 0; }
static int mergeable(int a, int b) { return !((a & 8) && (b & 8)) && foods_done(a) && foods_done(b); }
static int needs_chopped(int m) { return m == 1 || m == 2 || m == 4; }
static int deliverable(int m) { return __builtin_popcount(m & 15) > 1 && foods_done(m); }

static int lying_on(const gco_env* e, int x, int y) {
  for (int k = 0; k < e->n_objs; k++)
    if (e->ob[k].alive && e->ob[k].held_by < 0 && e->ob[k].x == x && e->ob[k].y == y) return k;
  return -1;
}

/* utils/interact.py:4-89 on the planning world */
static void plan_interact(const gco_level* lv, gco_env* e, int i, int action) {
  gco_agent* a = &e->ag[i];
  if (action == 4) return;
  int nx = clampi(a->x + DX[action], 0, lv->width - 1), ny = clampi(a->y + DY[action], 0, lv->height - 1);
  int ty = lv->type[ny][nx];
  if (ty == GCO_FLOOR) {
    a->x = nx;
    a->y = ny;
    return;
  }
  if (a->hold >= 0) {
    gco_obj* h = &e->ob[a->hold];
    if (ty == GCO_DELIVERY) {
      if (deliverable(h->mask)) {
        h->x = nx;
        h->y = ny;
        h->held_by = -1;
        a->hold = -1;
      }
      return;
    }
    int o = lying_on(e, nx, ny);
    if (o >= 0) {
      if (mergeable(h->mask, e->ob[o].mask)) {
        h->mask |= e->ob[o].mask;
        e->ob[o].alive = 0;
      }
    } else if (ty == GCO_CUTBOARD && needs_chopped(h->mask)) {
      h->mask |= (h->mask & 7) << 4;
    } else {
      h->x = nx;
      h->y = ny;
      h->held_by = -1;
      a->hold = -1;
    }
    return;
  }
  int o = lying_on(e, nx, ny);
  if (o >= 0 && ty != GCO_DELIVERY) {
    e->ob[o].held_by = i;
    a->hold = o;
  }
}

/* nav_utils.get_single_actions navigation_planner/utils.py:55-90; bit a set = action a valid */
static int single_actions(const gco_level* lv, const gco_env* e, int i) {
  int valid = 1 << 4; /* (0,0) always */
  const gco_agent* a = &e->ag[i];
  for (int act = 0; act < 4; act++) {
    int nx = clampi(a->x + DX[act], 0, lv->width - 1), ny = clampi(a->y + DY[act], 0, lv->height - 1);
    int blocked = 0;
    for (int j = 0; j < e->n_agents; j++) blocked |= e->ag[j].x == nx && e->ag[j].y == ny; /* :71 (incl. self) */
    if (blocked) continue;
    int ty = lv->type[ny][nx];
    if (ty == GCO_BLOCKED) continue; /* level 1: another agent stands there (:71) */
    if (ty == GCO_FLOOR || ty == GCO_DELIVERY) {
      valid |= 1 << act;
      continue;
    }
    int o = lying_on(e, nx, ny);
    if (o < 0 && a->hold >= 0) valid |= 1 << act;                                              /* :80-81 */
    else if (o >= 0 && a->hold < 0) valid |= 1 << act;                                         /* :82-83 */
    else if (o >= 0 && a->hold >= 0 && mergeable(e->ob[a->hold].mask, e->ob[o].mask)) valid |= 1 << act; /* :84-86 */
  }
  return valid;
}

/* env.is_collision :671-718 on the planning world; returns 1 when both may execute */
static int joint_ok(const gco_level* lv, const gco_env* e, int a1, int a2) {
  int x1 = e->ag[0].x, y1 = e->ag[0].y, x2 = e->ag[1].x, y2 = e->ag[1].y;
  int nx1 = x1 + DX[a1], ny1 = y1 + DY[a1], nx2 = x2 + DX[a2], ny2 = y2 + DY[a2];
  if (!is_floor(lv, nx1, ny1)) nx1 = x1, ny1 = y1;
  if (!is_floor(lv, nx2, ny2)) nx2 = x2, ny2 = y2;
  if (nx1 == nx2 && ny1 == ny2) return 0; /* every branch of :704-711 cancels at least one */
  if (x1 == nx2 && y1 == ny2 && x2 == nx1 && y2 == ny1) return 0;
  return 1;
}

/* e2e_brtdp._define_goal_state :435-566: number of goal objects (distinct locations, any
 * holder; Deliver: un-held on a Delivery square) */
static int goal_count(const plan_t* p, const gco_env* e) {
  int n = 0, seen_x[8], seen_y[8];
  for (int k = 0; k < e->n_objs; k++) {
    const gco_obj* o = &e->ob[k];
    if (!o->alive || o->mask != p->goal_mask) continue;
    int x = o->held_by >= 0 ? e->ag[o->held_by].x : o->x, y = o->held_by >= 0 ? e->ag[o->held_by].y : o->y;
    if (p->goal_kind == 3) {
      if (o->held_by < 0 && p->lv.type[y][x] == GCO_DELIVERY) n++;
    } else {
      int dup = 0;
      for (int q = 0; q < n; q++) dup |= seen_x[q] == x && seen_y[q] == y;
      if (!dup) {
        seen_x[n] = x;
        seen_y[n++] = y;
      }
    }
  }
  return n;
}

typedef struct {
  uint32_t w[4];
  int cost; /* tenths */
} qitem_t;

typedef struct {
  uint32_t (*keys)[4];
  int* best;
  size_t cap, used;
} table_t;

static size_t hash4(const uint32_t w[4]) {
  uint64_t h = ((uint64_t)w[0] << 32 | w[1]) * 0x9E3779B97F4A7C15ull;
  h ^= ((uint64_t)w[2] << 32 | w[3]) * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  return (size_t)(h * 0xBF58476D1CE4E5B9ull);
}

/* returns pointer to the best-cost slot of state w (inserting it with INT_MAX if new) */
static int* table_slot(table_t* t, const uint32_t w[4]) {
  size_t i = hash4(w) & (t->cap - 1);
  for (;;) {
    if (t->best[i] < 0) {
      memcpy(t->keys[i], w, 16);
      t->best[i] = 0x7fffffff;
      t->used++;
      return &t->best[i];
    }
    if (!memcmp(t->keys[i], w, 16)) return &t->best[i];
    i = (i + 1) & (t->cap - 1);
  }
}

/* planning world of the next gco_subtask_q calls: 0 = level 0 (other agents become Agent-Counters and
 * their held object is deleted, e2e_brtdp.py:386-406), 1 = level 1 (everybody stays; the other agents
 * are obstacles that can be neither entered nor used as counters, :379-381 + nav utils :62-71) */
static int g_planner_level = 0;
void gco_set_planner_level(int level) { g_planner_level = level; }

/* diagnostics: states inserted by the searches of the last gco_subtask_q call */
long long gco_last_search_states = 0;

/* V*(s0) in tenths, or -2 unreachable, -3 budget exceeded.  Bucketed Dijkstra (edge costs
 * 10..12). */
static int solve(const plan_t* p, const gco_env* s0, int n_ag, int max_states) {
  if (goal_count(p, s0) > p->base_count) return 0;
  size_t cap = 1;
  while (cap < (size_t)max_states * 2) cap <<= 1;
  table_t tb = {malloc(cap * 16), malloc(cap * sizeof(int)), cap, 0};
  memset(tb.best, 0xff, cap * sizeof(int));
  enum { NB = 16 };
  qitem_t* bucket[NB];
  size_t bn[NB], bc[NB];
  for (int b = 0; b < NB; b++) bucket[b] = malloc(sizeof(qitem_t) * (bc[b] = 1024)), bn[b] = 0;
  uint32_t w[4];
  gco_pack(s0, w);
  *table_slot(&tb, w) = 0;
  memcpy(bucket[0][0].w, w, 16);
  bucket[0][0].cost = 0;
  bn[0] = 1;
  int result = -2, cur = 0, empty_run = 0;
  while (empty_run < NB) {
    int b = cur % NB;
    if (bn[b] == 0) {
      cur++;
      empty_run++;
      continue;
    }
    empty_run = 0;
    qitem_t it = bucket[b][--bn[b]];
    if (it.cost != cur) { /* belongs to a later lap of the ring: cannot happen with costs <= 12 < NB */
      continue;
    }
    if (*table_slot(&tb, it.w) < it.cost) continue;
    gco_env e;
    gco_unpack(it.w, n_ag, &e);
    e.n_objs = GCO_MAX_OBJS;
    if (goal_count(p, &e) > p->base_count) {
      result = it.cost;
      break;
    }
    if (tb.used > (size_t)max_states) {
      result = -3;
      break;
    }
    int v1 = single_actions(&p->lv, &e, 0), v2 = n_ag == 2 ? single_actions(&p->lv, &e, 1) : (1 << 4);
    for (int a1 = 0; a1 < 5; a1++)
      for (int a2 = 0; a2 < 5; a2++) {
        if (!((v1 >> a1) & 1) || !((v2 >> a2) & 1)) continue;
        if (n_ag == 1 && a2 != 4) continue;
        if (n_ag == 2 && !joint_ok(&p->lv, &e, a1, a2)) continue;
        gco_env n = e;
        plan_interact(&p->lv, &n, 0, a1);
        if (n_ag == 2) plan_interact(&p->lv, &n, 1, a2);
        int c = it.cost + 10 + (a1 != 4) + (n_ag == 2 && a2 != 4); /* e2e_brtdp.cost :816-826 */
        uint32_t nw[4];
        gco_pack(&n, nw);
        int* best = table_slot(&tb, nw);
        if (c < *best) {
          *best = c;
          int nb = c % NB;
          if (bn[nb] == bc[nb]) bucket[nb] = realloc(bucket[nb], sizeof(qitem_t) * (bc[nb] *= 2));
          memcpy(bucket[nb][bn[nb]].w, nw, 16);
          bucket[nb][bn[nb]++].cost = c;
        }
      }
  }
  gco_last_search_states += (long long)tb.used;
  for (int b = 0; b < NB; b++) free(bucket[b]);
  free(tb.keys);
  free(tb.best);
  return result;
}

int gco_subtask_q(const gco_level* lv, const gco_env* e0, const gco_subtask* st, int ai, int aj, double* v,
                  double* q, int max_states) {
  plan_t p;
  p.lv = *lv;
  p.goal_kind = st->kind;
  p.goal_mask = st->goal;
  int ags[2] = {ai, aj}, n_ag = aj >= 0 ? 2 : 1;
  if (n_ag == 2 && aj < ai) ags[0] = aj, ags[1] = ai;
  /* planning env: only the subtask agents; frozen agents -> Agent-Counter, their object deleted */
  gco_env s;
  memset(&s, 0, sizeof(s));
  s.n_agents = n_ag;
  s.n_objs = e0->n_objs;
  for (int k = 0; k < e0->n_objs; k++) {
    s.ob[k] = e0->ob[k];
    if (s.ob[k].alive && s.ob[k].held_by >= 0) {
      int h = s.ob[k].held_by, idx = -1;
      for (int qq = 0; qq < n_ag; qq++)
        if (ags[qq] == h) idx = qq;
      if (idx < 0 && g_planner_level == 0) s.ob[k].alive = 0;
      else if (idx < 0) s.ob[k].held_by = 2; /* level 1: stays in the world, out of reach */
      else s.ob[k].held_by = idx;
    }
  }
  for (int qq = 0; qq < n_ag; qq++) {
    s.ag[qq] = e0->ag[ags[qq]];
    s.ag[qq].hold = -1;
  }
  for (int k = 0; k < s.n_objs; k++)
    if (s.ob[k].alive && s.ob[k].held_by >= 0 && s.ob[k].held_by < n_ag) s.ag[s.ob[k].held_by].hold = k;
  for (int i = 0; i < e0->n_agents; i++) {
    int in_set = 0;
    for (int qq = 0; qq < n_ag; qq++) in_set |= ags[qq] == i;
    if (!in_set) p.lv.type[e0->ag[i].y][e0->ag[i].x] = g_planner_level ? GCO_BLOCKED : GCO_COUNTER;
  }
  p.base_count = goal_count(&p, &s);
  gco_last_search_states = 0;
  for (int a = 0; a < 25; a++) q[a] = INFINITY;
  int v1 = single_actions(&p.lv, &s, 0), v2 = n_ag == 2 ? single_actions(&p.lv, &s, 1) : (1 << 4);
  double best = INFINITY;
  int budget = 0;
  for (int a1 = 0; a1 < 5; a1++)
    for (int a2 = 0; a2 < 5; a2++) {
      if (!((v1 >> a1) & 1) || !((v2 >> a2) & 1)) continue;
      if (n_ag == 1 && a2 != 4) continue;
      if (n_ag == 2 && !joint_ok(&p.lv, &s, a1, a2)) continue;
      gco_env n = s;
      plan_interact(&p.lv, &n, 0, a1);
      if (n_ag == 2) plan_interact(&p.lv, &n, 1, a2);
      int r = solve(&p, &n, n_ag, max_states);
      if (r == -3) budget = 1;
      if (r < 0) continue;
      double qa = 1.0 + 0.1 * ((a1 != 4) + (n_ag == 2 && a2 != 4)) + r / 10.0;
      q[n_ag == 1 ? a1 : 5 * a1 + a2] = qa;
      if (qa < best) best = qa;
    }
  *v = best;
  if (isinf(best)) return budget ? 3 : 2;
  return 0;
}
