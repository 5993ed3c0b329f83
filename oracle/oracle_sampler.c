/*
 * oracle_sampler.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, CPU restatement of the reference shaDow-GNN subgraph sampler
 * (facebookresearch/shaDow_GNN, para_graph_sampler/graph_engine/backend/).  It exists so the
 * CUDA path has a checker that travels to the GPU box (where /root/reference does not exist).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (shadow_gnn_b200/) never does.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors of its own
 * (SURVEY.md 4.1), so this restatement is pinned against the reference itself: the unmodified
 * reference sources are compiled into oracle/_ref/ (oracle/Makefile) and compared bit-for-bit with
 * this file by tests/test_oracle.py (when oracle/_ref has been built, i.e. where /root/reference is present) and through the
 * committed fixtures under tests/golden/ (generated from oracle/_ref by tests/golden/make_golden.py).
 *
 * Abbreviations for citations: PS.cpp = backend/ParallelSampler.cpp, PS.h = backend/ParallelSampler.h,
 * G.cpp / G.h = backend/Graph.cpp|h.   NodeType = uint32_t, PPRType = float (G.h:16-17).
 *
 * Deliberate, documented deviation: where the reference reads indices[nnz] (one word past the end
 * of the vector -- undefined behaviour, PS.cpp:401-405 on the last row) this file treats the slot
 * as "no edge".  Parity fixtures append a 2-node tail component so the reference never takes that read.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint32_t node_t;
#define NONE32 0xFFFFFFFFu

/* ------------------------------------------------------------------------------------------ */
/* glibc rand()/srand() (TYPE_3 additive feedback generator), restated from the published      */
/* algorithm (glibc stdlib/random_r.c); the reference draws from it at PS.cpp:534 and seeds it  */
/* at PS.h:49-53.  Checked against libc's own rand() in tests/test_oracle.py.                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t r[34];
  int f, b;          /* front / rear indices into the 31-word ring (r[0..30]) */
} orc_rand_t;

void orc_srand(orc_rand_t *st, uint32_t seed) {
  int32_t word;
  if (seed == 0) seed = 1;
  st->r[0] = seed;
  word = (int32_t)seed;
  for (int i = 1; i < 31; i++) {
    long hi = word / 127773, lo = word % 127773;
    word = (int32_t)(16807 * lo - 2836 * hi);
    if (word < 0) word += 2147483647;
    st->r[i] = (uint32_t)word;
  }
  st->f = 3; st->b = 0;
  for (int i = 0; i < 310; i++) {           /* glibc discards the first 310 outputs */
    st->r[st->f] += st->r[st->b];
    st->f = (st->f + 1) % 31; st->b = (st->b + 1) % 31;
  }
}

static inline uint32_t orc_rand_next(orc_rand_t *st) {
  st->r[st->f] += st->r[st->b];
  uint32_t out = st->r[st->f] >> 1;
  st->f = (st->f + 1) % 31; st->b = (st->b + 1) % 31;
  return out;
}
uint32_t orc_rand(orc_rand_t *st) { return orc_rand_next(st); }
void orc_rand_fill(orc_rand_t *st, uint32_t *out, int64_t n) {
  for (int64_t i = 0; i < n; i++) out[i] = orc_rand_next(st);
}

/* ------------------------------------------------------------------------------------------ */
/* sampler object: mirrors the state of class ParallelSampler (PS.h:144-157)                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const node_t *indptr, *indices;     /* GraphStruct (G.h:19-39); borrowed from the caller */
  node_t num_nodes, num_edges;
  node_t *nodes_target; node_t num_target;   /* PS.h:144 */
  node_t idx_root;                           /* PS.h:156 */
  int num_sampler_per_batch;                 /* PS.h:150 */
  int num_threads;
  orc_rand_t rng;                            /* stands in for the process-global rand() state */
  /* PPR tables (PS.h:145-146): CSR-like, row = node id */
  uint64_t *ppr_ptr;  /* [num_nodes+1] or NULL */
  node_t *ppr_neighs; float *ppr_scores;
  node_t prev_start, prev_end;               /* roots of the last call (shared by ensemble branches, PS.cpp:672) */
} orc_sampler;

/* method ids shared with the product's C ABI (include/shadow_b200.h) */
enum { ORC_KHOP = 0, ORC_PPR = 1, ORC_PPR_ST = 2, ORC_NODEIID = 3 };
enum { ORC_AUG_HOPS = 1, ORC_AUG_PPRS = 2, ORC_AUG_DRNLS = 4 };

typedef struct {
  int method, num_roots, depth, budget, k;
  float threshold;
  int add_self_edge, include_target_conn, return_target_only, aug;
  int fixed_mode;     /* 0 = bug-compatible with PS.cpp:401 (default), 1 = row bound fixed */
} orc_cfg;

/* one subgraph == SubgraphStruct (G.h:42-57) */
typedef struct {
  node_t n, m, nt;
  node_t *indptr, *indices, *orig_node, *orig_edge, *target, *hop, *drnl;
  float *ppr;
  int has_csr, has_hop, has_drnl;
} orc_subg;

/* a batch == SubgraphStructVec (G.h:59-97), flattened: arrays are concatenations, *_ptr delimit them */
typedef struct {
  int num_valid;
  int64_t *node_ptr, *edge_ptr, *indptr_ptr, *target_ptr, *hop_ptr, *drnl_ptr;   /* [num_valid+1] */
  node_t *indptr, *indices, *orig_node, *orig_edge, *target, *hop, *drnl;
  float *ppr;
  int64_t rand_draws;    /* number of rand() outputs this call consumed */
} orc_batch;

orc_sampler *orc_create(const node_t *indptr, const node_t *indices, node_t num_nodes, node_t num_edges,
                        int num_sampler_per_batch, int num_threads, int seed) {
  orc_sampler *s = (orc_sampler *)calloc(1, sizeof(orc_sampler));
  s->indptr = indptr; s->indices = indices; s->num_nodes = num_nodes; s->num_edges = num_edges;
  s->num_sampler_per_batch = num_sampler_per_batch;
  s->num_threads = num_threads;
  orc_srand(&s->rng, (uint32_t)seed);                    /* PS.h:49-53 (seed >= 0 branch) */
  return s;
}
void orc_destroy(orc_sampler *s) {
  if (!s) return;
  free(s->nodes_target); free(s->ppr_ptr); free(s->ppr_neighs); free(s->ppr_scores); free(s);
}
/* PS.cpp:36-43 (pre-shuffled branch; idx_root is NOT reset) */
void orc_shuffle_targets(orc_sampler *s, const node_t *t, node_t n) {
  free(s->nodes_target);
  s->nodes_target = (node_t *)malloc(sizeof(node_t) * (n ? n : 1));
  memcpy(s->nodes_target, t, sizeof(node_t) * n);
  s->num_target = n;
}
node_t orc_get_idx_root(orc_sampler *s) { return s->idx_root; }
void orc_reseed(orc_sampler *s, int seed) { orc_srand(&s->rng, (uint32_t)seed); }

/* install PPR tables (flattened top_ppr_neighs / top_ppr_scores) */
void orc_set_ppr(orc_sampler *s, const uint64_t *ptr, const node_t *neighs, const float *scores) {
  free(s->ppr_ptr); free(s->ppr_neighs); free(s->ppr_scores);
  uint64_t tot = ptr[s->num_nodes];
  s->ppr_ptr = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)s->num_nodes + 1));
  memcpy(s->ppr_ptr, ptr, sizeof(uint64_t) * ((size_t)s->num_nodes + 1));
  s->ppr_neighs = (node_t *)malloc(sizeof(node_t) * (tot ? tot : 1));
  s->ppr_scores = (float *)malloc(sizeof(float) * (tot ? tot : 1));
  memcpy(s->ppr_neighs, neighs, sizeof(node_t) * tot);
  memcpy(s->ppr_scores, scores, sizeof(float) * tot);
}

/* ------------------------------------------------------------------------------------------ */
/* small helpers: sorted-unique node sets (stand in for std::set / unordered_map keys)          */
/* ------------------------------------------------------------------------------------------ */
typedef struct { node_t *v; int64_t n, cap; } vec32;
static void v_push(vec32 *a, node_t x) {
  if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (node_t *)realloc(a->v, sizeof(node_t) * a->cap); }
  a->v[a->n++] = x;
}
static int cmp_u32(const void *a, const void *b) {
  node_t x = *(const node_t *)a, y = *(const node_t *)b; return (x > y) - (x < y);
}
static void v_sort_unique(vec32 *a) {
  if (a->n < 2) return;
  qsort(a->v, a->n, sizeof(node_t), cmp_u32);
  int64_t w = 1;
  for (int64_t i = 1; i < a->n; i++) if (a->v[i] != a->v[w - 1]) a->v[w++] = a->v[i];
  a->n = w;
}
static int64_t lower_bound32(const node_t *a, int64_t n, node_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return lo;
}
static int64_t find32(const node_t *a, int64_t n, node_t x) {   /* index or -1 */
  int64_t p = lower_bound32(a, n, x); return (p < n && a[p] == x) ? p : -1;
}

/* (node, ppr) pairs with "last write wins" map semantics (unordered_map operator[]=) */
typedef struct { node_t id; float ppr; int64_t seq; } touched_t;
static int cmp_touched(const void *a, const void *b) {
  const touched_t *x = (const touched_t *)a, *y = (const touched_t *)b;
  if (x->id != y->id) return (x->id > y->id) - (x->id < y->id);
  return (x->seq > y->seq) - (x->seq < y->seq);
}
typedef struct { touched_t *v; int64_t n, cap; } tvec;
static void t_push(tvec *a, node_t id, float p) {
  if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (touched_t *)realloc(a->v, sizeof(touched_t) * a->cap); }
  a->v[a->n].id = id; a->v[a->n].ppr = p; a->v[a->n].seq = a->n; a->n++;
}
/* collapse to sorted-unique ids keeping the LAST written value (insert_only: keep the FIRST) */
static void t_finalize(tvec *a, int keep_first) {
  if (a->n == 0) return;
  qsort(a->v, a->n, sizeof(touched_t), cmp_touched);
  int64_t w = 0;
  for (int64_t i = 0; i < a->n; i++) {
    if (w > 0 && a->v[w - 1].id == a->v[i].id) { if (!keep_first) a->v[w - 1] = a->v[i]; }
    else a->v[w++] = a->v[i];
  }
  a->n = w;
}

/* ------------------------------------------------------------------------------------------ */
/* compute_hops (G.cpp:32-64): BFS distance from one target inside the subgraph CSR             */
/* ------------------------------------------------------------------------------------------ */
static void compute_hops(const orc_subg *g, node_t t, node_t *hop) {
  node_t N = g->n;
  for (node_t i = 0; i < N; i++) hop[i] = NONE32;
  if (N == 0) return;
  uint8_t *vis = (uint8_t *)calloc(N, 1);
  node_t *q = (node_t *)malloc(sizeof(node_t) * N), *qh = (node_t *)malloc(sizeof(node_t) * N);
  node_t head = 0, tail = 0;
  vis[t] = 1; q[tail] = t; qh[tail] = 0; tail++;
  while (head < tail) {
    node_t cur = q[head], h = qh[head]; head++;
    hop[cur] = h;
    for (node_t i = g->indptr[cur]; i < g->indptr[cur + 1]; i++) {
      node_t u = g->indices[i];
      if (!vis[u]) { vis[u] = 1; q[tail] = u; qh[tail] = h + 1; tail++; }
    }
  }
  free(vis); free(q); free(qh);
}
/* compute_drnl_single (G.cpp:66-73), unsigned 32-bit arithmetic */
static node_t drnl_single(node_t dx, node_t dy) {
  if (dx >= 255 || dy >= 255) return 255;
  node_t d = dx + dy, mn = dx < dy ? dx : dy;
  return 1 + mn + (d / 2) * ((d / 2) + (d % 2) - 1);
}

/* ------------------------------------------------------------------------------------------ */
/* _node_induced_subgraph (PS.cpp:350-453)                                                      */
/* ------------------------------------------------------------------------------------------ */
static int in_targets(const node_t *targets, int nt, node_t v) {
  for (int i = 0; i < nt; i++) if (targets[i] == v) return 1;
  return 0;
}
static void induce(const orc_sampler *s, const touched_t *touched, int64_t n, const node_t *targets, int nt,
                   int include_self_conn, int include_target_conn, int aug, int fixed_mode, orc_subg *g) {
  const node_t *indptr = s->indptr, *indices = s->indices;
  if (nt == 1) include_target_conn = 1;                               /* PS.cpp:356-358 */
  g->n = (node_t)n; g->nt = (node_t)nt; g->has_csr = 1;
  g->orig_node = (node_t *)malloc(sizeof(node_t) * (n ? n : 1));
  g->ppr = (float *)malloc(sizeof(float) * (n ? n : 1));
  for (int64_t i = 0; i < n; i++) { g->orig_node[i] = touched[i].id; g->ppr[i] = touched[i].ppr; }   /* :359-366 */
  g->target = (node_t *)malloc(sizeof(node_t) * (nt ? nt : 1));
  for (int i = 0; i < nt; i++) {                                      /* :373-377; a missing key default-inserts 0 */
    int64_t f = find32(g->orig_node, n, targets[i]);
    g->target[i] = f < 0 ? 0 : (node_t)f;
  }
  g->indptr = (node_t *)calloc((size_t)n + 1, sizeof(node_t));
  vec32 cols = {0}, eids = {0};
  for (int64_t r = 0; r < n; r++) {                                   /* PS.cpp:381-427 */
    node_t v = g->orig_node[r];
    node_t idx_start = indptr[v], idx_end = indptr[v + 1];
    node_t idx_insert = NONE32;                                       /* NodeType idx_insert = -1  (:385) */
    if (include_self_conn) {                                          /* :386-400 (row assumed sorted) */
      int64_t lb = lower_bound32(indices + idx_start, idx_end - idx_start, v);
      int present = (lb < (int64_t)(idx_end - idx_start)) && indices[idx_start + lb] == v;
      if (!present) idx_insert = idx_start + (node_t)lb;
    }
    /* PS.cpp:401: `idx_insert >= 0` on an unsigned is always true => the bound is idx_end + 1 in BOTH
       cases.  With an insertion that is the intended extra slot; without, it is one slot too far. */
    node_t idx_end_adj = idx_end + 1;
    if (fixed_mode && idx_insert == NONE32) idx_end_adj = idx_end;
    int passed_self = 0;
    node_t cnt = 0;
    int v_is_target = in_targets(targets, nt, v);
    for (node_t e = idx_start; e < idx_end_adj; e++) {
      node_t e_adj = passed_self ? e - 1 : e;
      if (e == idx_insert) {                                          /* :406-411 */
        passed_self = 1;
        v_push(&cols, (node_t)r); v_push(&eids, NONE32); cnt++;
        continue;
      }
      if (e_adj >= s->num_edges) continue;      /* reference: UB read of indices[nnz]; defined here as "no edge" */
      node_t nb = indices[e_adj];
      int64_t sub = find32(g->orig_node, n, nb);
      if (sub >= 0 && (include_target_conn || !v_is_target || !in_targets(targets, nt, nb))) {   /* :412-419 */
        v_push(&cols, (node_t)sub); v_push(&eids, e_adj); cnt++;
      }
    }
    g->indptr[r + 1] = g->indptr[r] + cnt;                            /* :428-431 */
  }
  g->m = (node_t)cols.n;
  g->indices = cols.v ? cols.v : (node_t *)malloc(sizeof(node_t));
  g->orig_edge = eids.v ? eids.v : (node_t *)malloc(sizeof(node_t));
  if (aug & ORC_AUG_HOPS) {                                           /* :433-436 */
    g->hop = (node_t *)malloc(sizeof(node_t) * (n ? n : 1)); g->has_hop = 1;
    compute_hops(g, g->target[0], g->hop);
  }
  if (aug & ORC_AUG_DRNLS) {                                          /* :438-451 (hop is swapped out => empty) */
    node_t *dx = (node_t *)malloc(sizeof(node_t) * (n ? n : 1)), *dy = (node_t *)malloc(sizeof(node_t) * (n ? n : 1));
    compute_hops(g, g->target[0], dx);
    compute_hops(g, g->target[nt > 1 ? 1 : 0], dy);
    g->drnl = (node_t *)malloc(sizeof(node_t) * (n ? n : 1)); g->has_drnl = 1;
    for (int64_t i = 0; i < n; i++) g->drnl[i] = drnl_single(dx[i], dy[i]);
    free(dx); free(dy);
    if (g->has_hop) { free(g->hop); g->hop = NULL; g->has_hop = 0; }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* samplers                                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* khop (PS.cpp:510-556).  rng == NULL => draws come from the precomputed stream `draws` at *pos. */
static void khop_nodes(orc_sampler *s, const node_t *targets, int nt, int depth, int budget, tvec *out, int64_t *ndraw) {
  const node_t *indptr = s->indptr, *indices = s->indices;
  vec32 level = {0};
  for (int i = 0; i < nt; i++) v_push(&level, targets[i]);
  v_sort_unique(&level);                                   /* std::set (:518-522) */
  for (int64_t i = 0; i < level.n; i++) t_push(out, level.v[i], -1.0f);
  for (int lvl = 0; lvl < depth; lvl++) {                  /* :524-540 */
    vec32 frontier = {0};
    for (int64_t i = 0; i < level.n; i++) {                /* ascending node id: it is a std::set */
      node_t v = level.v[i];
      node_t deg = indptr[v + 1] - indptr[v];
      if (deg <= (node_t)budget || budget < 0) {           /* unsigned compare, as in :528 */
        for (node_t e = indptr[v]; e < indptr[v + 1]; e++) v_push(&frontier, indices[e]);
      } else {
        for (int b = 0; b < budget; b++) {
          node_t off = orc_rand_next(&s->rng) % deg;       /* rand()%deg (:534) */
          (*ndraw)++;
          v_push(&frontier, indices[indptr[v] + off]);
        }
      }
    }
    v_sort_unique(&frontier);
    for (int64_t i = 0; i < frontier.n; i++) t_push(out, frontier.v[i], -1.0f);
    free(level.v); level = frontier;
  }
  free(level.v);
  t_finalize(out, 1);                                      /* insert() keeps the first; all values are -1 anyway */
}

/* ppr (PS.cpp:565-595) */
static void ppr_nodes(const orc_sampler *s, const node_t *targets, int nt, int k, float threshold, tvec *out) {
  for (int it = 0; it < nt; it++) {
    node_t t = targets[it];
    t_push(out, t, -1.0f);                                            /* :574 */
    const node_t *nb = s->ppr_neighs + s->ppr_ptr[t];
    const float *sc = s->ppr_scores + s->ppr_ptr[t];
    int64_t size_all = (int64_t)(s->ppr_ptr[t + 1] - s->ppr_ptr[t]);
    int64_t size_neigh = k < size_all ? k : size_all;                 /* :576 */
    float max_ppr = 0;
    if (size_neigh > 1) max_ppr = sc[1];                              /* :578-579 */
    else if (size_all > 0) t_push(out, t, sc[0]);                     /* :581 (UB in the reference if the row is empty) */
    for (int64_t i = 0; i < size_neigh; i++) {                        /* :583-588 */
      if (max_ppr == 0 || sc[i] / max_ppr < threshold) break;
      t_push(out, nb[i], sc[i]);
    }
  }
  t_finalize(out, 0);                                                 /* operator[]= : last write wins */
}

/* ppr_stochastic (PS.cpp:603-650).  `rand() / RAND_MAX` is an INTEGER division (:635). */
typedef struct { double w; int idx; } wi_t;
static int cmp_wi(const void *a, const void *b) {
  const wi_t *x = (const wi_t *)a, *y = (const wi_t *)b;
  if (x->w < y->w) return -1; if (y->w < x->w) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}
static void ppr_st_nodes(orc_sampler *s, const node_t *targets, int nt, int k, float threshold, tvec *out, int64_t *ndraw) {
  for (int it = 0; it < nt; it++) {
    node_t t = targets[it];
    const node_t *nb = s->ppr_neighs + s->ppr_ptr[t];
    const float *sc = s->ppr_scores + s->ppr_ptr[t];
    int64_t size_all = (int64_t)(s->ppr_ptr[t + 1] - s->ppr_ptr[t]);
    int64_t size_neigh = k < size_all ? k : size_all;
    float max_ppr = size_neigh <= 1 ? 0 : sc[1];                      /* :615 */
    int cnt = 0;
    for (int64_t i = 0; i < size_neigh; i++) {                        /* :617-622: counts the failing entry too */
      cnt++;
      if (max_ppr == 0 || sc[i] / max_ppr < threshold) break;
    }
    wi_t *w = (wi_t *)malloc(sizeof(wi_t) * (size_all ? size_all : 1));
    for (int64_t i = 0; i < size_all; i++) {                          /* :634-638 over the WHOLE stored row */
      double u = (double)(orc_rand_next(&s->rng) / 2147483647u);      /* integer division => 0 (or 1) */
      (*ndraw)++;
      w[i].w = -pow(u, (double)(1 / sc[i])); w[i].idx = (int)i;
    }
    qsort(w, size_all, sizeof(wi_t), cmp_wi);                         /* nth_element: the `cnt` smallest pairs (:639) */
    for (int i = 0; i < cnt && i < size_all; i++) t_push(out, nb[w[i].idx], sc[w[i].idx]);
    free(w);
  }
  t_finalize(out, 0);
}

static void free_subg(orc_subg *g) {
  free(g->indptr); free(g->indices); free(g->orig_node); free(g->orig_edge); free(g->target);
  free(g->hop); free(g->drnl); free(g->ppr);
}

static void sample_one(orc_sampler *s, const orc_cfg *cfg, const node_t *targets, int nt, orc_subg *g, int64_t *ndraw) {
  memset(g, 0, sizeof(*g));
  if (cfg->return_target_only) {                                      /* dummy_sampler (PS.cpp:653-659) */
    g->n = (node_t)nt; g->orig_node = (node_t *)malloc(sizeof(node_t) * (nt ? nt : 1));
    memcpy(g->orig_node, targets, sizeof(node_t) * nt);
    return;
  }
  tvec touched = {0};
  int self_e = cfg->add_self_edge, tconn = cfg->include_target_conn;
  switch (cfg->method) {
    case ORC_KHOP:    khop_nodes(s, targets, nt, cfg->depth, cfg->budget, &touched, ndraw); break;
    case ORC_PPR:     ppr_nodes(s, targets, nt, cfg->k, cfg->threshold, &touched); break;
    case ORC_PPR_ST:  ppr_st_nodes(s, targets, nt, cfg->k, cfg->threshold, &touched, ndraw); break;
    case ORC_NODEIID:                                                 /* PS.cpp:498-508 */
      for (int i = 0; i < nt; i++) t_push(&touched, targets[i], -1.0f);
      t_finalize(&touched, 1); self_e = 0; tconn = 0; break;
    default: break;
  }
  induce(s, touched.v, touched.n, targets, nt, self_e, tconn, cfg->aug, cfg->fixed_mode, g);
  free(touched.v);
}

/* parallel_sampler_ensemble for ONE ensemble branch (PS.cpp:662-704) + _get_roots_p (PS.cpp:456-468).
 * advance_roots == 0 re-samples the same roots (used for the 2nd.. ensemble branches, which share roots). */
orc_batch *orc_sample(orc_sampler *s, const orc_cfg *cfg, int advance_roots) {
  node_t T = s->num_target, idx_start, idx_end;
  if (advance_roots) {
    idx_start = s->idx_root;
    uint64_t want = (uint64_t)idx_start + (uint64_t)cfg->num_roots * (uint64_t)s->num_sampler_per_batch;
    idx_end = (node_t)(want < T ? want : T);
    s->idx_root = (idx_end == T) ? 0 : idx_end;
    s->prev_start = idx_start; s->prev_end = idx_end;
  } else { idx_start = s->prev_start; idx_end = s->prev_end; }
  int P = (int)((idx_end - idx_start + cfg->num_roots - 1) / cfg->num_roots);
  orc_subg *subs = (orc_subg *)calloc(P ? P : 1, sizeof(orc_subg));
  int64_t ndraw = 0;
  int serial = (cfg->method == ORC_KHOP || cfg->method == ORC_PPR_ST) && !cfg->return_target_only;
  if (serial || s->num_threads == 1) {       /* rand() order == subgraph order only single-threaded */
    for (int p = 0; p < P; p++) {
      node_t a = idx_start + (node_t)p * cfg->num_roots;
      int nt = (int)((idx_end - a) < (node_t)cfg->num_roots ? (idx_end - a) : (node_t)cfg->num_roots);
      sample_one(s, cfg, s->nodes_target + a, nt, &subs[p], &ndraw);
    }
  } else {
#ifdef _OPENMP
    if (s->num_threads > 0) omp_set_num_threads(s->num_threads);
#endif
    #pragma omp parallel for schedule(dynamic, 4)
    for (int p = 0; p < P; p++) {
      node_t a = idx_start + (node_t)p * cfg->num_roots;
      int nt = (int)((idx_end - a) < (node_t)cfg->num_roots ? (idx_end - a) : (node_t)cfg->num_roots);
      int64_t dummy = 0;
      sample_one(s, cfg, s->nodes_target + a, nt, &subs[p], &dummy);
    }
  }
  /* flatten */
  orc_batch *b = (orc_batch *)calloc(1, sizeof(orc_batch));
  b->num_valid = P; b->rand_draws = ndraw;
  size_t np1 = (size_t)P + 1;
  b->node_ptr = (int64_t *)calloc(np1, 8); b->edge_ptr = (int64_t *)calloc(np1, 8); b->indptr_ptr = (int64_t *)calloc(np1, 8);
  b->target_ptr = (int64_t *)calloc(np1, 8); b->hop_ptr = (int64_t *)calloc(np1, 8); b->drnl_ptr = (int64_t *)calloc(np1, 8);
  for (int p = 0; p < P; p++) {
    orc_subg *g = &subs[p];
    b->node_ptr[p + 1] = b->node_ptr[p] + g->n;
    b->edge_ptr[p + 1] = b->edge_ptr[p] + g->m;
    b->indptr_ptr[p + 1] = b->indptr_ptr[p] + (g->has_csr ? g->n + 1 : 0);
    b->target_ptr[p + 1] = b->target_ptr[p] + (g->has_csr ? g->nt : 0);
    b->hop_ptr[p + 1] = b->hop_ptr[p] + (g->has_hop ? g->n : 0);
    b->drnl_ptr[p + 1] = b->drnl_ptr[p] + (g->has_drnl ? g->n : 0);
  }
#define ALLOC32(tot) ((node_t *)malloc(sizeof(node_t) * ((tot) ? (tot) : 1)))
  b->orig_node = ALLOC32(b->node_ptr[P]); b->ppr = (float *)malloc(sizeof(float) * (b->node_ptr[P] ? b->node_ptr[P] : 1));
  b->indices = ALLOC32(b->edge_ptr[P]); b->orig_edge = ALLOC32(b->edge_ptr[P]);
  b->indptr = ALLOC32(b->indptr_ptr[P]); b->target = ALLOC32(b->target_ptr[P]);
  b->hop = ALLOC32(b->hop_ptr[P]); b->drnl = ALLOC32(b->drnl_ptr[P]);
  for (int p = 0; p < P; p++) {
    orc_subg *g = &subs[p];
    memcpy(b->orig_node + b->node_ptr[p], g->orig_node, sizeof(node_t) * g->n);
    if (g->has_csr) {
      memcpy(b->ppr + b->node_ptr[p], g->ppr, sizeof(float) * g->n);
      memcpy(b->indices + b->edge_ptr[p], g->indices, sizeof(node_t) * g->m);
      memcpy(b->orig_edge + b->edge_ptr[p], g->orig_edge, sizeof(node_t) * g->m);
      memcpy(b->indptr + b->indptr_ptr[p], g->indptr, sizeof(node_t) * ((size_t)g->n + 1));
      memcpy(b->target + b->target_ptr[p], g->target, sizeof(node_t) * g->nt);
    }
    if (g->has_hop) memcpy(b->hop + b->hop_ptr[p], g->hop, sizeof(node_t) * g->n);
    if (g->has_drnl) memcpy(b->drnl + b->drnl_ptr[p], g->drnl, sizeof(node_t) * g->n);
    free_subg(g);
  }
  free(subs);
  return b;
}
void orc_batch_free(orc_batch *b) {
  if (!b) return;
  free(b->node_ptr); free(b->edge_ptr); free(b->indptr_ptr); free(b->target_ptr); free(b->hop_ptr); free(b->drnl_ptr);
  free(b->indptr); free(b->indices); free(b->orig_node); free(b->orig_edge); free(b->target); free(b->hop); free(b->drnl);
  free(b->ppr); free(b);
}

/* ------------------------------------------------------------------------------------------ */
/* preproc_ppr_approximate (PS.cpp:237-344): forward push in float32, no FMA contraction.       */
/* Output rows are written for the targets only, at out_neighs/out_scores[i*k .. i*k+out_len[i]) */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float negpi; node_t id; } pi_idx_t;
static int cmp_pi_idx(const void *a, const void *b) {      /* std::pair<float, NodeType> operator< */
  const pi_idx_t *x = (const pi_idx_t *)a, *y = (const pi_idx_t *)b;
  if (x->negpi < y->negpi) return -1; if (y->negpi < x->negpi) return 1;
  return (x->id > y->id) - (x->id < y->id);
}
typedef struct { node_t *h; int64_t n, cap; } heap_t;
static void heap_push(heap_t *hp, node_t x) {
  if (hp->n == hp->cap) { hp->cap = hp->cap ? hp->cap * 2 : 256; hp->h = (node_t *)realloc(hp->h, sizeof(node_t) * hp->cap); }
  int64_t i = hp->n++; hp->h[i] = x;
  while (i > 0) { int64_t p = (i - 1) >> 1; if (hp->h[p] <= hp->h[i]) break; node_t t = hp->h[p]; hp->h[p] = hp->h[i]; hp->h[i] = t; i = p; }
}
static void heap_pop(heap_t *hp) {
  hp->h[0] = hp->h[--hp->n];
  int64_t i = 0;
  for (;;) {
    int64_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < hp->n && hp->h[l] < hp->h[m]) m = l;
    if (r < hp->n && hp->h[r] < hp->h[m]) m = r;
    if (m == i) break;
    node_t t = hp->h[m]; hp->h[m] = hp->h[i]; hp->h[i] = t; i = m;
  }
}

void orc_ppr_push(const node_t *indptr, const node_t *indices, node_t num_nodes,
                  const node_t *targets, int64_t num_targets, int k, float alpha_in, float epsilon,
                  int num_threads, node_t *out_neighs, float *out_scores, node_t *out_len) {
  const float alpha = 1 - alpha_in;                                   /* PS.cpp:242 */
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
  #pragma omp parallel
  {
    /* dense pi / residue vectors as in the `use_map == false` branch (:266-269); the std::map branch
       (n > 5M) performs the same arithmetic on the same values (absent key == 0.f). */
    float *pi = (float *)calloc(num_nodes, sizeof(float));
    float *res = (float *)calloc(num_nodes, sizeof(float));
    uint8_t *in_set = (uint8_t *)calloc(num_nodes, 1);     /* membership of prop_set (std::set, :271) */
    uint8_t *seen = (uint8_t *)calloc(num_nodes, 1);       /* key present in touched_neigh_map */
    vec32 dirty = {0}, touched = {0};
    heap_t hp = {0};
    #pragma omp for schedule(dynamic, 8)
    for (int64_t it = 0; it < num_targets; it++) {
      node_t target = targets[it];
      res[target] = 1.0f; v_push(&dirty, target);
      heap_push(&hp, target); in_set[target] = 1;
      while (hp.n > 0) {
        node_t v = hp.h[0];                                           /* *(prop_set.begin()) : smallest id (:273) */
        if (!in_set[v]) { heap_pop(&hp); continue; }                  /* stale heap entry */
        float r0 = res[v];
        pi[v] += alpha * r0;                                          /* :284 */
        node_t degv = indptr[v + 1] - indptr[v];
        float m = (1 - alpha) * r0 / (float)(2 * degv);               /* :286 (2*deg in uint32, then converted) */
        for (node_t e = indptr[v]; e < indptr[v + 1]; e++) {          /* :287-303 */
          node_t u = indices[e];
          if (res[u] == 0.0f && pi[u] == 0.0f) v_push(&dirty, u);
          res[u] += m;
          node_t degu = indptr[u + 1] - indptr[u];
          if (res[u] > epsilon * (float)degu) { if (!in_set[u]) { in_set[u] = 1; heap_push(&hp, u); } }
        }
        res[v] = r0 * (1 - alpha) / 2;                                /* :312 */
        if (res[v] <= epsilon * (float)degv) {                        /* :313-316 */
          in_set[v] = 0;
          if (!seen[v]) { seen[v] = 1; v_push(&touched, v); }
        }
      }
      hp.n = 0;
      int64_t nt = touched.n;
      pi_idx_t *pv = (pi_idx_t *)malloc(sizeof(pi_idx_t) * (nt ? nt : 1));
      for (int64_t i = 0; i < nt; i++) { pv[i].negpi = -pi[touched.v[i]]; pv[i].id = touched.v[i]; }
      qsort(pv, nt, sizeof(pi_idx_t), cmp_pi_idx);                    /* nth_element + sort of the first _k (:326-327) */
      int64_t kk = k < nt ? k : nt;                                   /* :320 */
      for (int64_t i = 0; i < kk; i++) { out_neighs[it * k + i] = pv[i].id; out_scores[it * k + i] = -pv[i].negpi; }
      out_len[it] = (node_t)kk;
      free(pv);
      for (int64_t i = 0; i < dirty.n; i++) { pi[dirty.v[i]] = 0; res[dirty.v[i]] = 0; in_set[dirty.v[i]] = 0; }
      for (int64_t i = 0; i < touched.n; i++) { seen[touched.v[i]] = 0; pi[touched.v[i]] = 0; res[touched.v[i]] = 0; }
      dirty.n = 0; touched.n = 0;
    }
    free(pi); free(res); free(in_set); free(seen); free(dirty.v); free(touched.v); free(hp.h);
  }
}
