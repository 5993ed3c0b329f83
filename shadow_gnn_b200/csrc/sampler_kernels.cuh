// sampler_kernels.cuh -- fused subgraph extraction + node-induced CSR relabelling for sm_100a.
//
// One CTA builds one subgraph (root group) end to end, entirely on-chip except for the full-graph
// reads and the final output writes:
//   A. node-set selection     ppr: top-k row of the PPR table + relative threshold      (PS.cpp:565-595)
//                             khop: level-by-level expansion with per-level dedup        (PS.cpp:510-556)
//                             nodeIID: the roots                                         (PS.cpp:498-508)
//   B. sort ids, build the orig->sub hash in shared memory                               (PS.cpp:359-377)
//   C. count pass over the full-graph rows of the selected nodes (warp per row, coalesced),
//      block scan -> local indptr, decoupled look-back across CTAs -> batch-global offsets,
//      fill pass (rows re-read from L2) writing the block-diagonal CSR in place           (PS.cpp:379-431)
//   D. optional in-subgraph BFS for the hop / drnl labels                                (G.cpp:32-73)
// CTAs are persistent and take subgraphs from a ticket counter, so the look-back never waits on a
// CTA that has not started.  The batch layout written here is what the reference assembles on the host
// in Subgraph.cat_to_block_diagonal (frontend/graph.py:280-320).
#pragma once
#include "common.cuh"

#define SHADOW_MAX_ROOTS 4
#define SAMPLER_BLOCK 128
#ifndef SAMPLER_MIN_BLOCKS
#define SAMPLER_MIN_BLOCKS 10
#endif

struct WsLayout {          // byte offsets of the per-CTA workspace (shared memory, or global for huge scopes)
  uint32_t keys, cval, nodes, pprv, hkeys, st_row, row_s, row_e, row_cnt, row_ins, rowd, row_first, row_last, row_fgt, level, all, dist, fr_a, fr_b, st;
  uint32_t bytes;
};

struct WarpLayout {        // byte offsets of the per-warp workspace of ppr_induce_warp_kernel (ppr_warp_kernel.cuh)
  uint32_t nodes, rs, cp, rc, hkeys, ovf, bloom, queue, rlo, rins, rbug;
  uint32_t bytes;
};

struct SampleParams {
  // full graph (a1: GraphStruct, G.h:19-39)
  const uint32_t *indptr, *indices;
  uint32_t num_nodes, num_edges;
  // roots of this call: nodes_target[idx_start .. idx_end)
  const uint32_t *roots;
  int num_root_ids, num_subg;
  // config
  int method, num_roots, depth, budget, k;
  float threshold;
  int add_self, tconn, aug, fixed_mode, rng_mode;
  int count_only_last;               // glibc prepass: build levels < depth-1, only count the draws of the last
  // PPR tables (a7)
  const unsigned long long *ppr_ptr;
  const uint32_t *ppr_neighs;
  const float *ppr_scores;
  // the same rows re-ordered by node id at table-install time (+ each entry's position in score order): the single-root
  // PPR node set then needs no sort at all.  NULL => generic path.
  const uint32_t *ppr_sid;
  const float *ppr_sscore;
  const unsigned short *ppr_srank;
  const uint2 *ppr_srow;             // {indptr[id], degree(id)} of every entry of the id-sorted rows (warp fast path)
  // symmetric-graph variant of the warp fast path (verified once per graph, sym_build_kernel): upper part {first slot with neighbour >= id,
  // slots to the row end} of every table entry, and the reverse-slot index rev[slot of (u,v)] = slot of (v,u)
  const uint2 *ppr_supper;
  const uint32_t *sym_rev;
  // random streams
  const uint32_t *rand_stream;       // glibc replay: pre-generated rand() outputs
  long long *rand_off;               // [num_subg+1] stream offset of each subgraph (written by the prepass)
  uint32_t philox_seed, philox_epoch, root_slot_base;
  // per-subgraph capacities
  int ncap, ccap, ccap2, acap, acap2, hcap, hshift;
  int ecap;                          // staged edges per subgraph (0 = two-pass generic path)
  WsLayout L;
  unsigned char *gws;                // global workspace (GWS variant)
  unsigned long long gws_stride;
  // outputs (batch-global block-diagonal CSR)
  long long cap_nodes, cap_edges;
  int *node_ptr, *indices_out, *target, *num_target;
  int2 *row_span, *edge_span;        // raw layout: [start,end) of every row / subgraph inside indices_out (see shadow_b200.h)
  uint32_t *orig_node, *orig_edge, *hop, *drnl;
  float *ppr_out;
  // inter-CTA state
  unsigned long long *status_n;      // decoupled look-back words over the node counts: flag<<62 | value
  uint32_t *ticket;
  long long *totals;                 // [0]=nodes [1]=edge cursor (atomic) [2]=error bits
  // warp fast path (ppr_warp_kernel.cuh) and its hand-over to this file's kernel
  WarpLayout WL;
  int w_ecap, w_hbuckets, w_hshift;
  int *redo_list;                    // subgraphs the fast path could not hold on-chip; non-NULL in a redo launch of sample_induce_kernel
  uint32_t *redo_count, *ticket2;
  const unsigned short *w_cut;       // per-subgraph score-rank cut from ppr_count_kernel
  unsigned char *w_scratch;          // staged-edge scratch of the warp fast path: one region per resident warp (global memory, lives in L2)
  unsigned long long w_scratch_stride;
};

enum { ERR_WS_OVERFLOW = 1, ERR_OUT_OVERFLOW = 2 };

struct Ws {
  unsigned long long *keys; float *cval; uint32_t *nodes; float *pprv; uint32_t *hkeys, *row_s, *row_e, *row_cnt,
      *row_ins, *row_first, *row_last, *row_fgt, *level, *all, *dist, *fr_a, *fr_b;
  uint4 *rowd;
  uint2 *st;
  unsigned short *st_row;
};
__device__ __forceinline__ Ws make_ws(unsigned char *b, const WsLayout &L) {
  Ws w;
  w.keys = (unsigned long long *)(b + L.keys); w.cval = (float *)(b + L.cval);
  w.nodes = (uint32_t *)(b + L.nodes); w.pprv = (float *)(b + L.pprv);
  w.hkeys = (uint32_t *)(b + L.hkeys); w.st_row = (unsigned short *)(b + L.st_row);
  w.row_s = (uint32_t *)(b + L.row_s); w.row_e = (uint32_t *)(b + L.row_e);
  w.row_cnt = (uint32_t *)(b + L.row_cnt); w.row_ins = (uint32_t *)(b + L.row_ins);
  w.level = (uint32_t *)(b + L.level); w.all = (uint32_t *)(b + L.all);
  w.dist = (uint32_t *)(b + L.dist); w.fr_a = (uint32_t *)(b + L.fr_a); w.fr_b = (uint32_t *)(b + L.fr_b);
  w.rowd = (uint4 *)(b + L.rowd); w.row_first = (uint32_t *)(b + L.row_first); w.row_last = (uint32_t *)(b + L.row_last);
  w.row_fgt = (uint32_t *)(b + L.row_fgt); w.st = (uint2 *)(b + L.st);
  return w;
}

// ordered stream compaction over i in [0,n): emit(i, rank) for every i with pred(i); returns the count
template <class Pred, class Emit>
__device__ inline uint32_t block_ordered_compact(int n, Pred pred, Emit emit, uint32_t *warp_sums) {
  const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
  uint32_t carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    bool f = (i < n) && pred(i);
    uint32_t mask = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_sums[w] = __popc(mask);
    __syncthreads();
    uint32_t woff = 0, tot = 0;
    for (int j = 0; j < nw; j++) { uint32_t c = warp_sums[j]; if (j < w) woff += c; tot += c; }
    if (f) emit(i, carry + woff + __popc(mask & lanemask_lt()));
    carry += tot;
    __syncthreads();
  }
  return carry;
}

// orig -> sub map: open addressing over BUCKETS of 4 keys (one 16-byte shared-memory load tests 4 slots, so a probe
// almost never needs a second step and the warp-wide worst case stays at one iteration).  Buckets fill in slot order.
__device__ __forceinline__ uint32_t hash_bucket(uint32_t key, int shift) { return (key * 2654435761u) >> shift; }
// membership only: the sub id of a member is its rank in the sorted node list (sub_of), looked up on the rare hit path
__device__ __forceinline__ bool hash_contains(const uint32_t *hk, uint32_t bmask, int shift, uint32_t key) {
  uint32_t b = hash_bucket(key, shift);
  for (;;) {
    const uint4 kk = reinterpret_cast<const uint4 *>(hk)[b];
    if (kk.x == key || kk.y == key || kk.z == key || kk.w == key) return true;
    if (kk.w == NONE32) return false;
    b = (b + 1) & bmask;
  }
}
__device__ __forceinline__ void hash_insert(uint32_t *hk, uint32_t bmask, int shift, uint32_t key) {
  uint32_t b = hash_bucket(key, shift);
  for (;;) {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (atomicCAS(&hk[4 * b + j], NONE32, key) == NONE32) return;
    b = (b + 1) & bmask;
  }
}
__device__ __forceinline__ uint32_t sub_of(const uint32_t *nodes, int n, uint32_t key) {      // rank of a member in the sorted node list
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (nodes[mid] < key) lo = mid + 1; else hi = mid; }
  return (uint32_t)lo;
}
__device__ __forceinline__ uint32_t hash_lookup(const uint32_t *hk, const uint32_t *nodes, int n, uint32_t bmask, int shift, uint32_t key) {
  return hash_contains(hk, bmask, shift, key) ? sub_of(nodes, n, key) : NONE32;
}

// ------------------------------------------------------------------------------------------------
// A. node-set builders.  On return ws.nodes[0..n) is sorted ascending & unique, ws.pprv holds the ppr column.
// ------------------------------------------------------------------------------------------------
__device__ inline int finish_sorted_pairs(const Ws &ws, int ncand, int cap, uint32_t *warp_sums) {
  // ws.keys[0..ncand) = (id << 32 | seq); last write wins (unordered_map operator[]=, PS.cpp:574-587)
  int np2 = next_pow2(ncand);
  for (int i = ncand + threadIdx.x; i < np2; i += blockDim.x) ws.keys[i] = ~0ull;
  __syncthreads();
  block_bitonic_sort(ws.keys, np2);
  const unsigned long long *keys = ws.keys;
  return (int)block_ordered_compact(
      ncand, [&](int i) { return i == ncand - 1 || (uint32_t)(keys[i + 1] >> 32) != (uint32_t)(keys[i] >> 32); },
      [&](int i, uint32_t r) { if ((int)r < cap) { ws.nodes[r] = (uint32_t)(keys[i] >> 32); ws.pprv[r] = ws.cval[(uint32_t)keys[i]]; } }, warp_sums);
}

__device__ inline int build_nodes_ppr(const SampleParams &P, const Ws &ws, const uint32_t *roots, int nt, uint32_t *s_cut,
                                      uint32_t *warp_sums) {
  int ncand = 0;
  for (int it = 0; it < nt; it++) {
    const uint32_t t = roots[it];
    const unsigned long long off = P.ppr_ptr[t];
    const long long len_all = (long long)(P.ppr_ptr[t + 1] - off);
    const int size_neigh = (int)(len_all < (long long)P.k ? len_all : (long long)P.k);            // PS.cpp:576
    const float max_ppr = size_neigh > 1 ? P.ppr_scores[off + 1] : 0.f;                          // :578-579
    if (threadIdx.x == 0) {
      ws.keys[ncand] = ((unsigned long long)t << 32) | (uint32_t)ncand; ws.cval[ncand] = -1.f;   // :574
      *s_cut = (uint32_t)size_neigh;
    }
    ncand++;
    if (size_neigh <= 1 && len_all > 0) {                                                        // :581
      if (threadIdx.x == 0) { ws.keys[ncand] = ((unsigned long long)t << 32) | (uint32_t)ncand; ws.cval[ncand] = P.ppr_scores[off]; }
      ncand++;
    }
    __syncthreads();
    // first index that fails the relative threshold (:584); evaluated per entry, min over entries
    for (int i = threadIdx.x; i < size_neigh; i += blockDim.x) {
      float s = P.ppr_scores[off + i];
      if (max_ppr == 0.f || __fdiv_rn(s, max_ppr) < P.threshold) atomicMin(s_cut, (uint32_t)i);
    }
    __syncthreads();
    const int cut = (int)*s_cut;
    for (int i = threadIdx.x; i < cut; i += blockDim.x) {                                        // :587
      ws.keys[ncand + i] = ((unsigned long long)P.ppr_neighs[off + i] << 32) | (uint32_t)(ncand + i);
      ws.cval[ncand + i] = P.ppr_scores[off + i];
    }
    ncand += cut;
    __syncthreads();
  }
  return finish_sorted_pairs(ws, ncand, P.ncap, warp_sums);
}

// ppr_stochastic (PS.cpp:603-650).  `rand() / RAND_MAX` is an integer division (:635): u is 1 only when the draw equals
// RAND_MAX and 0 otherwise, so the weight -pow(u, 1/s) is -1 or -0 and nth_element keeps the entries whose draw hit RAND_MAX
// (by index) followed by the lowest indices.  One draw is consumed per stored entry of every target row.
__device__ inline int build_nodes_ppr_st(const SampleParams &P, const Ws &ws, const uint32_t *roots, int nt, int p, long long rand_base,
                                         long long *draws_out, uint32_t *s_cut, uint32_t *s_aux /*[32]*/, uint32_t *warp_sums) {
  int ncand = 0;
  long long rcur = rand_base;
  const uint32_t root_slot = P.root_slot_base + (uint32_t)p;
  for (int it = 0; it < nt; it++) {
    const uint32_t t = roots[it];
    const unsigned long long off = P.ppr_ptr[t];
    const int size_all = (int)(P.ppr_ptr[t + 1] - off);
    const int size_neigh = size_all < P.k ? size_all : P.k;
    const float max_ppr = size_neigh <= 1 ? 0.f : P.ppr_scores[off + 1];                           // :615
    if (threadIdx.x == 0) { *s_cut = (uint32_t)size_neigh; s_aux[0] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < size_neigh; i += blockDim.x)                                    // first failing index (:617-622)
      if (max_ppr == 0.f || __fdiv_rn(P.ppr_scores[off + i], max_ppr) < P.threshold) atomicMin(s_cut, (uint32_t)i);
    for (int i = threadIdx.x; i < size_all; i += blockDim.x) {                                    // entries whose draw is RAND_MAX (u == 1)
      const uint32_t d = (P.rng_mode == SHADOW_RNG_GLIBC) ? P.rand_stream[rcur + i]
                                                          : (philox4x32_10_x(t, 0x70707273u, (uint32_t)i, root_slot, P.philox_seed, P.philox_epoch) >> 1);
      if (d == 2147483647u) { const uint32_t q = atomicAdd(&s_aux[0], 1u); if (q < 30) s_aux[2 + q] = (uint32_t)i; }
    }
    __syncthreads();
    const int first_fail = (int)*s_cut;
    const int cnt = first_fail < size_neigh ? first_fail + 1 : size_neigh;                        // the failing entry is counted too
    const int nflag = (int)min(s_aux[0], 30u);
    // ordered append of the selected entries (their order inside one target does not matter: ids are distinct)
    const int nsel = (int)block_ordered_compact(
        size_all,
        [&](int i) { int less_f = 0; bool flagged = false; for (int q = 0; q < nflag; q++) { const int f = (int)s_aux[2 + q]; less_f += f < i; flagged |= f == i; }
                     return (flagged ? less_f : nflag + (i - less_f)) < cnt; },
        [&](int i, uint32_t r) { if (ncand + (int)r < P.ccap) { ws.keys[ncand + r] = ((unsigned long long)P.ppr_neighs[off + i] << 32) | (uint32_t)(ncand + r); ws.cval[ncand + r] = P.ppr_scores[off + i]; } },
        warp_sums);
    ncand += nsel;
    rcur += size_all;
    __syncthreads();
  }
  *draws_out = rcur - rand_base;
  if (ncand > P.ccap) return -1;
  if (ncand == 0) return 0;
  return finish_sorted_pairs(ws, ncand, P.ncap, warp_sums);
}

// ppr, single root, id-sorted table rows (PS.cpp:565-595 with the same outcome as build_nodes_ppr):
//   cut  = first position IN SCORE ORDER whose score fails the relative threshold (or size_neigh)
//   set  = {entries with score rank < cut}  U  {root}      (the root keeps -1, or scores[0] when the row has <= 1 entry, unless its own entry is selected)
__device__ inline int build_nodes_ppr_sorted(const SampleParams &P, const Ws &ws, uint32_t t, uint32_t *s_cut, uint32_t *s_aux, uint32_t *warp_sums) {
  const unsigned long long off = P.ppr_ptr[t];
  const int len_all = (int)(P.ppr_ptr[t + 1] - off);
  const int size_neigh = len_all < P.k ? len_all : P.k;
  const float max_ppr = size_neigh > 1 ? P.ppr_scores[off + 1] : 0.f;
  if (threadIdx.x == 0) { *s_cut = (uint32_t)size_neigh; s_aux[0] = 0; s_aux[1] = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < len_all; i += blockDim.x) {
    const uint32_t r = P.ppr_srank[off + i];
    if (r < (uint32_t)size_neigh) {
      const float sc = P.ppr_sscore[off + i];
      if (max_ppr == 0.f || __fdiv_rn(sc, max_ppr) < P.threshold) atomicMin(s_cut, r);
    }
  }
  __syncthreads();
  const uint32_t cut = *s_cut;
  // is the root's own entry selected, and how many selected ids are below the root?
  for (int base = 0; base < len_all; base += blockDim.x) {
    const int i = base + threadIdx.x;
    bool sel = false; uint32_t id = 0;
    if (i < len_all) { sel = P.ppr_srank[off + i] < cut; id = P.ppr_sid[off + i]; }
    const uint32_t below = __ballot_sync(0xffffffffu, sel && id < t), same = __ballot_sync(0xffffffffu, sel && id == t);
    if (lane_id() == 0 && (below | same)) { if (below) atomicAdd(&s_aux[0], (uint32_t)__popc(below)); if (same) s_aux[1] = 1; }
  }
  __syncthreads();
  const uint32_t n_below = s_aux[0];
  const bool root_in = s_aux[1] != 0;
  const int nsel = (int)block_ordered_compact(
      len_all, [&](int i) { return P.ppr_srank[off + i] < cut; },
      [&](int i, uint32_t r) {
        const uint32_t id = P.ppr_sid[off + i];
        const uint32_t at = r + ((!root_in && id > t) ? 1u : 0u);
        if ((int)at < P.ncap) { ws.nodes[at] = id; ws.pprv[at] = P.ppr_sscore[off + i]; }
      }, warp_sums);
  if (!root_in && threadIdx.x == 0 && (int)n_below < P.ncap) {
    ws.nodes[n_below] = t;
    ws.pprv[n_below] = (size_neigh <= 1 && len_all > 0) ? P.ppr_scores[off] : -1.f;                // PS.cpp:574,581
  }
  __syncthreads();
  return nsel + (root_in ? 0 : 1);
}

// sort + unique of a[0..n) (uint32), result written to out[0..cap); returns the (uncapped) count
__device__ inline int sort_unique_u32(uint32_t *a, int n, uint32_t *out, int cap, uint32_t *warp_sums) {
  int np2 = next_pow2(n);
  for (int i = n + threadIdx.x; i < np2; i += blockDim.x) a[i] = NONE32;
  __syncthreads();
  block_bitonic_sort(a, np2);
  return (int)block_ordered_compact(
      n, [&](int i) { return i == 0 || a[i - 1] != a[i]; }, [&](int i, uint32_t r) { if ((int)r < cap) out[r] = a[i]; }, warp_sums);
}

// khop (PS.cpp:510-556).  Returns n (or -1 on workspace overflow); *draws_out = rand() outputs consumed.
__device__ inline int build_nodes_khop(const SampleParams &P, const Ws &ws, const uint32_t *roots, int nt, int p,
                                       long long rand_base, long long *draws_out, uint32_t *warp_sums, int *s_n) {
  uint32_t *raw = (uint32_t *)ws.keys;     // the 64-bit sort buffer doubles as the raw frontier buffer
  if (threadIdx.x == 0) {                  // level 0 = std::set of the roots (:518-522)
    int n = 0;
    for (int i = 0; i < nt; i++) {
      uint32_t v = roots[i]; int j = 0;
      while (j < n && ws.level[j] < v) j++;
      if (j < n && ws.level[j] == v) continue;
      for (int q = n; q > j; q--) ws.level[q] = ws.level[q - 1];
      ws.level[j] = v; n++;
    }
    for (int i = 0; i < n; i++) ws.all[i] = ws.level[i];
    *s_n = n;
  }
  __syncthreads();
  int n_level = *s_n, n_all = n_level;
  long long rcur = rand_base, draws = 0;
  const uint32_t ubudget = (uint32_t)P.budget;
  const uint32_t root_slot = P.root_slot_base + (uint32_t)p;
  for (int lvl = 0; lvl < P.depth; lvl++) {                                                      // :524-540
    for (int i = threadIdx.x; i < n_level; i += blockDim.x) {
      uint32_t v = ws.level[i], s = P.indptr[v], deg = P.indptr[v + 1] - s;
      bool take_all = (P.budget < 0) || (deg <= ubudget);                                        // :528 (unsigned compare)
      ws.row_s[i] = s; ws.row_e[i] = deg;
      ws.row_cnt[i] = take_all ? deg : ubudget;
      ws.row_ins[i] = take_all ? 0u : ubudget;
    }
    __syncthreads();
    uint32_t total_raw = block_exclusive_scan(ws.row_cnt, n_level, warp_sums);
    uint32_t total_draws = block_exclusive_scan(ws.row_ins, n_level, warp_sums);
    draws += total_draws;
    if (P.count_only_last && lvl == P.depth - 1) break;
    if (total_raw > (uint32_t)P.ccap) return -1;
    for (int i = warp_id(); i < n_level; i += (blockDim.x >> 5)) {
      const uint32_t v = ws.level[i], s = ws.row_s[i], deg = ws.row_e[i], o = ws.row_cnt[i];
      const uint32_t cnt = ((i + 1 < n_level) ? ws.row_cnt[i + 1] : total_raw) - o;
      const bool take_all = (P.budget < 0) || (deg <= ubudget);
      for (uint32_t j = lane_id(); j < cnt; j += 32) {
        uint32_t idx = j;
        if (!take_all) {
          if (P.rng_mode == SHADOW_RNG_GLIBC) idx = P.rand_stream[rcur + ws.row_ins[i] + j] % deg;   // rand()%deg (:534)
          else idx = __umulhi(philox4x32_10_x(v, (uint32_t)lvl, j, root_slot, P.philox_seed, P.philox_epoch), deg);
        }
        raw[o + j] = P.indices[s + idx];
      }
    }
    rcur += total_draws;
    __syncthreads();
    n_level = sort_unique_u32(raw, (int)total_raw, ws.level, P.ncap, warp_sums);                          // std::set frontier
    if (n_level > P.ncap || n_all + n_level > P.acap) return -1;
    for (int i = threadIdx.x; i < n_level; i += blockDim.x) ws.all[n_all + i] = ws.level[i];
    n_all += n_level;
    __syncthreads();
  }
  *draws_out = draws;
  if (P.count_only_last) return 0;
  int n = sort_unique_u32(ws.all, n_all, ws.nodes, P.ncap, warp_sums);                                    // union of the levels (:542-547)
  if (n > P.ncap) return -1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ws.pprv[i] = -1.f;
  return n;
}

// ------------------------------------------------------------------------------------------------
// decoupled look-back over subgraph tickets: exclusive prefix of (n, m)
// ------------------------------------------------------------------------------------------------
#define LB_AGG  (1ull << 62)
#define LB_INCL (2ull << 62)
#define LB_VAL  ((1ull << 62) - 1)
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// The node count of a subgraph is known right after the node-set phase, long before its row scan ends, so it is
// published early (lookback_publish) and resolved late (lookback_resolve): in practice nobody ever spins.
__device__ __forceinline__ void lookback_publish(unsigned long long *status, int p, unsigned long long mine) {
  st_volatile_u64(&status[p], (p == 0 ? LB_INCL : LB_AGG) | mine);
}
// called by warp 0; returns the exclusive prefix in all lanes of warp 0
__device__ inline unsigned long long lookback_resolve(unsigned long long *status, int p, unsigned long long mine) {
  const int lane = lane_id();
  if (p == 0) return 0;
  unsigned long long excl = 0;
  int idx = p - 1;
  for (;;) {
    const int q = idx - lane;
    unsigned long long st = LB_INCL;     // lanes past the beginning act as a zero inclusive prefix
    if (q >= 0) { do { st = ld_volatile_u64(&status[q]); } while ((st >> 62) == 0); }
    const uint32_t incl_mask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
    const int first = incl_mask ? __ffs(incl_mask) - 1 : 32;    // nearest predecessor with an inclusive prefix
    unsigned long long v = (lane <= first) ? (st & LB_VAL) : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    excl += v;
    if (incl_mask) break;
    idx -= 32;
  }
  if (lane == 0) st_volatile_u64(&status[p], LB_INCL | (excl + mine));
  return excl;
}

// ------------------------------------------------------------------------------------------------
// D. BFS inside the just-written subgraph CSR (compute_hops, G.cpp:32-64); level-synchronous
// ------------------------------------------------------------------------------------------------
__device__ inline void subgraph_bfs(const SampleParams &P, const Ws &ws, int n, uint32_t src, long long node_base,
                                    const uint32_t *loc_indptr, long long edge_base, uint32_t *dist, uint32_t *s_cnt) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dist[i] = NONE32;
  __syncthreads();
  if (threadIdx.x == 0) { dist[src] = 0; ws.fr_a[0] = src; *s_cnt = 0; }
  __syncthreads();
  uint32_t *cur = ws.fr_a, *nxt = ws.fr_b;
  int ncur = 1;
  for (uint32_t lvl = 0; ncur > 0; lvl++) {
    for (int i = warp_id(); i < ncur; i += (blockDim.x >> 5)) {
      const uint32_t u = cur[i];
      const uint32_t s = loc_indptr[u], e = loc_indptr[u + 1];
      for (uint32_t j = s + lane_id(); j < e; j += 32) {
        uint32_t c = (uint32_t)(P.indices_out[edge_base + j] - (int)node_base);
        if (atomicCAS(&dist[c], NONE32, lvl + 1) == NONE32) nxt[atomicAdd(s_cnt, 1u)] = c;
      }
    }
    __syncthreads();
    ncur = (int)*s_cnt;
    __syncthreads();
    if (threadIdx.x == 0) *s_cnt = 0;
    uint32_t *t = cur; cur = nxt; nxt = t;
    __syncthreads();
  }
}
__device__ __forceinline__ uint32_t drnl_single(uint32_t dx, uint32_t dy) {      // G.cpp:66-73, uint32 arithmetic
  if (dx >= 255u || dy >= 255u) return 255u;
  uint32_t d = dx + dy, mn = dx < dy ? dx : dy;
  return 1u + mn + (d / 2u) * ((d / 2u) + (d % 2u) - 1u);
}


// ------------------------------------------------------------------------------------------------
// C (fast path): ONE pass over the full-graph rows, flat and chunk-granular.
// Every row is cut into "items" of 32 consecutive slots.  The item space of the subgraph is split evenly over the
// warps; a warp walks its range SCAN_U items at a time, issuing all SCAN_U coalesced 128-byte requests before the
// first use, probes the shared-memory hash and appends the kept edges -- in slot order, hence already in CSR order --
// to its staging region.  Hub rows are simply many items, shared by all warps; no row is read twice and no warp
// waits on a row-sized latency chain.  Everything that is per row (counts, self-edge position, the PS.cpp:401
// bug slot) is derived afterwards from the staged entries, so the inner loop is load -> probe -> ballot -> store.
// ------------------------------------------------------------------------------------------------
#define SCAN_U 8
struct KeepCtx { const uint32_t *hk, *nodes; int n; uint32_t hmask; int hshift; const uint32_t *roots; int nt; };
__device__ __forceinline__ uint32_t keep_lookup(const KeepCtx &K, uint32_t nb, bool v_is_t) {
  uint32_t sub = hash_lookup(K.hk, K.nodes, K.n, K.hmask, K.hshift, nb);
  if (sub != NONE32 && v_is_t) { bool nb_t = false; for (int j = 0; j < K.nt; j++) nb_t |= (K.roots[j] == nb); if (nb_t) sub = NONE32; }   // PS.cpp:412-418
  return sub;
}

// On entry ws.rowd[r] = {row start, row length, node id, first item}, ws.rowd[n].w = num_items.
// Returns the number of entries this warp staged (entries beyond its region are counted, not stored).
template <bool TCONN>
__device__ inline uint32_t scan_items_staged(const SampleParams &P, const Ws &ws, int n, uint32_t num_items, const KeepCtx &K) {
  const int lane = lane_id(), warp = warp_id(), nwarp = blockDim.x >> 5;
  const uint32_t rcap = (uint32_t)P.ecap / nwarp;              // staging region of this warp
  uint2 *st = ws.st + (size_t)warp * rcap;
  unsigned short *strow = ws.st_row + (size_t)warp * rcap;
  const uint32_t i_begin = (uint32_t)(((unsigned long long)num_items * warp) / nwarp);
  const uint32_t i_end = (uint32_t)(((unsigned long long)num_items * (warp + 1)) / nwarp);
  if (i_begin >= i_end) return 0;
  int row;
  { int lo = 0, hi = n; while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ws.rowd[mid].w <= i_begin) lo = mid; else hi = mid; } row = lo; }
  const uint4 *hb = reinterpret_cast<const uint4 *>(K.hk);
  uint32_t cnt = 0;
  for (uint32_t ib = i_begin; ib < i_end; ib += SCAN_U) {
    // lanes 0..SCAN_U-1 resolve one item each (row, first slot, slots left in the row), then broadcast
    int r = row;
    uint32_t cbase = 0, rem = 0;
    if (lane < SCAN_U && ib + lane < i_end) {
      const uint32_t it = ib + lane;
      while (it >= ws.rowd[r + 1].w) r++;
      const uint4 d = ws.rowd[r];
      const uint32_t j0 = (it - d.w) * 32u;
      cbase = d.x + j0; rem = d.y - j0;
    }
    row = __shfl_sync(0xffffffffu, r, (int)min((uint32_t)SCAN_U, i_end - ib) - 1);
    uint32_t nbv[SCAN_U];
#pragma unroll
    for (int u = 0; u < SCAN_U; u++) {
      const uint32_t rm = __shfl_sync(0xffffffffu, rem, u);
      const uint32_t c = __shfl_sync(0xffffffffu, cbase, u) + lane;
      nbv[u] = ((uint32_t)lane < rm) ? ldg_stream_u32(P.indices + c) : NONE32;
    }
#pragma unroll
    for (int u = 0; u < SCAN_U; u++) {
      const uint32_t nb = nbv[u];
      uint32_t b = hash_bucket(nb, K.hshift);
      uint4 kk = hb[b];
      bool hit = (kk.x == nb) || (kk.y == nb) || (kk.z == nb) || (kk.w == nb);
      while (!hit && kk.w != NONE32) {                         // bucket full and no match: next bucket (rare)
        b = (b + 1) & K.hmask; kk = hb[b];
        hit = (kk.x == nb) || (kk.y == nb) || (kk.z == nb) || (kk.w == nb);
      }
      hit = hit && (nb != NONE32);
      const uint32_t mk = __ballot_sync(0xffffffffu, hit);
      if (mk) {                                                // everything below is the rare path (~3 % of the slots are kept)
        const uint32_t rr = (uint32_t)__shfl_sync(0xffffffffu, r, u);
        const uint32_t c = __shfl_sync(0xffffffffu, cbase, u) + lane;
        if (!TCONN && hit) {                                   // multi-root groups: no target-target edges (PS.cpp:412-418)
          const uint32_t v = ws.nodes[rr];
          bool v_t = false, nb_t = false;
          for (int q = 0; q < K.nt; q++) { v_t |= (K.roots[q] == v); nb_t |= (K.roots[q] == nb); }
          hit = !(v_t && nb_t);
        }
        const uint32_t mk2 = TCONN ? mk : __ballot_sync(0xffffffffu, hit);
        if (hit) {                                             // stage (neighbour id, slot) + row; sub id and order flags are resolved at emit time
          const uint32_t at = cnt + __popc(mk2 & lanemask_lt());
          if (at < rcap) { st[at] = make_uint2(nb, c); strow[at] = (unsigned short)rr; }                          // :420-422
        }
        cnt += __popc(mk2);
      }
    }
  }
  return cnt;
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
template <bool GWS, bool REDO = false>
__global__ void __launch_bounds__(SAMPLER_BLOCK, SAMPLER_MIN_BLOCKS) sample_induce_kernel(const SampleParams P) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __shared__ uint32_t s_warp_sums[33];
  __shared__ int s_p, s_n;
  __shared__ uint32_t s_cnt, s_cut, s_scan[32];
  __shared__ long long s_base[2];
  __shared__ uint32_t s_roots[SHADOW_MAX_ROOTS], s_tl[SHADOW_MAX_ROOTS];

  unsigned char *wsb = GWS ? (P.gws + (size_t)blockIdx.x * P.gws_stride) : smem_dyn;
  const Ws ws = make_ws(wsb, P.L);
  const int lane = lane_id(), warp = warp_id(), nwarp = blockDim.x >> 5;
  const uint32_t hmask = ((uint32_t)P.hcap >> 2) - 1u;      // bucket mask

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      if (REDO) {                                            // redo launch: only the subgraphs the warp fast path handed over
        const uint32_t tk = atomicAdd(P.ticket2, 1u);
        s_p = tk < *P.redo_count ? P.redo_list[tk] : P.num_subg;
      } else s_p = (int)atomicAdd(P.ticket, 1u);
    }
    __syncthreads();
    const int p = s_p;
    if (p >= P.num_subg) break;
    const int r0 = p * P.num_roots;
    const int nt = min(P.num_roots, P.num_root_ids - r0);
    if (threadIdx.x < nt) s_roots[threadIdx.x] = P.roots[r0 + threadIdx.x];
    __syncthreads();

    // ---------------- A: node set ----------------
    int n = 0;
    long long draws = 0;
    if (P.method == SHADOW_PPR) {
      if (nt == 1 && P.ppr_sid) n = build_nodes_ppr_sorted(P, ws, s_roots[0], &s_cut, s_scan, s_warp_sums);
      else n = build_nodes_ppr(P, ws, s_roots, nt, &s_cut, s_warp_sums);
    } else if (P.method == SHADOW_PPR_ST) {
      long long rb = (P.rng_mode == SHADOW_RNG_GLIBC && P.rand_off) ? P.rand_off[p] : 0;
      n = build_nodes_ppr_st(P, ws, s_roots, nt, p, rb, &draws, &s_cut, s_scan, s_warp_sums);
    } else if (P.method == SHADOW_KHOP) {
      long long rb = (P.rng_mode == SHADOW_RNG_GLIBC && P.rand_off) ? P.rand_off[p] : 0;
      n = build_nodes_khop(P, ws, s_roots, nt, p, rb, &draws, s_warp_sums, &s_n);
    } else {   // nodeIID (PS.cpp:498-508)
      if (threadIdx.x == 0) {
        int m = 0;
        for (int i = 0; i < nt; i++) {
          uint32_t v = s_roots[i]; int j = 0;
          while (j < m && ws.nodes[j] < v) j++;
          if (j < m && ws.nodes[j] == v) continue;
          for (int q = m; q > j; q--) ws.nodes[q] = ws.nodes[q - 1];
          ws.nodes[j] = v; m++;
        }
        for (int i = 0; i < m; i++) ws.pprv[i] = -1.f;
        s_n = m;
      }
      __syncthreads();
      n = s_n;
    }
    bool ws_overflow = (n < 0) || (n > P.ncap);
    if (ws_overflow) n = 0;
    if (!REDO && threadIdx.x == 0) lookback_publish(P.status_n, p, (unsigned long long)n);
    __syncthreads();

    // ---------------- B: orig -> sub hash, targets, row extents ----------------
    for (int i = threadIdx.x; i < P.hcap; i += blockDim.x) ws.hkeys[i] = NONE32;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t v = ws.nodes[i];
      hash_insert(ws.hkeys, hmask, P.hshift, v);
      const uint32_t s = P.indptr[v], e = P.indptr[v + 1];
      ws.row_s[i] = s; ws.row_e[i] = e;
      if (!GWS && P.ecap > 0) {
        ws.row_cnt[i] = (e - s + 31u) >> 5;                  // items of 32 slots
        ws.row_first[i] = 0; ws.row_last[i] = 0; ws.row_fgt[i] = NONE32; ws.row_ins[i] = 0;
      }
    }
    __syncthreads();
    if (threadIdx.x < nt) {
      uint32_t t = hash_lookup(ws.hkeys, ws.nodes, n, hmask, P.hshift, s_roots[threadIdx.x]);
      s_tl[threadIdx.x] = (t == NONE32) ? 0u : t;          // operator[] default-inserts 0 (PS.cpp:375)
    }
    const bool tconn = P.tconn || nt == 1;                 // PS.cpp:356-358
    const bool add_self = P.add_self != 0;

    // ---------------- C: row scan ----------------
    bool staged = false;
    if (!GWS && P.ecap > 0) {                              // single pass, kept edges staged in shared memory
      const uint32_t num_items = block_exclusive_scan(ws.row_cnt, n, s_warp_sums);
      for (int i = threadIdx.x; i <= n; i += blockDim.x)
        ws.rowd[i] = (i < n) ? make_uint4(ws.row_s[i], ws.row_e[i] - ws.row_s[i], ws.nodes[i], ws.row_cnt[i]) : make_uint4(0, 0, 0, num_items);
      __syncthreads();
      KeepCtx KC;
      KC.hk = ws.hkeys; KC.nodes = ws.nodes; KC.n = n; KC.hmask = hmask; KC.hshift = P.hshift; KC.roots = s_roots; KC.nt = nt;
      const uint32_t wcnt = tconn ? scan_items_staged<true>(P, ws, n, num_items, KC) : scan_items_staged<false>(P, ws, n, num_items, KC);
      const uint32_t rcap = (uint32_t)P.ecap / nwarp;
      if (lane == 0) s_scan[warp] = wcnt;
      __syncthreads();
      uint32_t ktot = 0;
      staged = true;
      for (int w = 0; w < nwarp; w++) { ktot += s_scan[w]; staged &= (s_scan[w] <= rcap); }
      if (staged) {
        // row boundaries inside the staged stream (it is in CSR order): first/last entry of each row, first entry above v,
        // and whether the row already holds its self loop (a kept entry whose sub id is the row itself)
        for (uint32_t g = threadIdx.x; g < ktot; g += blockDim.x) {
          uint32_t wb = 0; int w = 0;
          while (g >= wb + s_scan[w]) { wb += s_scan[w]; w++; }
          const uint32_t nb = ws.st[(size_t)w * rcap + (g - wb)].x, row = ws.st_row[(size_t)w * rcap + (g - wb)];
          uint32_t rowp = NONE32, rown = NONE32, nbp = 0;
          if (g > 0) { const uint32_t gp = g - 1; uint32_t wb2 = 0; int w2 = 0; while (gp >= wb2 + s_scan[w2]) { wb2 += s_scan[w2]; w2++; } rowp = ws.st_row[(size_t)w2 * rcap + (gp - wb2)]; nbp = ws.st[(size_t)w2 * rcap + (gp - wb2)].x; }
          if (g + 1 < ktot) { const uint32_t gn = g + 1; uint32_t wb2 = 0; int w2 = 0; while (gn >= wb2 + s_scan[w2]) { wb2 += s_scan[w2]; w2++; } rown = ws.st_row[(size_t)w2 * rcap + (gn - wb2)]; }
          const uint32_t v = ws.nodes[row];
          const bool first = rowp != row, last = rown != row, gt = nb > v;
          if (first) ws.row_first[row] = g;
          if (last) ws.row_last[row] = g + 1;
          if (gt && (first || !(nbp > v))) ws.row_fgt[row] = g;
          if (nb == v) ws.row_ins[row] = 1;                                                           // self loop present
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const uint32_t v = ws.nodes[i], e = ws.row_e[i];
          const uint32_t kept = ws.row_last[i] - ws.row_first[i];
          bool v_is_t = false;
          if (!tconn) for (int j = 0; j < nt; j++) v_is_t |= (s_roots[j] == v);
          bool present = ws.row_ins[i] != 0;
          if (add_self && v_is_t && !present) {             // a dropped target-target self loop still counts as present: binary search (:386-400)
            uint32_t lo = ws.row_s[i], hi = e;
            while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (P.indices[mid] < v) lo = mid + 1; else hi = mid; }
            present = (lo < e) && (P.indices[lo] == v);
          }
          const bool inserting = add_self && !present;                                             // PS.cpp:386-400
          uint32_t bsub = NONE32;                          // PS.cpp:401: without an insertion, slot e (first slot of the next row) is tested too
          if (!inserting && !P.fixed_mode && e < P.num_edges) bsub = keep_lookup(KC, __ldg(P.indices + e), v_is_t);
          const uint32_t fgt = ws.row_fgt[i];
          const uint32_t less = (fgt == NONE32) ? kept : fgt - ws.row_first[i];                     // kept entries below v
          ws.row_fgt[i] = bsub;                              // reuse: bug-slot column (or NONE)
          ws.row_ins[i] = inserting ? less : NONE32;
          ws.row_cnt[i] = kept + (inserting ? 1u : 0u) + (bsub != NONE32 ? 1u : 0u);
        }
      }
      __syncthreads();
    }
    if (!staged) {
    // count pass of the generic two-pass path (warp per row)
    for (int r = warp; r < n; r += nwarp) {
      const uint32_t v = ws.nodes[r], s = ws.row_s[r], e = ws.row_e[r];
      bool v_is_t = false;
      if (!tconn) for (int j = 0; j < nt; j++) v_is_t |= (s_roots[j] == v);
      uint32_t kept = 0, kept_less = 0;
      bool present = false;
      for (uint32_t base = s; base < e; base += 32) {
        const uint32_t c = base + lane;
        bool keep = false, less = false, self = false;
        if (c < e) {
          const uint32_t nb = ldg_stream_u32(P.indices + c);
          keep = hash_lookup(ws.hkeys, ws.nodes, n, hmask, P.hshift, nb) != NONE32;
          if (keep && v_is_t) { bool nb_t = false; for (int j = 0; j < nt; j++) nb_t |= (s_roots[j] == nb); keep = !nb_t; }   // :412-418
          less = keep && nb < v; self = (nb == v);
        }
        kept += __popc(__ballot_sync(0xffffffffu, keep));
        if (add_self) { kept_less += __popc(__ballot_sync(0xffffffffu, less)); present |= __any_sync(0xffffffffu, self); }
      }
      const bool inserting = add_self && !present;         // :386-400
      if (!inserting && !P.fixed_mode && e < P.num_edges) {
        // PS.cpp:401: the row bound is idx_end + 1 even without an insertion => slot `e` (first slot of the next row) is tested too
        const uint32_t nb = P.indices[e];
        bool keep = hash_lookup(ws.hkeys, ws.nodes, n, hmask, P.hshift, nb) != NONE32;
        if (keep && v_is_t) { bool nb_t = false; for (int j = 0; j < nt; j++) nb_t |= (s_roots[j] == nb); keep = !nb_t; }
        kept += keep ? 1u : 0u;
      }
      if (lane == 0) { ws.row_cnt[r] = kept + (inserting ? 1u : 0u); ws.row_ins[r] = inserting ? kept_less : NONE32; }
    }
    }
    __syncthreads();
    const uint32_t m = block_exclusive_scan(ws.row_cnt, n, s_warp_sums);      // local indptr (:428-431)
    if (threadIdx.x == 0) ws.row_cnt[n] = m;

    // ---------------- look-back: batch-global offsets ----------------
    // rows stay in subgraph order (look-back over n, resolved long after it was published); the edge block of a
    // subgraph is placed wherever the cursor stands -- no CTA ever waits for another one's row scan
    if (warp == 0) {
      // redo launch: the fast path already resolved the look-back and reserved the rows (node_ptr[p])
      unsigned long long nb = REDO ? (unsigned long long)P.node_ptr[p] : lookback_resolve(P.status_n, p, (unsigned long long)n);
      if (lane == 0) { s_base[0] = (long long)nb; s_base[1] = (long long)atomicAdd((unsigned long long *)&P.totals[1], (unsigned long long)m); }
    }
    __syncthreads();
    const long long node_base = s_base[0], edge_base = s_base[1];
    const bool out_overflow = (node_base + n > P.cap_nodes) || (edge_base + (long long)m > P.cap_edges);
    if (threadIdx.x == 0) {
      if (ws_overflow) atomicOr((unsigned long long *)&P.totals[2], (unsigned long long)ERR_WS_OVERFLOW);
      if (out_overflow) atomicOr((unsigned long long *)&P.totals[2], (unsigned long long)ERR_OUT_OVERFLOW);
      if (p == P.num_subg - 1) P.totals[0] = node_base + n;
    }
    if (!out_overflow) {
      if (threadIdx.x == 0) {
        P.node_ptr[p] = (int)node_base; P.edge_span[p] = make_int2((int)edge_base, (int)(edge_base + m)); P.num_target[p] = nt;
        if (p == P.num_subg - 1) P.node_ptr[p + 1] = (int)(node_base + n);
      }
      if (threadIdx.x < P.num_roots) P.target[r0 + threadIdx.x] = (threadIdx.x < nt) ? (int)(node_base + s_tl[threadIdx.x]) : -1;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        P.orig_node[node_base + i] = ws.nodes[i];
        P.ppr_out[node_base + i] = ws.pprv[i];
        P.row_span[node_base + i] = make_int2((int)(edge_base + ws.row_cnt[i]), (int)(edge_base + ws.row_cnt[i + 1]));
      }
      // ---------------- emit the CSR ----------------
      if (staged) {                                        // staged rows -> CSR order already; no second read of the graph
        const uint32_t rcap = (uint32_t)P.ecap / nwarp;
        uint32_t ktot = 0;
        for (int w = 0; w < nwarp; w++) ktot += s_scan[w];
        for (uint32_t g = threadIdx.x; g < ktot; g += blockDim.x) {
          uint32_t wb = 0; int w = 0;
          while (g >= wb + s_scan[w]) { wb += s_scan[w]; w++; }
          const uint2 ent = ws.st[(size_t)w * rcap + (g - wb)];
          const uint32_t row = ws.st_row[(size_t)w * rcap + (g - wb)];
          const bool gt = ent.x > ws.nodes[row];
          const long long pos = edge_base + ws.row_cnt[row] + (g - ws.row_first[row]) + ((ws.row_ins[row] != NONE32 && gt) ? 1 : 0);
          P.indices_out[pos] = (int)(node_base + sub_of(ws.nodes, n, ent.x));
          P.orig_edge[pos] = ent.y;
        }
        for (int r = threadIdx.x; r < n; r += blockDim.x) {
          if (ws.row_ins[r] != NONE32) {                                                           // PS.cpp:406-411
            const long long pos = edge_base + ws.row_cnt[r] + ws.row_ins[r];
            P.indices_out[pos] = (int)(node_base + r); P.orig_edge[pos] = NONE32;
          }
          if (ws.row_fgt[r] != NONE32) {
            const long long pos = edge_base + ws.row_cnt[r + 1] - 1;
            P.indices_out[pos] = (int)(node_base + ws.row_fgt[r]); P.orig_edge[pos] = ws.row_e[r];
          }
        }
      } else
      // fill pass of the generic path (rows come back from L2)
      for (int r = warp; r < n; r += nwarp) {
        const uint32_t v = ws.nodes[r], s = ws.row_s[r], e = ws.row_e[r];
        bool v_is_t = false;
        if (!tconn) for (int j = 0; j < nt; j++) v_is_t |= (s_roots[j] == v);
        const uint32_t ins = ws.row_ins[r];
        const bool inserting = ins != NONE32;
        const long long obase = edge_base + ws.row_cnt[r];
        uint32_t kept = 0;
        for (uint32_t base = s; base < e; base += 32) {
          const uint32_t c = base + lane;
          uint32_t sub = NONE32, nb = 0;
          if (c < e) {
            nb = P.indices[c];
            sub = hash_lookup(ws.hkeys, ws.nodes, n, hmask, P.hshift, nb);
            if (sub != NONE32 && v_is_t) { bool nb_t = false; for (int j = 0; j < nt; j++) nb_t |= (s_roots[j] == nb); if (nb_t) sub = NONE32; }
          }
          const uint32_t mask = __ballot_sync(0xffffffffu, sub != NONE32);
          if (sub != NONE32) {
            const uint32_t pos = kept + __popc(mask & lanemask_lt()) + ((inserting && nb > v) ? 1u : 0u);
            P.indices_out[obase + pos] = (int)(node_base + sub);
            P.orig_edge[obase + pos] = c;                                                        // :422
          }
          kept += __popc(mask);
        }
        if (inserting) {
          if (lane == 0) { P.indices_out[obase + ins] = (int)(node_base + r); P.orig_edge[obase + ins] = NONE32; }   // :406-411
        } else if (!P.fixed_mode && e < P.num_edges) {
          const uint32_t nb = P.indices[e];
          uint32_t sub = hash_lookup(ws.hkeys, ws.nodes, n, hmask, P.hshift, nb);
          if (sub != NONE32 && v_is_t) { bool nb_t = false; for (int j = 0; j < nt; j++) nb_t |= (s_roots[j] == nb); if (nb_t) sub = NONE32; }
          if (sub != NONE32 && lane == 0) { P.indices_out[obase + kept] = (int)(node_base + sub); P.orig_edge[obase + kept] = e; }
        }
      }
      // ---------------- D: hop / drnl labels ----------------
      if ((P.aug & (SHADOW_AUG_HOPS | SHADOW_AUG_DRNLS)) && n > 0) {
        __syncthreads();       // this CTA's CSR writes are visible to the whole CTA
        if (P.aug & SHADOW_AUG_DRNLS) {                                                          // PS.cpp:438-451
          subgraph_bfs(P, ws, n, s_tl[0], node_base, ws.row_cnt, edge_base, ws.dist, &s_cnt);
          for (int i = threadIdx.x; i < n; i += blockDim.x) ws.pprv[i] = __uint_as_float(ws.dist[i]);   // dx parked in pprv (already written out)
          __syncthreads();
          subgraph_bfs(P, ws, n, s_tl[nt > 1 ? 1 : 0], node_base, ws.row_cnt, edge_base, ws.dist, &s_cnt);
          for (int i = threadIdx.x; i < n; i += blockDim.x)
            P.drnl[node_base + i] = drnl_single(__float_as_uint(ws.pprv[i]), ws.dist[i]);
        } else {                                                                                 // PS.cpp:433-436
          subgraph_bfs(P, ws, n, s_tl[0], node_base, ws.row_cnt, edge_base, ws.dist, &s_cnt);
          for (int i = threadIdx.x; i < n; i += blockDim.x) P.hop[node_base + i] = ws.dist[i];
        }
      }
    }
  }
}

// ppr_st glibc replay: the draw count of a subgraph is the total stored length of its target rows -> offsets by a serial scan
__global__ void ppr_st_rand_offsets_kernel(const SampleParams P) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long cursor = 0;
  for (int p = 0; p < P.num_subg; p++) {
    P.rand_off[p] = cursor;
    const int r0 = p * P.num_roots, nt = min(P.num_roots, P.num_root_ids - r0);
    for (int j = 0; j < nt; j++) { const uint32_t t = P.roots[r0 + j]; cursor += (long long)(P.ppr_ptr[t + 1] - P.ppr_ptr[t]); }
  }
  P.rand_off[P.num_subg] = cursor;
}

// glibc-replay prepass: ONE CTA walks the subgraphs in order and fixes the rand() offset of each
// (the stream position of subgraph p+1 depends on how many high-degree nodes subgraph p met)
__global__ void __launch_bounds__(SAMPLER_BLOCK) khop_rand_offsets_kernel(const SampleParams P) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  __shared__ uint32_t s_warp_sums[33];
  __shared__ int s_n;
  __shared__ uint32_t s_roots[SHADOW_MAX_ROOTS];
  unsigned char *wsb = P.gws ? P.gws : smem_dyn;
  const Ws ws = make_ws(wsb, P.L);
  long long cursor = 0;
  for (int p = 0; p < P.num_subg; p++) {
    __syncthreads();
    const int r0 = p * P.num_roots;
    const int nt = min(P.num_roots, P.num_root_ids - r0);
    if (threadIdx.x < nt) s_roots[threadIdx.x] = P.roots[r0 + threadIdx.x];
    if (threadIdx.x == 0) P.rand_off[p] = cursor;
    __syncthreads();
    long long draws = 0;
    int n = build_nodes_khop(P, ws, s_roots, nt, p, cursor, &draws, s_warp_sums, &s_n);
    if (n < 0) { if (threadIdx.x == 0) atomicOr((unsigned long long *)&P.totals[2], (unsigned long long)ERR_WS_OVERFLOW); }
    cursor += draws;
  }
  if (threadIdx.x == 0) P.rand_off[P.num_subg] = cursor;
}
