// ppr_warp_kernel.cuh -- the single-root PPR fast path of the fused select + induce step: ONE WARP per subgraph.
//
// Same contract as sample_induce_kernel (sampler_kernels.cuh) for  method == ppr, one root per subgraph, with or without hop labels
// (no drnl labels)  -- the configuration every PPR node-task run of the reference uses (shaDow/minibatch.py:373-375; PS.cpp:565-595
// + PS.cpp:350-453; `feature_augment: hops` in 46 of its 61 configs).  What is different is the mapping to the machine:
//   * a warp owns a subgraph end to end, so there is no CTA barrier and no cross-warp scan anywhere;
//   * the node set comes from the id-sorted PPR row, which also carries (row start, degree) of every neighbour -- captured once
//     when the tables are installed -- so the random 8-byte indptr reads (a 64-byte DRAM fetch each) disappear;
//   * the row scan runs over a FLAT space of chunks: all rows of the subgraph are cut into 16-byte aligned chunks of 4 slots, lane L of
//     window w takes chunk 32w+L whatever row it belongs to (rows are resolved 32 chunks at a time with one REDUX) and loads it with one
//     128-bit streaming load: short rows do not idle lanes;
//   * MEMBERSHIP IS A TWO-LEVEL TEST (round 2).  Level 1, once per scanned slot, is a 1024-bit blocked Bloom filter that lives in ONE
//     REGISTER per lane (lane w holds word w): the word is fetched with a single SHFL indexed by the slot's hash and two independent
//     bits of it are tested.  No shared-memory bank is touched and nothing is compacted per slot: a lane only ORs the verdicts of its 4
//     slots, and ONE ballot per window of 128 slots appends the (~30 %) chunks that may hold a member -- keys + packed (row, offset) --
//     to a small shared-memory queue.  Level 2 runs on full warps of 32 queued chunks: the exact 2-key-bucket probe, the row-range check
//     (alignment padding of a chunk belongs to the neighbouring rows) and the ordered append of round 1, now on a third of the chunks.
//     The round-1 kernel ran the exact probe (~6 shared-memory wavefronts, ~25 instructions) for every scanned slot and was bound by
//     issue slots and the shared-memory pipe (69 % each);
//   * with no self-edge insertion the PS.cpp:401 slot (one past the row) is simply one more slot of the row, and the staged stream
//     IS the CSR: emit is a straight copy with the sub id resolved per kept edge;
//   * kept edges are staged in a per-warp scratch region in GLOBAL memory that the warp reuses for every subgraph (it stays in L2), per-row
//     counts come from one shared-memory atomic per lane;
//   * the node count of a subgraph is a function of its table row alone, so a small count kernel + a one-block scan fix node_ptr[] before
//     this kernel starts: no decoupled look-back.
// A subgraph that does not fit the per-warp scratch (or whose hash overflow list is full) is not an error: its index goes on a
// redo list and the generic CTA kernel, launched right behind in redo mode, builds it into the rows this kernel reserved.
#pragma once
#include "sampler_kernels.cuh"

// tuning macros (measured on the S-products stand-in, scripts/explore_variants.py; defaults = the fastest build)
#ifndef WARP_U
#define WARP_U 6               // chunk windows in flight per warp (32 chunks each)
#endif
#ifndef WARP_U_SYM
#define WARP_U_SYM 4           // the same for the symmetric variant: half the slots per subgraph, shorter stages win (scripts/explore_variants.py)
#endif
#define WARP_U_MAX (WARP_U > WARP_U_SYM ? WARP_U : WARP_U_SYM)
#ifndef WARP_MIN_BLOCKS
#define WARP_MIN_BLOCKS 24     // one warp per CTA: resident warps per SM the register budget must allow
#endif
#ifndef WARP_DB
#define WARP_DB 0              // 1: the loads of stage i+1 are issued before stage i is processed (two stages of registers)
#endif
#define WARP_CS 4              // slots per chunk (one 128-bit load)
#define WARP_CSH 2
#define WARP_QCAP 64           // chunk queue: < 32 left over + at most 32 new per window
#ifndef WARP_OVF_CAP
#define WARP_OVF_CAP 16        // keys whose 2-key bucket was full
#endif

__device__ __forceinline__ uint32_t lanemask_le() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t x, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  return x;
}

// ---- level 1: blocked Bloom filter, NF words per lane (32 * NF words = 1024 * NF bits), one SHFL per word-register, 2 bits per key ----
__device__ __forceinline__ uint32_t bloom_hash(uint32_t key) { return key * 2654435761u; }
__device__ __forceinline__ uint32_t bloom_hash2(uint32_t key) { return __umulhi(key, 2654435761u); }     // high half of the same product: one IMAD.WIDE gives both
template <int NF>
__device__ __forceinline__ uint32_t bloom_word_index(uint32_t h) { return h >> (NF == 1 ? 27 : (NF == 2 ? 26 : 25)); }
__device__ __forceinline__ uint32_t bloom_bits(uint32_t h, uint32_t g) { return __funnelshift_l(0u, 1u, h) | __funnelshift_l(0u, 1u, g); }   // 1 << h % 32 | 1 << g % 32
// bit 0 of the result: both bits of `key` are set in its filter word
template <int NF>
__device__ __forceinline__ uint32_t bloom_test(const uint32_t (&f)[NF], uint32_t key) {
  const unsigned long long pr = (unsigned long long)key * 2654435761ull;
  const uint32_t h = (uint32_t)pr, g = (uint32_t)(pr >> 32), w = bloom_word_index<NF>(h);
  uint32_t fv = __shfl_sync(0xffffffffu, f[0], (int)w);       // source lane = w mod 32
  if (NF >= 2) { const uint32_t f1 = __shfl_sync(0xffffffffu, f[NF >= 2 ? 1 : 0], (int)w); fv = (w & 32u) ? f1 : fv; }
  if (NF == 4) {
    const uint32_t f2 = __shfl_sync(0xffffffffu, f[NF == 4 ? 2 : 0], (int)w), f3 = __shfl_sync(0xffffffffu, f[NF == 4 ? 3 : 0], (int)w);
    fv = (w & 64u) ? ((w & 32u) ? f3 : f2) : fv;
  }
  return __funnelshift_r(fv, 0u, h) & __funnelshift_r(fv, 0u, g);        // (fv >> h % 32) & (fv >> g % 32)
}

// ---- level 2: exact membership in buckets of 2 keys (one 8-byte shared-memory load); a key whose bucket was full sits in a short overflow list ----
__device__ __forceinline__ bool probe2(const uint2 *hb, const int hshift, const uint32_t *ovl, const uint32_t novf, const uint32_t key) {
  const uint2 kk = hb[bloom_hash(key) >> hshift];
  bool hit = (kk.x == key) | (kk.y == key);
#pragma unroll 1
  for (uint32_t j = 0; j < novf; j++) hit |= ovl[j] == key;
  return hit;
}

// ---- level 2 for large node sets (BIS): no hash table at all.  The 2-key-bucket table needs 3 buckets per key rounded up to a power of two:
// 16 KB of shared memory per warp at k = 400 (papers100M), i.e. 8 resident warps per SM instead of 23.  The sorted node list is already in
// shared memory; a branch-free lower bound over it (9 steps at n = 401, the 4 keys of a chunk interleaved) decides membership exactly, and
// only the chunks that passed the Bloom filter ever get here.
__device__ __forceinline__ void bisect4(const uint32_t *nodes, const int n, const uint32_t (&key)[4], bool (&hit)[4]) {
  uint32_t base[4] = {0u, 0u, 0u, 0u};
  int len = n;
  while (len > 1) {
    const int half = len >> 1;
#pragma unroll
    for (int e = 0; e < 4; e++) base[e] += nodes[base[e] + half - 1] < key[e] ? (uint32_t)half : 0u;
    len -= half;
  }
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const uint32_t a = nodes[base[e]];
    const uint32_t lb = base[e] + (a < key[e] ? 1u : 0u);
    hit[e] = lb < (uint32_t)n && (a == key[e] || (a < key[e] && nodes[min(lb, (uint32_t)n - 1u)] == key[e]));
  }
}
__device__ __forceinline__ bool bisect1(const uint32_t *nodes, const int n, const uint32_t key) {
  const uint32_t lb = sub_of(nodes, n, key);
  return lb < (uint32_t)n && nodes[lb] == key;
}

#define WARP_OFFBITS 19                                // candidate code = row << 19 | (slot - row start + 3): rows < 8192, scanned length < 2^19 - 8
#define WARP_OFFMASK ((1u << WARP_OFFBITS) - 1u)
template <int U>
struct ScanStage {
  uint4 q[U];
  uint32_t code[U];                             // code of the first slot of my chunk (WARP_OFFMASK for an idle lane: offset beyond any row)
};

// The aligned 16-byte chunk that holds the last slot of the array may reach up to 12 bytes past indices[E-1]: still inside the allocation
// (CUDA allocations are 256-byte granular and the array starts 16-byte aligned); those slots are >= E, outside every row range.
template <int U>
__device__ __forceinline__ void scan_load(ScanStage<U> &S, const uint32_t c0, const uint32_t total, int &rbase, const int n, const int lane,
                                          const uint32_t le, const uint32_t *cp, const uint2 *rs, const uint4 *ind4) {
#pragma unroll
  for (int u = 0; u < U; u++) {
    const uint32_t cw = c0 + 32u * u;
    S.q[u] = make_uint4(NONE32, NONE32, NONE32, NONE32);
    S.code[u] = WARP_OFFMASK;
    if (cw < total) {
      // rows that start inside this window: lane j looks at row rbase+1+j, one REDUX turns the starts into a bit mask
      const uint32_t rel = cp[min(rbase + 1 + lane, n)] - cw;
      const uint32_t mask = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
      const int row = min(rbase + __popc(mask & le), n - 1);
      rbase += __popc(mask);
      const uint32_t c = cw + lane;
      if (c < total) {
        const uint32_t rx = rs[row].x, cc = c - cp[row];                         // my chunk is the cc-th aligned chunk of its row
        S.code[u] = ((uint32_t)row << WARP_OFFBITS) | ((cc << WARP_CSH) + 3u - (rx & 3u));
        S.q[u] = ldg_stream_u4(ind4 + (rx >> WARP_CSH) + cc);
      }
    }
  }
}

// Level 2 on the queued chunks [first .. first+count) (count <= 32): exact probe + row-range check per slot, kept edges appended to the
// staged stream (a per-warp scratch region in global memory that lives in L2).  Stream order = queue order (window, lane), then slot
// = ascending full-graph slot = CSR order (PS.cpp:420-422).  Per-row facts (kept count; with a self-edge insertion also "kept entries
// below v" and "v itself kept") are accumulated with one shared-memory atomic per lane.
template <bool ADD_SELF, bool SYM, bool BIS>
__device__ __forceinline__ void drain_batch(const uint4 *cq_keys, const uint32_t *cq_code, const uint32_t first, const uint32_t count,
                                            const uint2 *hb, const int hshift, const uint32_t *ovl, const uint32_t novf, const uint32_t *nodes, const int n,
                                            const uint2 *rs, uint32_t *rc, uint2 *sc_ent, unsigned short *sc_row, uint32_t &cnt, const int lane,
                                            const uint32_t lt) {
  const bool active = (uint32_t)lane < count;
  uint4 q = make_uint4(NONE32, NONE32, NONE32, NONE32);
  uint32_t code = WARP_OFFMASK;
  if (active) { q = cq_keys[first + lane]; code = cq_code[first + lane]; }
  const uint32_t row = code >> WARP_OFFBITS, off0 = (code & WARP_OFFMASK) - 3u;
  uint2 r = rs[row];
  if (SYM) r.y &= 0x7fffffffu;                         // bit 31 marks a row whose PS.cpp:401 slot was appended (resolve pass)
  const uint32_t nb[WARP_CS] = {q.x, q.y, q.z, q.w};
  bool hit[WARP_CS];
  uint32_t bal[WARP_CS], at = cnt, tot = 0, inc = 0;
  uint32_t v = 0;
  if (ADD_SELF) v = nodes[row];
  if (BIS) bisect4(nodes, n, nb, hit);
  else {
#pragma unroll
    for (int e = 0; e < WARP_CS; e++) { const uint2 kk = hb[bloom_hash(nb[e]) >> hshift]; hit[e] = (kk.x == nb[e]) | (kk.y == nb[e]); }
#pragma unroll 1
    for (uint32_t j = 0; j < novf; j++) {              // keys that did not fit their bucket (1-2 per subgraph)
      const uint32_t kv = ovl[j];
#pragma unroll
      for (int e = 0; e < WARP_CS; e++) hit[e] |= kv == nb[e];
    }
  }
#pragma unroll
  for (int e = 0; e < WARP_CS; e++) {
    hit[e] = hit[e] && (off0 + (uint32_t)e) < r.y;    // member of the node set, slot inside [s, s + len)
    bal[e] = __ballot_sync(0xffffffffu, hit[e]);
    at += __popc(bal[e] & lt); tot += __popc(bal[e]);
    if (ADD_SELF) inc += hit[e] ? (1u + ((!SYM && nb[e] < v) ? (1u << 14) : 0u) + (nb[e] == v ? (1u << 28) : 0u)) : 0u;
    else inc += hit[e] ? 1u : 0u;
  }
#pragma unroll
  for (int e = 0; e < WARP_CS; e++) {
    if (hit[e]) { sc_ent[at] = make_uint2(nb[e], r.x + off0 + (uint32_t)e); if (ADD_SELF || SYM) sc_row[at] = (unsigned short)row; }
    at += hit[e] ? 1u : 0u;
  }
  if (inc) atomicAdd(&rc[row], inc);
  cnt += tot;
}

// Level 1 on one window: a lane ORs the Bloom verdicts of its 4 slots; ONE ballot appends the chunks that may hold a member to the queue,
// a full warp of queued chunks is drained at once.
template <bool ADD_SELF, int NF, bool SYM, bool BIS>
__device__ __forceinline__ void scan_window(const uint32_t any, const uint4 q, const uint32_t code, uint4 *cq_keys, uint32_t *cq_code,
                                            uint32_t &qn, const uint2 *hb, const int hshift, const uint32_t *ovl, const uint32_t novf,
                                            const uint32_t *nodes, const int n, const uint2 *rs, uint32_t *rc, uint2 *sc_ent, unsigned short *sc_row,
                                            uint32_t &cnt, const int lane, const uint32_t lt) {
  const uint32_t bal = __ballot_sync(0xffffffffu, any != 0u);
  if (any) { const uint32_t at = qn + __popc(bal & lt); cq_keys[at] = q; cq_code[at] = code; }
  qn += __popc(bal);
  if (qn >= 32u) {
    __syncwarp();
    drain_batch<ADD_SELF, SYM, BIS>(cq_keys, cq_code, 0u, 32u, hb, hshift, ovl, novf, nodes, n, rs, rc, sc_ent, sc_row, cnt, lane, lt);
    const uint32_t left = qn - 32u;                  // < 32: move the tail to the front
    uint4 tk = make_uint4(0u, 0u, 0u, 0u); uint32_t tc = 0;
    if ((uint32_t)lane < left) { tk = cq_keys[32 + lane]; tc = cq_code[32 + lane]; }
    __syncwarp();
    if ((uint32_t)lane < left) { cq_keys[lane] = tk; cq_code[lane] = tc; }
    qn = left;
    __syncwarp();
  }
}

// first position IN SCORE ORDER whose score fails the relative threshold (PS.cpp:584), or size_neigh; warp-uniform result
__device__ __forceinline__ uint32_t ppr_cut(const SampleParams &P, const unsigned long long off, const int len_all, const int size_neigh,
                                            const int lane) {
  const float max_ppr = size_neigh > 1 ? P.ppr_scores[off + 1] : 0.f;                          // :578-579
  uint32_t cut = (uint32_t)size_neigh;
#pragma unroll 1
  for (int base = 0; base < len_all; base += 32) {
    const int i = base + lane;
    if (i < len_all) {
      const uint32_t r = P.ppr_srank[off + i];
      if (r < (uint32_t)size_neigh && (max_ppr == 0.f || __fdiv_rn(P.ppr_sscore[off + i], max_ppr) < P.threshold)) cut = min(cut, r);
    }
  }
  return __reduce_min_sync(0xffffffffu, cut);
}

// Hop labels (compute_hops, G.cpp:32-64): level-synchronous BFS from the root over the subgraph's OWN rows, as just written (inserted self
// loops and PS.cpp:401 extras included), by the warp that built it.  The frontier's rows are flattened: 32 frontier nodes at a time, a warp
// scan of their row lengths, then lane k takes edge k of the concatenation (owner found by a 5-step bisection over the lanes' prefix
// values), so the root's ~140-edge row and a frontier of 140 two-edge rows both keep all lanes busy.  Distances are unique, so the
// result does not depend on the order in which a level claims its nodes.  Unreachable nodes keep 0xFFFFFFFF like the generic kernel.
__device__ __forceinline__ void warp_bfs_hops(const uint32_t *cp, const int *idx, const long long node_base, const int n, const uint32_t src, uint32_t *dist,
                                              uint32_t *cur, uint32_t *nxt, uint32_t *cnt, const int lane) {
  const uint32_t FULL = 0xffffffffu;
  for (int i = lane; i < n; i += 32) dist[i] = NONE32;
  if (lane == 0) *cnt = 0u;
  __syncwarp();
  if (lane == 0) { dist[src] = 0u; cur[0] = src; }
  __syncwarp();
  uint32_t ncur = 1;
#pragma unroll 1
  for (uint32_t lvl = 0; ncur > 0; lvl++) {
#pragma unroll 1
    for (uint32_t base = 0; base < ncur; base += 32) {
      const uint32_t i = base + (uint32_t)lane;
      uint32_t s = 0, len = 0;
      if (i < ncur) { const uint32_t u = cur[i]; s = cp[u]; len = cp[u + 1] - s; }
      const uint32_t incl = warp_incl_scan(len, lane), excl = incl - len;
      const uint32_t tot = __shfl_sync(FULL, incl, 31);
#pragma unroll 1
      for (uint32_t kb = 0; kb < tot; kb += 32) {
        const uint32_t k = kb + (uint32_t)lane;
        int j = 0;                                      // largest lane whose exclusive prefix is <= k: the owner of edge k
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          const uint32_t e = __shfl_sync(FULL, excl, (j + d) & 31);
          if (j + d < 32 && e <= k) j += d;
        }
        const uint32_t sj = __shfl_sync(FULL, s, j), ej = __shfl_sync(FULL, excl, j);
        if (k < tot) {
          const uint32_t c = (uint32_t)((long long)__ldcg(idx + sj + (k - ej)) - node_base);
          if (atomicCAS(&dist[c], NONE32, lvl + 1u) == NONE32) nxt[atomicAdd(cnt, 1u)] = c;
        }
      }
    }
    __syncwarp();
    ncur = *cnt;
    __syncwarp();
    if (lane == 0) *cnt = 0u;
    uint32_t *t = cur; cur = nxt; nxt = t;
    __syncwarp();
  }
}

#define PPR_COUNT_R 6          // table rows of up to 192 entries are counted from registers
// node count (and score-rank cut) of every subgraph of the launch: |{entries with rank < cut, id != root}| + 1
__global__ void __launch_bounds__(256) ppr_count_kernel(const SampleParams P, int *__restrict__ cnt, unsigned short *__restrict__ cut_out) {
  const int lane = threadIdx.x & 31;
  for (int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < P.num_subg; p += gridDim.x * (blockDim.x >> 5)) {
    const uint32_t t = P.roots[p];
    const unsigned long long off = P.ppr_ptr[t];
    const int len_all = (int)(P.ppr_ptr[t + 1] - off);
    const int size_neigh = len_all < P.k ? len_all : P.k;                                         // :576
    uint32_t cut, c = 0;
    if (len_all <= 32 * PPR_COUNT_R) {                   // the usual case: the whole row in registers, every load issued before the first use
      uint32_t rk[PPR_COUNT_R], id[PPR_COUNT_R];
      float sc[PPR_COUNT_R];
#pragma unroll
      for (int j = 0; j < PPR_COUNT_R; j++) {
        const int i = 32 * j + lane;
        rk[j] = NONE32; id[j] = t; sc[j] = 0.f;
        if (i < len_all) { rk[j] = P.ppr_srank[off + i]; id[j] = P.ppr_sid[off + i]; sc[j] = P.ppr_sscore[off + i]; }
      }
      const float max_ppr = size_neigh > 1 ? P.ppr_scores[off + 1] : 0.f;                        // :578-579
      cut = (uint32_t)size_neigh;
#pragma unroll
      for (int j = 0; j < PPR_COUNT_R; j++)
        if (rk[j] < (uint32_t)size_neigh && (max_ppr == 0.f || __fdiv_rn(sc[j], max_ppr) < P.threshold)) cut = min(cut, rk[j]);
      cut = __reduce_min_sync(0xffffffffu, cut);
#pragma unroll
      for (int j = 0; j < PPR_COUNT_R; j++) c += (rk[j] < cut && id[j] != t) ? 1u : 0u;
      c = __reduce_add_sync(0xffffffffu, c);
    } else {
      cut = ppr_cut(P, off, len_all, size_neigh, lane);
      for (int base = 0; base < len_all; base += 32) {
        const int i = base + lane;
        const bool selp = i < len_all && P.ppr_srank[off + i] < cut && P.ppr_sid[off + i] != t;
        c += __popc(__ballot_sync(0xffffffffu, selp));
      }
    }
    if (lane == 0) { cnt[p] = (int)c + 1; cut_out[p] = (unsigned short)cut; }
  }
}

// exclusive prefix sum of cnt[0..P) into out[0..P] (one block; P is 10^4 .. 10^5), total also to *total.  The array is walked in tiles of
// 1024 x 16 items: thread i owns 16 contiguous items of the tile (four 16-byte loads issued before the first add), one block scan per tile,
// the running total carries over.  cnt must be 16-byte aligned (the launcher pads P to a multiple of 4 ints in front of it).
#define SCAN_IPT 16
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int *__restrict__ cnt, const int P, int *__restrict__ out, long long *total) {
  __shared__ uint32_t sums[1024];
  __shared__ uint32_t warp_sums[33];
  uint32_t carry = 0;
  for (int t0 = 0; t0 < P; t0 += 1024 * SCAN_IPT) {
    const int b = t0 + threadIdx.x * SCAN_IPT;
    int4 v[SCAN_IPT / 4];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_IPT / 4; j++) {
      v[j] = make_int4(0, 0, 0, 0);
      const int i = b + 4 * j;
      if (i + 3 < P) v[j] = *reinterpret_cast<const int4 *>(cnt + i);
      else if (i < P) { v[j].x = cnt[i]; if (i + 1 < P) v[j].y = cnt[i + 1]; if (i + 2 < P) v[j].z = cnt[i + 2]; }
    }
#pragma unroll
    for (int j = 0; j < SCAN_IPT / 4; j++) s += (uint32_t)(v[j].x + v[j].y + v[j].z + v[j].w);
    sums[threadIdx.x] = s;
    __syncthreads();
    const uint32_t tot = block_exclusive_scan(sums, 1024, warp_sums);
    uint32_t run = carry + sums[threadIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_IPT / 4; j++) {
      const int i = b + 4 * j;
      const int4 o = make_int4((int)run, (int)run + v[j].x, (int)run + v[j].x + v[j].y, (int)run + v[j].x + v[j].y + v[j].z);
      if (i + 3 < P) *reinterpret_cast<int4 *>(out + i) = o;
      else if (i < P) { out[i] = o.x; if (i + 1 < P) out[i + 1] = o.y; if (i + 2 < P) out[i + 2] = o.z; }
      run += (uint32_t)(v[j].x + v[j].y + v[j].z + v[j].w);
    }
    carry += tot;
    __syncthreads();                                    // sums[] / warp_sums[] are reused by the next tile
  }
  if (threadIdx.x == 0) { out[P] = (int)carry; *total = (long long)carry; }
}

// NF = Bloom words per lane (1: ncap <= 192, 2: <= 448, 4 above)
// SYM = the graph was verified symmetric with strictly ascending rows when the sampler first used it (sym_build_kernel, sampler.cu).  Then
// v in adj(u) <=> u in adj(v), so only the UPPER part of every row (neighbours >= the row's own id; its extent travels with the table entry,
// ppr_supper) is scanned -- half the slots, half the probes.  A kept edge (u, v), u < v, found in row u also stands for the edge (v, u) of
// row v: its full-graph slot comes from the reverse-slot index sym_rev[] (one 4-byte read per kept edge), and because rows are scanned in
// ascending u the mirrored edges of a row arrive in ascending order, i.e. in CSR order, ahead of the row's own upper part (PS.cpp:420-422
// emits a row in slot order = ascending neighbour id).  The PS.cpp:401 slot one past a row is a DIRECTED quirk: it is kept for its row and
// never mirrored.  Output is bit-identical to the full scan (same parity tests); a graph that fails the check keeps the full scan.
// BIS = exact membership by bisection in the sorted node list instead of the 2-key-bucket table (large node sets, see bisect4).
template <bool ADD_SELF, int NF, bool SYM, bool BIS>
__global__ void __launch_bounds__(32, WARP_MIN_BLOCKS) ppr_induce_warp_kernel(const SampleParams P) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  uint32_t *const nodes = (uint32_t *)(smem_dyn + P.WL.nodes);
  uint2 *const rs = (uint2 *)(smem_dyn + P.WL.rs);                 // {row start, scanned length} of every node
  uint32_t *const cp = (uint32_t *)(smem_dyn + P.WL.cp);           // chunk prefix during the scan, local indptr afterwards
  uint32_t *const rc = (uint32_t *)(smem_dyn + P.WL.rc);           // per row: kept | kept below v << 14 | v kept << 28
  uint32_t *const hk = (uint32_t *)(smem_dyn + P.WL.hkeys);        // exact membership: buckets of 2 keys
  uint32_t *const ovf = (uint32_t *)(smem_dyn + P.WL.ovf);         // [0] = count, [1..WARP_OVF_CAP] = keys whose bucket was full
  uint32_t *const fw = (uint32_t *)(smem_dyn + P.WL.bloom);        // Bloom words while they are built (32 * NF)
  uint4 *const cq_keys = (uint4 *)(smem_dyn + P.WL.queue);         // chunk queue: the 4 keys of a chunk ...
  uint32_t *const cq_code = (uint32_t *)(cq_keys + WARP_QCAP);     // ... and row << 19 | offset of its first slot in the row + 3
  uint32_t *const rlo = (uint32_t *)(smem_dyn + P.WL.rlo);         // ADD_SELF: first staged entry / insert position / bug column per row; SYM: rlo = output offset of a row's staged run
  uint32_t *const rins = (uint32_t *)(smem_dyn + P.WL.rins);
  uint32_t *const rbug = (uint32_t *)(smem_dyn + P.WL.rbug);
  // staged kept edges {neighbour id, full-graph slot} (+ row), CSR order: per-warp scratch in global memory.  A warp reuses its few KB
  // for every subgraph, so the region stays in L2; keeping it out of shared memory is worth ~10 more resident warps per SM.
  uint2 *const sc_ent = (uint2 *)(P.w_scratch + (size_t)blockIdx.x * P.w_scratch_stride);
  unsigned short *const sc_row = (unsigned short *)(sc_ent + P.w_ecap);
  const uint2 *const hb = (const uint2 *)hk;
  const uint4 *const ind4 = (const uint4 *)P.indices;
  const int lane = threadIdx.x;
  const uint32_t FULL = 0xffffffffu;
  const uint32_t lt = lanemask_lt(), le = lanemask_le();
  const uint32_t nbuckets = (uint32_t)P.w_hbuckets;
  const int hshift = P.w_hshift;
  const uint32_t ecap = (uint32_t)P.w_ecap;
  const bool ext = !ADD_SELF && !P.fixed_mode;                     // PS.cpp:401: slot `e` is tested whenever no self edge is inserted
  const uint32_t E = P.num_edges;

  // The header of a subgraph (ticket -> root, rows reserved -> table extent) is a chain of dependent global round trips; it is fetched
  // one subgraph ahead, each link issued where the previous one has had a whole phase to arrive.
  int p = 0;
  if (lane == 0) p = (int)atomicAdd(P.ticket, 1u);
  p = __shfl_sync(FULL, p, 0);
  uint32_t t = 0, cut = 0;
  int node_base = 0;
  unsigned long long off = 0, row_end = 0;
  if (p < P.num_subg) { t = P.roots[p]; node_base = P.node_ptr[p]; cut = P.w_cut[p]; off = P.ppr_ptr[t]; row_end = P.ppr_ptr[t + 1]; }

  for (;;) {
    __syncwarp();
    if (p >= P.num_subg) break;
    uint32_t tk_next = 0;
    if (lane == 0) tk_next = atomicAdd(P.ticket, 1u);             // consumed after phase B
    int p_next = 0, node_base_next = 0;
    uint32_t t_next = 0, cut_next = 0;
    unsigned long long off_next = 0, row_end_next = 0;
    do {

    // ---------------- A: node set = {table entries with score rank < cut} U {root}, already in id order (PS.cpp:565-595) ----------------
    // (cut and the node count come from ppr_count_kernel; orig_node / ppr go straight to their reserved rows)
    const int len_all = (int)(row_end - off);
    const int size_neigh = len_all < P.k ? len_all : P.k;                                         // :576
    uint32_t run = 0, n_below = 0;
    bool root_in = false;
    // (the loads of round i+1 are issued before round i is consumed: the table row sits in L2, a round trip per round would be exposed)
    const uint2 *const trow = SYM ? P.ppr_supper : P.ppr_srow;
    uint32_t rk_n = NONE32, id_n = 0;
    float sc_n = 0.f;
    uint2 rw_n = make_uint2(0u, 0u);
    if (lane < len_all) { rk_n = P.ppr_srank[off + lane]; id_n = P.ppr_sid[off + lane]; sc_n = P.ppr_sscore[off + lane]; rw_n = trow[off + lane]; }
#pragma unroll 1
    for (int base = 0; base < len_all; base += 32) {
      const uint32_t rk = rk_n, id = id_n;
      const float scv = sc_n;
      const uint2 rw = rw_n;
      const int j = base + 32 + lane;
      rk_n = NONE32;
      if (j < len_all) { rk_n = P.ppr_srank[off + j]; id_n = P.ppr_sid[off + j]; sc_n = P.ppr_sscore[off + j]; rw_n = trow[off + j]; }
      const bool sel = rk < cut;                        // NONE32 (beyond the row) never is
      const bool selp = sel && id != t;
      const uint32_t m = __ballot_sync(FULL, selp), mb = __ballot_sync(FULL, selp && id < t), mr = __ballot_sync(FULL, sel && id == t);
      if (sel) {                                        // entries below the root keep their rank, the root's slot follows them, the rest shift by one
        const uint32_t at = run + __popc(m & lt) + ((selp && id > t) ? 1u : 0u);
        nodes[at] = id; rs[at] = rw;
        P.orig_node[node_base + at] = id; P.ppr_out[node_base + at] = scv;
      }
      run += __popc(m); n_below += __popc(mb); root_in |= (mr != 0);
    }
    if (!root_in && lane == 0) {                        // root not among the selected entries: -1, or scores[0] when the row has <= 1 entry (:574,:581)
      nodes[n_below] = t;
      P.orig_node[node_base + n_below] = t;
      P.ppr_out[node_base + n_below] = (size_neigh <= 1 && len_all > 0) ? P.ppr_scores[off] : -1.f;
      uint32_t s = P.indptr[t];
      const uint32_t e = P.indptr[t + 1];
      if (SYM) {                                        // upper part of the root's row: first slot whose neighbour is >= t
        uint32_t hi = e;
        while (s < hi) { const uint32_t mid = s + ((hi - s) >> 1); if (__ldg(P.indices + mid) < t) s = mid + 1; else hi = mid; }
      }
      rs[n_below] = make_uint2(s, e - s);
    }
    const int n = (int)run + 1;                         // == node_ptr[p+1] - node_ptr[p]

    // ---------------- B: Bloom filter + exact table, chunk prefix ----------------
    if (!BIS) {
#pragma unroll 1
      for (uint32_t i = lane; i < nbuckets / 2; i += 32) reinterpret_cast<uint4 *>(hk)[i] = make_uint4(NONE32, NONE32, NONE32, NONE32);
    }
#pragma unroll
    for (int j = 0; j < NF; j++) fw[32 * j + lane] = 0u;
    if (lane == 0) ovf[0] = 0;
    __syncwarp();
    uint32_t total = 0;                                 // chunks of the subgraph
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      uint32_t ch = 0;
      if (i < n) {
        const uint32_t v = nodes[i];
        const uint32_t h = bloom_hash(v);
        atomicOr(&fw[bloom_word_index<NF>(h)], bloom_bits(h, bloom_hash2(v)));
        if (!BIS) {
          const uint32_t b = h >> hshift;
          if (atomicCAS(&hk[2 * b], NONE32, v) != NONE32 && atomicCAS(&hk[2 * b + 1], NONE32, v) != NONE32) {
            const uint32_t q = atomicAdd(&ovf[0], 1u);
            if (q < WARP_OVF_CAP) ovf[1 + q] = v;
          }
        }
        uint2 r = rs[i];
        if (ext && r.x + r.y < E) { r.y += 1; rs[i] = make_uint2(r.x, SYM ? (r.y | 0x80000000u) : r.y); }      // the PS.cpp:401 slot joins the row
        ch = r.y ? (((r.x + r.y + (WARP_CS - 1)) >> WARP_CSH) - (r.x >> WARP_CSH)) : 1u;          // aligned chunks; every row owns >= 1
        rc[i] = 0;
      }
      const uint32_t x = warp_incl_scan(ch, lane);
      if (i < n) cp[i] = total + x - ch;
      total += __shfl_sync(FULL, x, 31);
    }
    if (lane == 0) cp[n] = total;
    __syncwarp();
    uint32_t f[NF];
#pragma unroll
    for (int j = 0; j < NF; j++) f[j] = fw[32 * j + lane];
    p_next = __shfl_sync(FULL, (int)tk_next, 0);
    if (p_next < P.num_subg) { t_next = P.roots[p_next]; node_base_next = P.node_ptr[p_next]; cut_next = P.w_cut[p_next]; }    // consumed after the scan
    const uint32_t novf = ovf[0];
    const uint32_t *const ovl = ovf + 1;
    bool bail = novf > WARP_OVF_CAP;

    // ---------------- C: one pass over the rows in chunk space ----------------
    uint32_t cnt = 0, qn = 0;
    if (!bail) {
      int rbase = 0;                                    // cp[rbase] <= first chunk of the window <= cp[rbase+1]
      constexpr int U = SYM ? WARP_U_SYM : WARP_U;
      const uint32_t step = 32u * U, room = 32u * (U + 1) * WARP_CS;      // what a stage can keep at most (incl. the queue's tail)
      ScanStage<U> SA;
#if WARP_DB
      ScanStage<U> SB;
      scan_load(SA, 0u, total, rbase, n, lane, le, cp, rs, ind4);
#endif
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < total; c0 += step) {
        if (cnt + room > ecap) { bail = true; break; }
#if WARP_DB
        if (c0 + step < total) scan_load(SB, c0 + step, total, rbase, n, lane, le, cp, rs, ind4);      // in flight while SA is processed
#else
        scan_load(SA, c0, total, rbase, n, lane, le, cp, rs, ind4);
#endif
        uint32_t any[U];                           // all Bloom tests of the stage first: 4 * WARP_U independent SHFLs in flight
#pragma unroll
        for (int u = 0; u < U; u++)
          any[u] = (bloom_test<NF>(f, SA.q[u].x) | bloom_test<NF>(f, SA.q[u].y) | bloom_test<NF>(f, SA.q[u].z) | bloom_test<NF>(f, SA.q[u].w)) & 1u;
#pragma unroll
        for (int u = 0; u < U; u++)
          scan_window<ADD_SELF, NF, SYM, BIS>(any[u], SA.q[u], SA.code[u], cq_keys, cq_code, qn, hb, hshift, ovl, novf, nodes, n, rs, rc, sc_ent, sc_row, cnt, lane, lt);
#if WARP_DB
#pragma unroll
        for (int u = 0; u < U; u++) { SA.q[u] = SB.q[u]; SA.code[u] = SB.code[u]; }
#endif
      }
      if (!bail && qn) { __syncwarp(); drain_batch<ADD_SELF, SYM, BIS>(cq_keys, cq_code, 0u, qn, hb, hshift, ovl, novf, nodes, n, rs, rc, sc_ent, sc_row, cnt, lane, lt); }
    }
    __syncwarp();
    if (p_next < P.num_subg) { off_next = P.ppr_ptr[t_next]; row_end_next = P.ppr_ptr[t_next + 1]; }      // consumed before the emit
    // launch statistics for the host's choice between the full and the symmetric variant: staged edges, scanned chunks
    if (!bail && lane == 0) { atomicAdd((unsigned long long *)&P.totals[4], (unsigned long long)cnt); atomicAdd((unsigned long long *)&P.totals[5], (unsigned long long)total); }
    if (SYM && !bail) {
      // resolve pass over the staged (upper-part) edges, all lanes busy: sub id of the neighbour, "does this edge have a mirror image",
      // mirror count of the neighbour's row (bits 14..27 of rc[]); the entry is rewritten as row | sub << 13 | mirrored << 31
      // (two entries per lane and round: both scratch reads and both binary searches overlap)
#pragma unroll 1
      for (uint32_t g0 = lane; g0 < cnt; g0 += 64) {
        const uint32_t g1 = g0 + 32;
        const bool has1 = g1 < cnt;
        const uint2 ent0 = __ldcg(sc_ent + g0);
        const uint32_t row0 = __ldcg(sc_row + g0);
        uint2 ent1 = ent0;
        uint32_t row1 = row0;
        if (has1) { ent1 = __ldcg(sc_ent + g1); row1 = __ldcg(sc_row + g1); }
        int lo0 = 0, hi0 = n, lo1 = 0, hi1 = n;         // sub_of for both keys, interleaved
        while (lo0 < hi0 || lo1 < hi1) {
          if (lo0 < hi0) { const int mid = (lo0 + hi0) >> 1; if (nodes[mid] < ent0.x) lo0 = mid + 1; else hi0 = mid; }
          if (lo1 < hi1) { const int mid = (lo1 + hi1) >> 1; if (nodes[mid] < ent1.x) lo1 = mid + 1; else hi1 = mid; }
        }
        const uint2 r0 = rs[row0], r1 = rs[row1];
        const bool ext0 = (r0.y >> 31) && ent0.y == r0.x + (r0.y & 0x7fffffffu) - 1u;             // the PS.cpp:401 slot: directed, no mirror image
        const bool ext1 = (r1.y >> 31) && ent1.y == r1.x + (r1.y & 0x7fffffffu) - 1u;
        const bool mir0 = !ext0 && ent0.x > nodes[row0];                                        // a self loop is its own mirror image
        const bool mir1 = has1 && !ext1 && ent1.x > nodes[row1];
        if (mir0) { atomicAdd(&rc[lo0], 1u << 14); prefetch_l2(P.sym_rev + ent0.y); }           // the emit reads sym_rev[slot]: pull it into L2 now
        if (mir1) { atomicAdd(&rc[lo1], 1u << 14); prefetch_l2(P.sym_rev + ent1.y); }
        sc_ent[g0].x = row0 | ((uint32_t)lo0 << 13) | (mir0 ? 0x80000000u : 0u);
        if (has1) sc_ent[g1].x = row1 | ((uint32_t)lo1 << 13) | (mir1 ? 0x80000000u : 0u);
      }
      __syncwarp();
    }

    // ---------------- per-row counts -> local indptr (PS.cpp:428-431) ----------------
    uint32_t m = cnt;
    if (!bail) {
      uint32_t carry = 0, carry_k = 0;
#pragma unroll 1
      for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        uint32_t c_i = 0, k_i = 0;
        if (i < n) {
          const uint32_t w = rc[i];
          if (!ADD_SELF && SYM) { k_i = w & 0x3fffu; c_i = k_i + (w >> 14); }      // own upper part + mirrored lower part
          else if (!ADD_SELF) c_i = w;                  // nothing is inserted: the staged stream is the CSR
          else {
            k_i = w & 0x3fffu;
            const bool present = (w >> 28) != 0;        // the row already holds its self loop (:386-400)
            uint32_t bsub = NONE32;
            if (present && !P.fixed_mode) {             // no insertion => slot e is tested too (:401)
              const uint32_t e = rs[i].x + rs[i].y;
              if (e < E) {
                const uint32_t nb = __ldg(P.indices + e);
                if (BIS ? bisect1(nodes, n, nb) : probe2(hb, hshift, ovl, novf, nb)) bsub = sub_of(nodes, n, nb);
              }
            }
            rins[i] = present ? NONE32 : ((w >> 14) & 0x3fffu); rbug[i] = bsub;                // SYM: bits 14..27 = mirrored edges, all below v
            c_i = k_i + (present ? 0u : 1u) + (bsub != NONE32 ? 1u : 0u) + (SYM ? ((w >> 14) & 0x3fffu) : 0u);
          }
        }
        const uint32_t x = warp_incl_scan(c_i, lane);
        if (i < n) cp[i] = carry + x - c_i;
        if (ADD_SELF || SYM) {
          const uint32_t y = warp_incl_scan(k_i, lane);
          if (i < n) {
            if (SYM) {                                  // staged entry g of this row goes to rlo + g; the mirrored edges fill the row's head, rc = their cursor
              const uint32_t mcount = (rc[i] >> 14) & 0x3fffu;
              rlo[i] = (carry + x - c_i) + mcount + ((ADD_SELF && rins[i] != NONE32) ? 1u : 0u) - (carry_k + y - k_i);
              rc[i] = carry + x - c_i;
            } else rlo[i] = carry_k + y - k_i;
          }
          carry_k += __shfl_sync(FULL, y, 31);
        }
        carry += __shfl_sync(FULL, x, 31);
      }
      if (lane == 0) cp[n] = carry;
      m = carry;
    }
    __syncwarp();

    // ---------------- the edge block goes wherever the cursor stands ----------------
    if (bail) {                                         // hand the subgraph to the generic kernel (redo mode); its rows are reserved already
      if (lane == 0) P.redo_list[atomicAdd(P.redo_count, 1u)] = p;
      break;
    }
    long long edge_base = 0;
    if (lane == 0) edge_base = (long long)atomicAdd((unsigned long long *)&P.totals[1], (unsigned long long)m);
    if (p_next < P.num_subg) {                          // meanwhile: pull the next subgraph's table row into L2
      const unsigned long long ln = row_end_next - off_next;
      const unsigned long long o128 = (unsigned long long)lane * 128ull;
      if (o128 < ln * 4ull + 128ull) prefetch_l2((const char *)(P.ppr_sid + off_next) + o128);
      if (o128 < ln * 4ull + 128ull) prefetch_l2((const char *)(P.ppr_sscore + off_next) + o128);
      if (o128 < ln * 2ull + 128ull) prefetch_l2((const char *)(P.ppr_srank + off_next) + o128);
      if (o128 < ln * 8ull + 128ull) prefetch_l2((const char *)((SYM ? P.ppr_supper : P.ppr_srow) + off_next) + o128);
    }
    edge_base = __shfl_sync(FULL, edge_base, 0);
    if (edge_base + (long long)m > P.cap_edges) {
      if (lane == 0) atomicOr((unsigned long long *)&P.totals[2], (unsigned long long)ERR_OUT_OVERFLOW);
      break;
    }
    if (lane == 0) {
      P.edge_span[p] = make_int2((int)edge_base, (int)(edge_base + m)); P.num_target[p] = 1;
      P.target[p] = (int)(node_base + sub_of(nodes, n, t));
    }
#pragma unroll 1
    for (int i = lane; i < n; i += 32) P.row_span[node_base + i] = make_int2((int)(edge_base + cp[i]), (int)(edge_base + cp[i + 1]));
    // ---------------- emit ----------------
    if (SYM) {
      // (the staged entries of round i+1 and their reverse slots are requested before round i is written out)
      uint2 ent_n = make_uint2(0u, 0u);
      uint32_t rv_n = 0;
      if ((uint32_t)lane < cnt) { ent_n = __ldcg(sc_ent + lane); if (ent_n.x >> 31) rv_n = __ldg(P.sym_rev + ent_n.y); }
#pragma unroll 1
      for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t g = base + lane;
        const bool act = g < cnt;
        const uint2 ent = ent_n;
        const uint32_t rv = rv_n;                       // full-graph slot of the mirror image (v, u)
        ent_n = make_uint2(0u, 0u);
        if (g + 32u < cnt) { ent_n = __ldcg(sc_ent + g + 32u); if (ent_n.x >> 31) rv_n = __ldg(P.sym_rev + ent_n.y); }
        const uint32_t row = ent.x & 0x1fffu, sv = (ent.x >> 13) & 0x1fffu;
        const bool mir = act && (ent.x >> 31);
        if (act) {                                      // the edge itself: row u, in staged (= slot) order behind the row's mirrored head
          const long long pos = edge_base + (long long)(rlo[row] + g);
          P.indices_out[pos] = (int)(node_base + sv); P.orig_edge[pos] = ent.y;
        }
        if (__any_sync(FULL, mir)) {                    // mirror images: row v; entries of one row arrive in ascending u = CSR order
          const uint32_t peers = __match_any_sync(FULL, mir ? sv : (0x2000u + (uint32_t)lane));
          const uint32_t rank = __popc(peers & lt);
          uint32_t cur = 0;
          if (mir) cur = rc[sv];
          __syncwarp();
          if (mir && rank == 0) rc[sv] = cur + __popc(peers);
          __syncwarp();
          if (mir) {
            const long long pos = edge_base + (long long)(cur + rank);
            P.indices_out[pos] = (int)(node_base + row); P.orig_edge[pos] = rv;
          }
        }
      }
    } else {
#pragma unroll 1
    for (uint32_t g = lane; g < cnt; g += 32) {
      const uint2 ent = __ldcg(sc_ent + g);
      long long pos = edge_base + g;
      if (ADD_SELF) {
        const uint32_t row = __ldcg(sc_row + g);
        pos = edge_base + cp[row] + (g - rlo[row]) + ((rins[row] != NONE32 && ent.x > nodes[row]) ? 1 : 0);
      }
      P.indices_out[pos] = (int)(node_base + sub_of(nodes, n, ent.x));
      P.orig_edge[pos] = ent.y;                                                                   // :422
    }
    }
    if (ADD_SELF) {
#pragma unroll 1
      for (int r = lane; r < n; r += 32) {
        if (rins[r] != NONE32) {                                                                  // inserted self edge (:406-411)
          const long long pos = edge_base + cp[r] + rins[r];
          P.indices_out[pos] = (int)(node_base + r); P.orig_edge[pos] = NONE32;
        }
        if (rbug[r] != NONE32) {
          const long long pos = edge_base + cp[r + 1] - 1;
          P.indices_out[pos] = (int)(node_base + rbug[r]); P.orig_edge[pos] = rs[r].x + rs[r].y;
        }
      }
    }
    if (P.aug & SHADOW_AUG_HOPS) {                      // PS.cpp:433-436; rs[] / nodes[] / the overflow list are dead by now
      __syncwarp();                                     // the warp's own stores to indices_out are ordered before its reads
      const uint32_t tl = sub_of(nodes, n, t);
      uint32_t *const dist = reinterpret_cast<uint32_t *>(rs);
      __syncwarp();
      warp_bfs_hops(cp, P.indices_out + edge_base, (long long)node_base, n, tl, dist, dist + n, nodes, ovf, lane);
#pragma unroll 1
      for (int i = lane; i < n; i += 32) P.hop[node_base + i] = dist[i];
    }
    } while (0);
    p = p_next; t = t_next; node_base = node_base_next; cut = cut_next; off = off_next; row_end = row_end_next;
  }
}
