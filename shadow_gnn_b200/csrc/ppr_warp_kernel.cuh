// ppr_warp_kernel.cuh -- the single-root PPR fast path of the fused select + induce step: ONE WARP per subgraph.
//
// Same contract as sample_induce_kernel (sampler_kernels.cuh) for  method == ppr, one root per subgraph, no hop/drnl
// labels  -- the configuration every PPR node-task run of the reference uses (shaDow/minibatch.py:373-375; PS.cpp:565-595
// + PS.cpp:350-453).  What is different is the mapping to the machine:
//   * a warp owns a subgraph end to end, so there is no CTA barrier and no cross-warp scan anywhere;
//   * the node set comes from the id-sorted PPR row, which also carries (row start, degree) of every neighbour -- captured once
//     when the tables are installed -- so the random 8-byte indptr reads (a 64-byte DRAM fetch each) disappear;
//   * the row scan runs over a FLAT space of chunks: all rows of the subgraph are cut into aligned chunks of 4*WARP_CH slots, lane L of
//     window w takes chunk 32w+L whatever row it belongs to (rows are resolved 32 chunks at a time with one REDUX), loads it with
//     WARP_CH 128-bit streaming loads and probes the shared-memory hash once per slot.  Short rows no longer idle lanes (a
//     row-granular 128-byte item keeps 82 % of the lanes busy on the S-products stand-in, a 512-byte one only 43 %);
//   * with no self-edge insertion the PS.cpp:401 slot (one past the row) is simply one more slot of the row, and the staged stream
//     IS the CSR: emit is a straight copy with the sub id resolved per kept edge;
//   * kept edges are staged in a per-warp scratch region in GLOBAL memory that the warp reuses for every subgraph (it stays in L2), per-row
//     counts come from one shared-memory atomic per lane: 7 KB of shared memory per warp instead of 12, i.e. 28 instead of 18 warps/SM;
//   * the node count of a subgraph is a function of its table row alone, so a small count kernel + a one-block scan fix node_ptr[] before
//     this kernel starts: no decoupled look-back (with ~2,500 subgraphs in flight in lock-step the nearest inclusive prefix was ~75
//     windows of 32 tickets back: 17 % of all instructions of the first version were that walk).
// A subgraph that does not fit the per-warp scratch (or whose hash overflow list is full) is not an error: its index goes on a
// redo list and the generic CTA kernel, launched right behind in redo mode, builds it into the rows this kernel reserved.
#pragma once
#include "sampler_kernels.cuh"

// tuning macros (measured on the S-products stand-in, scripts/explore_variants.py; defaults = the fastest build)
#ifndef WARP_U
#define WARP_U 4               // chunk windows in flight per warp (32 chunks each)
#endif
#ifndef WARP_CH
#define WARP_CH 1              // 16-byte loads per lane and window: a chunk is 4*WARP_CH slots
#endif
#ifndef WARP_MIN_BLOCKS
#define WARP_MIN_BLOCKS 28     // one warp per CTA: resident warps per SM the register budget must allow
#endif
#ifndef WARP_DB
#define WARP_DB 0              // 1: scan stages double-buffered in registers
#endif
#ifndef WARP_OVK
#define WARP_OVK 3             // overflow keys the straight-line scan variant checks from registers (more overflow keys: looping variant)
#endif
#ifndef WARP_BK
#define WARP_BK 2              // keys per hash bucket: 2 (8-byte probe, ~6 shared-memory wavefronts per warp-wide probe, 4 buckets per key:
                               // 1-2 keys per subgraph overflow their bucket and ride in registers) or 4 (16-byte probe, ~10 wavefronts,
                               // 1 bucket per key, overflow in one subgraph out of ten).  65,536 subgraphs: 1.40 ms vs 1.51 ms
#endif
#define WARP_OVF_CAP 8
#define WARP_CS (4 * WARP_CH)                                  // slots per chunk
#define WARP_CSH (WARP_CH == 1 ? 2 : (WARP_CH == 2 ? 3 : 4))   // log2(WARP_CS)

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
// membership: buckets of WARP_BK keys, one shared-memory load per probe.  A key whose bucket is full goes to a short overflow list: up to
// WARP_OVK of them are checked from registers by the straight-line scan variant, longer lists select the looping variant
#if WARP_BK == 4
typedef uint4 bucket_t;
__device__ __forceinline__ bool probe4(const bucket_t *hb, const int hshift, const uint32_t key) {
  const uint4 kk = hb[(key * 2654435761u) >> hshift];
  return (kk.x == key) | (kk.y == key) | (kk.z == key) | (kk.w == key);
}
#else
typedef uint2 bucket_t;   // 2-key buckets: ~6 instead of ~10 shared-memory wavefronts per warp-wide probe, but twice the table for the same overflow rate
__device__ __forceinline__ bool probe4(const bucket_t *hb, const int hshift, const uint32_t key) {
  const uint2 kk = hb[(key * 2654435761u) >> hshift];
  return (kk.x == key) | (kk.y == key);
}
#endif

// One stage of the scan = WARP_U windows of 32 chunks, all loads issued before the first use.
struct ScanStage {
  uint4 q[WARP_U][WARP_CH];
  uint32_t jb[WARP_U], s[WARP_U], d[WARP_U], rw[WARP_U];
};

__device__ __forceinline__ void scan_load(ScanStage &S, const uint32_t c0, const uint32_t total, int &rbase, const int n, const int lane,
                                          const uint32_t le, const uint32_t *cp, const uint2 *rs, const uint4 *ind4, const uint32_t *indices,
                                          const uint32_t E, const uint32_t E_al) {
#pragma unroll
  for (int u = 0; u < WARP_U; u++) {
    const uint32_t cw = c0 + 32u * u;
#pragma unroll
    for (int h = 0; h < WARP_CH; h++) S.q[u][h] = make_uint4(0u, 0u, 0u, 0u);
    S.jb[u] = 0; S.s[u] = 0; S.d[u] = 0; S.rw[u] = 0;
    if (cw < total) {
      // rows that start inside this window: lane j looks at row rbase+1+j, one REDUX turns the starts into a bit mask
      const uint32_t rel = cp[min(rbase + 1 + lane, n)] - cw;
      const uint32_t mask = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
      const int row = min(rbase + __popc(mask & le), n - 1);
      rbase += __popc(mask);
      const uint32_t c = cw + lane;
      if (c < total) {
        const uint2 r = rs[row];
        const uint32_t b = ((r.x >> WARP_CSH) + (c - cp[row])) << WARP_CSH;      // first slot of my chunk (16*WARP_CH-byte aligned)
        S.jb[u] = b; S.s[u] = r.x; S.d[u] = r.y; S.rw[u] = (uint32_t)row;
        if (b < E_al) {
#pragma unroll
          for (int h = 0; h < WARP_CH; h++) S.q[u][h] = ldg_stream_u4(ind4 + (b >> 2) + h);
        } else {                                                               // the last, partial chunk of the array
          uint32_t tmp[WARP_CS];
#pragma unroll
          for (int e = 0; e < WARP_CS; e++) tmp[e] = (b + e < E) ? indices[b + e] : 0u;
#pragma unroll
          for (int h = 0; h < WARP_CH; h++) S.q[u][h] = make_uint4(tmp[4 * h], tmp[4 * h + 1], tmp[4 * h + 2], tmp[4 * h + 3]);
        }
      }
    }
  }
}

// Probe one stage and append the kept slots to the staged stream (a per-warp scratch region in global memory that lives in L2).
// Stream order = window, lane, slot = ascending full-graph slot = CSR order (PS.cpp:420-422).  Per-row facts (kept count; with a
// self-edge insertion also "kept entries below v" and "v itself kept") are accumulated with one shared-memory atomic per lane.
// The caller guarantees room for a whole stage, so the OVF == false variant is straight-line code: the windows' probes interleave.
template <bool ADD_SELF, bool OVF>
__device__ __forceinline__ void scan_process(const ScanStage &S, uint32_t &cnt, const bucket_t *hb, const int hshift, const uint32_t novf,
                                             const uint32_t (&ovk)[WARP_OVK > 0 ? WARP_OVK : 1],
                                             const uint32_t *ovf, const uint32_t *nodes, uint32_t *rc, uint2 *sc_ent, unsigned short *sc_row,
                                             const uint32_t lt) {
#pragma unroll
  for (int u = 0; u < WARP_U; u++) {
    const uint32_t o = S.jb[u] - S.s[u];            // slots outside [s, s+len) (alignment padding, idle lanes) never match
    uint32_t nb[WARP_CS];
    bool hit[WARP_CS];
#pragma unroll
    for (int h = 0; h < WARP_CH; h++) { nb[4 * h] = S.q[u][h].x; nb[4 * h + 1] = S.q[u][h].y; nb[4 * h + 2] = S.q[u][h].z; nb[4 * h + 3] = S.q[u][h].w; }
#pragma unroll
    for (int e = 0; e < WARP_CS; e++) hit[e] = probe4(hb, hshift, nb[e]);
    if (!OVF && WARP_OVK > 0) {                                        // up to WARP_OVK overflow keys ride in registers (NONE32 = unused)
#pragma unroll
      for (int j = 0; j < WARP_OVK; j++)
#pragma unroll
        for (int e = 0; e < WARP_CS; e++) hit[e] |= ovk[j] == nb[e];
    }
    if (OVF) {                                                         // keys that did not fit their bucket
#pragma unroll 1
      for (uint32_t j = 1; j <= novf; j++) {
        const uint32_t kv = ovf[j];
#pragma unroll
        for (int e = 0; e < WARP_CS; e++) hit[e] |= kv == nb[e];
      }
    }
    uint32_t bal[WARP_CS], at = cnt, tot = 0, inc = 0;
    uint32_t v = 0;
    if (ADD_SELF) v = nodes[S.rw[u]];
#pragma unroll
    for (int e = 0; e < WARP_CS; e++) {
      hit[e] = hit[e] && (o + (uint32_t)e < S.d[u]);
      bal[e] = __ballot_sync(0xffffffffu, hit[e]);
      at += __popc(bal[e] & lt); tot += __popc(bal[e]);
      if (ADD_SELF) inc += hit[e] ? (1u + (nb[e] < v ? (1u << 14) : 0u) + (nb[e] == v ? (1u << 28) : 0u)) : 0u;
      else inc += hit[e] ? 1u : 0u;
    }
#pragma unroll
    for (int e = 0; e < WARP_CS; e++) {
      if (hit[e]) { sc_ent[at] = make_uint2(nb[e], S.jb[u] + (uint32_t)e); if (ADD_SELF) sc_row[at] = (unsigned short)S.rw[u]; }
      at += hit[e] ? 1u : 0u;
    }
    if (inc) atomicAdd(&rc[S.rw[u]], inc);
    cnt += tot;
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

// node count (and score-rank cut) of every subgraph of the launch: |{entries with rank < cut, id != root}| + 1
__global__ void __launch_bounds__(256) ppr_count_kernel(const SampleParams P, int *__restrict__ cnt, unsigned short *__restrict__ cut_out) {
  const int lane = threadIdx.x & 31;
  for (int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < P.num_subg; p += gridDim.x * (blockDim.x >> 5)) {
    const uint32_t t = P.roots[p];
    const unsigned long long off = P.ppr_ptr[t];
    const int len_all = (int)(P.ppr_ptr[t + 1] - off);
    const int size_neigh = len_all < P.k ? len_all : P.k;                                         // :576
    const uint32_t cut = ppr_cut(P, off, len_all, size_neigh, lane);
    uint32_t c = 0;
    for (int base = 0; base < len_all; base += 32) {
      const int i = base + lane;
      const bool selp = i < len_all && P.ppr_srank[off + i] < cut && P.ppr_sid[off + i] != t;
      c += __popc(__ballot_sync(0xffffffffu, selp));
    }
    if (lane == 0) { cnt[p] = (int)c + 1; cut_out[p] = (unsigned short)cut; }
  }
}

// exclusive prefix sum of cnt[0..P) into out[0..P] (one block; P is a few 10^4), total also to *total.  Thread i owns the contiguous
// items [i*ipt, (i+1)*ipt), ipt a multiple of 4: all loads of a thread are issued as 16-byte vectors before the first add.
#define SCAN_MAX_IPT 64
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int *__restrict__ cnt, const int P, int *__restrict__ out, long long *total) {
  __shared__ uint32_t sums[1024];
  __shared__ uint32_t warp_sums[33];
  const int ipt = ((P + 1023) / 1024 + 3) & ~3, b = threadIdx.x * ipt;
  uint32_t s = 0;
  if (ipt <= SCAN_MAX_IPT) {
    int4 v[SCAN_MAX_IPT / 4];
#pragma unroll
    for (int j = 0; j < SCAN_MAX_IPT / 4; j++) {
      v[j] = make_int4(0, 0, 0, 0);
      const int i = b + 4 * j;
      if (4 * j < ipt && i < P) {
        if (i + 3 < P) v[j] = *reinterpret_cast<const int4 *>(cnt + i);
        else { v[j].x = cnt[i]; if (i + 1 < P) v[j].y = cnt[i + 1]; if (i + 2 < P) v[j].z = cnt[i + 2]; }
      }
    }
#pragma unroll
    for (int j = 0; j < SCAN_MAX_IPT / 4; j++) s += (uint32_t)(v[j].x + v[j].y + v[j].z + v[j].w);
    sums[threadIdx.x] = s;
    __syncthreads();
    const uint32_t tot = block_exclusive_scan(sums, 1024, warp_sums);
    uint32_t run = sums[threadIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_MAX_IPT / 4; j++) {
      const int i = b + 4 * j;
      if (4 * j < ipt && i < P) {
        const int4 o = make_int4((int)run, (int)run + v[j].x, (int)run + v[j].x + v[j].y, (int)run + v[j].x + v[j].y + v[j].z);
        if (i + 3 < P) *reinterpret_cast<int4 *>(out + i) = o;
        else { out[i] = o.x; if (i + 1 < P) out[i + 1] = o.y; if (i + 2 < P) out[i + 2] = o.z; }
        run += (uint32_t)(v[j].x + v[j].y + v[j].z + v[j].w);
      }
    }
    if (threadIdx.x == 0) { out[P] = (int)tot; *total = (long long)tot; }
  } else {                                              // very large launches: plain loops
    const int e = min(P, b + ipt);
    for (int i = b; i < e; i++) s += (uint32_t)cnt[i];
    sums[threadIdx.x] = s;
    __syncthreads();
    const uint32_t tot = block_exclusive_scan(sums, 1024, warp_sums);
    uint32_t run = sums[threadIdx.x];
    for (int i = b; i < e; i++) { out[i] = (int)run; run += (uint32_t)cnt[i]; }
    if (threadIdx.x == 0) { out[P] = (int)tot; *total = (long long)tot; }
  }
}

template <bool ADD_SELF>
__global__ void __launch_bounds__(32, WARP_MIN_BLOCKS) ppr_induce_warp_kernel(const SampleParams P) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  uint32_t *const nodes = (uint32_t *)(smem_dyn + P.WL.nodes);
  uint2 *const rs = (uint2 *)(smem_dyn + P.WL.rs);                 // {row start, scanned length} of every node
  uint32_t *const cp = (uint32_t *)(smem_dyn + P.WL.cp);           // chunk prefix during the scan, local indptr afterwards
  uint32_t *const rc = (uint32_t *)(smem_dyn + P.WL.rc);           // per row: kept | kept below v << 14 | v kept << 28
  uint32_t *const hk = (uint32_t *)(smem_dyn + P.WL.hkeys);
  uint32_t *const ovf = (uint32_t *)(smem_dyn + P.WL.ovf);         // [0] = count, [1..WARP_OVF_CAP] = keys whose bucket was full
  uint32_t *const rlo = (uint32_t *)(smem_dyn + P.WL.rlo);         // ADD_SELF only: first staged entry / insert position / bug column per row
  uint32_t *const rins = (uint32_t *)(smem_dyn + P.WL.rins);
  uint32_t *const rbug = (uint32_t *)(smem_dyn + P.WL.rbug);
  // staged kept edges {neighbour id, full-graph slot} (+ row), CSR order: per-warp scratch in global memory.  A warp reuses its few KB
  // for every subgraph, so the region stays in L2; keeping it out of shared memory is worth ~10 more resident warps per SM.
  uint2 *const sc_ent = (uint2 *)(P.w_scratch + (size_t)blockIdx.x * P.w_scratch_stride);
  unsigned short *const sc_row = (unsigned short *)(sc_ent + P.w_ecap);
  const bucket_t *const hb = (const bucket_t *)hk;
  const uint4 *const ind4 = (const uint4 *)P.indices;
  const int lane = threadIdx.x;
  const uint32_t FULL = 0xffffffffu;
  const uint32_t lt = lanemask_lt(), le = lanemask_le();
  const uint32_t nbuckets = (uint32_t)P.w_hbuckets;
  const int hshift = P.w_hshift;
  const uint32_t ecap = (uint32_t)P.w_ecap;
  const bool ext = !ADD_SELF && !P.fixed_mode;                     // PS.cpp:401: slot `e` is tested whenever no self edge is inserted
  const uint32_t E = P.num_edges, E_al = E & ~(uint32_t)(WARP_CS - 1);

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
#pragma unroll 1
    for (int base = 0; base < len_all; base += 32) {
      const int i = base + lane;
      bool sel = false;
      uint32_t id = 0;
      if (i < len_all) { sel = P.ppr_srank[off + i] < cut; id = P.ppr_sid[off + i]; }
      const bool selp = sel && id != t;
      const uint32_t m = __ballot_sync(FULL, selp), mb = __ballot_sync(FULL, selp && id < t), mr = __ballot_sync(FULL, sel && id == t);
      if (sel) {                                        // entries below the root keep their rank, the root's slot follows them, the rest shift by one
        const uint32_t at = run + __popc(m & lt) + ((selp && id > t) ? 1u : 0u);
        nodes[at] = id; rs[at] = P.ppr_srow[off + i];
        P.orig_node[node_base + at] = id; P.ppr_out[node_base + at] = P.ppr_sscore[off + i];
      }
      run += __popc(m); n_below += __popc(mb); root_in |= (mr != 0);
    }
    if (!root_in && lane == 0) {                        // root not among the selected entries: -1, or scores[0] when the row has <= 1 entry (:574,:581)
      nodes[n_below] = t;
      P.orig_node[node_base + n_below] = t;
      P.ppr_out[node_base + n_below] = (size_neigh <= 1 && len_all > 0) ? P.ppr_scores[off] : -1.f;
      const uint32_t s = P.indptr[t];
      rs[n_below] = make_uint2(s, P.indptr[t + 1] - s);
    }
    const int n = (int)run + 1;                         // == node_ptr[p+1] - node_ptr[p]

    // ---------------- B: membership hash, chunk prefix ----------------
#pragma unroll 1
    for (uint32_t i = lane; i < nbuckets * WARP_BK / 4; i += 32) reinterpret_cast<uint4 *>(hk)[i] = make_uint4(NONE32, NONE32, NONE32, NONE32);
    if (lane == 0) ovf[0] = 0;
    __syncwarp();
    uint32_t total = 0;                                 // chunks of the subgraph
#pragma unroll 1
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      uint32_t ch = 0;
      if (i < n) {
        const uint32_t v = nodes[i];
        const uint32_t b = (v * 2654435761u) >> hshift;
        bool done = false;
#pragma unroll
        for (int j = 0; j < WARP_BK; j++)
          if (!done && atomicCAS(&hk[WARP_BK * b + j], NONE32, v) == NONE32) done = true;
        if (!done) { const uint32_t q = atomicAdd(&ovf[0], 1u); if (q < WARP_OVF_CAP) ovf[1 + q] = v; }
        uint2 r = rs[i];
        if (ext && r.x + r.y < E) { r.y += 1; rs[i] = r; }                                       // the PS.cpp:401 slot joins the row
        ch = r.y ? (((r.x + r.y + (WARP_CS - 1)) >> WARP_CSH) - (r.x >> WARP_CSH)) : 1u;          // aligned chunks; every row owns >= 1
        rc[i] = 0;
      }
      const uint32_t x = warp_incl_scan(ch, lane);
      if (i < n) cp[i] = total + x - ch;
      total += __shfl_sync(FULL, x, 31);
    }
    if (lane == 0) cp[n] = total;
    __syncwarp();
    p_next = __shfl_sync(FULL, (int)tk_next, 0);
    if (p_next < P.num_subg) { t_next = P.roots[p_next]; node_base_next = P.node_ptr[p_next]; cut_next = P.w_cut[p_next]; }    // consumed after the scan
    const uint32_t novf = ovf[0];
    bool bail = novf > WARP_OVF_CAP;

    // ---------------- C: one pass over the rows in chunk space ----------------
    uint32_t cnt = 0;
    if (!bail) {
      int rbase = 0;                                    // cp[rbase] <= first chunk of the window <= cp[rbase+1]
      const uint32_t step = 32u * WARP_U, room = 32u * WARP_U * WARP_CS;      // a stage can keep at most `room` entries
      ScanStage SA;
      uint32_t ovk[WARP_OVK > 0 ? WARP_OVK : 1];
#pragma unroll
      for (int j = 0; j < (WARP_OVK > 0 ? WARP_OVK : 1); j++) ovk[j] = (WARP_OVK > 0 && (uint32_t)j < novf) ? ovf[1 + j] : NONE32;
      if (novf <= WARP_OVK) {
#if WARP_DB
        // stages double-buffered in registers: the loads of stage i+1 are in flight while stage i is probed
        ScanStage SB;
        scan_load(SA, 0u, total, rbase, n, lane, le, cp, rs, ind4, P.indices, E, E_al);
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < total; c0 += 2u * step) {
          if (cnt + 2u * room > ecap) { bail = true; break; }
          const bool more = c0 + step < total;
          if (more) scan_load(SB, c0 + step, total, rbase, n, lane, le, cp, rs, ind4, P.indices, E, E_al);
          scan_process<ADD_SELF, false>(SA, cnt, hb, hshift, 0u, ovk, ovf, nodes, rc, sc_ent, sc_row, lt);
          if (!more) break;
          if (c0 + 2u * step < total) scan_load(SA, c0 + 2u * step, total, rbase, n, lane, le, cp, rs, ind4, P.indices, E, E_al);
          scan_process<ADD_SELF, false>(SB, cnt, hb, hshift, 0u, ovk, ovf, nodes, rc, sc_ent, sc_row, lt);
        }
#else
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < total; c0 += step) {
          if (cnt + room > ecap) { bail = true; break; }
          scan_load(SA, c0, total, rbase, n, lane, le, cp, rs, ind4, P.indices, E, E_al);
          scan_process<ADD_SELF, false>(SA, cnt, hb, hshift, 0u, ovk, ovf, nodes, rc, sc_ent, sc_row, lt);
        }
#endif
      } else {
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < total; c0 += step) {
          if (cnt + room > ecap) { bail = true; break; }
          scan_load(SA, c0, total, rbase, n, lane, le, cp, rs, ind4, P.indices, E, E_al);
          scan_process<ADD_SELF, true>(SA, cnt, hb, hshift, novf, ovk, ovf, nodes, rc, sc_ent, sc_row, lt);
        }
      }
    }
    __syncwarp();
    if (p_next < P.num_subg) { off_next = P.ppr_ptr[t_next]; row_end_next = P.ppr_ptr[t_next + 1]; }      // consumed before the emit

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
          if (!ADD_SELF) c_i = w;                       // nothing is inserted: the staged stream is the CSR
          else {
            k_i = w & 0x3fffu;
            const bool present = (w >> 28) != 0;        // the row already holds its self loop (:386-400)
            uint32_t bsub = NONE32;
            if (present && !P.fixed_mode) {             // no insertion => slot e is tested too (:401)
              const uint32_t e = rs[i].x + rs[i].y;
              if (e < E) {
                const uint32_t nb = __ldg(P.indices + e);
                bool h = probe4(hb, hshift, nb);
                for (uint32_t j = 1; j <= novf; j++) h |= ovf[j] == nb;
                if (h) bsub = sub_of(nodes, n, nb);
              }
            }
            rins[i] = present ? NONE32 : ((w >> 14) & 0x3fffu); rbug[i] = bsub;
            c_i = k_i + (present ? 0u : 1u) + (bsub != NONE32 ? 1u : 0u);
          }
        }
        const uint32_t x = warp_incl_scan(c_i, lane);
        if (i < n) cp[i] = carry + x - c_i;
        carry += __shfl_sync(FULL, x, 31);
        if (ADD_SELF) {
          const uint32_t y = warp_incl_scan(k_i, lane);
          if (i < n) rlo[i] = carry_k + y - k_i;
          carry_k += __shfl_sync(FULL, y, 31);
        }
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
      if (o128 < ln * 8ull + 128ull) prefetch_l2((const char *)(P.ppr_srow + off_next) + o128);
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
    } while (0);
    p = p_next; t = t_next; node_base = node_base_next; cut = cut_next; off = off_next; row_end = row_end_next;
  }
}
