// ppr_warp_kernel.cuh -- the single-root PPR fast path of the fused select + induce step: ONE WARP per subgraph.
//
// Same contract as sample_induce_kernel (sampler_kernels.cuh) for  method == ppr, one root per subgraph, no hop/drnl
// labels  -- the configuration every PPR node-task run of the reference uses (shaDow/minibatch.py:373-375; PS.cpp:565-595
// + PS.cpp:350-453).  What is different is the mapping to the machine:
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
#define WARP_U 4               // chunk windows (32 chunks each) whose Bloom tests are issued together
#endif
#ifndef WARP_MIN_BLOCKS
#define WARP_MIN_BLOCKS 20     // one warp per CTA: resident warps per SM the register budget must allow
#endif
#ifndef WARP_K
#define WARP_K 128             // chunks per TMA stage (2 KB); a multiple of 32 * WARP_U
#endif
#define WARP_NST 2             // stages: one being scanned, one in flight
#define WARP_CS 4              // slots per chunk (16 bytes)
#define WARP_CSH 2
#define WARP_QCAP 192          // candidate queue: < 32 left over + at most 128 new per window (+ head room)
#define WARP_CBITS 19          // chunks of one subgraph < 2^19 (plan: ncap * (dmax / 4 + 2) must fit)

__device__ __forceinline__ uint32_t lanemask_le() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
  return m;
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

// ---- 1-D TMA: cp.async.bulk global -> shared, completion counted in bytes on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

// ---- level 1: blocked Bloom filter, NF words per lane (32 * NF words = 1024 * NF bits), one SHFL per word-register, 2 bits per key ----
__device__ __forceinline__ uint32_t bloom_hash(uint32_t key) { return key * 2654435761u; }
__device__ __forceinline__ uint32_t bloom_hash2(uint32_t key) { return __umulhi(key, 0x85EBCA6Bu); }
template <int NF>
__device__ __forceinline__ uint32_t bloom_word_index(uint32_t h) { return h >> (NF == 1 ? 27 : (NF == 2 ? 26 : 25)); }
__device__ __forceinline__ uint32_t bloom_bits(uint32_t h, uint32_t g) { return __funnelshift_l(0u, 1u, h) | __funnelshift_l(0u, 1u, g); }   // 1 << h % 32 | 1 << g % 32
// bit 0 of the result: both bits of `key` are set in its filter word
template <int NF>
__device__ __forceinline__ uint32_t bloom_test(const uint32_t (&f)[NF], uint32_t key) {
  const uint32_t h = bloom_hash(key), g = bloom_hash2(key), w = bloom_word_index<NF>(h);
  uint32_t fv = __shfl_sync(0xffffffffu, f[0], (int)w);       // source lane = w mod 32
  if (NF >= 2) { const uint32_t f1 = __shfl_sync(0xffffffffu, f[NF >= 2 ? 1 : 0], (int)w); fv = (w & 32u) ? f1 : fv; }
  if (NF == 4) {
    const uint32_t f2 = __shfl_sync(0xffffffffu, f[NF == 4 ? 2 : 0], (int)w), f3 = __shfl_sync(0xffffffffu, f[NF == 4 ? 3 : 0], (int)w);
    fv = (w & 64u) ? ((w & 32u) ? f3 : f2) : fv;
  }
  return __funnelshift_r(fv, 0u, h) & __funnelshift_r(fv, 0u, g);        // (fv >> h % 32) & (fv >> g % 32)
}

// ---- level 2, on 32 queued candidate SLOTS at a time (one per lane): two interleaved bisections -- the row that owns the slot's chunk (over
// the chunk prefix cp[]) and the key's place in the sorted node list (exact membership AND the sub id, no hash table) -- then the row-range
// check (the alignment padding of a chunk belongs to the neighbouring rows).  Kept edges are appended to the staged stream {sub id, full-graph
// slot} (a per-warp scratch region in global memory that lives in L2).  Queue order = (window, lane, slot) = ascending full-graph slot inside
// every row = CSR order (PS.cpp:420-422).  Per-row facts (kept count; with a self-edge insertion also "kept entries below v" and "v itself
// kept") are accumulated with one shared-memory atomic per kept edge. ----
template <bool ADD_SELF>
__device__ __forceinline__ void drain_batch(const uint2 *cq, const uint32_t count, const uint32_t *nodes, const uint32_t *cp, const int n,
                                            const uint32_t topstep, const uint2 *rs, uint32_t *rc, uint2 *sc_ent, unsigned short *sc_row,
                                            uint32_t &cnt, const int lane, const uint32_t lt) {
  const bool active = (uint32_t)lane < count;
  uint2 ent = make_uint2(NONE32, 0u);                  // {key, flat slot = chunk * 4 + position in the chunk}
  if (active) ent = cq[lane];
  const uint32_t key = ent.x, c = ent.y >> WARP_CSH;
  uint32_t row = 0, sub = 0, kv = nodes[0];            // last row with cp[row] <= c (cp[n] = total > c); last node with nodes[sub] <= key
#pragma unroll 1
  for (uint32_t step = topstep; step; step >>= 1) {
    const uint32_t ir = min(row + step, (uint32_t)n), is = min(sub + step, (uint32_t)n - 1u);
    const uint32_t cv = cp[ir], nv = nodes[is];
    if (cv <= c) row = ir;
    if (nv <= key) { sub = is; kv = nv; }
  }
  const uint2 r = rs[row];
  const uint32_t off = ent.y - (cp[row] << WARP_CSH) - (r.x & 3u);        // slot relative to the row start (wraps for the padding in front of it)
  const bool hit = active && kv == key && off < r.y;                       // member of the node set, slot inside [s, s + len)
  const uint32_t bal = __ballot_sync(0xffffffffu, hit);
  if (hit) {
    const uint32_t at = cnt + __popc(bal & lt);
    sc_ent[at] = make_uint2(sub, r.x + off);
    if (ADD_SELF) { sc_row[at] = (unsigned short)row; atomicAdd(&rc[row], 1u + (sub < row ? (1u << 14) : 0u) + (sub == row ? (1u << 28) : 0u)); }   // ids sorted: sub < row <=> id < v
    else atomicAdd(&rc[row], 1u);
  }
  cnt += __popc(bal);
}

// sub id of `key` in the sorted node list, or NONE32
__device__ __forceinline__ uint32_t find_sub(const uint32_t *nodes, const int n, const uint32_t topstep, const uint32_t key) {
  uint32_t sub = 0, kv = nodes[0];
#pragma unroll 1
  for (uint32_t step = topstep; step; step >>= 1) {
    const uint32_t is = min(sub + step, (uint32_t)n - 1u), nv = nodes[is];
    if (nv <= key) { sub = is; kv = nv; }
  }
  return kv == key ? sub : NONE32;
}

// Level 1 on one window: the Bloom verdict of each of a lane's 4 slots; 4 ballots append the slots that may hold a member to the queue in
// (lane, slot) order; full warps of queued candidates are drained at once.
template <bool ADD_SELF>
__device__ __forceinline__ void scan_window(const uint32_t v0, const uint32_t v1, const uint32_t v2, const uint32_t v3, const uint4 q,
                                            const uint32_t code, uint2 *cq, uint32_t &qn, const uint32_t *nodes, const uint32_t *cp, const int n,
                                            const uint32_t topstep, const uint2 *rs, uint32_t *rc, uint2 *sc_ent, unsigned short *sc_row,
                                            uint32_t &cnt, const int lane, const uint32_t lt) {
  const uint32_t b0 = __ballot_sync(0xffffffffu, v0 != 0u), b1 = __ballot_sync(0xffffffffu, v1 != 0u);
  const uint32_t b2 = __ballot_sync(0xffffffffu, v2 != 0u), b3 = __ballot_sync(0xffffffffu, v3 != 0u);
  uint32_t at = qn + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
  if (v0) cq[at] = make_uint2(q.x, code);
  at += v0;
  if (v1) cq[at] = make_uint2(q.y, code + 1u);
  at += v1;
  if (v2) cq[at] = make_uint2(q.z, code + 2u);
  at += v2;
  if (v3) cq[at] = make_uint2(q.w, code + 3u);
  qn += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
  if (qn >= 32u) {
    __syncwarp();
    uint32_t head = 0;
#pragma unroll 1
    do {
      drain_batch<ADD_SELF>(cq + head, 32u, nodes, cp, n, topstep, rs, rc, sc_ent, sc_row, cnt, lane, lt);
      head += 32u; qn -= 32u;
    } while (qn >= 32u);
    uint2 tk = make_uint2(0u, 0u);                   // < 32 left: move the tail to the front
    if ((uint32_t)lane < qn) tk = cq[head + lane];
    __syncwarp();
    if ((uint32_t)lane < qn) cq[lane] = tk;
    __syncwarp();
  }
}

// One TMA stage = chunks [c0, hi_c) of the subgraph's flat chunk space.  The rows tile that space back to back, so every row (piece) that
// falls into the stage is ONE cp.async.bulk issued by one lane: source = the row's aligned chunks in the full-graph CSR, destination = its place
// in the stage buffer.  rb = the row that owns chunk c0 on entry, the row that owns chunk hi_c on exit.
__device__ __forceinline__ void issue_stage(const uint32_t c0, const uint32_t hi_c, int &rb, const int n, const uint32_t total, const int lane,
                                            const uint32_t *cp, const uint2 *rs, const uint4 *ind4, const uint32_t last_chunk, uint4 *buf, uint64_t *bar) {
  if (lane == 0) mbar_expect_tx(bar, (hi_c - c0) * 16u);
  __syncwarp();
  int rb_next = rb;
#pragma unroll 1
  for (int r0 = rb;; r0 += 32) {
    const int r = r0 + lane;
    uint32_t lo = total, hi = total;
    if (r < n) { lo = cp[r]; hi = cp[r + 1]; }
    if (lo < hi_c) {                                    // (hi > c0 holds for every row from rb on; every row owns >= 1 chunk)
      const uint32_t a = max(lo, c0), b = min(hi, hi_c);
      // (an empty row behind the last edge owns a chunk too: its source is clamped to the last chunk of the array)
      bulk_g2s(buf + (a - c0), ind4 + min(rs[r].x >> WARP_CSH, last_chunk) + (a - lo), (b - a) * 16u, bar);
    }
    const uint32_t own = __ballot_sync(0xffffffffu, lo <= hi_c && hi_c < hi);       // at most one row owns chunk hi_c
    if (own) rb_next = r0 + __ffs(own) - 1;
    if (!__shfl_sync(0xffffffffu, lo < hi_c ? 1 : 0, 31)) break;                     // rows are ordered by their first chunk
  }
  rb = rb_next;
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

// NF = Bloom words per lane (1: ncap <= 192, 2: <= 448, 4 above)
template <bool ADD_SELF, int NF>
__global__ void __launch_bounds__(32, WARP_MIN_BLOCKS) ppr_induce_warp_kernel(const SampleParams P) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  uint32_t *const nodes = (uint32_t *)(smem_dyn + P.WL.nodes);
  uint2 *const rs = (uint2 *)(smem_dyn + P.WL.rs);                 // {row start, scanned length} of every node
  uint32_t *const cp = (uint32_t *)(smem_dyn + P.WL.cp);           // chunk prefix during the scan, local indptr afterwards
  uint32_t *const rc = (uint32_t *)(smem_dyn + P.WL.rc);           // per row: kept | kept below v << 14 | v kept << 28
  uint32_t *const fw = (uint32_t *)(smem_dyn + P.WL.bloom);        // Bloom words while they are built (32 * NF)
  uint2 *const cq = (uint2 *)(smem_dyn + P.WL.queue);              // candidate queue: {key, flat slot} of every slot that passed the Bloom filter
  uint4 *const stage = (uint4 *)(smem_dyn + P.WL.stage);           // WARP_NST TMA stage buffers of WARP_K chunks
  uint64_t *const bars = (uint64_t *)(smem_dyn + P.WL.bars);       // one mbarrier per stage
  uint32_t *const rlo = (uint32_t *)(smem_dyn + P.WL.rlo);         // ADD_SELF only: first staged entry / insert position / bug column per row
  uint32_t *const rins = (uint32_t *)(smem_dyn + P.WL.rins);
  uint32_t *const rbug = (uint32_t *)(smem_dyn + P.WL.rbug);
  // staged kept edges {sub id, full-graph slot} (+ row), CSR order: per-warp scratch in global memory.  A warp reuses its few KB
  // for every subgraph, so the region stays in L2.
  uint2 *const sc_ent = (uint2 *)(P.w_scratch + (size_t)blockIdx.x * P.w_scratch_stride);
  unsigned short *const sc_row = (unsigned short *)(sc_ent + P.w_ecap);
  const uint4 *const ind4 = (const uint4 *)P.indices;
  const int lane = threadIdx.x;
  const uint32_t FULL = 0xffffffffu;
  const uint32_t lt = lanemask_lt();
  const uint32_t ecap = (uint32_t)P.w_ecap;
  const bool ext = !ADD_SELF && !P.fixed_mode;                     // PS.cpp:401: slot `e` is tested whenever no self edge is inserted
  const uint32_t E = P.num_edges;

  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < WARP_NST; j++) mbar_init(bars + j, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t ph = 0;                                                  // bit j = parity the next wait on stage j expects

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
    const int n = (int)run + 1;                         // == node_ptr[p+1] - node_ptr[p]; the root is node n_below

    // ---------------- B: Bloom filter, chunk prefix ----------------
#pragma unroll
    for (int j = 0; j < NF; j++) fw[32 * j + lane] = 0u;
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
    uint32_t f[NF];
#pragma unroll
    for (int j = 0; j < NF; j++) f[j] = fw[32 * j + lane];
    p_next = __shfl_sync(FULL, (int)tk_next, 0);
    if (p_next < P.num_subg) { t_next = P.roots[p_next]; node_base_next = P.node_ptr[p_next]; cut_next = P.w_cut[p_next]; }    // consumed after the scan
    bool bail = total >= (1u << WARP_CBITS);
    uint32_t topstep = 0;                               // bisections over cp[0..n-1] / nodes[0..n-1]: largest power of two <= n - 1
    if (n > 1) topstep = 1u << (31 - __clz(n - 1));

    // ---------------- C: one pass over the rows in chunk space, the rows staged in shared memory by TMA ----------------
    uint32_t cnt = 0, qn = 0;
    if (!bail) {
      const uint32_t nstages = (total + WARP_K - 1) / WARP_K;
      const uint32_t room = 32u * (WARP_U * WARP_CS + 1);          // what a group of windows can keep at most (incl. the queue's tail)
      int rb = 0;                                                   // row that owns the first chunk of the next stage to issue
      uint32_t issued = 0;
      for (; issued < nstages && issued < WARP_NST; issued++)
        issue_stage(issued * WARP_K, min(total, (issued + 1) * WARP_K), rb, n, total, lane, cp, rs, ind4, (E - 1u) >> WARP_CSH, stage + (issued % WARP_NST) * WARP_K, bars + (issued % WARP_NST));
      uint32_t i = 0;
#pragma unroll 1
      for (; i < nstages && !bail; i++) {
        const uint32_t st = i % WARP_NST, c0 = i * WARP_K, hi_c = min(total, c0 + WARP_K);
        mbar_wait(bars + st, (ph >> st) & 1u);
        ph ^= 1u << st;
        const uint4 *const buf = stage + st * WARP_K;
#pragma unroll 1
        for (uint32_t w0 = 0; c0 + w0 < hi_c; w0 += 32u * WARP_U) {
          if (cnt + room > ecap) { bail = true; break; }
          uint4 q[WARP_U];
          uint32_t cc[WARP_U], v[WARP_U][WARP_CS];
#pragma unroll
          for (int u = 0; u < WARP_U; u++) {
            const uint32_t ci = w0 + 32u * u + lane;
            cc[u] = c0 + ci;
            q[u] = buf[ci];                             // (chunks beyond hi_c hold stale data: masked below)
          }
#pragma unroll
          for (int u = 0; u < WARP_U; u++) {            // all Bloom tests of the group first: 4 * WARP_U independent SHFLs in flight
            const uint32_t ok = cc[u] < hi_c ? 1u : 0u;
            v[u][0] = bloom_test<NF>(f, q[u].x) & ok; v[u][1] = bloom_test<NF>(f, q[u].y) & ok;
            v[u][2] = bloom_test<NF>(f, q[u].z) & ok; v[u][3] = bloom_test<NF>(f, q[u].w) & ok;
          }
#pragma unroll
          for (int u = 0; u < WARP_U; u++)
            scan_window<ADD_SELF>(v[u][0], v[u][1], v[u][2], v[u][3], q[u], cc[u] << WARP_CSH, cq, qn, nodes, cp, n, topstep, rs, rc, sc_ent, sc_row, cnt, lane, lt);
        }
        __syncwarp();                                   // every lane is done reading the buffer before TMA refills it
        if (!bail && issued < nstages) {
          issue_stage(issued * WARP_K, min(total, (issued + 1) * WARP_K), rb, n, total, lane, cp, rs, ind4, (E - 1u) >> WARP_CSH, stage + st * WARP_K, bars + st);
          issued++;
        }
      }
      if (bail) {                                       // copies still in flight must land before the buffers are reused
        for (; i < issued; i++) { const uint32_t st = i % WARP_NST; mbar_wait(bars + st, (ph >> st) & 1u); ph ^= 1u << st; }
      } else if (qn) {
        __syncwarp();
        drain_batch<ADD_SELF>(cq, qn, nodes, cp, n, topstep, rs, rc, sc_ent, sc_row, cnt, lane, lt);
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
              if (e < E) bsub = find_sub(nodes, n, topstep, __ldg(P.indices + e));
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
      P.target[p] = (int)(node_base + n_below);
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
        pos = edge_base + cp[row] + (g - rlo[row]) + ((rins[row] != NONE32 && ent.x > row) ? 1 : 0);
      }
      P.indices_out[pos] = (int)(node_base + ent.x);
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
