// sampler.cu -- host side of the sampler C ABI (include/shadow_b200.h) + kernel launches.
// Mirrors class ParallelSampler (backend/ParallelSampler.h:25-158) with the graph, the PPR tables, the target
// list and the results all resident in HBM.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "sampler_kernels.cuh"
#include "ppr_warp_kernel.cuh"

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void shadow_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char *shadow_last_error(void) { return g_err; }
extern "C" int shadow_version(void) { return 100; }

// glibc rand() TYPE_3 generator (published algorithm, glibc stdlib/random_r.c); the reference's khop draws
// from the process-global instance (PS.cpp:534) seeded in the ctor (PS.h:49-53).
struct GlibcRand {
  uint32_t r[31];
  int f, b;
  void seed(uint32_t s) {
    if (s == 0) s = 1;
    r[0] = s;
    int32_t w = (int32_t)s;
    for (int i = 1; i < 31; i++) {
      long hi = w / 127773, lo = w % 127773;
      w = (int32_t)(16807 * lo - 2836 * hi);
      if (w < 0) w += 2147483647;
      r[i] = (uint32_t)w;
    }
    f = 3; b = 0;
    for (int i = 0; i < 310; i++) next();
  }
  inline uint32_t next() {
    r[f] += r[b];
    uint32_t o = r[f] >> 1;
    f = (f + 1) % 31; b = (b + 1) % 31;
    return o;
  }
};

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    // 25 % head room: batch sizes fluctuate by a few percent from call to call, and every cudaFree / cudaMalloc pair is a device-wide
    // synchronisation (it used to drain the training pipeline for 4-13 ms whenever a super-batch set a new maximum)
    size_t want = std::max(need + need / 4, (size_t)256);
    if (cudaMalloc(&p, want) != cudaSuccess) {
      cudaGetLastError();
      want = std::max(need, (size_t)256);
      if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return -1; }
    }
    bytes = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct Result {              // one ensemble branch of one call: SubgraphStructVec (G.h:59-97) in HBM
  DevBuf node_ptr, edge_ptr, rowptr, indices, target, num_target, orig_node, orig_edge, hop, drnl, ppr, sync;
  DevBuf row_span, edge_span, indices_raw, orig_edge_raw, redo;
  bool canon_valid = false;
  long long cap_nodes = 0, cap_edges = 0;
  int cap_subg = 0;
  long long *totals_host = nullptr;   // pinned mirror of totals[0..5] ([3] = subgraphs the warp fast path handed to the generic kernel, [4] / [5] = its staged edges / scanned chunks)
  bool fast_launch = false;
  // description of the launch (for validation / re-run)
  shadow_sampler_cfg cfg;
  uint32_t idx_start = 0, idx_end = 0;
  int num_subg = 0;
  bool pending = false, valid = false;
  long long total_nodes = 0, total_edges = 0, rand_draws = 0;
  uint32_t philox_epoch = 0;
};

struct KernCfg { const void *f; int threads; size_t smem; int bps; };

struct shadow_sampler {
  int device = 0, num_sms = 0;
  cudaStream_t stream = nullptr;
  uint32_t *indptr = nullptr, *indices = nullptr;
  bool owns_graph = false, graph_dropped = false;
  uint32_t N = 0, E = 0;
  long long dmax = -1;
  DevBuf targets;
  bool owns_targets = true;
  const uint32_t *targets_ptr = nullptr;
  uint32_t T = 0, idx_root = 0, epoch = 0;
  int per_batch = 0, num_ens = 1, num_ring = 1, cur = -1;
  int seed = 0;
  GlibcRand rng;
  // ppr tables
  DevBuf ppr_ptr, ppr_neighs, ppr_scores, ppr_sid, ppr_sscore, ppr_srank, ppr_srow;
  bool has_ppr = false, ppr_sorted = false;
  long long ppr_maxlen = 0;
  int warp_ecap_mult = 16;                 // staged edges per node of the PPR fast path (grows when launches need many redos)
  DevBuf wscratch;                         // its per-warp staging scratch
  // symmetric-graph variant of the fast path: 0 = not checked yet, 1 = graph is symmetric with strictly ascending rows (sym_rev valid), -1 = it is not
  int sym_state = 0;
  bool supper_valid = false;               // ppr_supper matches the installed tables
  DevBuf sym_rev, ppr_supper;
  // Every mirrored edge costs a sub-id search, a cursor update and a random 4-byte read of sym_rev[]: about twice a directly kept edge, while
  // a scanned slot costs a sixth of one.  Measured break-even (S-products stand-ins, uniform vs planted partition): ~14 % of the scanned
  // slots kept; the symmetric variant runs while the previous launch stayed below SHADOW_SYM_RATIO (default 0.12).
  double kept_ratio = 0.0;
  long long last_redo = 0;                 // subgraphs of the last validated launch that went through the redo kernel
  bool last_sym = false;                   // the last launch ran the symmetric (upper-triangle) variant of the fast path
  cudaEvent_t kev0 = nullptr, kev1 = nullptr;   // around the last fast-path main kernel (diagnostics: shadow_sampler_last_kernel_ms)
  cudaEvent_t sev0 = nullptr, sev1 = nullptr;   // around all GPU work of that launch (reset, count, scan, main kernel, redo)
  bool kev_valid = false;
  std::vector<KernCfg> kcache;
  std::vector<std::vector<Result>> ring;   // [num_ring][num_ens]
  DevBuf rand_stream, rand_off, gws;
  std::vector<uint32_t> rand_host;
};

static inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

// cudaFuncSetAttribute(max dynamic shared memory) + resident blocks per SM of a kernel, asked once per (kernel, block size, bytes): both are
// host-side driver calls of several microseconds, and a sampler launch is a handful of short kernels the GPU would otherwise wait between
static int kern_cfg(std::vector<KernCfg> &cache, const void *f, int threads, size_t smem, int *bps) {
  for (const KernCfg &c : cache)
    if (c.f == f && c.threads == threads && c.smem == smem) { if (bps) *bps = c.bps; return 0; }
  CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int b = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, f, threads, smem));
  cache.push_back(KernCfg{f, threads, smem, b});
  if (bps) *bps = b;
  return 0;
}

struct Caps { int ncap, ccap, ccap2, acap, acap2, hcap, hshift, ecap; WsLayout L; long long max_draws; };

static int graph_dmax(shadow_sampler *s, long long *out);

static int plan_caps(shadow_sampler *s, const shadow_sampler_cfg &c, Caps *o) {
  const long long N = s->N, nr = c.num_roots;
  long long ncap = nr, ccap = nr, acap = nr, max_draws = 0;
  if (c.method == SHADOW_PPR || c.method == SHADOW_PPR_ST) {
    ncap = nr * ((long long)c.k + 1);
    ccap = nr * ((long long)c.k + 2);
    acap = 1;
    if (c.method == SHADOW_PPR_ST) max_draws = nr * s->ppr_maxlen;
  } else if (c.method == SHADOW_KHOP) {
    long long fan;
    if (c.budget < 0) { long long d; int rc = graph_dmax(s, &d); if (rc) return rc; fan = d; }
    else fan = c.budget;
    long long lvl = nr; acap = nr; ccap = nr;
    for (int l = 0; l < c.depth; l++) {
      long long raw = lvl * fan;
      if (c.budget < 0) raw = std::min<long long>(raw, (long long)s->E);
      ccap = std::max(ccap, raw);
      if (c.budget >= 0) max_draws += lvl * fan;
      lvl = std::min<long long>(raw, N);
      acap += lvl;
    }
    ncap = std::min<long long>(acap, N);
    ncap = std::max<long long>(ncap, nr);
  }
  if (ncap > (1ll << 28) || ccap > (1ll << 28) || acap > (1ll << 28)) FAIL(SHADOW_ECAP, "sampler scope too large (ncap=%lld)", ncap);
  o->ncap = (int)ncap; o->ccap = (int)ccap; o->acap = (int)acap;
  o->ccap2 = next_pow2((int)ccap); o->acap2 = next_pow2((int)acap);
  o->hcap = std::max(64, next_pow2(4 * (int)ncap));          // slots; buckets of 4
  int lg = 0; while ((1 << lg) < o->hcap / 4) lg++;
  o->hshift = 32 - lg;                                     // hash -> bucket index
  o->max_draws = max_draws;
  WsLayout &L = o->L;
  uint32_t off = 0;
  auto take = [&](size_t bytes) { uint32_t r = off; off = align16(off + (uint32_t)bytes); return r; };
  // region 1: node-set scratch (dead once the node set is final) -- the edge staging area of the scan overlays it
  L.keys = take((size_t)o->ccap2 * 8);
  L.cval = take((size_t)o->ccap * 4);
  L.level = take((size_t)o->ncap * 4);
  L.all = take((size_t)o->acap2 * 4);
  const uint32_t region1 = off;
  // staged-edge capacity of the single-pass path (15-bit sub ids); 0 => generic two-pass path
  {
    const char *env = getenv("SHADOW_ECAP_MULT");
    const int mult = env ? atoi(env) : 6;
    o->ecap = (o->ncap <= 65535 && mult > 0) ? std::min(16384, std::max(512, mult * o->ncap)) : 0;
    L.st = 0;
    off = std::max(region1, align16((uint32_t)o->ecap * 8));
  }
  L.nodes = take((size_t)o->ncap * 4);
  L.pprv = take((size_t)o->ncap * 4);
  L.hkeys = take((size_t)o->hcap * 4);
  L.st_row = take((size_t)o->ecap * 2);
  L.row_s = take((size_t)o->ncap * 4);
  L.row_e = take((size_t)o->ncap * 4);
  L.row_cnt = take(((size_t)o->ncap + 1) * 4);
  L.row_ins = take((size_t)o->ncap * 4);
  L.rowd = take(((size_t)o->ncap + 1) * 16);
  L.row_first = take((size_t)o->ncap * 4);
  L.row_last = take((size_t)o->ncap * 4);
  L.row_fgt = take((size_t)o->ncap * 4);
  const bool bfs = (c.aug & (SHADOW_AUG_HOPS | SHADOW_AUG_DRNLS)) != 0;
  L.dist = take(bfs ? (size_t)o->ncap * 4 : 0);
  L.fr_a = take(bfs ? (size_t)o->ncap * 4 : 0);
  L.fr_b = take(bfs ? (size_t)o->ncap * 4 : 0);
  if (off > 200 * 1024) o->ecap = 0;               // too big for shared memory anyway: global-workspace variant, generic path
  L.bytes = off;
  return 0;
}

// per-warp workspace of the single-root PPR fast path (ppr_warp_kernel.cuh)
static void plan_warp(const Caps &caps, const shadow_sampler_cfg &c, bool sym, bool bis, int ecap_mult, WarpLayout *W, int *w_ecap, int *w_nf, int *w_hbuckets, int *w_hshift) {
  uint32_t off = 0;
  auto take = [&](size_t bytes) { uint32_t r = off; off = align16(off + (uint32_t)bytes); return r; };
  // staged edges per warp (global scratch, L2-resident): ecap_mult (16 to start with) per node + one full stage of head room.  A subgraph
  // that overflows is not an error, it takes the redo launch; the multiplier grows when more than 2 % of a launch had to be redone
  const char *env = getenv("SHADOW_WARP_ECAP_MULT");
  *w_ecap = (env ? std::max(0, atoi(env)) : ecap_mult) * caps.ncap + 32 * ((sym ? WARP_U_SYM : WARP_U) + 1) * WARP_CS;
  // Bloom words per lane: 1,024 bits per word-register, 2 bits per key: ~7 bits per key keep the false-positive rate near 6 %
  const char *envf = getenv("SHADOW_WARP_NF");
  // (the symmetric variant scans half the slots, its exact stage weighs more: one more word per lane pays off already at k = 150)
  int nf = caps.ncap <= (sym ? 96 : 192) ? 1 : (caps.ncap <= 448 ? 2 : 4);
  if (envf && (atoi(envf) == 1 || atoi(envf) == 2 || atoi(envf) == 4)) nf = atoi(envf);
  *w_nf = nf;
  // exact table: 2-key buckets, >= 3 buckets per key (4 KB for k = 150): 1-2 keys per subgraph overflow their bucket
  const char *envb = getenv("SHADOW_WARP_BUCKET_MULT");
  *w_hbuckets = std::max(16, next_pow2((envb ? std::max(1, atoi(envb)) : 3) * caps.ncap));
  int lg = 0; while ((1 << lg) < *w_hbuckets) lg++;
  *w_hshift = 32 - lg;
  W->hkeys = take(bis ? 16 : (size_t)*w_hbuckets * 8);        // bisection variant: no exact table
  W->queue = take((size_t)WARP_QCAP * 20);                   // chunk queue: uint4 keys + packed (row, offset) codes
  W->rs = take((size_t)caps.ncap * 8);
  W->nodes = take((size_t)caps.ncap * 4);
  W->cp = take(((size_t)caps.ncap + 1) * 4);
  W->rc = take((size_t)caps.ncap * 4);
  W->ovf = take((1 + WARP_OVF_CAP) * 4);
  W->bloom = take((size_t)32 * nf * 4);                      // Bloom words while they are built
  const bool ins = c.add_self_edge != 0;
  // rlo is written after the scan, when the chunk queue is dead: it lives there when it fits (one more resident warp per SM at k = 150)
  if ((ins || sym) && (size_t)caps.ncap * 4 <= (size_t)WARP_QCAP * 20) W->rlo = W->queue;
  else W->rlo = take((ins || sym) ? (size_t)caps.ncap * 4 : 0);
  W->rins = take(ins ? (size_t)caps.ncap * 4 : 0);
  W->rbug = take(ins ? (size_t)caps.ncap * 4 : 0);
  W->bytes = off;
}

typedef void (*warp_kernel_t)(const SampleParams);
template <bool A, bool S, bool B>
static warp_kernel_t pick_warp_kernel(int nf) {
  return nf == 1 ? ppr_induce_warp_kernel<A, 1, S, B> : (nf == 2 ? ppr_induce_warp_kernel<A, 2, S, B> : ppr_induce_warp_kernel<A, 4, S, B>);
}
template <bool A, bool S>
static warp_kernel_t pick_warp_kernel(int nf, bool bis) { return bis ? pick_warp_kernel<A, S, true>(nf) : pick_warp_kernel<A, S, false>(nf); }

// ------------------------------------------------------------------------------------------------
__global__ void max_degree_kernel(const uint32_t *indptr, uint32_t n, unsigned long long *out) {
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, indptr[i + 1] - indptr[i]);
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)m);
}
static int graph_dmax(shadow_sampler *s, long long *out) {
  if (s->dmax < 0) {
    unsigned long long *d;
    CUDA_TRY(cudaMalloc(&d, 8));
    CUDA_TRY(cudaMemsetAsync(d, 0, 8, s->stream));
    max_degree_kernel<<<s->num_sms * 4, 256, 0, s->stream>>>(s->indptr, s->N, d);
    unsigned long long h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    cudaFree(d);
    s->dmax = (long long)h;
  }
  *out = s->dmax;
  return 0;
}

static int read_bin_u32(const char *path, std::vector<uint32_t> &v) {     // read_array_from_bin (PS.cpp:70-86)
  FILE *f = fopen(path, "rb");
  if (!f) FAIL(SHADOW_EIO, "cannot open %s", path);
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize((size_t)sz / 4);
  size_t got = fread(v.data(), 4, v.size(), f);
  fclose(f);
  if (got != v.size()) FAIL(SHADOW_EIO, "short read on %s", path);
  return 0;
}

static int sampler_common_init(shadow_sampler *s, int per_batch, int num_ens, int seed, int device, int num_ring) {
  if (per_batch <= 0 || num_ens <= 0 || num_ring <= 0) FAIL(SHADOW_EINVAL, "num_sampler_per_batch, num_subgraphs_ensemble and num_ring must be > 0");
  s->device = device; s->per_batch = per_batch; s->num_ens = num_ens; s->num_ring = num_ring; s->seed = seed;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  s->num_sms = prop.multiProcessorCount;
  if (seed < 0) seed = (int)((uint32_t)time(nullptr) & 0x7fffffffu);     // PS.h:49-53: srand(time) -- resolved ONCE, so glibc replay and Philox agree
  s->seed = seed;
  s->rng.seed((uint32_t)seed);
  s->ring.assign(num_ring, std::vector<Result>(num_ens));
  for (auto &slot : s->ring)
    for (auto &r : slot) CUDA_TRY(cudaMallocHost(&r.totals_host, 6 * sizeof(long long)));
  return 0;
}

extern "C" int shadow_sampler_create(const uint32_t *indptr_host, const uint32_t *indices_host, uint32_t num_nodes,
                                     uint32_t num_edges, const char *path_indptr, const char *path_indices,
                                     int num_sampler_per_batch, int num_subgraphs_ensemble, int seed, int device,
                                     int num_ring, shadow_sampler **out) {
  if (!out) FAIL(SHADOW_EINVAL, "out is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); FAIL(SHADOW_ECUDA, "no CUDA device: libshadow_b200 has no CPU fallback"); }
  CUDA_TRY(cudaSetDevice(device));
  std::vector<uint32_t> vp, vi;
  if (!indptr_host) {
    if (!path_indptr || !*path_indptr) FAIL(SHADOW_EINVAL, "neither indptr nor path_indptr given");
    int rc = read_bin_u32(path_indptr, vp); if (rc) return rc;
    indptr_host = vp.data(); num_nodes = (uint32_t)vp.size() - 1;
  }
  if (!indices_host) {
    if (!path_indices || !*path_indices) FAIL(SHADOW_EINVAL, "neither indices nor path_indices given");
    int rc = read_bin_u32(path_indices, vi); if (rc) return rc;
    indices_host = vi.data(); num_edges = (uint32_t)vi.size();
  }
  if (indptr_host[0] != 0 || indptr_host[num_nodes] != num_edges) FAIL(SHADOW_EINVAL, "indptr[0] != 0 or indptr[N] != nnz (G.h:29-30)");
  shadow_sampler *s = new shadow_sampler();
  int rc = sampler_common_init(s, num_sampler_per_batch, num_subgraphs_ensemble, seed, device, num_ring);
  if (rc) { shadow_sampler_destroy(s); return rc; }
  s->N = num_nodes; s->E = num_edges; s->owns_graph = true;
  if (cudaMalloc(&s->indptr, ((size_t)num_nodes + 1) * 4) != cudaSuccess || cudaMalloc(&s->indices, ((size_t)num_edges + 1) * 4) != cudaSuccess ||
      cudaMemcpy(s->indptr, indptr_host, ((size_t)num_nodes + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(s->indices, indices_host, (size_t)num_edges * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    shadow_set_error("CUDA error uploading the graph: %s", cudaGetErrorString(cudaGetLastError()));
    shadow_sampler_destroy(s);
    return SHADOW_ECUDA;
  }
  *out = s;
  return 0;
}

extern "C" int shadow_sampler_create_dev(const uint32_t *indptr_dev, const uint32_t *indices_dev, uint32_t num_nodes,
                                         uint32_t num_edges, int num_sampler_per_batch, int num_subgraphs_ensemble, int seed,
                                         int device, int num_ring, shadow_sampler **out) {
  if (!out || !indptr_dev || !indices_dev) FAIL(SHADOW_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(device));
  shadow_sampler *s = new shadow_sampler();
  int rc = sampler_common_init(s, num_sampler_per_batch, num_subgraphs_ensemble, seed, device, num_ring);
  if (rc) { shadow_sampler_destroy(s); return rc; }
  s->N = num_nodes; s->E = num_edges; s->owns_graph = false;
  s->indptr = const_cast<uint32_t *>(indptr_dev); s->indices = const_cast<uint32_t *>(indices_dev);
  *out = s;
  return 0;
}

static void result_release(Result &r) {
  DevBuf *all[] = {&r.node_ptr, &r.edge_ptr, &r.rowptr, &r.indices, &r.target, &r.num_target, &r.orig_node, &r.orig_edge, &r.hop, &r.drnl, &r.ppr, &r.sync,
                   &r.row_span, &r.edge_span, &r.indices_raw, &r.orig_edge_raw, &r.redo};
  for (auto b : all) b->release();
  if (r.totals_host) cudaFreeHost(r.totals_host);
  r.totals_host = nullptr;
}

extern "C" int shadow_sampler_destroy(shadow_sampler *s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  if (s->owns_graph) { cudaFree(s->indptr); if (s->indices) cudaFree(s->indices); }
  s->targets.release(); s->ppr_ptr.release(); s->ppr_neighs.release(); s->ppr_scores.release();
  s->ppr_sid.release(); s->ppr_sscore.release(); s->ppr_srank.release(); s->ppr_srow.release();
  s->rand_stream.release(); s->rand_off.release(); s->gws.release(); s->wscratch.release(); s->sym_rev.release(); s->ppr_supper.release();
  for (auto &slot : s->ring) for (auto &r : slot) result_release(r);
  if (s->kev0) { cudaEventDestroy(s->kev0); cudaEventDestroy(s->kev1); cudaEventDestroy(s->sev0); cudaEventDestroy(s->sev1); }
  delete s;
  return 0;
}
extern "C" int shadow_sampler_set_stream(shadow_sampler *s, void *st) {
  if (!s) FAIL(SHADOW_EINVAL, "NULL sampler");
  if (s->stream != (cudaStream_t)st) {                 // launches still in flight on the old stream are waited for before anything runs on the new one
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->stream = (cudaStream_t)st;
  }
  return 0;
}
extern "C" uint32_t shadow_sampler_num_nodes(const shadow_sampler *s) { return s->N; }
extern "C" uint32_t shadow_sampler_num_edges(const shadow_sampler *s) { return s->graph_dropped ? 0 : s->E; }
extern "C" uint32_t shadow_sampler_num_nodes_target(const shadow_sampler *s) { return s->T; }
extern "C" uint32_t shadow_sampler_get_idx_root(const shadow_sampler *s) { return s->idx_root; }
extern "C" int shadow_sampler_set_num_per_batch(shadow_sampler *s, int n) { if (n <= 0) FAIL(SHADOW_EINVAL, "num_sampler_per_batch must be > 0"); s->per_batch = n; return 0; }
extern "C" int shadow_sampler_reseed(shadow_sampler *s, int seed) { s->seed = seed; s->rng.seed((uint32_t)seed); s->rand_host.clear(); return 0; }

extern "C" int shadow_sampler_shuffle_targets(shadow_sampler *s, const uint32_t *t, uint32_t n) {
  if (!s || (!t && n)) FAIL(SHADOW_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(s->device));
  for (uint32_t i = 0; i < n; i++) if (t[i] >= s->N) FAIL(SHADOW_EINVAL, "target %u out of range", t[i]);
  CUDA_TRY(cudaStreamSynchronize(s->stream));      // earlier calls may still read the old list
  if (s->targets.ensure((size_t)std::max(n, 1u) * 4)) FAIL(SHADOW_ECUDA, "cudaMalloc(targets) failed");
  CUDA_TRY(cudaMemcpyAsync(s->targets.p, t, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  s->targets_ptr = (const uint32_t *)s->targets.p;
  if (s->T != n) s->idx_root = 0;      // the reference asserts equal sizes (PS.cpp:40); a new-size list restarts the traversal
  s->T = n;
  return 0;
}
extern "C" int shadow_sampler_shuffle_targets_dev(shadow_sampler *s, const uint32_t *t, uint32_t n) {
  if (!s || (!t && n)) FAIL(SHADOW_EINVAL, "NULL argument");
  s->targets_ptr = t;
  if (s->T != n) s->idx_root = 0;
  s->T = n;
  return 0;
}
extern "C" int shadow_sampler_drop_full_graph_info(shadow_sampler *s) {      // PS.cpp:22-34
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (s->owns_graph && s->indices) { cudaFree(s->indices); }
  s->indices = nullptr; s->graph_dropped = true;
  s->ppr_neighs.release(); s->ppr_scores.release(); s->ppr_ptr.release(); s->has_ppr = false;
  s->ppr_sid.release(); s->ppr_sscore.release(); s->ppr_srank.release(); s->ppr_srow.release(); s->ppr_sorted = false;
  s->sym_rev.release(); s->ppr_supper.release(); s->sym_state = 0; s->supper_valid = false;
  return 0;
}


// id-sorted copy of every PPR row (+ position in score order), built once when the tables are installed
#define PPR_SORT_CAP 1024
__global__ void __launch_bounds__(128) ppr_sort_rows_kernel(const unsigned long long *__restrict__ ptr, const uint32_t *__restrict__ neighs,
                                                            const float *__restrict__ scores, uint32_t num_nodes, uint32_t *__restrict__ sid,
                                                            float *__restrict__ sscore, unsigned short *__restrict__ srank,
                                                            const uint32_t *__restrict__ indptr, uint2 *__restrict__ srow, int *too_long) {
  __shared__ unsigned long long keys[PPR_SORT_CAP];
  for (uint32_t v = blockIdx.x; v < num_nodes; v += gridDim.x) {
    const unsigned long long off = ptr[v];
    const int len = (int)(ptr[v + 1] - off);
    if (len == 0) continue;
    if (len > PPR_SORT_CAP) { if (threadIdx.x == 0) *too_long = 1; continue; }
    const int np2 = next_pow2(len);
    for (int i = threadIdx.x; i < np2; i += blockDim.x) keys[i] = i < len ? (((unsigned long long)neighs[off + i] << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    block_bitonic_sort(keys, np2);
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const uint32_t r = (uint32_t)keys[i];
      const uint32_t id = (uint32_t)(keys[i] >> 32);
      sid[off + i] = id; sscore[off + i] = scores[off + r]; srank[off + i] = (unsigned short)r;
      if (id < num_nodes) { const uint32_t rs = indptr[id]; srow[off + i] = make_uint2(rs, indptr[id + 1] - rs); }   // row extent travels with the entry
      else srow[off + i] = make_uint2(0u, 0u);
    }
    __syncthreads();
  }
}

extern "C" int shadow_sampler_set_ppr_tables(shadow_sampler *s, const uint64_t *ptr, const uint32_t *neighs, const float *scores) {
  if (!s || !ptr) FAIL(SHADOW_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  const uint64_t tot = ptr[s->N];
  s->ppr_maxlen = 0;
  for (uint32_t v = 0; v < s->N; v++) s->ppr_maxlen = std::max<long long>(s->ppr_maxlen, (long long)(ptr[v + 1] - ptr[v]));
  if (s->ppr_ptr.ensure(((size_t)s->N + 1) * 8) || s->ppr_neighs.ensure((size_t)std::max<uint64_t>(tot, 1) * 4) ||
      s->ppr_scores.ensure((size_t)std::max<uint64_t>(tot, 1) * 4))
    FAIL(SHADOW_ECUDA, "cudaMalloc(ppr tables) failed");
  CUDA_TRY(cudaMemcpy(s->ppr_ptr.p, ptr, ((size_t)s->N + 1) * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(s->ppr_neighs.p, neighs, (size_t)tot * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(s->ppr_scores.p, scores, (size_t)tot * 4, cudaMemcpyHostToDevice));
  s->has_ppr = true;
  s->ppr_sorted = false;
  s->supper_valid = false;
  if (!getenv("SHADOW_NO_SORTED_PPR")) {
    if (s->ppr_sid.ensure((size_t)std::max<uint64_t>(tot, 1) * 4) || s->ppr_sscore.ensure((size_t)std::max<uint64_t>(tot, 1) * 4) ||
        s->ppr_srank.ensure((size_t)std::max<uint64_t>(tot, 1) * 2) || s->ppr_srow.ensure((size_t)std::max<uint64_t>(tot, 1) * 8))
      FAIL(SHADOW_ECUDA, "cudaMalloc(sorted ppr tables) failed");
    int *flag; CUDA_TRY(cudaMalloc(&flag, 4)); CUDA_TRY(cudaMemsetAsync(flag, 0, 4, s->stream));
    ppr_sort_rows_kernel<<<s->num_sms * 16, 128, 0, s->stream>>>((const unsigned long long *)s->ppr_ptr.p, (const uint32_t *)s->ppr_neighs.p,
                                                                 (const float *)s->ppr_scores.p, s->N, (uint32_t *)s->ppr_sid.p, (float *)s->ppr_sscore.p,
                                                                 (unsigned short *)s->ppr_srank.p, s->indptr, (uint2 *)s->ppr_srow.p, flag);
    int too_long = 0;
    CUDA_TRY(cudaMemcpyAsync(&too_long, flag, 4, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    cudaFree(flag);
    s->ppr_sorted = !too_long;
  }
  return 0;
}

extern "C" int shadow_sampler_get_ppr_row(shadow_sampler *s, uint32_t v, uint32_t cap, uint32_t *neighs, float *scores, uint32_t *len_out) {
  if (!s->has_ppr) FAIL(SHADOW_ESTATE, "PPR tables not installed");
  if (v >= s->N) FAIL(SHADOW_EINVAL, "node out of range");
  CUDA_TRY(cudaSetDevice(s->device));
  uint64_t pr[2];
  CUDA_TRY(cudaMemcpy(pr, (uint64_t *)s->ppr_ptr.p + v, 16, cudaMemcpyDeviceToHost));
  uint32_t len = (uint32_t)(pr[1] - pr[0]);
  *len_out = len;
  uint32_t c = std::min(len, cap);
  if (c) {
    CUDA_TRY(cudaMemcpy(neighs, (uint32_t *)s->ppr_neighs.p + pr[0], (size_t)c * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(scores, (float *)s->ppr_scores.p + pr[0], (size_t)c * 4, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// symmetric-graph variant of the warp fast path: one-time check + reverse-slot index, upper part of every table entry
// ------------------------------------------------------------------------------------------------
// One warp per row u: every slot (u, v) must have its image (v, u) (binary search in row v; rows strictly ascending, checked here too);
// rev[slot of (u, v)] = slot of (v, u).  *bad != 0 => the graph is not symmetric / not strictly sorted: the full-scan kernels are used.
__global__ void __launch_bounds__(256) sym_build_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices, uint32_t N,
                                                        uint32_t *__restrict__ rev, int *bad) {
  const int lane = threadIdx.x & 31;
  const uint32_t wpg = gridDim.x * (blockDim.x >> 5);
  for (uint32_t u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < N; u += wpg) {
    if ((u & 1023u) < wpg && *(volatile int *)bad) return;          // somebody found a counter-example: stop early
    const uint32_t s = indptr[u], e = indptr[u + 1];
    for (uint32_t slot = s + lane; slot < e; slot += 32) {
      const uint32_t v = indices[slot];
      bool ok = v < N && (slot + 1 >= e || indices[slot + 1] > v);
      if (ok) {
        uint32_t lo = indptr[v];
        const uint32_t end = indptr[v + 1];
        uint32_t hi = end;
        while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (indices[mid] < u) lo = mid + 1; else hi = mid; }
        ok = lo < end && indices[lo] == u;
        if (ok) rev[slot] = lo;
      }
      if (!ok) *bad = 1;
    }
  }
}
// upper part of the row of every entry of the id-sorted PPR tables: {first slot whose neighbour is >= id, slots from there to the row end}
__global__ void __launch_bounds__(256) ppr_upper_kernel(const uint32_t *__restrict__ sid, const uint2 *__restrict__ srow, unsigned long long total,
                                                        const uint32_t *__restrict__ indices, uint2 *__restrict__ supper) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t id = sid[i];
    const uint2 r = srow[i];
    uint32_t lo = r.x, hi = r.x + r.y;
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (indices[mid] < id) lo = mid + 1; else hi = mid; }
    supper[i] = make_uint2(lo, r.x + r.y - lo);
  }
}
// true when the symmetric variant can run: graph verified (once per graph), tables' upper parts built (once per table install)
static bool ensure_sym(shadow_sampler *s) {
  if (s->sym_state == 0) {
    s->sym_state = -1;
    if (getenv("SHADOW_NO_SYM") || s->E == 0) return false;
    if (s->sym_rev.ensure((size_t)s->E * 4)) { cudaGetLastError(); return false; }      // no room for the index: keep the full scan
    int *flag = nullptr;
    if (cudaMalloc(&flag, 4) != cudaSuccess) { cudaGetLastError(); s->sym_rev.release(); return false; }
    cudaMemsetAsync(flag, 0, 4, s->stream);
    sym_build_kernel<<<s->num_sms * 8, 256, 0, s->stream>>>(s->indptr, s->indices, s->N, (uint32_t *)s->sym_rev.p, flag);
    int bad = 1;
    cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, s->stream);
    const bool okc = cudaStreamSynchronize(s->stream) == cudaSuccess;
    cudaFree(flag);
    if (!okc || bad) { cudaGetLastError(); s->sym_rev.release(); return false; }
    s->sym_state = 1;
  }
  if (s->sym_state != 1) return false;
  if (!s->supper_valid) {
    if (!s->ppr_sorted) return false;
    unsigned long long tot = 0;
    if (cudaMemcpy(&tot, (const unsigned long long *)s->ppr_ptr.p + s->N, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return false; }
    if (s->ppr_supper.ensure((size_t)std::max<unsigned long long>(tot, 1) * 8)) { cudaGetLastError(); return false; }
    if (tot) ppr_upper_kernel<<<s->num_sms * 8, 256, 0, s->stream>>>((const uint32_t *)s->ppr_sid.p, (const uint2 *)s->ppr_srow.p, tot, s->indices, (uint2 *)s->ppr_supper.p);
    if (cudaGetLastError() != cudaSuccess) return false;
    s->supper_valid = true;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// launch of one ensemble branch
// ------------------------------------------------------------------------------------------------
static int ensure_result_caps(Result &r, int P, int num_roots, long long cap_nodes, long long cap_edges) {
  if (cap_nodes >= (1ll << 31) || cap_edges >= (1ll << 31)) FAIL(SHADOW_ECAP, "batch exceeds 2^31 nodes or edges (%lld, %lld): lower num_sampler_per_batch", cap_nodes, cap_edges);
  bool bad = false;
  bad |= r.node_ptr.ensure(((size_t)P + 1) * 4) != 0;
  bad |= r.edge_span.ensure((size_t)std::max(P, 1) * 8) != 0;
  bad |= r.num_target.ensure((size_t)P * 4) != 0;
  bad |= r.target.ensure((size_t)P * num_roots * 4) != 0;
  bad |= r.sync.ensure(64 + (size_t)P * 8) != 0;
  bad |= r.row_span.ensure((size_t)std::max<long long>(cap_nodes, 1) * 8) != 0;
  bad |= r.orig_node.ensure((size_t)cap_nodes * 4) != 0;
  bad |= r.ppr.ensure((size_t)cap_nodes * 4) != 0;
  bad |= r.hop.ensure((size_t)cap_nodes * 4) != 0;
  bad |= r.drnl.ensure((size_t)cap_nodes * 4) != 0;
  bad |= r.indices_raw.ensure((size_t)cap_edges * 4) != 0;
  bad |= r.orig_edge_raw.ensure((size_t)cap_edges * 4) != 0;
  bad |= r.redo.ensure((size_t)std::max(P, 1) * 10 + 64) != 0;      // redo list int[P4] | node counts int[P4] | score-rank cuts u16[P]   (P4 = P rounded up to 4)
  if (bad) FAIL(SHADOW_ECUDA, "cudaMalloc(result buffers) failed");
  r.cap_nodes = cap_nodes; r.cap_edges = cap_edges; r.cap_subg = P;
  return 0;
}

static int launch_branch(shadow_sampler *s, Result &r) {
  const shadow_sampler_cfg &c = r.cfg;
  const int P = r.num_subg;
  const int nids = (int)(r.idx_end - r.idx_start);
  if (c.return_target_only) {                         // dummy_sampler (PS.cpp:653-659): origNodeID = roots, nothing else
    if (r.orig_node.ensure((size_t)std::max(nids, 1) * 4)) FAIL(SHADOW_ECUDA, "cudaMalloc failed");
    CUDA_TRY(cudaMemcpyAsync(r.orig_node.p, s->targets_ptr + r.idx_start, (size_t)nids * 4, cudaMemcpyDeviceToDevice, s->stream));
    r.total_nodes = nids; r.total_edges = 0; r.pending = false; r.valid = true; r.rand_draws = 0;
    return 0;
  }
  if (s->graph_dropped) FAIL(SHADOW_ESTATE, "full graph was dropped (drop_full_graph_info); only return_target_only sampling is possible");
  if ((c.method == SHADOW_PPR || c.method == SHADOW_PPR_ST) && !s->has_ppr) FAIL(SHADOW_ESTATE, "ppr sampler used before preproc_ppr_approximate / set_ppr_tables");
  Caps caps;
  int rc = plan_caps(s, c, &caps);
  if (rc) return rc;
  // output capacities: nodes are bounded by P * ncap; edges start from a heuristic and grow on overflow
  long long cap_nodes = std::max<long long>(r.cap_nodes, std::min<long long>((long long)P * caps.ncap, (1ll << 31) - 2));
  long long cap_edges = std::max<long long>(r.cap_edges, std::min<long long>((long long)P * caps.ncap * 24 + 1024, (1ll << 31) - 2));
  rc = ensure_result_caps(r, std::max(P, r.cap_subg), c.num_roots, cap_nodes, cap_edges);
  if (rc) return rc;

  SampleParams K;
  memset(&K, 0, sizeof(K));
  K.indptr = s->indptr; K.indices = s->indices; K.num_nodes = s->N; K.num_edges = s->E;
  K.roots = s->targets_ptr + r.idx_start; K.num_root_ids = nids; K.num_subg = P;
  K.method = c.method; K.num_roots = c.num_roots; K.depth = c.depth; K.budget = c.budget; K.k = c.k; K.threshold = c.threshold;
  K.add_self = (c.method == SHADOW_NODEIID) ? 0 : c.add_self_edge;
  K.tconn = (c.method == SHADOW_NODEIID) ? 0 : c.include_target_conn;
  K.aug = c.aug; K.fixed_mode = c.fixed_mode; K.rng_mode = c.rng_mode;
  K.ppr_ptr = (const unsigned long long *)s->ppr_ptr.p; K.ppr_neighs = (const uint32_t *)s->ppr_neighs.p; K.ppr_scores = (const float *)s->ppr_scores.p;
  if (s->ppr_sorted) { K.ppr_sid = (const uint32_t *)s->ppr_sid.p; K.ppr_sscore = (const float *)s->ppr_sscore.p; K.ppr_srank = (const unsigned short *)s->ppr_srank.p; K.ppr_srow = (const uint2 *)s->ppr_srow.p; }
  K.philox_seed = (uint32_t)s->seed; K.philox_epoch = r.philox_epoch; K.root_slot_base = r.idx_start / (uint32_t)c.num_roots;
  K.ecap = caps.ecap;
  K.ncap = caps.ncap; K.ccap = caps.ccap; K.ccap2 = caps.ccap2; K.acap = caps.acap; K.acap2 = caps.acap2; K.hcap = caps.hcap; K.hshift = caps.hshift;
  K.L = caps.L;
  K.cap_nodes = r.cap_nodes; K.cap_edges = r.cap_edges;
  K.node_ptr = (int *)r.node_ptr.p; K.row_span = (int2 *)r.row_span.p; K.edge_span = (int2 *)r.edge_span.p; K.indices_out = (int *)r.indices_raw.p;
  K.target = (int *)r.target.p; K.num_target = (int *)r.num_target.p;
  K.orig_node = (uint32_t *)r.orig_node.p; K.orig_edge = (uint32_t *)r.orig_edge_raw.p; K.hop = (uint32_t *)r.hop.p; K.drnl = (uint32_t *)r.drnl.p;
  K.ppr_out = (float *)r.ppr.p;
  unsigned char *sync = (unsigned char *)r.sync.p;
  K.ticket = (uint32_t *)sync; K.totals = (long long *)(sync + 8);
  K.status_n = (unsigned long long *)(sync + 64);
  K.ticket2 = (uint32_t *)(sync + 4); K.redo_count = (uint32_t *)(sync + 32);      // redo_count == low word of totals[3]

  const size_t smem_limit = 200 * 1024;
  const bool use_gws = caps.L.bytes > smem_limit;
  int grid;
  if (use_gws) {
    grid = std::min(P, std::max(1, std::min(s->num_sms, (int)((size_t)(1ull << 30) / std::max<size_t>(caps.L.bytes, 1)))));
    K.gws_stride = (caps.L.bytes + 255) & ~255ull;
    if (s->gws.ensure((size_t)K.gws_stride * grid)) FAIL(SHADOW_ECUDA, "cudaMalloc(global workspace) failed");
    K.gws = (unsigned char *)s->gws.p;
  } else {
    int bps = 0;
    rc = kern_cfg(s->kcache, (const void *)sample_induce_kernel<false>, SAMPLER_BLOCK, caps.L.bytes, &bps); if (rc) return rc;
    grid = std::min(P, std::max(1, bps) * s->num_sms);
  }

  // ---- glibc replay: generate the stream, fix per-subgraph offsets with the serial prepass ----
  const bool glibc_khop = (c.method == SHADOW_KHOP && c.rng_mode == SHADOW_RNG_GLIBC && c.budget >= 0 && c.depth > 0);
  const bool glibc_st = (c.method == SHADOW_PPR_ST && c.rng_mode == SHADOW_RNG_GLIBC);
  const bool glibc = glibc_khop || glibc_st;
  if (glibc) {
    const long long need = (long long)P * caps.max_draws;
    if (need > (1ll << 28)) FAIL(SHADOW_ECAP, "glibc-replay stream of %lld draws is too long; use SHADOW_RNG_PHILOX for super-batches", need);
    while ((long long)s->rand_host.size() < need) s->rand_host.push_back(s->rng.next());
    if (s->rand_stream.ensure((size_t)std::max<long long>(need, 1) * 4) || s->rand_off.ensure(((size_t)P + 1) * 8)) FAIL(SHADOW_ECUDA, "cudaMalloc(rand stream) failed");
    CUDA_TRY(cudaMemcpyAsync(s->rand_stream.p, s->rand_host.data(), (size_t)need * 4, cudaMemcpyHostToDevice, s->stream));
    K.rand_stream = (const uint32_t *)s->rand_stream.p; K.rand_off = (long long *)s->rand_off.p;
  }
  // single-root PPR without hop/drnl labels: one warp per subgraph (ppr_warp_kernel.cuh); whatever does not fit its on-chip
  // staging is rebuilt by the generic kernel in redo mode, launched right behind (it exits at once when the list is empty)
  // (hop labels are a warp-level BFS at the end of the fast path: 46 of the reference's 61 configs use `feature_augment: hops`; drnl labels -- one
  // config, two BFS passes -- stay on the generic kernel)
  bool fast = c.method == SHADOW_PPR && c.num_roots == 1 && s->ppr_sorted && !(c.aug & SHADOW_AUG_DRNLS) && !use_gws &&
              (((uintptr_t)s->indices & 15) == 0) && caps.ncap <= 8191 && !getenv("SHADOW_NO_WARP_PPR");
  int w_ecap = 0;
  int w_nf = 1;
  if (fast) { long long d; int rcd = graph_dmax(s, &d); if (rcd) return rcd; fast = d + 8 < (1ll << WARP_OFFBITS); }      // packed candidate codes
  bool sym = false, bis = false;
  if (fast) {
    const char *envr = getenv("SHADOW_SYM_RATIO");
    sym = s->kept_ratio < (envr ? atof(envr) : 0.12) && ensure_sym(s);
    // exact membership by bisection instead of the hash table once the table would cost resident warps (> 4 KB: more than ~170 nodes)
    const char *envb = getenv("SHADOW_WARP_BISECT");
    bis = envb ? atoi(envb) != 0 : caps.ncap > 200;
    plan_warp(caps, c, sym, bis, s->warp_ecap_mult, &K.WL, &w_ecap, &w_nf, &K.w_hbuckets, &K.w_hshift); K.w_ecap = w_ecap; fast = K.WL.bytes <= 96 * 1024;
    if (sym) { K.ppr_supper = (const uint2 *)s->ppr_supper.p; K.sym_rev = (const uint32_t *)s->sym_rev.p; }
  }
  s->last_sym = fast && sym;
  warp_kernel_t kern = nullptr;
  int gridw = 0;
  if (fast && P > 0) {                                  // every host-side query happens BEFORE the first GPU operation of the launch
    kern = sym ? (K.add_self ? pick_warp_kernel<true, true>(w_nf, bis) : pick_warp_kernel<false, true>(w_nf, bis))
               : (K.add_self ? pick_warp_kernel<true, false>(w_nf, bis) : pick_warp_kernel<false, false>(w_nf, bis));
    int wps = 0;
    rc = kern_cfg(s->kcache, (const void *)kern, 32, K.WL.bytes, &wps); if (rc) return rc;
    rc = kern_cfg(s->kcache, (const void *)sample_induce_kernel<false, true>, SAMPLER_BLOCK, caps.L.bytes, nullptr); if (rc) return rc;
    gridw = std::min(P, std::max(1, wps) * s->num_sms);
    K.w_scratch_stride = ((unsigned long long)w_ecap * 10ull + 255ull) & ~255ull;
    if (s->wscratch.ensure((size_t)K.w_scratch_stride * gridw)) FAIL(SHADOW_ECUDA, "cudaMalloc(warp scratch) failed");
    K.w_scratch = (unsigned char *)s->wscratch.p;
    if (!s->kev0) { CUDA_TRY(cudaEventCreate(&s->kev0)); CUDA_TRY(cudaEventCreate(&s->kev1)); CUDA_TRY(cudaEventCreate(&s->sev0)); CUDA_TRY(cudaEventCreate(&s->sev1)); }
    CUDA_TRY(cudaEventRecord(s->sev0, s->stream));
  }
  CUDA_TRY(cudaMemsetAsync(sync, 0, 64 + (size_t)P * 8, s->stream));
  if (glibc_st) {
    ppr_st_rand_offsets_kernel<<<1, 32, 0, s->stream>>>(K);
    CUDA_TRY(cudaGetLastError());
  }
  if (glibc_khop) {
    SampleParams Kp = K;
    Kp.count_only_last = 1;
    if (use_gws) { khop_rand_offsets_kernel<<<1, SAMPLER_BLOCK, 0, s->stream>>>(Kp); }
    else {
      Kp.gws = nullptr;
      CUDA_TRY(cudaFuncSetAttribute(khop_rand_offsets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)caps.L.bytes));
      khop_rand_offsets_kernel<<<1, SAMPLER_BLOCK, caps.L.bytes, s->stream>>>(Kp);
    }
    CUDA_TRY(cudaGetLastError());
  }
  if (P > 0) {
    if (use_gws) sample_induce_kernel<true><<<grid, SAMPLER_BLOCK, 0, s->stream>>>(K);
    else if (fast) {
      K.redo_list = (int *)r.redo.p;
      const size_t P4 = ((size_t)P + 3) & ~(size_t)3;      // keeps the count array 16-byte aligned for scan_counts_kernel
      int *cnt = (int *)r.redo.p + P4;
      K.w_cut = (const unsigned short *)((int *)r.redo.p + 2 * P4);
      // node_ptr[] up front: the node count of a PPR subgraph depends on its table row only (no look-back in the main kernel)
      ppr_count_kernel<<<(P + 7) / 8, 256, 0, s->stream>>>(K, cnt, (unsigned short *)K.w_cut);
      scan_counts_kernel<<<1, 1024, 0, s->stream>>>(cnt, P, K.node_ptr, K.totals);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaEventRecord(s->kev0, s->stream));
      kern<<<gridw, 32, K.WL.bytes, s->stream>>>(K);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaEventRecord(s->kev1, s->stream));
      s->kev_valid = true;
      sample_induce_kernel<false, true><<<std::min(P, 2 * s->num_sms), SAMPLER_BLOCK, caps.L.bytes, s->stream>>>(K);
      CUDA_TRY(cudaEventRecord(s->sev1, s->stream));
    } else sample_induce_kernel<false><<<grid, SAMPLER_BLOCK, caps.L.bytes, s->stream>>>(K);
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaMemcpyAsync(r.totals_host, K.totals, 6 * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
  r.fast_launch = fast && P > 0;
  r.pending = true; r.valid = false; r.rand_draws = 0; r.canon_valid = false;
  if (glibc) {                 // the host generator must advance by what this call consumed before the next call
    long long used = 0;
    CUDA_TRY(cudaMemcpyAsync(&used, (long long *)s->rand_off.p + P, 8, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (used > (long long)s->rand_host.size()) FAIL(SHADOW_ECUDA, "internal: rand stream underestimated");
    r.rand_draws = used;
  }
  return 0;
}

// wait for the launch, grow buffers and re-run if the edge capacity was exceeded
static int validate_branch(shadow_sampler *s, Result &r) {
  for (int attempt = 0; r.pending; attempt++) {
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    const long long tn = r.num_subg ? r.totals_host[0] : 0, te = r.num_subg ? r.totals_host[1] : 0, err = r.totals_host[2];
    if (err & ERR_WS_OVERFLOW) FAIL(SHADOW_ECAP, "internal: per-subgraph workspace overflow");
    if (err & ERR_OUT_OVERFLOW) {
      if (attempt >= 3) FAIL(SHADOW_ECAP, "internal: output capacity did not converge");
      r.cap_nodes = std::max(r.cap_nodes, tn); r.cap_edges = std::max(r.cap_edges, te + te / 8 + 1024);
      // the replayed glibc draws of this call are still at the head of rand_host: re-running is deterministic
      int rc = launch_branch(s, r);
      if (rc) return rc;
      continue;
    }
    r.total_nodes = tn; r.total_edges = te; r.pending = false; r.valid = true;
    s->last_redo = r.num_subg ? (long long)(r.totals_host[3] & 0xffffffffll) : 0;
    if (s->last_redo * 50 > r.num_subg && s->warp_ecap_mult < 128) s->warp_ecap_mult *= 2;
    // kept / scanned slots of this launch (the same for both variants: each sees half of both) steers the next launch's choice
    if (r.fast_launch && r.num_subg >= 64 && r.totals_host[5] > 0) s->kept_ratio = (double)r.totals_host[4] / (4.0 * (double)r.totals_host[5]);
  }
  return 0;
}

extern "C" int shadow_sampler_sample(shadow_sampler *s, const shadow_sampler_cfg *cfgs, int num_cfgs) {
  if (!s || !cfgs) FAIL(SHADOW_EINVAL, "NULL argument");
  if (num_cfgs != s->num_ens) FAIL(SHADOW_EINVAL, "got %d sampler configs for %d ensemble branches (PS.cpp:666)", num_cfgs, s->num_ens);
  CUDA_TRY(cudaSetDevice(s->device));
  const int num_roots = cfgs[0].num_roots;
  if (num_roots < 1 || num_roots > SHADOW_MAX_ROOTS) FAIL(SHADOW_EINVAL, "num_roots must be in [1,%d]", SHADOW_MAX_ROOTS);
  for (int i = 0; i < num_cfgs; i++) {
    if (cfgs[i].num_roots != num_roots) FAIL(SHADOW_EINVAL, "all ensemble branches must use the same num_roots (PS.cpp:668-671)");
    if (cfgs[i].method < 0 || cfgs[i].method > SHADOW_NODEIID) FAIL(SHADOW_EINVAL, "unknown sampler method %d", cfgs[i].method);
    if (cfgs[i].method == SHADOW_KHOP && cfgs[i].depth < 0) FAIL(SHADOW_EINVAL, "khop depth must be >= 0");
    if ((cfgs[i].method == SHADOW_PPR || cfgs[i].method == SHADOW_PPR_ST) && cfgs[i].k < 0) FAIL(SHADOW_EINVAL, "ppr k must be >= 0");
  }
  // _get_roots_p, sequential branch (PS.cpp:458-468)
  if (s->idx_root > s->T) s->idx_root = 0;
  const uint32_t T = s->T, idx_start = s->idx_root;
  const uint64_t want = (uint64_t)idx_start + (uint64_t)num_roots * (uint64_t)s->per_batch;
  const uint32_t idx_end = (uint32_t)std::min<uint64_t>(want, T);
  const uint32_t epoch = s->epoch;
  s->idx_root = (idx_end == T) ? 0 : idx_end;
  if (idx_end == T) s->epoch++;
  s->cur = (s->cur + 1) % s->num_ring;
  const int P = (int)((idx_end - idx_start + num_roots - 1) / num_roots);
  for (int b = 0; b < num_cfgs; b++) {
    Result &r = s->ring[s->cur][b];
    if (r.pending) CUDA_TRY(cudaStreamSynchronize(s->stream));
    r.cfg = cfgs[b]; r.idx_start = idx_start; r.idx_end = idx_end; r.num_subg = P; r.philox_epoch = epoch;
    int rc = launch_branch(s, r);
    if (rc) return rc;
    const bool glibc = !r.cfg.return_target_only && r.cfg.rng_mode == SHADOW_RNG_GLIBC &&
                       ((r.cfg.method == SHADOW_KHOP && r.cfg.budget >= 0) || r.cfg.method == SHADOW_PPR_ST);
    if (glibc) {             // keep the stream position exact: validate now, then drop the consumed prefix
      rc = validate_branch(s, r);
      if (rc) return rc;
      s->rand_host.erase(s->rand_host.begin(), s->rand_host.begin() + r.rand_draws);
    }
  }
  return 0;
}

extern "C" int shadow_sampler_batch_info(shadow_sampler *s, int branch, shadow_batch_info *info) {
  if (!s || !info) FAIL(SHADOW_EINVAL, "NULL argument");
  if (s->cur < 0) FAIL(SHADOW_ESTATE, "no sampler call yet");
  if (branch < 0 || branch >= s->num_ens) FAIL(SHADOW_EINVAL, "branch out of range");
  CUDA_TRY(cudaSetDevice(s->device));
  Result &r = s->ring[s->cur][branch];
  int rc = validate_branch(s, r);
  if (rc) return rc;
  info->num_subg = r.num_subg; info->num_roots = r.cfg.num_roots;
  info->total_nodes = r.total_nodes; info->total_edges = r.total_edges;
  info->has_csr = r.cfg.return_target_only ? 0 : 1;
  info->has_hop = (!r.cfg.return_target_only && (r.cfg.aug & SHADOW_AUG_HOPS) && !(r.cfg.aug & SHADOW_AUG_DRNLS)) ? 1 : 0;
  info->has_drnl = (!r.cfg.return_target_only && (r.cfg.aug & SHADOW_AUG_DRNLS)) ? 1 : 0;
  info->has_ppr = info->has_csr;
  info->rand_draws = r.rand_draws;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// RAW -> CANONICAL CSR (see shadow_b200.h): exclusive scan of the per-subgraph edge counts, then one streaming copy
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_edge_counts_kernel(const int2 *__restrict__ edge_span, int P, int *__restrict__ edge_ptr) {
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t chunk[1024];
  uint32_t carry = 0;
  for (int base = 0; base < P; base += 1024) {
    const int i = base + threadIdx.x;
    chunk[threadIdx.x] = (i < P) ? (uint32_t)(edge_span[i].y - edge_span[i].x) : 0u;
    __syncthreads();
    const uint32_t tot = block_exclusive_scan(chunk, 1024, warp_sums);
    if (i < P) edge_ptr[i] = (int)(carry + chunk[threadIdx.x]);
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) edge_ptr[P] = (int)carry;
}
__global__ void __launch_bounds__(256) canonicalize_kernel(const int *__restrict__ node_ptr, const int2 *__restrict__ edge_span,
                                                           const int2 *__restrict__ row_span, const int *__restrict__ idx_raw,
                                                           const uint32_t *__restrict__ eid_raw, const int *__restrict__ edge_ptr, int P,
                                                           int *__restrict__ rowptr, int *__restrict__ idx, uint32_t *__restrict__ eid) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < P; p += gridDim.x * wpb) {
    const int2 es = edge_span[p];
    const int dst = edge_ptr[p], n0 = node_ptr[p], n1 = node_ptr[p + 1];
    for (int j = lane; j < es.y - es.x; j += 32) { idx[dst + j] = idx_raw[es.x + j]; eid[dst + j] = eid_raw[es.x + j]; }
    for (int i = n0 + lane; i < n1; i += 32) rowptr[i] = dst + (row_span[i].x - es.x);
    if (p == P - 1 && lane == 0) rowptr[n1] = edge_ptr[P];
  }
}
static int ensure_canonical(shadow_sampler *s, Result &r) {
  if (r.canon_valid || r.cfg.return_target_only) return 0;
  const int P = r.num_subg;
  if (r.edge_ptr.ensure(((size_t)P + 1) * 4) || r.rowptr.ensure(((size_t)r.total_nodes + 1) * 4) ||
      r.indices.ensure((size_t)std::max<long long>(r.total_edges, 1) * 4) || r.orig_edge.ensure((size_t)std::max<long long>(r.total_edges, 1) * 4))
    FAIL(SHADOW_ECUDA, "cudaMalloc(canonical CSR) failed");
  if (P == 0) { CUDA_TRY(cudaMemsetAsync(r.edge_ptr.p, 0, 4, s->stream)); CUDA_TRY(cudaMemsetAsync(r.rowptr.p, 0, 4, s->stream)); r.canon_valid = true; return 0; }
  scan_edge_counts_kernel<<<1, 1024, 0, s->stream>>>((const int2 *)r.edge_span.p, P, (int *)r.edge_ptr.p);
  canonicalize_kernel<<<std::min((P + 7) / 8, s->num_sms * 8), 256, 0, s->stream>>>(
      (const int *)r.node_ptr.p, (const int2 *)r.edge_span.p, (const int2 *)r.row_span.p, (const int *)r.indices_raw.p,
      (const uint32_t *)r.orig_edge_raw.p, (const int *)r.edge_ptr.p, P, (int *)r.rowptr.p, (int *)r.indices.p, (uint32_t *)r.orig_edge.p);
  CUDA_TRY(cudaGetLastError());
  r.canon_valid = true;
  return 0;
}

static int field_lookup(shadow_sampler *s, int branch, int field, void **ptr, int64_t *count) {
  if (s->cur < 0) FAIL(SHADOW_ESTATE, "no sampler call yet");
  if (branch < 0 || branch >= s->num_ens) FAIL(SHADOW_EINVAL, "branch out of range");
  Result &r = s->ring[s->cur][branch];
  int rc = validate_branch(s, r);
  if (rc) return rc;
  const bool csr = !r.cfg.return_target_only;
  const int P = r.num_subg;
  if (field == SHADOW_F_EDGE_PTR || field == SHADOW_F_ROWPTR || field == SHADOW_F_INDICES || field == SHADOW_F_ORIG_EDGE) {
    rc = ensure_canonical(s, r);
    if (rc) return rc;
  }
  switch (field) {
    case SHADOW_F_NODE_PTR: *ptr = r.node_ptr.p; *count = csr ? P + 1 : 0; break;
    case SHADOW_F_EDGE_PTR: *ptr = r.edge_ptr.p; *count = csr ? P + 1 : 0; break;
    case SHADOW_F_ROWPTR: *ptr = r.rowptr.p; *count = csr ? r.total_nodes + 1 : 0; break;
    case SHADOW_F_INDICES: *ptr = r.indices.p; *count = csr ? r.total_edges : 0; break;
    case SHADOW_F_ORIG_NODE: *ptr = r.orig_node.p; *count = r.total_nodes; break;
    case SHADOW_F_ORIG_EDGE: *ptr = r.orig_edge.p; *count = csr ? r.total_edges : 0; break;
    case SHADOW_F_TARGET: *ptr = r.target.p; *count = csr ? (int64_t)P * r.cfg.num_roots : 0; break;
    case SHADOW_F_NUM_TARGET: *ptr = r.num_target.p; *count = csr ? P : 0; break;
    case SHADOW_F_PPR: *ptr = r.ppr.p; *count = csr ? r.total_nodes : 0; break;
    case SHADOW_F_HOP: *ptr = r.hop.p; *count = (csr && (r.cfg.aug & SHADOW_AUG_HOPS) && !(r.cfg.aug & SHADOW_AUG_DRNLS)) ? r.total_nodes : 0; break;
    case SHADOW_F_ROW_SPAN: *ptr = r.row_span.p; *count = csr ? 2 * r.total_nodes : 0; break;
    case SHADOW_F_EDGE_SPAN: *ptr = r.edge_span.p; *count = csr ? 2 * (int64_t)P : 0; break;
    case SHADOW_F_INDICES_RAW: *ptr = r.indices_raw.p; *count = csr ? r.total_edges : 0; break;
    case SHADOW_F_ORIG_EDGE_RAW: *ptr = r.orig_edge_raw.p; *count = csr ? r.total_edges : 0; break;
    case SHADOW_F_DRNL: *ptr = r.drnl.p; *count = (csr && (r.cfg.aug & SHADOW_AUG_DRNLS)) ? r.total_nodes : 0; break;
    default: FAIL(SHADOW_EINVAL, "unknown field %d", field);
  }
  return 0;
}
extern "C" int shadow_sampler_batch_field_dev(shadow_sampler *s, int branch, int field, void **ptr_dev, int64_t *count) {
  if (!s || !ptr_dev || !count) FAIL(SHADOW_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(s->device));
  return field_lookup(s, branch, field, ptr_dev, count);
}
extern "C" int shadow_sampler_batch_field_host(shadow_sampler *s, int branch, int field, void *dst, int64_t count) {
  if (!s || (!dst && count)) FAIL(SHADOW_EINVAL, "NULL argument");
  CUDA_TRY(cudaSetDevice(s->device));
  void *p; int64_t n;
  int rc = field_lookup(s, branch, field, &p, &n);
  if (rc) return rc;
  if (count != n) FAIL(SHADOW_EINVAL, "field %d has %lld elements, caller asked for %lld", field, (long long)n, (long long)count);
  if (n) { CUDA_TRY(cudaMemcpyAsync(dst, p, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream)); CUDA_TRY(cudaStreamSynchronize(s->stream)); }
  return 0;
}

extern "C" int64_t shadow_sampler_last_redo_count(const shadow_sampler *s) { return s ? s->last_redo : -1; }
extern "C" int shadow_sampler_last_sym(const shadow_sampler *s) { return s ? (s->last_sym ? 1 : 0) : -1; }
extern "C" float shadow_sampler_last_sequence_ms(shadow_sampler *s) {
  if (!s || !s->kev_valid) return -1.f;
  float ms = -1.f;
  if (cudaSetDevice(s->device) != cudaSuccess || cudaEventSynchronize(s->sev1) != cudaSuccess || cudaEventElapsedTime(&ms, s->sev0, s->sev1) != cudaSuccess) { cudaGetLastError(); return -1.f; }
  return ms;
}
extern "C" float shadow_sampler_last_kernel_ms(shadow_sampler *s) {
  if (!s || !s->kev_valid) return -1.f;
  float ms = -1.f;
  if (cudaSetDevice(s->device) != cudaSuccess || cudaEventSynchronize(s->kev1) != cudaSuccess || cudaEventElapsedTime(&ms, s->kev0, s->kev1) != cudaSuccess) { cudaGetLastError(); return -1.f; }
  return ms;
}

// hook for ppr_push.cu
int shadow_internal_graph(shadow_sampler *s, const uint32_t **indptr, const uint32_t **indices, uint32_t *N, uint32_t *E,
                          cudaStream_t *stream, int *num_sms) {
  if (!s) FAIL(SHADOW_EINVAL, "NULL sampler");
  if (s->graph_dropped) FAIL(SHADOW_ESTATE, "full graph was dropped (drop_full_graph_info)");
  CUDA_TRY(cudaSetDevice(s->device));
  *indptr = s->indptr; *indices = s->indices; *N = s->N; *E = s->E; *stream = s->stream; *num_sms = s->num_sms;
  return 0;
}
