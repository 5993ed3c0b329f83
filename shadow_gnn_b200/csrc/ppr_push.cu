// ppr_push.cu -- preproc_ppr_approximate (PS.cpp:237-344) and its binary cache (PS.cpp:94-231).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

// internal hooks implemented in sampler.cu
struct shadow_sampler;
int shadow_internal_graph(shadow_sampler *s, const uint32_t **indptr, const uint32_t **indices, uint32_t *N, uint32_t *E,
                          cudaStream_t *stream, int *num_sms);

// ------------------------------------------------------------------------------------------------
// binary cache: header {float alpha(=1-alpha_in), float epsilon, int32 k, uint32 count(=num_nodes)}
// then per node {uint32 len, len x (uint32 | float)}                         (PS.cpp:105-136)
// ------------------------------------------------------------------------------------------------
static bool read_ppr_file(const char *fn, bool is_float, int k, float alpha, float epsilon, uint32_t N,
                          std::vector<uint64_t> &ptr, std::vector<uint32_t> &vals, bool fill_ptr) {
  FILE *f = fopen(fn, "rb");
  if (!f) return false;
  float a = -1, e = -1; int kk = -1; uint32_t cnt = 0;
  bool ok = fread(&a, 4, 1, f) == 1 && fread(&e, 4, 1, f) == 1 && fread(&kk, 4, 1, f) == 1;
  // validity rule of read_PPR_from_binary_file (PS.cpp:166): alpha equal, epsilon within +-10 %, stored k >= requested
  if (!ok || a != alpha || e > 1.1 * epsilon || e < 0.9 * epsilon || kk < k) { fclose(f); return false; }
  if (fread(&cnt, 4, 1, f) != 1 || cnt != N) { fclose(f); return false; }
  if (fill_ptr) ptr.assign((size_t)N + 1, 0);
  vals.clear();
  std::vector<uint32_t> row;
  for (uint32_t i = 0; i < cnt; i++) {
    uint32_t len = 0;
    if (fread(&len, 4, 1, f) != 1) { fclose(f); return false; }
    row.resize(len);
    if (len && fread(row.data(), 4, len, f) != len) { fclose(f); return false; }
    uint32_t clip = std::min<uint32_t>(len, (uint32_t)k);                       // rows truncated to the requested k (:176-183)
    vals.insert(vals.end(), row.begin(), row.begin() + clip);
    if (fill_ptr) ptr[i + 1] = ptr[i] + clip;
    else if (ptr[i + 1] - ptr[i] != clip) { fclose(f); return false; }
  }
  (void)is_float;
  fclose(f);
  return true;
}

static void write_ppr_file(const char *fn, int k, float alpha, float epsilon, uint32_t N, const std::vector<uint64_t> &ptr,
                           const uint32_t *vals) {
  FILE *f = fopen(fn, "wb");
  if (!f) return;                                  // the reference silently skips an unopenable file (:106,123)
  fwrite(&alpha, 4, 1, f); fwrite(&epsilon, 4, 1, f); fwrite(&k, 4, 1, f); fwrite(&N, 4, 1, f);
  for (uint32_t i = 0; i < N; i++) {
    uint32_t len = (uint32_t)(ptr[i + 1] - ptr[i]);
    fwrite(&len, 4, 1, f);
    if (len) fwrite(vals + ptr[i], 4, len, f);
  }
  fclose(f);
}

int shadow_ppr_push_gpu(shadow_sampler *s, const uint32_t *targets_host, uint64_t T, int k, float alpha1m, float epsilon,
                        std::vector<uint32_t> &nb, std::vector<float> &sc, std::vector<uint32_t> &ln);

extern "C" int shadow_sampler_set_ppr_tables(shadow_sampler *, const uint64_t *, const uint32_t *, const float *);

extern "C" int shadow_sampler_preproc_ppr_approximate(shadow_sampler *s, const uint32_t *targets, uint64_t num_targets, int k,
                                                      float alpha, float epsilon, const char *fname_neighs, const char *fname_scores) {
  if (!s || (!targets && num_targets) || k <= 0) FAIL(SHADOW_EINVAL, "bad argument to preproc_ppr_approximate");
  const uint32_t *indptr, *indices; uint32_t N, E; cudaStream_t stream; int sms;
  int rc = shadow_internal_graph(s, &indptr, &indices, &N, &E, &stream, &sms);
  if (rc) return rc;
  alpha = 1 - alpha;                                                            // PS.cpp:242
  const bool named = fname_neighs && *fname_neighs && fname_scores && *fname_scores;
  if (named) {                                                                  // PS.cpp:244-247
    std::vector<uint64_t> ptr; std::vector<uint32_t> nbv, scv;
    if (read_ppr_file(fname_neighs, false, k, alpha, epsilon, N, ptr, nbv, true) &&
        read_ppr_file(fname_scores, true, k, alpha, epsilon, N, ptr, scv, false))
      return shadow_sampler_set_ppr_tables(s, ptr.data(), nbv.data(), (const float *)scv.data());
  }
  for (uint64_t i = 0; i < num_targets; i++) if (targets[i] >= N) FAIL(SHADOW_EINVAL, "ppr target %u out of range", targets[i]);
  std::vector<uint32_t> nb, ln; std::vector<float> sc;
  rc = shadow_ppr_push_gpu(s, targets, num_targets, k, alpha, epsilon, nb, sc, ln);
  if (rc) return rc;
  // rows by node id (top_ppr_neighs[target] = ..., PS.cpp:338-339); a repeated target keeps one identical row
  std::vector<uint64_t> len_of(N, 0), src_of(N, 0);
  for (uint64_t i = 0; i < num_targets; i++) { len_of[targets[i]] = ln[i]; src_of[targets[i]] = i; }
  std::vector<uint64_t> ptr((size_t)N + 1, 0);
  for (uint32_t v = 0; v < N; v++) ptr[v + 1] = ptr[v] + len_of[v];
  std::vector<uint32_t> fn(std::max<uint64_t>(ptr[N], 1)); std::vector<float> fs(std::max<uint64_t>(ptr[N], 1));
  for (uint32_t v = 0; v < N; v++) if (len_of[v]) {
    memcpy(&fn[ptr[v]], &nb[src_of[v] * k], len_of[v] * 4);
    memcpy(&fs[ptr[v]], &sc[src_of[v] * k], len_of[v] * 4);
  }
  rc = shadow_sampler_set_ppr_tables(s, ptr.data(), fn.data(), fs.data());
  if (rc) return rc;
  if (named) {                                                                  // PS.cpp:343
    write_ppr_file(fname_neighs, k, alpha, epsilon, N, ptr, fn.data());
    write_ppr_file(fname_scores, k, alpha, epsilon, N, ptr, (const uint32_t *)fs.data());
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// GPU forward push.  One warp owns one target and replays the reference's strictly ordered loop
// (always push the smallest active node id, std::set at PS.cpp:271-273) so that every float32 sum is
// formed in the same order; the parallelism is across targets and, inside one push, across the row
// (each neighbour of a row is distinct, so `r[u] += m` has no intra-row dependence).  No FMA contraction,
// denormals kept.
//   state per warp (global memory, L2 resident): open-addressing table {node, r, pi, flags}, active list.
// ------------------------------------------------------------------------------------------------
struct PushEntry { uint32_t key; float r; float pi; uint32_t flags; };   // flags: 1 = in prop_set, 2 = in touched map
#define PUSH_EMPTY 0xFFFFFFFFu

__device__ __forceinline__ uint32_t push_find_or_insert(PushEntry *tab, uint32_t mask, int shift, uint32_t key, bool *inserted) {
  uint32_t h = (key * 2654435761u) >> shift;
  *inserted = false;
  for (;;) {
    uint32_t k = tab[h].key;
    if (k == key) return h;
    if (k == PUSH_EMPTY) {
      uint32_t old = atomicCAS(&tab[h].key, PUSH_EMPTY, key);
      if (old == PUSH_EMPTY) { *inserted = true; return h; }
      if (old == key) return h;
    }
    h = (h + 1) & mask;
  }
}

__global__ void __launch_bounds__(128) ppr_push_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                                                       const uint32_t *__restrict__ targets, long long T, int k, float alpha,
                                                       float epsilon, PushEntry *tabs, uint32_t tab_cap, uint32_t *actives,
                                                       uint32_t act_cap, unsigned long long *sortbuf, uint32_t *out_nb,
                                                       float *out_sc, uint32_t *out_len, int *status, unsigned int *ticket) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  PushEntry *tab = tabs + (size_t)gw * tab_cap;
  uint32_t *act = actives + (size_t)gw * act_cap;
  unsigned long long *sb = sortbuf + (size_t)gw * tab_cap;
  const uint32_t mask = tab_cap - 1;
  const int shift = __clz(tab_cap) + 1;            // 32 - log2(tab_cap)
  const float c = __fsub_rn(1.f, alpha);                      // (1 - alpha)
  for (;;) {
    long long it = 0;
    if (lane == 0) it = (long long)atomicAdd(ticket, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= T) break;
    if (status[it] == 1) continue;                            // finished in an earlier (smaller-table) pass
    for (uint32_t i = lane; i < tab_cap; i += 32) { tab[i].key = PUSH_EMPTY; tab[i].r = 0.f; tab[i].pi = 0.f; tab[i].flags = 0; }
    __syncwarp();
    const uint32_t target = targets[it];
    uint32_t n_act = 1, n_used = 1;
    bool overflow = false;
    if (lane == 0) {
      bool isnew;
      uint32_t h = push_find_or_insert(tab, mask, shift, target, &isnew);
      tab[h].r = 1.f; tab[h].flags = 1; act[0] = target;
    }
    __syncwarp();
    while (n_act > 0 && !overflow) {
      // v = *(prop_set.begin()): the smallest active id (warp min-reduction over the active list)
      uint32_t best = 0xFFFFFFFFu, best_pos = 0;
      for (uint32_t i = lane; i < n_act; i += 32) { uint32_t a = act[i]; if (a < best) { best = a; best_pos = i; } }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        uint32_t ob = __shfl_xor_sync(0xffffffffu, best, d), op = __shfl_xor_sync(0xffffffffu, best_pos, d);
        if (ob < best) { best = ob; best_pos = op; }
      }
      const uint32_t v = best;
      bool isnew_v;
      const uint32_t hv = push_find_or_insert(tab, mask, shift, v, &isnew_v);
      const float r0 = tab[hv].r;
      const uint32_t s = indptr[v], e = indptr[v + 1], degv = e - s;
      const float m = __fdiv_rn(__fmul_rn(c, r0), (float)(2u * degv));          // (1-alpha)*r0 / (2*deg)   (:286)
      __syncwarp();
      if (lane == 0) tab[hv].pi = __fadd_rn(tab[hv].pi, __fmul_rn(alpha, r0));  // :284
      for (uint32_t base = s; base < e; base += 32) {                          // :287-303
        const uint32_t j = base + lane;
        bool ins = false, isnew = false; uint32_t u = 0;
        if (j < e) {
          u = indices[j];
          const uint32_t hu = push_find_or_insert(tab, mask, shift, u, &isnew);
          const float ru = __fadd_rn(tab[hu].r, m);
          tab[hu].r = ru;
          const uint32_t degu = indptr[u + 1] - indptr[u];
          if (ru > __fmul_rn(epsilon, (float)degu) && !(tab[hu].flags & 1u)) { tab[hu].flags |= 1u; ins = true; }
        }
        const uint32_t bm = __ballot_sync(0xffffffffu, ins);
        const uint32_t cntnew = __popc(__ballot_sync(0xffffffffu, isnew));
        if (ins) { uint32_t pos = n_act + __popc(bm & ((1u << lane) - 1u)); if (pos < act_cap) act[pos] = u; }
        n_act += __popc(bm);
        n_used += cntnew;                                                      // occupied slots
        if (n_act > act_cap || n_used > (tab_cap >> 1) + (tab_cap >> 2)) { overflow = true; break; }   // before the table can fill up
      }
      __syncwarp();
      if (overflow) break;
      // r[v] = r0*(1-alpha)/2 ; deactivate + record pi when below the threshold  (:312-316)
      const float rv = __fdiv_rn(__fmul_rn(r0, c), 2.f);
      const bool done = rv <= __fmul_rn(epsilon, (float)degv);
      if (lane == 0) {
        tab[hv].r = rv;
        if (done) { tab[hv].flags = (tab[hv].flags & ~1u) | 2u; }
      }
      if (done) {                       // erase v from the active list: v may have moved (appends never move entries) -> swap with last
        __syncwarp();
        if (lane == 0) act[best_pos] = act[n_act - 1];
        n_act--;
      }
      __syncwarp();
    }
    if (overflow) { if (lane == 0) status[it] = 2; continue; }
    // top-k of (-pi, id) ascending over the touched map (:320-339): pack (~bits(pi) , id) so that an unsigned sort does it.
    uint32_t cnt = 0;
    for (uint32_t base = 0; base < tab_cap; base += 32) {
      const uint32_t i = base + lane;
      const bool t = (tab[i].key != PUSH_EMPTY) && (tab[i].flags & 2u);
      const uint32_t bm = __ballot_sync(0xffffffffu, t);
      if (t) {
        // pi >= 0: larger pi <=> larger bit pattern; key = (~bits << 32) | id  sorts by pi desc, id asc; -0.0/+0.0 both map to +0
        uint32_t bits = __float_as_uint(tab[i].pi) & 0x7FFFFFFFu;
        sb[cnt + __popc(bm & ((1u << lane) - 1u))] = ((unsigned long long)(~bits) << 32) | tab[i].key;
      }
      cnt += __popc(bm);
    }
    __syncwarp();
    // selection of the k smallest keys by repeated warp min (k and cnt are small next to the push itself)
    const uint32_t kk = min((uint32_t)k, cnt);
    for (uint32_t o = 0; o < kk; o++) {
      unsigned long long best = ~0ull; uint32_t bp = 0;
      for (uint32_t i = lane; i < cnt; i += 32) { unsigned long long x = sb[i]; if (x < best) { best = x; bp = i; } }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, d); uint32_t op = __shfl_xor_sync(0xffffffffu, bp, d);
        if (ob < best) { best = ob; bp = op; }
      }
      if (lane == 0) {
        out_nb[it * k + o] = (uint32_t)best;
        out_sc[it * k + o] = __uint_as_float(~(uint32_t)(best >> 32));
        sb[bp] = ~0ull;
      }
      __syncwarp();
    }
    if (lane == 0) { out_len[it] = kk; status[it] = 1; }
  }
}

int shadow_ppr_push_gpu(shadow_sampler *s, const uint32_t *targets_host, uint64_t T, int k, float alpha1m, float epsilon,
                        std::vector<uint32_t> &nb, std::vector<float> &sc, std::vector<uint32_t> &ln) {
  const uint32_t *indptr, *indices; uint32_t N, E; cudaStream_t stream; int sms;
  int rc = shadow_internal_graph(s, &indptr, &indices, &N, &E, &stream, &sms);
  if (rc) return rc;
  nb.assign(std::max<uint64_t>(T * k, 1), 0); sc.assign(std::max<uint64_t>(T * k, 1), 0.f); ln.assign(std::max<uint64_t>(T, 1), 0);
  if (T == 0) return 0;
  uint32_t *d_t = nullptr, *d_nb = nullptr, *d_len = nullptr; float *d_sc = nullptr; int *d_status = nullptr; unsigned int *d_ticket = nullptr;
  CUDA_TRY(cudaMalloc(&d_t, T * 4)); CUDA_TRY(cudaMalloc(&d_nb, T * k * 4)); CUDA_TRY(cudaMalloc(&d_sc, T * k * 4));
  CUDA_TRY(cudaMalloc(&d_len, T * 4)); CUDA_TRY(cudaMalloc(&d_status, T * 4)); CUDA_TRY(cudaMalloc(&d_ticket, 4));
  CUDA_TRY(cudaMemcpyAsync(d_t, targets_host, T * 4, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaMemsetAsync(d_status, 0, T * 4, stream));
  CUDA_TRY(cudaMemsetAsync(d_len, 0, T * 4, stream));
  std::vector<int> status(T);
  uint64_t remaining = T;
  // pass p uses tables of 2^(14+3p) entries per warp; targets whose scope does not fit are retried in the next pass
  for (int pass = 0; pass < 5 && remaining > 0; pass++) {
    const uint32_t tab_cap = 1u << (14 + 3 * pass);
    const uint32_t act_cap = tab_cap / 2;
    const size_t per_warp = (size_t)tab_cap * (sizeof(PushEntry) + 8) + (size_t)act_cap * 4;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    long long warps = std::min<long long>({(long long)sms * 16, (long long)((free_b / 2) / per_warp), (long long)remaining});
    if (warps < 1) FAIL(SHADOW_ECAP, "not enough device memory for the PPR push tables");
    const int blocks = (int)((warps + 3) / 4);
    warps = (long long)blocks * 4;
    PushEntry *tabs; uint32_t *acts; unsigned long long *sbuf;
    CUDA_TRY(cudaMalloc(&tabs, (size_t)warps * tab_cap * sizeof(PushEntry)));
    CUDA_TRY(cudaMalloc(&acts, (size_t)warps * act_cap * 4));
    CUDA_TRY(cudaMalloc(&sbuf, (size_t)warps * tab_cap * 8));
    CUDA_TRY(cudaMemsetAsync(d_ticket, 0, 4, stream));
    ppr_push_kernel<<<blocks, 128, 0, stream>>>(indptr, indices, d_t, (long long)T, k, alpha1m, epsilon, tabs, tab_cap, acts, act_cap,
                                                sbuf, d_nb, d_sc, d_len, d_status, d_ticket);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(status.data(), d_status, T * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    cudaFree(tabs); cudaFree(acts); cudaFree(sbuf);
    remaining = 0;
    for (uint64_t i = 0; i < T; i++) remaining += (status[i] != 1);
    if (getenv("SHADOW_DEBUG")) fprintf(stderr, "[ppr_push] pass %d tab_cap=%u warps=%lld remaining=%llu\n", pass, tab_cap, warps, (unsigned long long)remaining);
  }
  if (remaining) FAIL(SHADOW_ECAP, "%llu PPR targets exceed the largest push table", (unsigned long long)remaining);
  CUDA_TRY(cudaMemcpy(nb.data(), d_nb, T * k * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(sc.data(), d_sc, T * k * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(ln.data(), d_len, T * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_t); cudaFree(d_nb); cudaFree(d_sc); cudaFree(d_len); cudaFree(d_status); cudaFree(d_ticket);
  return 0;
}
