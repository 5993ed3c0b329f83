/*
 * shadow_b200.h -- C ABI of the B200-native shaDow-GNN hot path (libshadow_b200.so).
 *
 * Plain pointers and sizes only; no torch / pybind types.  Each entry point names the reference
 * interface it replaces (paths relative to the reference tree; PS.cpp / PS.h =
 * para_graph_sampler/graph_engine/backend/ParallelSampler.{cpp,h}, G.h = backend/Graph.h,
 * layers.py = shaDow/layers.py, graph_utils.py = para_graph_sampler/graph_engine/frontend/graph_utils.py).
 *
 * Conventions
 *   - every function returns 0 on success, a negative SHADOW_E* code on failure;
 *     shadow_last_error() gives the message of the last failure on the calling thread.
 *   - "dev" pointers are CUDA device pointers on the sampler's device; "host" pointers are host memory.
 *   - all kernels are enqueued on the stream given (a cudaStream_t passed as void*; NULL = default stream).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with SHADOW_ECUDA.
 */
#ifndef SHADOW_B200_H
#define SHADOW_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHADOW_EINVAL   (-1)   /* bad argument (the reference would throw std::out_of_range / invalid_argument) */
#define SHADOW_ECUDA    (-2)   /* CUDA runtime error / no device */
#define SHADOW_EIO      (-3)   /* file missing / short */
#define SHADOW_ESTATE   (-4)   /* call order violated (e.g. sampling ppr before tables are installed) */
#define SHADOW_ECAP     (-5)   /* a size exceeded a 31-bit offset; use a smaller super-batch */

const char *shadow_last_error(void);
int shadow_version(void);

/* ---------------------------------------------------------------------------------------------- */
/* sampler object  == class ParallelSampler (PS.h:25-158; python surface PS.cpp:707-734)          */
/* ---------------------------------------------------------------------------------------------- */
typedef struct shadow_sampler shadow_sampler;

/* sampler methods: cfg.at("method") dispatch in parallel_sampler_ensemble (PS.cpp:685-693) */
enum { SHADOW_KHOP = 0, SHADOW_PPR = 1, SHADOW_PPR_ST = 2, SHADOW_NODEIID = 3 };
/* feature-augmentation flags: configs_aug sets {"hops","pprs","drnls"} (PS.cpp:433-451) */
enum { SHADOW_AUG_HOPS = 1, SHADOW_AUG_PPRS = 2, SHADOW_AUG_DRNLS = 4 };
/* random streams for khop / ppr_st */
enum { SHADOW_RNG_GLIBC = 0,    /* replays glibc rand() in the reference's single-thread order (PS.cpp:534, PS.h:49-53) */
       SHADOW_RNG_PHILOX = 1 }; /* counter-based Philox4x32-10 keyed (seed, call, root slot, level, node, draw): parallel */

/* one entry of `configs_samplers` (std::unordered_map<string,string>, PS.cpp:662-664) parsed to numbers */
typedef struct {
  int32_t method;               /* "method" */
  int32_t num_roots;            /* "num_roots" (1 node task, 2 link task) */
  int32_t depth, budget;        /* khop: "depth", "budget" (budget < 0 => all neighbours) */
  int32_t k;                    /* ppr / ppr_st: "k" */
  float   threshold;            /* ppr / ppr_st: "threshold" (stod narrowed to float, PS.cpp:571) */
  int32_t add_self_edge;        /* "add_self_edge" */
  int32_t include_target_conn;  /* "include_target_conn" */
  int32_t return_target_only;   /* "return_target_only" => dummy_sampler (PS.cpp:653-659) */
  int32_t aug;                  /* SHADOW_AUG_* bits */
  int32_t fixed_mode;           /* 0 = bug-compatible with the always-true compare at PS.cpp:401 (default),
                                   1 = row bound repaired */
  int32_t rng_mode;             /* SHADOW_RNG_* */
} shadow_sampler_cfg;

/* ParallelSampler ctor (PS.h:27-69).  indptr/indices are HOST arrays (copied to HBM) or NULL to load the raw
 * little-endian uint32 files path_indptr/path_indices (read_array_from_bin, PS.cpp:70-86).  `data`,
 * `edge_reweighted` of the reference are unused by it (PS.h:48) and have no counterpart.
 * num_ring >= 1 result buffers are kept alive: the output of call i stays valid until call i + num_ring. */
int shadow_sampler_create(const uint32_t *indptr_host, const uint32_t *indices_host,
                          uint32_t num_nodes, uint32_t num_edges,
                          const char *path_indptr, const char *path_indices,
                          int num_sampler_per_batch, int num_subgraphs_ensemble, int seed,
                          int device, int num_ring, shadow_sampler **out);
/* variant for a CSR that is already resident in HBM (the sampler borrows the pointers) */
int shadow_sampler_create_dev(const uint32_t *indptr_dev, const uint32_t *indices_dev,
                              uint32_t num_nodes, uint32_t num_edges,
                              int num_sampler_per_batch, int num_subgraphs_ensemble, int seed,
                              int device, int num_ring, shadow_sampler **out);
int shadow_sampler_destroy(shadow_sampler *s);
int shadow_sampler_set_stream(shadow_sampler *s, void *cuda_stream);

uint32_t shadow_sampler_num_nodes(const shadow_sampler *s);            /* PS.cpp:49  */
uint32_t shadow_sampler_num_edges(const shadow_sampler *s);            /* PS.cpp:51  */
uint32_t shadow_sampler_num_nodes_target(const shadow_sampler *s);     /* PS.cpp:53  */
uint32_t shadow_sampler_get_idx_root(const shadow_sampler *s);         /* PS.cpp:45  */
int shadow_sampler_set_num_per_batch(shadow_sampler *s, int num_sampler_per_batch);
/* shuffle_targets (PS.cpp:36-43), pre-shuffled branch: installs the epoch's target order. idx_root is not reset. */
int shadow_sampler_shuffle_targets(shadow_sampler *s, const uint32_t *targets_host, uint32_t n);
int shadow_sampler_shuffle_targets_dev(shadow_sampler *s, const uint32_t *targets_dev, uint32_t n);
/* srand() of the ctor (PS.h:49-53): restart the replayed glibc stream */
int shadow_sampler_reseed(shadow_sampler *s, int seed);
/* drop_full_graph_info (PS.cpp:22-34) */
int shadow_sampler_drop_full_graph_info(shadow_sampler *s);

/* PPR tables == top_ppr_neighs / top_ppr_scores (PS.h:145-146), CSR-flattened by node id:
 * row of node v = [ptr[v], ptr[v+1]).  Host arrays; copied to HBM. */
int shadow_sampler_set_ppr_tables(shadow_sampler *s, const uint64_t *ptr_host, const uint32_t *neighs_host,
                                  const float *scores_host);
/* preproc_ppr_approximate (PS.cpp:237-344): loads the binary cache if its header matches (PS.cpp:145-231),
 * otherwise runs the float32 forward push on the GPU (bit-exact with the reference order of operations)
 * and writes the cache (PS.cpp:94-139).  Empty / NULL file names => no cache. */
int shadow_sampler_preproc_ppr_approximate(shadow_sampler *s, const uint32_t *targets_host, uint64_t num_targets,
                                           int k, float alpha, float epsilon,
                                           const char *fname_neighs, const char *fname_scores);
/* read back the installed tables (row of node v) -- for tests and cache inspection */
int shadow_sampler_get_ppr_row(shadow_sampler *s, uint32_t v, uint32_t cap, uint32_t *neighs_host, float *scores_host,
                               uint32_t *len_out);

/* parallel_sampler_ensemble (PS.cpp:662-704): advances the root cursor (_get_roots_p, PS.cpp:456-468), samples
 * every ensemble branch on the same roots and leaves the results in HBM.  Asynchronous w.r.t. the host. */
int shadow_sampler_sample(shadow_sampler *s, const shadow_sampler_cfg *cfgs, int num_cfgs);

/* ---------------------------------------------------------------------------------------------- */
/* results == SubgraphStructVec (G.h:59-97) of the latest call, one per ensemble branch, kept in   */
/* HBM as ONE block-diagonal batch (what Subgraph.cat_to_block_diagonal, frontend/graph.py:280-320, */
/* builds on the host in the reference).  Rows (nodes) are always in subgraph order:               */
/*   node_ptr[P+1]                  exclusive prefix sum of |V_s|                                  */
/*   orig_node[N_tot], ppr[N_tot], hop[N_tot], drnl[N_tot]                                         */
/*   target[P*num_roots]            BATCH-global row of each root (-1 pads a short last group)     */
/* The edges exist in two layouts.  RAW is what the sampling kernel writes: the edge block of a    */
/* subgraph is contiguous and in CSR order, but blocks are placed in completion order so that no   */
/* CTA waits for another one's row scan:                                                           */
/*   row_span[N_tot]  (int32 pairs) [start,end) of every row inside indices_raw / orig_edge_raw    */
/*   edge_span[P]     (int32 pairs) [start,end) of every subgraph's edge block                     */
/*   indices_raw[E_tot]             BATCH-global column ids (local id + node_ptr[s])               */
/*   orig_edge_raw[E_tot]                                                                          */
/* CANONICAL is the standard CSR, produced on first request by one extra streaming kernel:         */
/*   edge_ptr[P+1], rowptr[N_tot+1] (== concatenated indptr), indices[E_tot], orig_edge[E_tot]     */
/* ---------------------------------------------------------------------------------------------- */
enum { SHADOW_F_NODE_PTR = 0, SHADOW_F_EDGE_PTR = 1, SHADOW_F_ROWPTR = 2, SHADOW_F_INDICES = 3,
       SHADOW_F_ORIG_NODE = 4, SHADOW_F_ORIG_EDGE = 5, SHADOW_F_TARGET = 6, SHADOW_F_PPR = 7,
       SHADOW_F_HOP = 8, SHADOW_F_DRNL = 9, SHADOW_F_NUM_TARGET = 10,
       SHADOW_F_ROW_SPAN = 11, SHADOW_F_EDGE_SPAN = 12, SHADOW_F_INDICES_RAW = 13, SHADOW_F_ORIG_EDGE_RAW = 14,
       SHADOW_NUM_FIELDS = 15 };

typedef struct {
  int32_t num_subg;        /* get_num_valid_subg() (G.cpp:92-94) */
  int32_t num_roots;
  int64_t total_nodes, total_edges;
  int32_t has_csr;         /* 0 for return_target_only */
  int32_t has_hop, has_ppr, has_drnl;
  int64_t rand_draws;      /* rand() outputs consumed by this branch (glibc mode) */
} shadow_batch_info;

/* synchronises the sampler's stream, validates capacities (re-running with larger buffers if needed) */
int shadow_sampler_batch_info(shadow_sampler *s, int branch, shadow_batch_info *info);
/* diagnostics: how many subgraphs of the last validated launch did not fit the one-warp-per-subgraph PPR fast path's on-chip
 * staging and were rebuilt by the generic kernel (no reference counterpart; results are identical either way) */
int64_t shadow_sampler_last_redo_count(const shadow_sampler *s);
/* diagnostics: 1 when the last launch ran the symmetric-graph variant of the PPR fast path.  The first fast-path launch on a graph verifies
 * that every edge has its reverse and that rows are strictly ascending (what fe/graph_utils.py:19-45 to_undirected guarantees); on such a
 * graph only the upper part of every row is scanned and each kept edge also yields its mirror image through a reverse-slot index built by
 * the same check.  Results are bit-identical to the full scan; SHADOW_NO_SYM=1 in the environment keeps the full scan. */
int shadow_sampler_last_sym(const shadow_sampler *s);
/* diagnostics: duration (ms, CUDA events on the sampler's stream) of ppr_induce_warp_kernel alone in the last fast-path launch, without
 * its helper launches (count, scan, redo); waits for that kernel; -1 when no fast-path launch has run */
float shadow_sampler_last_kernel_ms(shadow_sampler *s);
/* the same for all GPU work of that launch: state reset, ppr_count_kernel, scan_counts_kernel, the main kernel, the redo launch */
float shadow_sampler_last_sequence_ms(shadow_sampler *s);
/* device pointer + count of 4-byte elements of one field of the latest batch (valid until num_ring further calls) */
int shadow_sampler_batch_field_dev(shadow_sampler *s, int branch, int field, void **ptr_dev, int64_t *count);
/* copy one field to the host; 4 bytes per element */
int shadow_sampler_batch_field_host(shadow_sampler *s, int branch, int field, void *dst_host, int64_t count);

/* ---------------------------------------------------------------------------------------------- */
/* feature gather: feat_full[subgs.node] (shaDow/minibatch.py:469).  out[i,:] = feat[ids[i],:]      */
/* ---------------------------------------------------------------------------------------------- */
int shadow_gather_rows_f32(const float *feat_dev, int64_t num_rows, int32_t dim, const uint32_t *ids_dev,
                           int64_t n, float *out_dev, void *cuda_stream);
/* one training batch [row lo, lo+n) x [edge e0, e0+e) of a super-batch's canonical CSR -> static buffers of capacity row_cap rows (a slice in
 * the sense of OneBatchSubgraph, shaDow/minibatch.py:428-487, rebased to row 0 / edge 0; rows >= n get empty adjacency rows): rowptr_dst
 * int32[row_cap+1], span_dst int32[row_cap][2], col_dst int32[>= e], feat_dst float[row_cap][F] (first n rows written), tgt_dst int64[B] */
int shadow_load_batch(const int32_t *rowptr_src, int32_t e0, int32_t n, int32_t e, int32_t row_cap, int32_t *rowptr_dst, int32_t *span_dst,
                      const int32_t *idx_src, int32_t lo, int32_t *col_dst, const float *feat_src, int32_t F, float *feat_dst,
                      const int32_t *tgt_src, int32_t B, int64_t *tgt_dst, void *cuda_stream);

/* ---------------------------------------------------------------------------------------------- */
/* message-passing layers (shaDow/layers.py) on the RAW batch layout.  row_span = int32 pairs      */
/* [start,end) per row into col[] / val[]; column id = col[p] - col_off; fp32; all pointers device. */
/* ---------------------------------------------------------------------------------------------- */
/* adjacency values: data = 1 (PS.cpp:411,423) */
int shadow_edge_vals_fill(const int32_t *row_span, int32_t n, float *val, float x, void *cuda_stream);
/* dropedge: int(e*p) indices drawn with replacement are zeroed (graph_utils.py:86-88, layers.py:516-519,593-596);
 * row_ord[n+1] = exclusive prefix sum of the row lengths (ordinal position of every row's first edge); e = row_ord[n], the number
 * of draws int(e*p) and the Philox stream position *step_dev are read on the device (CUDA-graph friendly) */
int shadow_edge_vals_dropedge(const int32_t *row_span, const int32_t *row_ord, int32_t n, float p, uint32_t seed,
                              const uint32_t *step_dev, float *val, void *cuda_stream);
/* mode 0: adj_norm_rw (graph_utils.py:89-94); mode 1: GIN rescale deg_orig/deg_dropped (layers.py:520-522) */
int shadow_edge_vals_row_normalize(const int32_t *row_span, int32_t n, int32_t mode, float *val, void *cuda_stream);
/* adj_norm_sym (graph_utils.py:109-145): symmetric survival of `mask` (when dropedge > 0) then D^-1/2 A D^-1/2 */
int shadow_edge_vals_sym_normalize(const int32_t *row_span, const int32_t *col, int32_t col_off, int32_t n, int32_t symmetric_survival,
                                   const float *mask, float *val, float *deg_scratch, void *cuda_stream);
/* nn.Linear (layers.py:421,451-452,501-505,556,670; models.py:81) and its two backward products on the tensor cores, fp32 in / out:
 *   C[M,N] (+)= op(A)[M,K] op(B)[K,N] (+ bias[N]), every product an error-compensated 3xTF32 sum accumulated in fp32 (fp32-accurate).
 *   trans_a = 0: A[m][k] = A[m*lda + k];  1: A[m][k] = A[k*lda + m].      b_kn = 0: B[k][n] = B[n*ldb + k] (a Linear weight);  1: B[k][n] = B[k*ldb + n].
 *   accumulate = 1: C += ... with fp32 atomics (C initialised by the caller; required for split_k > 1, which cuts K into split_k slices). */
int shadow_gemm_tf32x3_f32(const float *A, int32_t lda, int32_t trans_a, const float *B, int32_t ldb, int32_t b_kn, float *C, int32_t ldc,
                           const float *bias, int32_t M, int32_t N, int32_t K, int32_t accumulate, int32_t split_k, void *cuda_stream);
/* two products of identical shape and layout in one launch (the self and the neighbour branch of a GraphSAGE layer, layers.py:474-483) */
int shadow_gemm_tf32x3_pair_f32(const float *A0, const float *A1, int32_t lda, int32_t trans_a, const float *B0, const float *B1, int32_t ldb,
                                int32_t b_kn, float *C0, float *C1, int32_t ldc, const float *bias0, const float *bias1, int32_t M, int32_t N,
                                int32_t K, int32_t accumulate, int32_t split_k, void *cuda_stream);
/* The same three products on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators, TMA-staged operands; fp32 operands split
 * into 3 bf16 terms in the kernel, nine-term products accumulated in fp32 => fp32-level accuracy).  csrc/gemm_umma.cu.  All leading
 * dimensions, K and N must be multiples of 4 floats and the pointers 16-byte aligned (TMA); otherwise SHADOW_EINVAL and the caller uses
 * shadow_gemm_tf32x3_f32.
 *   fwd    Z[M,N]  = X[M,K] W[N,K]^T + bias[N]                 (bias may be NULL)
 *   dgrad  dX[M,N] = dZ[M,K] W[K,N]
 *   wgrad  part[l][N_out,K_in] = dZ[l*r:(l+1)*r, :N_out]^T X[l*r:(l+1)*r, :K_in],  l < slices, r = rows_per_slice (caller sums the slices) */
int shadow_linear_umma_fwd_f32(const float *X, int64_t ldx, const float *W, int64_t ldw, const float *bias, float *Z, int64_t ldz, int32_t M, int32_t N,
                               int32_t K, void *cuda_stream);
int shadow_linear_umma_dgrad_f32(const float *dZ, int64_t lddz, const float *W, int64_t ldw, float *dX, int64_t lddx, int32_t M, int32_t N, int32_t K,
                                 void *cuda_stream);
int shadow_linear_umma_wgrad_f32(const float *dZ, int64_t lddz, const float *X, int64_t ldx, float *part, int32_t rows_per_slice, int32_t slices,
                                 int32_t N_out, int32_t K_in, void *cuda_stream);
/* torch.sparse.mm(adj, X) (layers.py:326-327,433,475,523): Y = beta*Y + A X ; backward dX += A^T dY (dX pre-zeroed by the caller) */
int shadow_spmm_csr_fwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *X, float *Y,
                            int32_t n, int32_t F, float beta, void *cuda_stream);
int shadow_spmm_csr_bwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *dY, float *dX,
                            int32_t n, int32_t F, void *cuda_stream);
/* The Linear of a shaDow layer with its epilogue, hand-written for the 5th-generation tensor cores (csrc/linear_tc.cu: TMA-staged operands,
 * tcgen05.mma.kind::tf32 with the accumulator in TMEM, error-compensated 3xTF32 products = fp32-level accuracy, epilogue out of tcgen05.ld).
 * Replaces nn.Linear + F_ACT + norm_feat of shaDow/layers.py:329-338,421,451-452,474-483 (one or two branches per launch: the self and the
 * neighbour branch of GraphSAGE) and, with act = 1 (identity), do_norm = 0 and W = the transposed weight, the input-gradient product.
 *   Z   = X[M,K] W[N,K]^T + bias           (saved when Z != NULL)
 *   o   = norm_feat(act(Z)) over the N <= 256 features (mean / rstd saved when do_norm), or act(Z)
 *   out = o (out_mode 0) | out + o (1) | atomically added into a zeroed buffer (2: both branches write one tensor)
 * All pointers 16-byte aligned, leading dimensions and N multiples of 4 floats; otherwise SHADOW_EINVAL. */
typedef struct shadow_linear_branch {
  const float *X, *W, *bias, *scale, *offset;
  float *Z, *out, *mean, *rstd;
  const float *W_lo;   /* NULL: W is the fp32 weight, split inside the kernel.  Else W = its TF32 head and W_lo = the exact remainder
                          (shadow_tf32_split_f32 / shadow_tf32_split_transpose_f32, refreshed once per optimizer step) */
} shadow_linear_branch;
int shadow_linear_tc_f32(const shadow_linear_branch *br, int32_t nbranch, int64_t ldx, int64_t ldw, int64_t ldz, int64_t ldo, int32_t M,
                         int32_t N, int32_t K, int32_t act, int32_t do_norm, int32_t out_mode, void *cuda_stream);
/* Weight gradient on the same tensor-core path:  grad[N_out, K_in] += dZ[M, N_out]^T X[M, K_in]  for one or two (dZ, X, grad) triples
 * (dZ1 NULL = one).  The batch is cut into 128-row slices, every slice's partial product goes to scratch (shadow_wgrad_tc_scratch_floats
 * floats per triple) and a second kernel adds the slices in slice order: deterministic.  Row-major contiguous operands (ld = N_out / K_in),
 * N_out and K_in multiples of 4 in [8, 256], 16-byte aligned pointers. */
int64_t shadow_wgrad_tc_scratch_floats(int32_t M, int32_t N_out, int32_t K_in);
int shadow_wgrad_tc_f32(const float *dZ0, const float *X0, float *grad0, float *scratch0, const float *dZ1, const float *X1, float *grad1,
                        float *scratch1, int32_t M, int32_t N_out, int32_t K_in, void *cuda_stream);
/* hi = src with the low 13 mantissa bits cleared (what a TF32 tensor core reads), lo = src - hi (exact in fp32) */
int shadow_tf32_split_f32(const float *src, int64_t n, float *hi, float *lo, void *cuda_stream);
/* the same for the TRANSPOSE of every 2-D weight inside one flat buffer: table_dev[e] = {src offset, rows, cols, dst offset} (int64, device) */
int shadow_tf32_split_transpose_f32(const float *src, const int64_t *table_dev, int32_t num_entries, float *t_hi, float *t_lo, void *cuda_stream);
/* act + norm_feat (layers.py:329-338; F_ACT layers.py:26-39): act ids 0 relu, 1 I, 2 elu, 3 tanh, 4 leakyrelu(0.2) */
int shadow_act_norm_fwd_f32(const float *Z, int32_t ldz, const float *scale, const float *offset, float *out, int32_t ldo,
                            float *mean, float *rstd, int32_t n, int32_t D, int32_t act, int32_t do_norm, int32_t accumulate,
                            void *cuda_stream);
int shadow_act_norm_bwd_f32(const float *dOut, int32_t ldo, const float *Z, int32_t ldz, const float *scale, const float *mean,
                            const float *rstd, float *dZ, int32_t lddz, float *dscale, float *doffset, float *dbias, int32_t n,
                            int32_t D, int32_t act, int32_t do_norm, void *cuda_stream);   /* dscale/doffset/dbias (column sums of dZ, may be NULL) are ACCUMULATED */
/* the same backward for one or two branches that share dOut (GraphSAGE: out = norm0(act(Z0)) + norm1(act(Z1)), layers.py:474-483; Z1 NULL = one
 * branch).  Column sums are reduced in two deterministic stages through `scratch` (>= shadow_act_norm_bwd_pair_nparts(n) * branches * 3 * D floats; 2 CTAs per SM by default): no atomics. D <= 256. */
int shadow_act_norm_bwd_pair_f32(const float *dOut, int32_t ldo, const float *Z0, const float *Z1, int32_t ldz, const float *scale0,
                                 const float *scale1, const float *mean0, const float *rstd0, const float *mean1, const float *rstd1,
                                 float *dZ0, float *dZ1, int32_t lddz, float *dscale0, float *doffset0, float *dbias0, float *dscale1,
                                 float *doffset1, float *dbias1, int32_t n, int32_t D, int32_t act, int32_t do_norm, float *scratch,
                                 int64_t scratch_floats, void *cuda_stream);
/* the same without the final column-sum launch (two branches with norm_feat only): the caller runs shadow_colsum_finish_f32 over `scratch`
 * (shadow_act_norm_bwd_pair_nparts(n) parts) on a stream of its own choice -- parameter gradients are off the backward pass's critical path */
int shadow_act_norm_bwd_pair_nofinish_f32(const float *dOut, int32_t ldo, const float *Z0, const float *Z1, int32_t ldz, const float *scale0,
                                          const float *scale1, const float *mean0, const float *rstd0, const float *mean1, const float *rstd1,
                                          float *dZ0, float *dZ1, int32_t lddz, int32_t n, int32_t D, int32_t act, int32_t do_norm, float *scratch,
                                          int64_t scratch_floats, void *cuda_stream);
int32_t shadow_act_norm_bwd_pair_nparts(int32_t n);
/* GAT._aggregate_attention for all heads (layers.py:560-582): a_self/a_neigh [n,heads] already through LeakyReLU(0.2) */
int shadow_gat_fwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *a_self,
                       const float *a_neigh, const float *H, float *out, float *rowmax, float *denom, int32_t n, int32_t heads,
                       int32_t d, void *cuda_stream);
int shadow_gat_bwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *a_self,
                       const float *a_neigh, const float *H, const float *out, const float *rowmax, const float *denom,
                       const float *dOut, float *dH, float *da_self, float *da_neigh, int32_t n, int32_t heads, int32_t d,
                       void *cuda_stream);
/* The GAT layer between its two Linear products, all heads per launch (csrc/gat.cu; shaDow/layers.py:539-645).  Shapes: heads in {1,2,4,8},
 * head width d in {16,32,64,128,256}, heads * d <= 256 (shadow_gat_supported); att / scale / offset are [2][heads * d].
 *   logits    s_b[i,k] = sum_f att[b,k,f] h_b[i,k,f], a_b = LeakyReLU_0.2(s_b)       (b = 0 self, 1 neighbour)
 *   agg       softmax-weighted neighbour sum of h_neigh per head (rowmax / denom saved), backward scatters into dHn / da_self / da_neigh (pre-zeroed)
 *   headnorm  out = (norm_feat_{1,k}(h_self) + norm_feat_{0,k}(agg)) / 2 per head; mean / rstd are [2][n][heads]
 *   pre_bwd   attention-logit backward + activation backward: dZ_b = (dh_b + ds_b att_b) act'(.), d att and d bias ACCUMULATED
 * Column sums are reduced in two deterministic stages through `scratch` (shadow_gat_scratch_floats floats). */
int shadow_gat_supported(int32_t heads, int32_t d);
int64_t shadow_gat_scratch_floats(int32_t heads, int32_t d);
int shadow_gat_logits_fwd_f32(const float *h_self, const float *h_neigh, const float *att, float *s_self, float *s_neigh, float *a_self, float *a_neigh,
                              int32_t n, int32_t heads, int32_t d, void *cuda_stream);
int shadow_gat_agg_fwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *a_self, const float *a_neigh,
                           const float *Hn, float *out, float *rowmax, float *denom, int32_t n, int32_t heads, int32_t d, void *cuda_stream);
int shadow_gat_agg_bwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *a_self, const float *a_neigh,
                           const float *Hn, const float *out, const float *rowmax, const float *denom, const float *dOut, float *dHn, float *da_self,
                           float *da_neigh, int32_t n, int32_t heads, int32_t d, void *cuda_stream);
int shadow_gat_headnorm_fwd_f32(const float *h_self, const float *agg, const float *scale, const float *offset, float *out, float *mean, float *rstd,
                                int32_t n, int32_t heads, int32_t d, int32_t do_norm, void *cuda_stream);
int shadow_gat_headnorm_bwd_f32(const float *dOut, const float *h_self, const float *agg, const float *scale, const float *mean, const float *rstd,
                                float *dh_self, float *dagg, float *dscale, float *doffset, int32_t n, int32_t heads, int32_t d, int32_t do_norm,
                                float *scratch, int64_t scratch_floats, void *cuda_stream);
int shadow_gat_pre_bwd_f32(const float *dh_self, const float *dh_neigh, const float *h_self, const float *h_neigh, const float *s_self, const float *s_neigh,
                           const float *da_self, const float *da_neigh, const float *att, float *dZ_self, float *dZ_neigh, float *datt, float *dbias_self,
                           float *dbias_neigh, int32_t n, int32_t heads, int32_t d, int32_t act, float *scratch, int64_t scratch_floats, void *cuda_stream);
int shadow_colsum_finish_f32(const float *partials, int32_t nparts, int32_t D, float *d00, float *d01, float *d02, float *d10, float *d11, float *d12,
                             void *cuda_stream);
/* per-subgraph pooling (F.embedding_bag in ResPool, layers.py:168-184): mode 0 sum, 1 mean, 2 max; seg = node_ptr slice */
int shadow_segment_pool_fwd_f32(const float *X, const int32_t *seg, int32_t seg_off, int32_t S, int32_t F, int32_t mode, float *out,
                                int32_t *argmax, void *cuda_stream);
int shadow_segment_pool_bwd_f32(const float *dOut, const int32_t *seg, int32_t seg_off, int32_t S, int32_t F, int32_t mode,
                                const int32_t *argmax, float *dX, void *cuda_stream);
/* clip_grad_norm_(params, max_norm) + Adam.step (models.py:223-224) over one flat fp32 buffer; grads are pre-scaled by
 * grad_scale (1/world_size after the NCCL sum); *step_dev is the device-side step counter (incremented here) */
int shadow_adam_clip_step_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float grad_scale,
                              float max_norm, float lr, float beta1, float beta2, float eps, int32_t *step_dev,
                              float *sqnorm_scratch, void *cuda_stream);
/* Data-parallel exchange over NVLink peer memory, fused with the optimizer's norm pass (the all-reduce of the flat gradient bucket that follows
 * loss.backward() in a DistributedDataParallel run of shaDow/models.py:209-237).  shadow_p2p_alloc: cudaMalloc'd, zeroed memory + its 64-byte
 * CUDA IPC handle; shadow_p2p_open: a peer's allocation mapped into this process (peer access enabled lazily).  grad_ptrs / flag_ptrs: HOST
 * arrays of `world` device pointers (index = rank; this rank's own allocations at [rank]); flags = uint32[32] per rank, zero-initialised;
 * state = uint32[516] in local device memory, zero-initialised.  shadow_p2p_zero_grad_f32 waits until every peer has finished reading the
 * previous step's gradients, then clears this rank's buffer; shadow_p2p_adam_clip_step_f32 = barrier + one-shot all-reduce into `gsum`
 * (local, n floats) + squared norm, then the step of shadow_adam_clip_step_f32 on gsum (grad_scale = 1/world gives the mean).  Polls give
 * up after ~10 s and set state[2] instead of hanging the GPU.  world <= 16. */
int shadow_p2p_alloc(int64_t bytes, void **ptr_dev, unsigned char *handle64);
int shadow_p2p_open(const unsigned char *handle64, void **ptr_dev);
int shadow_p2p_close(void *ptr_dev);
int shadow_p2p_free(void *ptr_dev);
int shadow_p2p_zero_grad_f32(const uint64_t *grad_ptrs, const uint64_t *flag_ptrs, int32_t world, int32_t rank, int64_t n, uint32_t *state, void *cuda_stream);
int shadow_p2p_adam_clip_step_f32(const uint64_t *grad_ptrs, const uint64_t *flag_ptrs, int32_t world, int32_t rank, float *param, float *gsum,
                                  float *exp_avg, float *exp_avg_sq, int64_t n, float grad_scale, float max_norm, float lr, float beta1, float beta2,
                                  float eps, int32_t *step_dev, float *sqnorm_scratch, uint32_t *state, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* SHADOW_B200_H */
