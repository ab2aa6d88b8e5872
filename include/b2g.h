/*
 * b2g.h — C ABI of libb2g.so: the B200-native (sm_100a) message-passing hot path that replaces,
 * behind the reference's own Python API, what Caesar3142/GNN-BFS-RANS executes through
 * torch-geometric and Python loops.  Every entry point cites the reference interface it replaces
 * (paths relative to /root/reference).
 *
 * Conventions
 *   - plain `extern "C"`, plain pointers and sizes, no torch types.
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`.
 *     Outputs and workspaces are pre-allocated by the caller (`*_workspace_bytes` queries).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 = ok, <0 = invalid argument (B2G_E_*), >0 = a cudaError_t.
 *   - feature matrices are row-major [rows, width] with an explicit row stride `ld` in ELEMENTS;
 *     rows must be 16-byte aligned (ld*sizeof(elem) % 16 == 0, base 16-byte aligned).
 *   - index arrays of the CSR are int32 (E < 2^31); edge_index stays int64 as in PyG.
 *   - no global mutable state except lazily initialised per-device attributes.
 */
#ifndef B2G_H_
#define B2G_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2G_VERSION 100

typedef enum { B2G_F32 = 0, B2G_BF16 = 1 } b2g_dtype;

enum {
  B2G_OK = 0,
  B2G_E_ARG = -1,        /* null pointer / negative size / bad enum                         */
  B2G_E_ALIGN = -2,      /* pointer or row stride not 16-byte aligned                       */
  B2G_E_SHAPE = -3,      /* unsupported width / heads combination                           */
  B2G_E_RANGE = -4,      /* size does not fit the int32 CSR index type                      */
  B2G_E_UNSUPPORTED = -5 /* this build has no kernel for the request                        */
};

int b2g_version(void);
const char* b2g_error_string(int code);
/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
int64_t b2g_launch_count(void);
void b2g_launch_count_reset(void);

/* ===================================================================================== K0
 * Graph builder.  Replaces GraphConstructor.build_edge_index (graph_constructor.py:28-56) and the
 * edge part of GraphConstructor.build_graph (graph_constructor.py:109-187, 220-227).
 * Bit-exact int64 [2,E] row-major output. */

/* build_edge_index(): E = 2*n_nei + (n_owner - n_nei); out_ei is int64 [2,E]. */
int b2g_build_edge_index(const int32_t* owner, const int32_t* neighbour, int64_t n_owner,
                         int64_t n_nei, int64_t* out_ei, void* stream);

/* internal_mask (uint8 [n_cells]) -> old_to_new (int32 [n_cells], rank among set entries or -1)
 * (graph_constructor.py:116-129).  n_set_out: device int64[1]. */
int64_t b2g_mask_to_map_workspace_bytes(int64_t n_cells);
int b2g_mask_to_map(const uint8_t* mask, int64_t n_cells, int32_t* old_to_new, int64_t* n_set_out,
                    void* ws, void* stream);

/* build_graph edge list in two phases (output size is data dependent).
 *   mode 0 ("C", graph_constructor.py:156,168-173): build_edge_index() then drop edges with an
 *          endpoint >= n_nodes.  old_to_new must be NULL.
 *   mode 1 ("A"/"B", graph_constructor.py:137-154): keep internal faces whose two cells both map
 *          to >= 0 through old_to_new[n_cells]; endpoints are remapped.  For mode A
 *          (n_internal_cells = n) pass old_to_new = NULL and n_nodes = n (identity map on [0,n)).
 * Then (graph_constructor.py:176-187, 220-227) one (v,v) per node v in [0,n_nodes) that appears in
 * no kept edge, ascending, appended after the kept edges.
 * counts_out: device int64[3] = {E_kept, n_isolated, n_bad} (n_bad = cell ids outside
 * [0,n_cells) met in mode 1: the reference raises IndexError there).
 * Phase 2 needs E_total = E_kept + n_isolated known on the host; out_ei is int64 [2,E_total]. */
int64_t b2g_build_graph_workspace_bytes(int64_t n_owner, int64_t n_nei, int64_t n_nodes);
int b2g_build_graph_count(const int32_t* owner, const int32_t* neighbour, int64_t n_owner,
                          int64_t n_nei, int mode, const int32_t* old_to_new, int64_t n_cells,
                          int64_t n_nodes, void* ws, int64_t* counts_out, void* stream);
int b2g_build_graph_fill(const int32_t* owner, const int32_t* neighbour, int64_t n_owner,
                         int64_t n_nei, int mode, const int32_t* old_to_new, int64_t n_cells,
                         int64_t n_nodes, const void* ws, int64_t E_total, int64_t* out_ei,
                         void* stream);

/* Edge attributes [dir_x,dir_y,dir_z,dist]: float64 arithmetic, fp32 output, zeros for self loops
 * and out-of-range endpoints.  Replaces graph_constructor.py:58-90 and :190-219.
 * cell_centers: float64 [n_centers,3]; edge_index int64 [2,E]; out fp32 [E,4]. */
int b2g_edge_attr(const double* cell_centers, int64_t n_centers, const int64_t* edge_index,
                  int64_t E, float* out, void* stream);

/* ===================================================================================== K1
 * CSR of the edge list a layer aggregates over.  No reference counterpart (PyG rebuilds gcn_norm /
 * self loops on every forward and scatters by edge); defined as the STABLE sort of the effective
 * edge list by the grouping endpoint.
 *   self_loops: 0 = raw list (GINConv, TransformerConv);
 *               1 = remove_self_loops + add_self_loops appended at the end (GCNConv gcn_norm,
 *                   GATConv; torch_geometric.utils.loop).
 *   by_source : 0 = rows are targets (edge_index[1]), col = source   (forward aggregation)
 *               1 = rows are sources (edge_index[0]), col = target   (backward / transposed)
 * Outputs: rowptr int32 [N+1]; col int32 [nnz]; eid int32 [nnz] (edge id: the position e in
 * edge_index for a kept edge, E+v for the appended loop of node v); dinv fp32 [N] = deg^-1/2 (0 if deg==0), may be
 * NULL.  nnz = b2g_csr_nnz result: E for self_loops=0, (E - #loops) + N for self_loops=1; the
 * #loops count needs a device pass: b2g_csr_count writes {nnz} to nnz_out (device int64[1]).
 * Edges with an endpoint outside [0,N) are an error (counted in nnz_out[1]). */
int64_t b2g_csr_workspace_bytes(int64_t E, int64_t N);
int b2g_csr_count(const int64_t* edge_index, int64_t E, int64_t N, int self_loops, int by_source,
                  int32_t* rowptr, int64_t* nnz_out, void* ws, void* stream);
int b2g_csr_fill(const int64_t* edge_index, int64_t E, int64_t N, int self_loops, int by_source,
                 const int32_t* rowptr, int64_t nnz, int32_t* col, int32_t* eid, float* dinv,
                 void* ws, void* stream);
/* perm[q] = position in `csr_a` of the edge stored at position q of `csr_b` (two CSRs of the same
 * effective edge list).  scratch: int32 [E+N] (edge-id space). */
int b2g_csr_perm(const int32_t* eid_a, const int32_t* eid_b, int64_t nnz, int32_t* scratch,
                 int32_t* perm, void* stream);

/* ===================================================================================== K2/K3
 * Deterministic CSR segment-sum.  Replaces MessagePassing.propagate's index_select + scatter_add_
 *  for GCNConv (gnn_model.py:63,166) and GINConv (gnn_model.py:75,166):
 *   out[i,:] = row_scale[i] * sum_{p in row i} col_scale[col[p]] * x[col[p],:]
 *              + self_coef * x[i,:] + bias[:]
 * row_scale / col_scale / bias may be NULL (=1 / =1 / =0); accumulation in fp32 in CSR order.
 * x: [*, F] dtype `dt`, row stride ldx; out: [n_rows, F] dtype `dt`, row stride ldo.
 * `x_self` (may be NULL -> x) is the matrix the self term reads (rows indexed by i).
 * relu != 0 applies max(.,0) last.
 * The fast paths prefetch col[clamp(position)] with unconditional loads: `col` (and `perm` where one is taken) must
 * point at >= 1 readable int32 even when the CSR has no entries at all (the Python host passes a dummy word). */
int b2g_seg_sum(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                int64_t ldo, int64_t n_rows, int F, int dt, const int32_t* rowptr,
                const int32_t* col, const float* row_scale, const float* col_scale,
                float self_coef, const float* bias, int relu, void* stream);

/* Same, with the graph's index band (max |col - row| over all entries) as a hint: for band-structured meshes the
 * rows are processed panel by panel so the +-band neighbours stay L2-resident (see aggregate.cu RowOrder).
 * band <= 0 = linear order.  Results are independent of the hint (every row is computed exactly once, in the
 * same per-row summation order). */
int b2g_seg_sum_banded(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                       int64_t ldo, int64_t n_rows, int F, int dt, const int32_t* rowptr,
                       const int32_t* col, const float* row_scale, const float* col_scale,
                       float self_coef, const float* bias, int relu, int64_t band, void* stream);

/* Same with a second hint: the longest row of the CSR (0 = unknown).  Reserved for kernel selection; results do
 * not depend on it. */
int b2g_seg_sum_hinted(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                       int64_t ldo, int64_t n_rows, int F, int dt, const int32_t* rowptr,
                       const int32_t* col, const float* row_scale, const float* col_scale,
                       float self_coef, const float* bias, int relu, int64_t band, int64_t max_row_len,
                       void* stream);

/* Same with explicit tuning parameters, PER CALL (the library holds no mutable tuning state):
 *   impl        0 = auto, 1 = the generic register-gather kernel (any 16-byte-multiple width)
 *   chunk_rows  rows the co-resident CTAs take per step (power of two >= 8; <= 0 = default 32)
 *   panel_rows  rows per panel of the band order (power of two >= chunk_rows; <= 0 = default 8192)
 * Results are bit-identical for every setting (every row is computed exactly once, in CSR order). */
int b2g_seg_sum_tuned(const void* x, int64_t ldx, const void* x_self, int64_t ldxs, void* out,
                      int64_t ldo, int64_t n_rows, int F, int dt, const int32_t* rowptr,
                      const int32_t* col, const float* row_scale, const float* col_scale,
                      float self_coef, const float* bias, int relu, int64_t band, int64_t max_row_len, int impl,
                      int chunk_rows, int panel_rows, void* stream);

/* ===================================================================================== K2f (fused)
 * The CSR segment-sum FUSED with the layer's (first) Linear, bf16, F = 256 (csrc/gcn_fused.cu): the aggregated rows go from the
 * gather warps into the shared-memory A operand of a tcgen05 GEMM against W (resident in shared memory); the [n, F]
 * intermediate of the unfused K6 + K2 pair never exists in HBM.
 *   out[i, :] = act( row_scale[i] * ( sum_{p in row i} col_scale[col_p] x[col_p, :] + self_coef x[i, :] ) W^T + bias )
 * GCNConv (gnn_model.py:63,166): col_scale = row_scale = deg^-1/2 over the self-loop-replaced list, self_coef = 0;
 * GINConv + the first Linear / ReLU of its MLP (gnn_model.py:70-75,166; inference): scales NULL, self_coef = 1 + eps, relu = 1.
 * x: bf16 [*, F]; w: bf16 [C, F]; out: bf16 [n_rows, C]; C a multiple of 64, <= 256. */
int b2g_segw_gemm_supported(int64_t n, int F, int C, int dt);
int b2g_segw_gemm(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const float* col_scale,
                  const float* row_scale, float self_coef, const void* w, int64_t ldw, const float* bias, int relu, void* out,
                  int64_t ldo, int64_t n_rows, int F, int C, int dt, int64_t band, void* stream);

/* ===================================================================================== K4
 * GATConv (gnn_model.py:65-68,168) fused edge-score + segment-softmax + aggregate + head-mean +
 * bias.  xw: [N, H*C] (= lin(x)); a_src/a_dst: fp32 [N,H] with row stride
 * lda (they are the two halves of the aux output of b2g_linear_fwd, see K6).
 *   s_p = leaky_relu(a_src[col[p],h] + a_dst[i,h], slope); alpha = softmax over row i;
 *   out[i,c] = (concat ? per head : mean over h) sum_p alpha_{p,h} xw[col[p],h,c] + bias
 * smax/ssum: fp32 [N,H] saved row max and sum(exp)+1e-16 (may be NULL in inference).
 * dropout: p_drop in [0,1) with Philox seed/offset (0 -> no dropout); mask regenerated in bwd.
 * max_degree: largest row length of the CSR if the caller knows it (0 = unknown); <= 32 with concat == 0 selects
 * the single-accumulator small-degree kernel (same result up to fp32 rounding order). */
int b2g_gat_fwd(const void* xw, int64_t ldxw, const float* a_src, const float* a_dst, int64_t lda,
                void* out, int64_t ldo, int64_t n_rows, int H, int C, int dt, int concat, float slope,
                const int32_t* rowptr, const int32_t* col, const float* bias, float* smax,
                float* ssum, float p_drop, uint64_t seed, int max_degree, void* stream);
/* Backward, target-major pass: recomputes alpha, produces per-edge alpha_e / dscore_e (fp32
 * [nnz,H], CSR order) and d a_dst [N,H].  gout: [N, concat?H*C:C]. */
int b2g_gat_bwd_dst(const void* xw, int64_t ldxw, const float* a_src, const float* a_dst, int64_t lda,
                    const void* gout, int64_t ldg, int64_t n_rows, int H, int C, int dt, int concat,
                    float slope, const int32_t* rowptr, const int32_t* col, const float* smax,
                    const float* ssum, float p_drop, uint64_t seed, float* alpha_e, float* ds_e,
                    float* d_a_dst, int64_t ldda, void* stream);
/* Backward, source-major pass over the transposed CSR (perm maps its positions to CSR positions):
 *   d xw[j,h,:] = sum_p alpha_{p,h} * gout[col_t[p], (h,):] (/H if mean);  d a_src[j,h] = sum_p ds_{p,h} */
int b2g_gat_bwd_src(const void* gout, int64_t ldg, const float* alpha_e, const float* ds_e,
                    void* d_xw, int64_t ldd, float* d_a_src, int64_t ldda, int64_t n_rows, int H, int C, int dt,
                    int concat, const int32_t* rowptr_t, const int32_t* col_t, const int32_t* perm,
                    void* stream);

/* ------------------------------------------------------------------------------------- K4, aggregate-first path
 * GATConv(heads = 4, concat = False) (gnn_model.py:65-68,168) for feature rows of 512 or 1024 bytes.  By linearity
 *   out_i = 1/H sum_h W_h (sum_j alpha_ijh x_j) + b = z_i Wc^T + b,  z_i = [sum_j alpha_ij1 x_j | ... | sum_j alpha_ijH x_j]
 * so the kernels gather the F-wide INPUT rows once per edge for all heads instead of the H*C-wide projected rows
 * (gat_rows.cu).  b2g_gatz_supported: 1 when these entry points cover (n, H, F, dt). */
int b2g_gatz_supported(int64_t n, int H, int F, int dt);
/* out[n, 0..7] = x[n, F] . V[8, F]^T (fp32 V, fp32 out with row stride ldo >= 8): the attention logits
 * a_src | a_dst of all heads in one pass over x (V = [W_h^T att_src_h ; W_h^T att_dst_h]). */
int b2g_rowdot8(const void* x, int64_t ldx, const float* V, int64_t ldv, float* out, int64_t ldo, int64_t n, int F,
                int dt, void* stream);
/* z [n, H*F] from x [*, F] and a = [a_src | a_dst] fp32 [*, >= 2H] (row stride lda floats): LeakyReLU(slope) scores,
 * exact max-subtracted segment softmax over each target's CSR row (PyG softmax, + 1e-16 in the denominator), optional
 * attention dropout (Philox keyed by the CSR position), weighted sum per head.  smax/ssum [n,H] (both or neither)
 * receive the statistics the backward pass needs. */
int b2g_gatz_fwd(const void* x, int64_t ldx, const float* a_srcdst, int64_t lda, void* z, int64_t ldz, int64_t n, int H,
                 int F, int dt, float slope, const int32_t* rowptr, const int32_t* col, float* smax, float* ssum,
                 float p_drop, uint64_t seed, int64_t band, void* stream);
/* Target side of the backward pass: given dz [n, H*F] writes the per-edge attention weights alpha_e [nnz,H] (after
 * dropout), the per-edge logit gradients de_e [nnz,H] (both in target-major CSR order) and d a_dst into columns
 * H..2H-1 of d_a (row stride ldda floats).  max_row_len: the longest row of (rowptr, col) when the caller knows it (0 = not
 * known): bf16 F = 256 runs rows of <= 8 entries on the tensor cores and needs the second pass over the longer rows only when
 * there are any. */
int b2g_gatz_bwd_dst(const void* x, int64_t ldx, const float* a_srcdst, int64_t lda, const void* dz, int64_t lddz,
                     int64_t n, int H, int F, int dt, float slope, const int32_t* rowptr, const int32_t* col,
                     const float* smax, const float* ssum, float p_drop, uint64_t seed, float* alpha_e, float* de_e,
                     void* d_a, int64_t ldda, int d_a_dt, const float* edge_bias, int64_t band, int64_t max_row_len, void* stream);
/* edge_bias (may be NULL): GATConv(edge_dim) — fp32 [nnz, H], the edge term of the logits (added in front of the LeakyReLU,
 * as b2g_gat_alpha does); de_e is then also the gradient of that term.
 * d_a / d_a_dt: the [n, >= 2H] logit-gradient block, fp32 (B2G_F32) or bf16 (B2G_BF16: the column block of the bf16 dgrad
 * operand [y | d a], written in place); ldda = its row stride in elements of that type.
 * Source side: over the transposed CSR (rowptr_t, col_t, perm = position of each entry in the target-major CSR)
 * y[j] = [sum_i alpha_ij1 g_i | ... | sum_i alpha_ijH g_i] ([n, H*C], g = d out [*, C]) and d a_src into columns
 * 0..H-1 of d_a. */
int b2g_gatz_bwd_src(const void* g, int64_t ldg, const float* alpha_e, const float* de_e, void* y, int64_t ldy,
                     void* d_a, int64_t ldda, int d_a_dt, int64_t n, int H, int C, int dt, const int32_t* rowptr_t,
                     const int32_t* col_t, const int32_t* perm, int64_t band, void* stream);

/* TransformerConv(heads = 4, concat = False) (gnn_model.py:77-80,170), aggregate-first (gat_rows.cu): the logits are
 * e_ijh = u_ih . x_j with u = x Mq + cq (a K6 GEMM; the key bias drops out of the softmax), and
 *   z_aug[i] = [sum_j alpha_ij1 x_j | ... | sum_j alpha_ijH x_j | s_i1 .. s_iH 0 0 0 0 | x_i]   (s_ih = sum_j alpha_ijh)
 * so that value projection, value bias and skip connection are one K6 GEMM with k = H*F + 8 + F.  alpha_e (may be
 * NULL) receives the pre-dropout attention weights [nnz, H] for the backward pass.  edge_bias (may be NULL): fp32
 * [nnz, H] in target-major CSR order, added to the logits — the edge-feature term of TransformerConv(edge_dim)
 * (b2g_edge_dot4 below). */
int b2g_tz_fwd(const void* x, int64_t ldx, const void* x_self, const void* u, int64_t ldu, void* z_aug, int64_t ldz, int64_t n,
               int H, int F, int dt, const int32_t* rowptr, const int32_t* col, float* alpha_e, const float* edge_bias,
               float p_drop, uint64_t seed, int64_t band, void* stream);
/* x_self (NULL = x): the matrix whose row i is target row i's own feature row (copied into z_aug).  It differs from x — the
 * matrix the column indices address — when the call covers a row range [r0, r1) of a larger problem (rowptr + r0, u + r0 rows,
 * z_aug + r0 rows, x_self = x + r0 rows): the host-buffer pipeline of streaming.py. */
/* Attention weights only (first half of b2g_tz_fwd; fused TransformerConv forward): alpha_pre [nnz, H] = softmax over the
 * entries of a row of u_ih . x_j (+ edge_bias), alpha_post (may be NULL when p_drop == 0: identical) = alpha_pre times the
 * attention-dropout keep scale, ssum fp32 [n, H] = per-head sums of the post-dropout weights.  impl 0 = default (bf16,
 * F = 256: the logits on the tensor cores), 1 = the SIMT dot products for every shape (A/B runs). */
int b2g_tz_alpha(const void* x, int64_t ldx, const void* u, int64_t ldu, int64_t n, int H, int F, int dt, const int32_t* rowptr,
                 const int32_t* col, float* alpha_pre, float* alpha_post, float* ssum, const float* edge_bias, float p_drop,
                 uint64_t seed, int64_t band, int impl, void* stream);
/* Target side of its backward pass: dz_aug [n, >= H*F + 8] (columns H*F .. H*F+H-1 = d s), alpha_in = the forward
 * pass's alpha_e; writes alpha_e (after dropout) and de_e [nnz, H] (gradients of the logits), target-major.  The
 * `du` (may be NULL) [n, H*F] receives d u_i = [sum_j de_ij1 x_j | ...] from the same gather.  The sums over the
 * transposed CSR (y, w) are b2g_gatz_bwd_src calls: perm == NULL there means the weights are already in the order of
 * the CSR passed in, d_a == NULL skips the logit-gradient row sums. */
int b2g_tz_bwd_dst(const void* x, int64_t ldx, const void* dz_aug, int64_t lddz, const float* alpha_in, int64_t n, int H,
                   int F, int dt, const int32_t* rowptr, const int32_t* col, float p_drop, uint64_t seed, float* alpha_e,
                   float* de_e, void* du, int64_t lddu, const float* edge_bias, int64_t band, int64_t max_row_len, void* stream);
/* Edge features of TransformerConv(edge_dim = 4) (PyG: key_j + lin_edge(edge_attr), value_j + the same),
 * aggregate-first: with r_ih = We_h^T q_ih / sqrt(C) (16 more columns of the u GEMM) the logit term is r_ih . a_ij, and
 * sum_j alpha'_ijh We_h a_ij = We_h m_ih with m_ih = sum_j alpha'_ijh a_ij (16 more columns of the output GEMM): the
 * [E, H*C] edge embedding never exists.  ea_csr: fp32 [nnz, 4] edge attributes in target-major CSR order.
 *   b2g_edge_dot4 : out[p, h] = v[i(p), 4h..4h+3] . ea_csr[p]       v fp32 [n, ldv >= 4H], out fp32 [nnz, H]
 *                   (edge_bias of b2g_tz_fwd from v = r; edge_bias of b2g_tz_bwd_dst, the d alpha term, from v = d m)
 *   b2g_edge_wsum4: out[i, 4h + c] = sum_p w[p, h] keep(p, h) ea_csr[p, c]   out [n, 4H] of dtype dt, row stride ldo
 *                   (m from w = alpha with the attention-dropout mask of (p_drop, seed); d r from w = de_e, p_drop = 0) */
int b2g_edge_dot4(const float* v, int64_t ldv, const float* ea_csr, const int32_t* rowptr, int64_t n, int H, float* out,
                  void* stream);
int b2g_edge_wsum4(const float* w, const float* ea_csr, const int32_t* rowptr, int64_t n, int H, float p_drop, uint64_t seed,
                   void* out, int64_t ldo, int dt, void* stream);

/* ===================================================================================== K4f (fused)
 * GATConv(heads = 4, concat = False), bf16, F = 256: the attention-weighted aggregation FUSED with the output projection
 * (csrc/gat_fused.cu) — z [n, H*F] never exists in HBM.  Replaces, like K4, PyG's softmax + [E,H,C] messages + scatter
 *  and the projection (row 10) of gnn_model.py:65-68,168.
 *
 * b2g_gat_alpha: alpha[p, h] (fp32 [nnz, 4], target-major CSR order) = dropout(segment_softmax(leaky_relu(a_src[col_p, h]
 *   + a_dst[i, h] (+ edge_bias[p, h])))) with PyG's max-subtraction and + 1e-16; a_srcdst fp32 [n, lda >= 8] = [a_src | a_dst];
 *   smax / ssum (fp32 [n, 4], both or neither) receive the softmax statistics the backward kernels read.
 * b2g_gatw_gemm: out[i, :] = sum_h (sum_{p in row i} alpha[q(p), h] x[col_p, :]) Wc_h^T + bias, q(p) = perm ? perm[p] : p.
 *   x: bf16 [*, F] (row stride ldx); wp: bf16 [C, H*F] = Wc with its columns permuted by 64-feature chunk:
 *   wp[c, kc*256 + h*64 + f] = Wc[c, h*256 + kc*64 + f]; out: bf16 [n_rows, C]; C a multiple of 64, <= 256.
 *   With the source-major CSR, perm and g in place of x it is the backward's y-aggregation + dgrad GEMM. */
int b2g_gatw_gemm_supported(int64_t n, int H, int F, int C, int dt);
int b2g_gat_alpha(const float* a_srcdst, int64_t lda, const int32_t* rowptr, const int32_t* col, const float* edge_bias,
                  int64_t row0, int64_t n, int H, float slope, float p_drop, uint64_t seed, float* alpha, float* smax, float* ssum,
                  void* stream);   /* rows [row0, row0 + n); every array is indexed by the global row / CSR position */
int b2g_gatw_gemm(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                  const float* alpha, const void* wp, int64_t ldw, const float* bias, void* out, int64_t ldo, int64_t n_rows,
                  int H, int F, int C, int dt, int64_t band, void* stream);

/* The same kernel with the epilogue terms of TransformerConv(heads = 4, concat = False), whose output is
 *   out_i = sum_h (z_ih Wv_h^T + s_ih bv_h) / H + x_i Ws^T + bs:   srow fp32 [n_rows, 4] = s_ih (per-head sums of the post-dropout
 * attention weights), bvh fp32 [4, C] = bv_h / H (both or neither), addend bf16 [n_rows, C] (row stride ldadd) = the skip
 * projection; any of them may be NULL.  With b2g_tz_alpha (K5) this is the fused TransformerConv forward: z_aug never exists. */
/* GATConv forward with the segment softmax INSIDE the fused kernel (every row <= 8 entries): the gather warps derive the attention
 * weights of their rows from a_src (fp32, indexed by global source node, row stride lda) and a_dst (fp32, indexed by the rows of
 * this call, same stride) in the per-tile prologue — no alpha round trip through HBM, no b2g_gat_alpha launch.  edge_bias fp32
 * [nnz, 4] or NULL; smax / ssum fp32 [n_rows, 4] (both or neither): the statistics b2g_gatz_bwd_dst needs; dropout as
 * b2g_gat_alpha.  max_row_len = the longest row of (rowptr, col); > 8 (or unknown, 0) -> B2G_E_UNSUPPORTED. */
int b2g_gatw_gemm_sm(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const float* a_src, const float* a_dst,
                     int64_t lda, const float* edge_bias, float slope, float p_drop, uint64_t seed, float* smax, float* ssum,
                     const void* wp, int64_t ldw, const float* bias, void* out, int64_t ldo, int64_t n_rows, int64_t max_row_len,
                     int H, int F, int C, int dt, int64_t band, void* stream);
int b2g_gatw_gemm_ex(const void* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                     const float* alpha, const void* wp, int64_t ldw, const float* bias, const float* srow, const float* bvh,
                     const void* addend, int64_t ldadd, void* out, int64_t ldo, int64_t n_rows, int H, int F, int C, int dt,
                     int64_t band, void* stream);

/* GATConv(edge_dim = 4): edge attributes fp32 [E, 4] -> fp32 [nnz, 4] in the order of the self-loop-replaced
 * target-major CSR (eid from b2g_csr_fill): dropped loops lose their attributes, every node's new loop gets the mean
 * attribute of its incoming edges (PyG fill_value = 'mean'; 0 without any).  The logit term is then
 * b2g_edge_dot4(v = We_h^T att_edge_h with ldv = 0, ...) -> the edge_bias input of b2g_gat_alpha / b2g_gatz_bwd_dst. */
int b2g_edge_rows_sl(const float* edge_attr, int64_t E, const int32_t* eid, const int32_t* rowptr, int64_t n, float* out,
                     void* stream);

/* ===================================================================================== K5
 * TransformerConv (gnn_model.py:77-80,170) fused q.k score + segment-softmax + aggregate +
 * head-mean + skip.  q,k,v: [N,H*C]; skip: [N, concat?H*C:C] or NULL.
 *   alpha = softmax_row( <q[i,h,:], k[col[p],h,:]> / sqrt(C) );  out = mean_h sum alpha v + skip */
int b2g_tconv_fwd(const void* q, const void* k, const void* v, int64_t ldqkv, const void* skip,
                  int64_t lds, void* out, int64_t ldo, int64_t n_rows, int H, int C, int dt,
                  int concat, const int32_t* rowptr, const int32_t* col, float* smax, float* ssum,
                  float p_drop, uint64_t seed, void* stream);
int b2g_tconv_bwd_dst(const void* q, const void* k, const void* v, int64_t ldqkv, const void* gout,
                      int64_t ldg, int64_t n_rows, int H, int C, int dt, int concat,
                      const int32_t* rowptr, const int32_t* col, const float* smax,
                      const float* ssum, float p_drop, uint64_t seed, float* alpha_e, float* ds_e,
                      void* dq, int64_t lddq, void* stream);
int b2g_tconv_bwd_src(const void* q, int64_t ldq, const void* gout, int64_t ldg,
                      const float* alpha_e, const float* ds_e, void* dk, void* dv, int64_t ldd,
                      int64_t n_rows, int H, int C, int dt, int concat, const int32_t* rowptr_t,
                      const int32_t* col_t, const int32_t* perm, void* stream);

/* ===================================================================================== K6
 * Dense per-node Linear (PyG Linear / torch nn.Linear inside the conv layers).
 *   fwd  : Y[n,m] = act( row_scale[n] * (sum_k X[n,k] W[m,k]) + bias[m] )        (F.linear)
 *          columns [0,m_main) go to Y (dtype dt); columns [m_main,m) go to `aux` as fp32 with row
 *          stride ldaux (GATConv's per-head attention logits ride along as extra columns of the
 *          same GEMM); m_main == m and aux == NULL for a plain Linear.
 *   dgrad: dX[n,k] = sum_m dY[n,m] W[m,k]
 *   wgrad: dW[m,k] = sum_n dY[n,m] X[n,k];  db[m] = sum_n dY[n,m]   (fp32 outputs)
 * X,Y,dX,dY dtype `dt`; W dtype `dt` for fwd/dgrad; accumulate fp32.  act: 0 none, 1 relu.
 * impl: 0 = auto, 1 = SIMT (FFMA) kernel, 2 = tcgen05 tensor-core kernel.  b2g_linear_fwd: bits 8..15 of impl = the number of
 * SMs the persistent tensor-core kernel leaves idle (its CTAs hold a whole SM's shared memory; a halo exchange running on another
 * stream gets no SM otherwise and waits for the GEMM to finish). */
int64_t b2g_linear_workspace_bytes(int64_t n, int m, int k, int dt, int which /*0 fwd 1 dgrad 2 wgrad*/);
/* Which kernel impl=0 (auto) picks for this shape: 1 = SIMT, 2 = tcgen05. */
int b2g_linear_impl(int64_t n, int m, int k, int dt, int which);
int b2g_linear_fwd(const void* X, int64_t ldx, const void* W, int64_t ldw, const float* bias,
                   const float* row_scale, void* Y, int64_t ldy, float* aux, int64_t ldaux, int64_t n,
                   int m, int m_main, int k, int dt, int act, int impl, void* ws, void* stream);
/* Y[n,m] = sum_k X[n,k] W[m,k] where mask[n,m] > 0, else 0 (bf16, tcgen05 path; B2G_E_UNSUPPORTED otherwise): the ReLU backward
 * (aten.threshold_backward) fused into the dgrad GEMM of the Linear behind the ReLU — X = that Linear's dY, W = its weight
 * transposed, mask = the ReLU's output.  Used by the Linear-ReLU-Linear MLP of GINConv (gnn_model.py:70-75). */
int b2g_linear_fwd_masked(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* mask, int64_t ldmask, void* Y,
                          int64_t ldy, int64_t n, int m, int k, int dt, void* ws, void* stream);
int b2g_linear_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw, void* dX,
                     int64_t lddx, int64_t n, int m, int k, int dt, int impl, void* ws,
                     void* stream);
int b2g_linear_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW,
                     int64_t lddw, float* db, int64_t n, int m, int k, int dt, int impl, void* ws,
                     void* stream);

/* Column sums: out[c] = sum_r x[r,c] (fp32, fixed order -> deterministic): bias gradients.
 * ws: fp32 [b2g_colsum_workspace_bytes(F)/4]. */
int64_t b2g_colsum_workspace_bytes(int F);
int b2g_colsum(const void* x, int64_t ldx, int64_t n_rows, int F, int dt, float* out, void* ws,
               void* stream);

/* ===================================================================================== halo
 * Multi-GPU halo plumbing (no reference counterpart).  pack: out[r,:] = x[idx[r],:];
 * unpack_add: x[idx[r],:] += in[r,:] (idx unique within one call).  b2g_rows_gather with idx == NULL copies rows
 * 0 .. n_idx-1 (any two row strides: a matrix into a column block of a wider one). */
int b2g_rows_gather(const void* x, int64_t ldx, const int32_t* idx, int64_t n_idx, void* out,
                    int64_t ldo, int F, int dt, void* stream);
int b2g_rows_scatter_add(void* x, int64_t ldx, const int32_t* idx, int64_t n_idx, const void* in,
                         int64_t ldi, int F, int dt, void* stream);

/* ------------------------------------------------------------------------------------- dropout epoch
 * Dropout seeds are call arguments, so a captured CUDA graph would replay identical masks.  Every dropout kernel of the
 * library (attention dropout of GATConv / TransformerConv, the fused BatchNorm-ReLU-dropout) mixes a per-device 64-bit
 * epoch, kept in device memory, into its seed.  b2g_dropout_epoch_advance increments it with one tiny launch on `stream`
 * (capture it at the top of a training-step graph: graphs.GraphedTrainStep); b2g_dropout_epoch_set writes it.  The epoch
 * is 0 until advanced: eager runs draw exactly the masks their seeds define. */
int b2g_dropout_epoch_advance(void* stream);
int b2g_dropout_epoch_set(uint64_t value, void* stream);

/* ===================================================================================== BatchNorm (+ fused glue)
 * torch_geometric.nn.BatchNorm over node features [n, C] (gnn_model.py:9,87,188) and, optionally fused around it, the
 * caller's residual add / ReLU / dropout (gnn_model.py:184-192):  s = x (+ r),
 *   y = dropout(relu(gamma (s - mean) rstd + beta)).   C * sizeof(dt) must be a multiple of 16 and <= 4096 bytes.
 * b2g_bn_stats: stats[0..C) = mean, [C..2C) = 1/sqrt(biased var + eps), [2C..3C) = biased variance of s (training).
 * b2g_bn_apply: mean / rstd = batch statistics (training) or running statistics (eval); s_out (may be NULL) receives
 *   s when r != NULL (kept for the backward pass); relu / p_drop fuse the activation and an inverted dropout
 *   (Philox keyed by the element index).
 * b2g_bn_bwd_stats: sums[0..C) = sum dz, [C..2C) = sum dz * xhat with dz = dy (* drop_scale where y > 0 when relu).
 * b2g_bn_bwd_apply: ds = gamma rstd (dz - sums0/n - xhat sums1/n) (training) or gamma rstd dz (eval). */
int64_t b2g_bn_workspace_bytes(int C);
int b2g_bn_stats(const void* x, int64_t ldx, const void* r, int64_t ldr, int64_t n, int C, int dt, float eps,
                 float* stats, void* ws, void* stream);
int b2g_bn_apply(const void* x, int64_t ldx, const void* r, int64_t ldr, void* y, int64_t ldy, void* s_out, int64_t lds,
                 int64_t n, int C, int dt, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 int relu, float p_drop, uint64_t seed, void* stream);
int b2g_bn_bwd_stats(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* s, int64_t lds, int64_t n,
                     int C, int dt, const float* mean, const float* rstd, int relu, float drop_scale, float* sums,
                     void* ws, void* stream);
int b2g_bn_bwd_apply(const void* dy, int64_t lddy, const void* y, int64_t ldy, const void* s, int64_t lds, void* ds,
                     int64_t ldds, int64_t n, int C, int dt, const float* mean, const float* rstd, const float* gamma,
                     const float* sums, int relu, float drop_scale, int training, void* stream);

/* ===================================================================================== mesh ingest
 * The arrays /root/reference/openfoam_loader.py derives from the polyMesh connectivity with Python loops over every
 * face, computed on the device from the flattened mesh (faces[i] = face_pts[face_off[i] .. face_off[i+1])):
 *   b2g_mesh_num_cells      n_cells = max(max(owner), max(neighbour)) + 1            (openfoam_loader.py:197, :236);
 *                           n_cells_out: device int64[1]; ws: >= 256 device bytes.
 *   b2g_mesh_cell_centers   get_cell_centers (:191-227): centers [n_cells,3] fp64 = mean of the unique vertices of the
 *                           faces a cell owns (face i of owner[i], i < n_owner) or neighbours (face i of neighbour[i],
 *                           i < n_nb); cells without a face stay 0.  Vertices are added in ascending id (the reference
 *                           adds in CPython set order: parity is 1e-13 absolute).  n_slots = face_off[n_owner] +
 *                           face_off[n_nb]; ws of b2g_mesh_workspace_bytes(n_cells, n_slots).
 *   b2g_mesh_internal_cells get_internal_cells (:229-248): mask [n_cells] uint8; ws: >= 256 device bytes.
 * n_bad_out (device int64[1]) counts out-of-range cell / vertex ids (the reference raises IndexError or wraps negative
 * ids; the host wrapper raises).  n_owner or n_nb > n_faces and n_nb > n_owner are B2G_E_ARG (IndexError upstream). */
int b2g_mesh_num_cells(const int32_t* owner, int64_t n_owner, const int32_t* neighbour, int64_t n_nb, int64_t* n_cells_out,
                       void* ws, void* stream);
int64_t b2g_mesh_workspace_bytes(int64_t n_cells, int64_t n_slots);
int b2g_mesh_cell_centers(const double* points, int64_t n_points, const int32_t* owner, int64_t n_owner,
                          const int32_t* neighbour, int64_t n_nb, const int64_t* face_off, const int32_t* face_pts,
                          int64_t n_faces, int64_t n_slots, int64_t n_cells, double* centers, int64_t* n_bad_out,
                          void* ws, int64_t ws_bytes, void* stream);
int b2g_mesh_internal_cells(const int32_t* owner, int64_t n_owner, const int32_t* neighbour, int64_t n_nb, int64_t n_cells,
                            uint8_t* mask, int64_t* n_bad_out, void* ws, void* stream);

/* ===================================================================================== training glue
 * Device side of torch_geometric.data.Batch.from_data_list (train.py:155): after the B samples' edge_index blocks have been
 * copied into place ([2, e_tot] row-major, sample s at columns edge_ptr[s]..edge_ptr[s+1]), add sample s's node offset
 * node_ptr[s] to its entries and write the per-node graph id vector batch[n_tot] (batch may be NULL).  edge_ptr / node_ptr:
 * int64 [B + 1] on the device. */
int b2g_batch_finalize(int64_t* edge_index, int64_t e_tot, const int64_t* edge_ptr, const int64_t* node_ptr, int n_graphs,
                       int64_t* batch, int64_t n_tot, void* stream);

/* The reference's field-wise weighted MSE loss (normalization.py:177-236, use_fieldwise=True): columns U(3) p k epsilon nut,
 * loss = sum_f w_f mean_f((pred - target)^2) + w_p * prw * (mean pred_p - mean target_p)^2.  field_weights: fp32 [5] (device);
 * loss: fp32 [1]; coef: fp32 [8] handed to b2g_wmse_bwd, which writes d loss / d pred * grad_out[0] (grad_out NULL = 1). */
int64_t b2g_wmse_workspace_bytes(void);
int b2g_wmse_fwd(const void* pred, int64_t ldp, const void* target, int64_t ldt, int64_t n, int dt, const float* field_weights,
                 float pressure_ref_weight, float* loss, float* coef, void* ws, void* stream);
int b2g_wmse_bwd(const void* pred, int64_t ldp, const void* target, int64_t ldt, int64_t n, int dt, const float* coef,
                 const float* grad_out, void* dpred, int64_t ldd, void* stream);

/* clip_grad_norm_(max_norm) + torch.optim.Adam(lr, betas, eps, weight_decay).step() (train.py:188-189) over ONE flat fp32
 * parameter / gradient / moment buffer: gradient norm (two-stage, deterministic), clip coefficient, Adam update, step counter.
 * state: fp32 [2] on the device = (steps taken, gradient norm of the last step); max_norm <= 0 = no clipping. */
int64_t b2g_adam_workspace_bytes(void);
int b2g_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float max_norm, float lr,
                       float beta1, float beta2, float eps, float weight_decay, float* state, void* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2G_H_ */
