/* gnnome_b200 -- C ABI of the B200-native GatedGCN + edge-score path.
 *
 * The reference (lbcb-sci/GNNome) is pure Python and has no FFI layer; its "operator API" for this
 * path is the nn.Module surface of layers/ and models/ (SURVEY.md section 8b).  The entry points
 * below are what a ctypes binding inside those modules calls; each cites the reference lines whose
 * arithmetic it replaces.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch); the library allocates nothing
 *    and keeps no state between calls.  All launches go to the caller's stream; no implicit sync.
 *  - return 0 on success, a cudaError_t (> 0) or a negative GNB_E_* code otherwise; the message is
 *    in gnb_last_error() (thread local).  Nothing aborts or throws across the ABI.
 *  - dense matrices are row-major fp32.  "k-major" weight = the transpose of nn.Linear.weight,
 *    i.e. Wt[k][n] = weight[n][k], so that consecutive outputs n are contiguous.
 *  - edge state lives in DST-SORTED order (position p), see gnb_graph_t; gnb_gather_rows converts.
 *  - H (hidden_features) must be one of 32, 64, 128, 256.
 */
#ifndef GNNOME_B200_H
#define GNNOME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNB_ABI_VERSION 11

#define GNB_E_INVALID   (-1)  /* bad argument (shape, alignment, unsupported H) */
#define GNB_E_WORKSPACE (-2)  /* workspace too small */
#define GNB_E_ARCH      (-3)  /* device is not sm_100 */

/* flags for gnb_edge_forward / gnb_node_update */
#define GNB_F_SYMMETRIC 1  /* SymGatedGCN: also aggregate over out-edges with A3h (gated_gcn_full.py:117-127) */
#define GNB_F_RESIDUAL  2  /* in_channels == out_channels (gated_gcn_full.py:24-25,108,136) */

/* Staged graph: the two CSR views the kernels walk.  Replaces what DGL builds behind
 * g.apply_edges / g.update_all / dgl.reverse (gated_gcn_full.py:99,104,112-113,117,125-126).
 * Position p in [0,E) enumerates edges sorted by (dst, original id); q enumerates them sorted by
 * (src, p).  All arrays int32, caller-allocated, filled by gnb_graph_stage. */
typedef struct gnb_graph {
  int64_t num_nodes;
  int64_t num_edges;
  int32_t* in_ptr;   /* [N+1] in-edges of node i are positions in_ptr[i] .. in_ptr[i+1]        */
  int32_t* in_src;   /* [E]   source node of the edge at position p                              */
  int32_t* in_dst;   /* [E]   destination node of the edge at position p (non-decreasing)        */
  int32_t* in_eid;   /* [E]   original (DGL) edge id of the edge at position p                   */
  int32_t* out_ptr;  /* [N+1] out-edges of node i are q in out_ptr[i] .. out_ptr[i+1]            */
  int32_t* out_pos;  /* [E]   position p of the q-th edge in src-sorted order                    */
  int32_t* out_dst;  /* [E]   destination node of that edge                                      */
} gnb_graph_t;

int gnb_abi_version(void);
const char* gnb_last_error(void);

/* Bytes of scratch gnb_graph_stage needs for (E, N). */
int gnb_graph_stage_workspace(int64_t num_edges, int64_t num_nodes, size_t* bytes);

/* Build both CSR views from the COO edge list (src[k], dst[k]) in original edge-id order.
 * g->num_nodes / num_edges and the seven array pointers must be set by the caller. */
int gnb_graph_stage(const int32_t* src, const int32_t* dst, gnb_graph_t* g,
                    void* workspace, size_t workspace_bytes, void* stream);

/* out[r][:] = W2 * relu(W1 * in[idx ? idx[r] : r][:] + b1) + b2
 * The two-layer encoders: models/full_graph.py:26-27, layers/node_encoder.py:28-33,
 * layers/edge_encoder.py:27-32.  W1 is [hid][in_f] (nn.Linear layout), W2t is k-major [hid][H]. */
int gnb_encode(const float* in, const int32_t* idx, int64_t rows, int in_f, int hid, int H,
               const float* W1, const float* b1, const float* W2t, const float* b2,
               float* out, void* stream);

/* out[r][0:M] = A[r][0:K] * Wt + bias, Wt k-major [K][M].  The node-side nn.Linear calls
 * A_1,A_2,A_3,B_1,B_2 (gated_gcn_full.py:91-96) as ONE product with the concatenated weight, and
 * the node half of ScorePredictor.W1 (score_predictor.py:13-14).  K % 16 == 0, M % 4 == 0. */
int gnb_node_linear(const float* A, int64_t rows, int K, const float* Wt, const float* bias, int M,
                    float* out, int64_t ld_out, void* stream);

/* ---- tensor-core (tcgen05) variants ---------------------------------------------------------------
 * The dense products run on the 5th-generation tensor cores with fp32 operands split on the fly into
 * fp16 (hi, lo) pairs and three MMAs accumulated in fp32 (TMEM): accuracy indistinguishable from the
 * fp32 FFMA kernels above.  Weights are pre-split once per parameter update. */

/* Bytes of the packed image of an [M][K] nn.Linear weight (ceil(M/128) blocks x (hi, lo) x 128 x K fp16). */
size_t gnb_packed_linear_bytes(int M, int K);

/* W[M][K] (nn.Linear layout, fp32) -> packed fp16 (hi, lo) blocks at Wp (16-byte aligned). K % 16 == 0. */
int gnb_pack_linear_tc(const float* W, int M, int K, void* Wp, void* stream);

/* Spin watchdog of the tensor-core kernels.  Every device-side barrier wait is bounded (default 10 s of wall
 * clock per wait; GNB_SPIN_TIMEOUT_MS in the environment or gnb_set_spin_timeout_ms override it, 0 = unbounded):
 * a wait that expires writes {kernel, block, thread, warp role, barrier, index, parity, tile iteration} into
 * host-mapped memory and traps, so a lost barrier phase becomes a CUDA launch failure at the next
 * synchronisation instead of a silent spin.  gnb_hang_report formats that record into buf (returns 1) or
 * returns 0 when no wait has expired in this process; it needs no working CUDA context. */
void gnb_set_spin_timeout_ms(long long ms);
int gnb_hang_report(char* buf, size_t cap);

/* Number of consecutive edge positions one aggregation chunk covers (carry granularity). */
int gnb_edge_chunk(int H);

/* Fused edge pass of one (Sym)GatedGCN layer over the dst-CSR (gated_gcn_full.py:97,104-114):
 *   z_p   = P[src_p][B1h] + P[dst_p][B2h] + e_p * We_t
 *   e'_p  = relu(z_p * scale_e + shift_e) (+ e_p if GNB_F_RESIDUAL)       -> written over e in place
 *   s_p   = sigmoid(e'_p)
 *   F_i   = sum_{p: dst_p = i} s_p * P[src_p][A2h] / (sum_p s_p + 1e-6)
 * P is the node table [N][ldP] written by gnb_node_linear with column blocks
 *   [0,2H): (B1h[c], A2h[c]) interleaved per channel c | [2H,3H): B2h | [3H,4H): A3h | [4H,5H): A1h
 * (non-symmetric layers have no A3h block: [3H,4H) is A1h).  b_B3 is folded into shift_e by the
 * caller (shift_e = bn shift + scale * b_B3); in eval mode scale/shift are the BatchNorm affine.
 * Segments that straddle a chunk boundary leave partial sums in carry[chunk][4][H]
 * (head num, head den, tail num, tail den); gnb_node_update resolves them, so F alone is only
 * meaningful together with carry.  carry has ceil(E / gnb_edge_chunk(H)) entries. */
int gnb_edge_forward(const gnb_graph_t* g, int H, const float* P, int64_t ldP,
                     const float* We_t, const float* scale_e, const float* shift_e,
                     float* e, float* F, float* carry, int flags, void* stream);

/* Reverse aggregation over the src-CSR fused with the node update (gated_gcn_full.py:124-137):
 *   Bk_i = sum_{q: src_q = i} sigmoid(e'_q) * P[dst_q][A3h] / (sum_q sigmoid(e'_q) + 1e-6)   (symmetric only)
 *   h'_i = relu((P[i][A1h] + F_i + Bk_i) * scale_h + shift_h) (+ h_i if GNB_F_RESIDUAL)
 * chunk = carry granularity of the edge pass that produced F / carry (gnb_edge_chunk or gnb_edge_chunk_tc2).
 * Only nodes node_begin <= i < node_end are updated (single GPU: 0, num_nodes; h_in / h_out need rows
 * for those nodes only).  Multi-GPU (graph partitioned by destination range, the staged graph holds the
 * owned nodes followed by halo source nodes): xp_ptr[i] .. xp_ptr[i+1] index xp_row, whose entries are
 * rows of xp_buf[.][2H] = (num | den) partial sums that OTHER ranks computed for node i with
 * gnb_reverse_partial; they are added, in that order, before the division.  Pass NULLs when unused.
 * Dropout (gated_gcn_full.py:139) is left to the caller (torch RNG semantics). */
int gnb_node_update(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const float* e,
                    const float* F, const float* carry, const float* h_in,
                    const float* scale_h, const float* shift_h, float* h_out, int flags,
                    int chunk, int64_t node_begin, int64_t node_end, const int32_t* xp_ptr,
                    const int32_t* xp_row, const float* xp_buf, void* stream);

/* The un-normalised halves of the reverse aggregation (gated_gcn_full.py:125-126) for the nodes
 * node_begin <= i < node_end of the staged (local) graph:
 *   out[i - node_begin][0:H]  = sum_{q: src_q = i} sigmoid(e'_q) * P[dst_q][A3h]
 *   out[i - node_begin][H:2H] = sum_{q: src_q = i} sigmoid(e'_q)
 * Used for halo source nodes, whose owner rank finishes the sum (gnb_node_update xp_*). */
int gnb_reverse_partial(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const float* e,
                        int64_t node_begin, int64_t node_end, float* out, void* stream);

/* ScorePredictor (score_predictor.py:12-24) with W1 = [W1s | W1d | W1e] split so that the node
 * halves are projected once per node:  S[n] = [x_n * W1s^T | x_n * W1d^T + b1]  ([N][2*hs], from
 * gnb_node_linear), then per edge
 *   score = W3 * relu(W2 * relu(S[src][0:hs] + S[dst][hs:2hs] + e_p * W1e_t) + b2) + b3
 * written to scores[in_eid[p]] (original edge order, [E] == the reference's (E,1)).
 * hs must be 32, 64 or 128; W2 is [32][hs] (nn.Linear layout), W3 is [32]. */
int gnb_score_forward(const gnb_graph_t* g, int H, int hs, const float* S, const float* W1e_t,
                      const float* W2, const float* b2, const float* W3, const float* b3,
                      const float* e, float* scores, void* stream);

/* out[r][0:W] = in[idx[r]][0:W]  (W % 4 == 0).  Moves edge rows between original order and
 * position order for the layer-level API. */
int gnb_gather_rows(const float* in, const int32_t* idx, int64_t rows, int W, float* out,
                    void* stream);
/* Same with explicit row strides (in floats, multiples of 4): packs W columns of selected rows of a wider
 * table, e.g. the (B1h, A2h) block of the node table for the multi-GPU halo exchange. */
int gnb_gather_rows_ld(const float* in, int64_t ld_in, const int32_t* idx, int64_t rows, int W,
                       float* out, int64_t ld_out, void* stream);
/* out[idx[r]][0:W] = in[r][0:W] */
int gnb_scatter_rows(const float* in, const int32_t* idx, int64_t rows, int W, float* out,
                     void* stream);

/* ---- split16 path: TMA-fed tcgen05 kernels (the product path for H in {64, 128, 256}) ------------------
 * Edge and node state X[rows][K] is held in HBM as one row-major fp16 matrix [rows][2K], row r =
 *   [ hi[r][0..K) | lo[r][0..K) ],  hi = fp16(x / 16),  lo = fp16(x / 16 - hi)     (x ~ 16 * (hi + lo), 22 significant bits):
 * 4 bytes per element like fp32 and a row is still contiguous, but the two halves are the tensor-core operands,
 * so tiles go HBM -> shared memory -> MMA through the TMA engine with no conversion instructions, and the updated
 * state goes back with TMA stores. */

/* Bytes of the split16 form of a [rows][K] matrix. */
size_t gnb_split16_bytes(int64_t rows, int K);

/* out16[r] = split(scale * in[idx ? idx[r] : r])   (fp32 rows -> images; idx = gnb_graph_t.in_eid moves edge rows from
 * edge-id order to dst-sorted position order).  K % 8 == 0.  scale: optional DEVICE scalar (NULL = 1); the training
 * path passes a power of two that brings a gradient tensor into the range the fp16 pair resolves (|x| ~ 2^-4 .. 2^16)
 * and undoes it with gnb_node_linear_tc2's out_scale -- gradients are ~1 / E small, activations are O(1). */
int gnb_split_rows(const float* in, const int32_t* idx, int64_t rows, int K, void* out16, const float* scale,
                   void* stream);
/* out[idx ? idx[r] : r] = merge(in16[r])   (images -> fp32 rows). */
int gnb_merge_rows(const void* in16, const int32_t* idx, int64_t rows, int K, float* out, void* stream);

/* gnb_encode with split16 and / or fp32 output (either may be NULL): models/full_graph.py:26-27.
 * in_f <= 4, hid <= 64. */
int gnb_encode2(const float* in, const int32_t* idx, int64_t rows, int in_f, int hid, int H,
                const float* W1, const float* b1, const float* W2t, const float* b2, void* out16,
                float* out32, void* stream);

/* gnb_node_linear on the tensor cores: X given as split16 images, the weight as gnb_pack_linear_tc(W[M][K]) output
 * (gated_gcn_full.py:91-96, score_predictor.py:13-14).  K in {64, 128, 256}.  out = out_scale * (X W^T) + bias with
 * out_scale an optional DEVICE scalar (NULL = 1). */
int gnb_node_linear_tc2(const void* X16, int64_t rows, int K, const void* Wp, const float* bias, int M,
                        float* out, int64_t ld_out, const float* out_scale, void* stream);

/* Edges per tile and carry granularity (edges per aggregation chunk) of gnb_edge_forward_tc2. */
int gnb_edge_tile_tc2(int H);
int gnb_edge_chunk_tc2(int H);

/* Tensor-core edition of gnb_edge_forward (gated_gcn_full.py:97,104-114) with the edge state e16 in split16 format
 * (updated in place) and carry sized ceil(E / gnb_edge_chunk_tc2(H)) x 4 x H.  The eval-mode norm affine is FOLDED
 * INTO THE OPERANDS by the caller, so the kernel has no scale / shift arguments:
 *   Wp = gnb_pack_linear_tc(diag(scale_e) * B_3.weight),   P[.][B1h block] = scale_e * B1h,
 *   P[.][B2h block] = scale_e * B2h + shift_e  (shift_e includes scale_e * b_B3);  the A2h / A3h / A1h blocks unchanged.
 *   e'_p = relu(e_p * Wp^T + P[src_p][B1h] + P[dst_p][B2h]) (+ e_p if GNB_F_RESIDUAL), F and carry as gnb_edge_forward.
 * H in {64, 128, 256}; for H = 256 the two 128-channel halves of a tile run on the two CTAs of a cluster that share
 * the tile through TMA multicast. */
int gnb_edge_forward_tc2(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const void* Wp,
                         void* e16, float* F, float* carry, int flags, void* stream);

/* Debugging aid: buf = device uint64[SMs][32 warps][5] (or NULL to switch off); every epilogue warp of
 * gnb_edge_forward_tc2 then leaves its cycle accounting there (full wait, accumulator wait, compute, hand-off, tiles). */
void gnb_debug_edge_timing(void* buf);

/* Fault injection for the liveness tests: the TMA-store thread of gnb_edge_forward_tc2 sleeps ns nanoseconds after
 * every tile it stores (0 = off).  The pipeline must terminate with unchanged results however slow that thread is.
 * ns < 0: the thread never releases a stage -- a deliberate dead-lock that the spin watchdog must turn into a
 * launch failure with a record (run it in a throw-away process: the CUDA context does not survive the trap). */
void gnb_debug_store_delay_ns(int ns);

/* L2 management of gnb_edge_forward_tc2, a bit mask: 1 = the TMA loads of the e tiles carry an evict-first policy,
 * 2 = the TMA stores of e' do, 4 = the node rows of a tile are prefetched into L2 when its input stage is free instead
 * of a tile period earlier, 8 = no prefetch; a negative value restores the library's default (8: what measured best at
 * BASELINE config 3; GNB_EDGE_MODE in the environment sets it at the first launch).  Results
 * do not depend on it; tools/edge_modes.py compares the settings. */
void gnb_debug_edge_mode(int mode);

/* gnb_node_update with e' read from split16 images; writes h' of the nodes node_begin .. node_end as fp32 rows
 * (h_out, row i) and, if h16_out is not NULL, as split16 images (h16_out, row i as well: both outputs are indexed
 * by the absolute node id) for the next layer's gnb_node_linear_tc2 (gated_gcn_full.py:117-137). */
int gnb_node_update2(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const void* e16,
                     const float* F, const float* carry, const float* h_in, const float* scale_h,
                     const float* shift_h, float* h_out, void* h16_out, int flags, int chunk,
                     int64_t node_begin, int64_t node_end, const int32_t* xp_ptr, const int32_t* xp_row,
                     const float* xp_buf, void* stream);
/* gnb_reverse_partial with e' read from split16 images. */
int gnb_reverse_partial2(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const void* e16,
                         int64_t node_begin, int64_t node_end, float* out, void* stream);

/* ScorePredictor on the tensor cores (score_predictor.py:12-24): the e * W1e^T product runs on tcgen05 with the
 * e tile brought in by TMA; Wp = gnb_pack_linear_tc(W1e [hs][H]) (zero-padded to 128 rows).  Other arguments
 * as gnb_score_forward. */
int gnb_score_forward_tc2(const gnb_graph_t* g, int H, int hs, const float* S, const void* Wp,
                          const float* W2, const float* b2, const float* W3, const float* b3,
                          const void* e16, float* scores, void* stream);

/* gnb_score_forward (CUDA cores) with the edge state given as split16 images: cross-check of the above. */
int gnb_score_forward2(const gnb_graph_t* g, int H, int hs, const float* S, const float* W1e_t,
                       const float* W2, const float* b2, const float* W3, const float* b3,
                       const void* e16, float* scores, void* stream);

/* ---- training-mode primitives (fp32 rows, edge rows in position order) ---------------------------------
 * train.py runs the model under autograd (train.py:141-183,329,346).  gnnome_b200/autograd.py composes these graph
 * primitives and their adjoints into the layer as layers/gated_gcn_full.py:82-142 does; nn.Linear products stay plain
 * library GEMMs.  All sums walk CSR ranges in a fixed order (no atomics).  H % 4 == 0; tables may have a row pitch. */

/* z[p] = A[in_src[p]] + B[in_dst[p]] (+ C[p])  -- apply_edges(u_add_v) + B_3(e), gated_gcn_full.py:104-105 */
int gnb_t_gather_add3(const gnb_graph_t* g, int H, const float* A, int64_t ldA, const float* B, int64_t ldB,
                      const float* C, float* out, void* stream);
/* out[i] = sum of X[p] over the in-edges (mode 0) / out-edges (mode 1) of node i: adjoint of the gathers */
int gnb_t_seg_sum(const gnb_graph_t* g, int H, const float* X, int mode, float* out, int64_t ldo, void* stream);
/* out[i] = sum sigma[p] * A[nbr_p] / (den[i] + 1e-6), den[i] = sum sigma[p]; mode 0: in-edges, nbr = src
 * (gated_gcn_full.py:112-114); mode 1: out-edges, nbr = dst (:125-127).  mode | 2 ("raw"): out[i] = the UN-NORMALISED
 * sum (the sharded training step adds the partial sums of several ranks before the division); in the two adjoints
 * below raw mode takes gout = d/d num, out = d/d den (per node) and ignores den (may be NULL):
 *   gsigma[p] (+)= gnum[i_p] * A[nbr_p] + gden[i_p],   gA[n] = sum gnum[i_p] * sigma[p]. */
int gnb_t_agg_fwd(const gnb_graph_t* g, int H, const float* A, int64_t ldA, const float* sigma, int mode,
                  float* den, float* out, void* stream);
/* gsigma[p] (+)= gout[i_p] / (den[i_p] + 1e-6) * (A[nbr_p] - out[i_p]) */
int gnb_t_agg_bwd_edge(const gnb_graph_t* g, int H, const float* gout, const float* out, const float* den,
                       const float* A, int64_t ldA, int mode, float* gsigma, int accumulate, void* stream);
/* gA[n] = sum over the edges whose neighbour end is n of gout[i_p] / (den[i_p] + 1e-6) * sigma[p] */
int gnb_t_agg_bwd_node(const gnb_graph_t* g, int H, const float* gout, const float* den, const float* sigma,
                       int mode, float* gA, int64_t ldg, void* stream);
/* e' = relu(ehat) (+ e_in if not NULL), sigma = sigmoid(e')  -- gated_gcn_full.py:107-111 */
int gnb_t_gate_fwd(const float* ehat, const float* e_in, int64_t rows, int H, float* e_out, float* sigma, void* stream);
/* t = g_e + g_sigma * sigma * (1 - sigma); g_ehat = t * [ehat > 0]; g_ein = t (if not NULL) */
int gnb_t_gate_bwd(const float* g_e, const float* g_sigma, const float* ehat, const float* sigma, int64_t rows,
                   int H, float* g_ehat, float* g_ein, void* stream);
/* out = a * (x - shift_x) (+ b * (y - shift_y)) + c with per-channel a, b, c and shifts (NULL = 0): BatchNorm normalise
 * (y NULL) and its input gradient.  The shifts carry the column means, so that a channel whose mean is large against its
 * spread keeps its digits (torch centres before scaling in the backward too). */
int gnb_t_affine2(const float* x, const float* y, const float* a, const float* b, const float* c, const float* shift_x,
                  const float* shift_y, int64_t rows, int H, float* out, void* stream);
/* nn.LayerNorm over the W channels of each row (gated_gcn_full.py:39-42): y = gamma * xhat + beta with
 * xhat = (x - mean_row) / sqrt(var_row + eps) (biased variance); xhat [rows][W] and rstd [rows] are kept for the backward */
int gnb_t_layer_norm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int W, float eps,
                         float* y, float* xhat, float* rstd, void* stream);
/* gx = rstd * (g*gamma - mean_row(g*gamma) - xhat * mean_row(g*gamma*xhat)); d gamma / d beta are the column sums of
 * g * xhat and g (gnb_t_col_stats(g, xhat)) */
int gnb_t_layer_norm_bwd(const float* g, const float* xhat, const float* rstd, const float* gamma, int64_t rows, int W,
                         float* gx, void* stream);
/* With a' = a - shift_a and b' = b - shift_b per channel (NULL shifts = 0): out[0:H] = column sums of a',
 * out[H:2H] = column sums of a' * b' (b NULL: a' * a'), in fp64, deterministic; workspace of
 * gnb_t_col_stats_workspace(rows, H) bytes.  BatchNorm1d batch statistics over all E / N rows: mean from a first
 * pass, variance about that mean from a second. */
size_t gnb_t_col_stats_workspace(int64_t rows, int H);
int gnb_t_col_stats(const float* a, const float* b, const float* shift_a, const float* shift_b, int64_t rows, int H,
                    double* out, void* workspace, void* stream);

/* ---- input features of the scoring pass (the callers' host-side torch code, moved to the device) ---------------------
 * x[i] = (in_degree(i), out_degree(i)) as floats -- g.in_degrees() / g.out_degrees() of utils/data_utils.py:50-51 --
 * in the column order of inference.py:420 / train.py:120; swap != 0 gives (out, in), the order train.py:117-118 uses for
 * the reversed graph.  x is [N][2], 8-byte aligned. */
int gnb_degree_rows(const gnb_graph_t* g, int swap, float* x, void* stream);
/* In place on x [rows][cols] (cols <= 4), for every column c whose bit is set in col_mask:
 *   x[r][c] = (x[r][c] - mean_c) / std_c,  std unbiased (rows - 1)
 * i.e. torch's (v - v.mean()) / v.std() of inference.py:416-419, train.py:113-116 and utils/data_utils.py:36.  Mean and
 * variance (about the mean, second pass) are accumulated in fp64 in a fixed order; the normalisation itself is fp32 like
 * the reference's.  stats (device, may be NULL) receives double[cols][2] = (mean, std).  A constant column gives nan
 * (0 / 0), one row gives nan, as in the reference.  workspace: gnb_zscore_workspace() bytes, 8-byte aligned. */
size_t gnb_zscore_workspace(void);
int gnb_zscore_cols(float* x, int64_t rows, int cols, int col_mask, double* stats, void* workspace, void* stream);

/* ---- node-induced subgraph: dgl.node_subgraph(g, keep, store_ids=True) (train.py:91-100 strand-wise masking; the
 * '_ID' maps the mini-batch code reads, train.py:125-135) --------------------------------------------------------------
 * keep [N] bytes (non-zero = kept); src / dst the parent's edge list in edge-id order.  Kept nodes are renumbered in
 * increasing order of their id; an edge is induced when both endpoints are kept, and induced edges keep the order of
 * their ids.  Two passes sharing one workspace of gnb_subgraph_workspace bytes (256-byte aligned):
 *   gnb_subgraph_count  counts[0] = kept nodes, counts[1] = induced edges (device int64[2]); the caller reads them to
 *                       size the outputs
 *   gnb_subgraph_fill   node_id [n'] / edge_id [e'] = the '_ID' maps, sub_src / sub_dst [e'] = renumbered endpoints
 *                       (an output may be NULL when its count is 0)
 * keep must be 4-byte, src / dst 16-byte aligned. */
int gnb_subgraph_workspace(int64_t num_nodes, int64_t num_edges, size_t* bytes);
int gnb_subgraph_count(const uint8_t* keep, const int32_t* src, const int32_t* dst, int64_t num_nodes,
                       int64_t num_edges, void* workspace, size_t workspace_bytes, int64_t* counts, void* stream);
int gnb_subgraph_fill(const uint8_t* keep, const int32_t* src, const int32_t* dst, int64_t num_nodes, int64_t num_edges,
                      void* workspace, int32_t* node_id, int32_t* edge_id, int32_t* sub_src, int32_t* sub_dst,
                      void* stream);

/* ---- greedy decoder walks: HOST functions, host pointers (inference.py:70-164, 231-305) ---------------------------------
 * The step after the scoring pass.  Successor lists as a CSR in the order of the reference's succs[i] lists (the order
 * decides exact ties), succ_edge[q] = edges[(i, succ_node[q])]; nodes come in strand pairs (2k, 2k + 1). */
typedef struct gnb_walk_graph {
  int64_t num_nodes;
  const int64_t* succ_ptr;   /* [N+1] */
  const int32_t* succ_node;  /* [sum of list lengths] */
  const int32_t* succ_edge;  /* [same] edge id of (i, succ_node[q]) */
} gnb_walk_graph_t;
/* run_greedy_both_ways (inference.py:164-168) for n_cand start edges against one shared visited map (bytes, NULL =
 * nothing visited), RANDOM / early_stopping off as shipped (:25-28): walk k = walk_buf[walk_off[k] .. walk_off[k+1]) =
 * walk_b + walk_f, back_len[k] = len(walk_b), sum_logp[2k], [2k+1] = sumLogProb_f, sumLogProb_b accumulated in fp32 in
 * the reference's order (bit-equal).  log_probs[eid] = log(sigmoid(score)) (:184).  threads <= 0: all hardware threads.
 * Returns GNB_E_WORKSPACE with walk_off filled when walk_cap < walk_off[n_cand]: size the buffer and call again. */
int gnb_greedy_walks(const gnb_walk_graph_t* g, const float* log_probs, const uint8_t* visited, int64_t n_cand,
                     const int32_t* cand_src, const int32_t* cand_dst, int threads, int32_t* walk_buf, int64_t walk_cap,
                     int64_t* walk_off, int64_t* back_len, float* sum_logp);
/* get_contig_length (inference.py:30-37): sum of prefix_length over the walk's edges + read_length of its last node */
int gnb_walk_contig_length(const gnb_walk_graph_t* g, const int64_t* prefix_length, const int64_t* read_length,
                           const int32_t* walk, int64_t len, int64_t* out);
/* The "jumped-over" nodes of an accepted walk (inference.py:316-322): for consecutive (ss, dd), every t in
 * succs[ss] & preds[dd] and its complement t ^ 1 get mark[t] = 1 (mark [N] bytes, not cleared).  pred: the CSR of the
 * predecessor lists (succ_edge unused). */
int gnb_walk_jumped_nodes(const gnb_walk_graph_t* succ, const gnb_walk_graph_t* pred, const int32_t* walk, int64_t len,
                          uint8_t* mark);

#ifdef __cplusplus
}
#endif
#endif /* GNNOME_B200_H */
