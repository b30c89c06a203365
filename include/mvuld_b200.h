/* mvuld_b200 -- C ABI of the B200-native MVulD forward hot path (libmvuld_b200.so).
 *
 * The reference (jacknichao/MVulD) has no FFI / operator table: its hot path is plain PyTorch + DGL + HF
 * transformers Python.  These entry points are what a reference-side binding would load
 * (ctypes stub in INTEGRATION.md) to replace, call site by call site, the library kernels the reference launches
 * implicitly.  Conventions:
 *   - every pointer is a DEVICE pointer unless stated; bf16/fp16 tensors are passed as void*;
 *   - sizes are plain ints; `stream` is a cudaStream_t (the caller's current stream); nothing allocates;
 *   - return 0 on success, a cudaError_t (>0) or -1/-2/-3 (bad argument / driver entry missing / TMA encode
 *     failure) otherwise; mvuld_last_error() returns the message for the calling thread;
 *   - file:line citations are relative to /root/reference.
 */
#ifndef MVULD_B200_H_
#define MVULD_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mvuld_stream_t; /* == cudaStream_t */

const char* mvuld_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Dense layers (tcgen05 GEMM, TMA-fed, fp32 accumulate in TMEM).
 * out[M,N] = act(A[M,K] @ W[N,K]^T + bias) + res      A, W bf16 row-major (nn.Linear weight layout [out, in]).
 * act: 0 none, 1 GELU(erf), 2 ELU.  Either/both of out_bf16 / out_f32 may be given; res_f32 (fp32 [M, ldr]) may
 * alias out_f32.  Replaces F.linear / nn.Linear / Conv1d(k=1) at: mvuld/models/swin_transformer_v2.py:26-32,177,361;
 * mvuld/models/GraphModel.py:153,159,171,176,186; mvuld/models/Rs_GCN.py:57,60,62,71; HF RobertaModel dense layers
 * (mvuld/models/unixcoder.py:36); DGL GATConv.fc, GatedGraphConv.linears / GRUCell (baselines/models/reveal/ggnn/
 * model.py:15-16).
 * ---------------------------------------------------------------------------------------------------------- */
int mvuld_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, int act,
                    const float* res_f32, int ldr, void* out_bf16, float* out_f32, int ldc, mvuld_stream_t stream);

/* Dense layer fused with a full-row LayerNorm and the res-post-norm residual (tcgen05 GEMM whose fp32 accumulator row
 * lives in the 512 TMEM columns; N in {128, 256, 512}):
 *   x = shortcut + LayerNorm(A W^T + bias) * gamma + beta      shortcut (fp32 [M,N], may alias x32) and bias optional;
 * outputs fp32 x32 and/or bf16 xb.  Replaces proj + norm1 + residual and fc2 + norm2 + residual of
 * mvuld/models/swin_transformer_v2.py:177,301 and :30,304, and PatchMerging reduction + norm (:361-362). */
int mvuld_gemm_ln_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                       const float* gamma, const float* beta, float eps, const float* shortcut_f32, float* x32,
                       void* xb, mvuld_stream_t stream);
/* Same contract for rows wider than one CTA can double-buffer in TMEM: N in {512, 768, 1024}; a thread-block cluster of
 * N / 256 CTAs owns a 128-row tile, each CTA one 256-column block, row statistics exchanged through distributed shared
 * memory (swin_transformer_v2.py:301,304; shortcut may be null: PatchMerging :361-362). */
int mvuld_gemm_ln_wide_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                            const float* gamma, const float* beta, float eps, const float* shortcut_f32, float* x32,
                            void* xb, mvuld_stream_t stream);

/* SwinV2 Mlp + res-post-norm in ONE kernel for the narrow stages (C in {128, 256}):
 *   x = shortcut + LayerNorm(fc2(GELU(fc1(X) + b1)) + b2) * gamma + beta
 * X bf16 [M, C] contiguous, W1 bf16 [4C, C], W2 bf16 [C, 4C]; shortcut fp32 [M, C] (may alias x32); outputs fp32 x32
 * and/or bf16 xb (xb may alias X).  The hidden activation [M, 4C] stays in shared memory / TMEM (csrc/mlp_ln.cu);
 * bit-identical to mvuld_gemm_bf16(act = GELU) followed by mvuld_gemm_ln_bf16.
 * Replaces mvuld/models/swin_transformer_v2.py:26-32 (Mlp.forward) + :304 (norm2 + residual). */
int mvuld_mlp_ln_bf16(const void* X, const void* W1, const float* b1, const void* W2, const float* b2,
                      const float* gamma, const float* beta, float eps, const float* shortcut_f32, float* x32, void* xb,
                      int M, int C, mvuld_stream_t stream);

/* SwinV2 qkv projection fused with: cat(q_bias, 0, v_bias), per-head L2 normalisation of q and k, the learnable
 * logit scale (qscale[h] = exp(min(logit_scale_h, ln 100)) * log2 e, folded into q), window partition and cyclic
 * shift (pure index math).  X bf16 [B*H*W, C]; q,k fp16 and v bf16, each [B*nW, nH, ws*ws, 32].
 * Replaces mvuld/models/swin_transformer_v2.py:147-157 and :276-286. */
int mvuld_swin_qkv(const void* X, const void* Wqkv, const float* q_bias, const float* v_bias, const float* qscale,
                   void* q, void* k, void* v, int B, int H, int W, int C, int nH, int ws, int shift,
                   mvuld_stream_t stream);

/* Training variant of mvuld_swin_qkv: also keeps rq / rk = 1 / max(|q|, eps), 1 / max(|k|, eps) per (token, head), fp32
 * in the window-major order of q / k, for the backward of F.normalize (swin_transformer_v2.py:155). */
int mvuld_swin_qkv_train(const void* X, const void* Wqkv, const float* q_bias, const float* v_bias, const float* qscale,
                         void* q, void* k, void* v, float* rq, float* rk, int B, int H, int W, int C, int nH, int ws,
                         int shift, mvuld_stream_t stream);
/* RoBERTa qkv projection, head-major outputs bf16 [B, nH, L, hd]; q multiplied by qmul (= log2 e / sqrt(hd)).
 * Wqkv = cat(query.weight, key.weight, value.weight) [3*Hd, Hd].  HF RobertaSelfAttention via unixcoder.py:36. */
int mvuld_heads_qkv(const void* X, const void* Wqkv, const float* bias, void* q, void* k, void* v, int B, int L,
                    int Hd, int nH, float qmul, mvuld_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Attention (tcgen05 QK^T and PV, softmax on the TMEM accumulator, one CTA per (window|sequence, head)).
 * ---------------------------------------------------------------------------------------------------------- */
/* Continuous position bias table, once per weight version: table_ref[h, (2ws-1)^2] = 16*sigmoid(cpb_mlp(coords))
 * in the reference's order; table_rev = same * log2 e with the w axis reversed (kernel layout); table_max[h].
 * w1 [512,2], b1 [512], w2 [nH,512].  swin_transformer_v2.py:98-111,159,163. */
int mvuld_cpb_table(const float* w1, const float* b1, const float* w2, int nH, int ws, int pretrained_ws,
                    float* table_rev, float* table_ref, float* table_max, mvuld_stream_t stream);

/* softmax(q k^T + bias[rel_pos_index] + shift_mask) v, written token-major bf16 [B*H*W, C] with window_reverse and
 * the inverse cyclic shift folded into the store.  ws in {7, 14, 28}; shift in {0, ws/2}.
 * swin_transformer_v2.py:155-176 and :292-299.  q_norm (optional, fp32 [nH]): the norm the qkv epilogue gave each
 * head's queries (exp(min(logit_scale, ln 100)) * log2 e); heads with 2 q_norm + bias_max <= 100 use a constant softmax
 * reference (no maximum pass, no rescaling), the others -- and all heads when q_norm is null -- a running maximum. */
int mvuld_swin_window_attention(const void* q, const void* k, const void* v, const float* bias_rev,
                                const float* bias_max, const float* q_norm, void* out, int B, int H, int W, int C,
                                int nH, int ws, int shift, mvuld_stream_t stream);
/* The same operation for launches in which EVERY head meets the constant-reference condition 2 q_norm + bias_max <= 100
 * (checked by the caller once per weight version; a violating head traps) and ws == 28: three softmax warpgroups on
 * (query tile, key tile) units, P aliased onto S in TMEM, the row sum taken by the PV product (N = 48). */
int mvuld_swin_window_attention_fixed(const void* q, const void* k, const void* v, const float* bias_rev,
                                      const float* bias_max, const float* q_norm, void* out, int B, int H, int W,
                                      int C, int nH, int ws, int shift, mvuld_stream_t stream);

/* Training forward of the two entry points above: also writes lse, the log2-domain log-sum-exp of every score row
 * (fp32 [windows * nH, ws^2], window-major like q), which mvuld_swin_attention_bwd recomputes the probabilities from.
 * fixed != 0: the constant-reference kernel (ws == 28, every head checked by the caller). */
int mvuld_swin_window_attention_train(const void* q, const void* k, const void* v, const float* bias_rev,
                                      const float* bias_max, const float* q_norm, void* out, float* lse, int fixed,
                                      int B, int H, int W, int C, int nH, int ws, int shift, mvuld_stream_t stream);
/* Backward of the window attention (autograd through swin_transformer_v2.py:155-176, trained by mvuld/main.py:251-300).
 * prep: gathers dO (bf16 token-major [B*H*W, C], gradient of the attention output before proj) into the window-major
 * head-major order of q / k / v (dOw), packs ld[row] = (lse, D = rowsum(dO o O)) and makes bf16 copies qb / kb of the
 * fp16 q^ / k^.  bwd: per (window, head) recomputes P from lse, G = P o (dO V^T - D), and writes fp32 [windows*nH, ws^2, 32]
 *   dq = G k^,  dk = G^T q^ (q^ as stored: logit scale and log2 e folded in),  dv = P^T dO,
 * plus gt = G^T as bf16 [windows*nH, ws^2, ntok_pad] (key major; ntok_pad = ws^2 rounded up to 8; null to skip) for
 * mvuld_swin_bias_grad.  G = dL / d(natural-unit logits).  ws in {7, 14, 28}.  Deterministic (no atomics). */
int mvuld_swin_attention_bwd_prep(const void* dO, const void* O, const float* lse, const void* qh, const void* kh,
                                  void* dOw, void* ld, void* qb, void* kb, int B, int H, int W, int C, int nH, int ws,
                                  int shift, mvuld_stream_t stream);
int mvuld_swin_attention_bwd(const void* qh, const void* qb, const void* kh, const void* kb, const void* v,
                             const void* dOw, const void* ld, const float* bias_rev, float* dq, float* dk, float* dv,
                             void* gt, int ntok_pad, int B, int H, int W, int nH, int ws, int shift,
                             mvuld_stream_t stream);

/* Bias-table gradient from the G^T matrices of mvuld_swin_attention_bwd: dtab[h, (qy-ky+ws-1) * (2ws-1) + (qx-kx+ws-1)]
 * += sum over the n_win window instances and all (query, key) pairs of that offset (the reference's table order,
 * natural units).  partial: fp32 workspace [mvuld_swin_bias_grad_splits(n_win, nH, ws), nH, ws, ws, 2ws-1] (the window
 * instances are split over blocks; splits are summed in split order).  Fixed summation order. */
int mvuld_swin_bias_grad(const void* gt, int n_win, int nH, int ws, int npad, float* partial, float* dtab,
                         mvuld_stream_t stream);
int mvuld_swin_bias_grad_splits(int n_win, int nH, int ws);
/* cpb_mlp backward (swin_transformer_v2.py:98-111,159,163): tab = table_ref of mvuld_cpb_table, dtab as above;
 * accumulates dw1 [512,2], db1 [512], dw2 [nH,512]. */
int mvuld_cpb_mlp_bwd(const float* w1, const float* b1, const float* w2, const float* tab, const float* dtab, int nH,
                      int ws, int pretrained_ws, float* dw1, float* db1, float* dw2, mvuld_stream_t stream);
/* (dq, dk, dv) of mvuld_swin_attention_bwd -> d qkv, bf16 token-major [B*H*W, 3C] (q | k | v columns): backward of
 * F.normalize with the saved inverse norms rq / rk, window_reverse and the inverse cyclic shift; accumulates
 * dlogit_scale [nH] (zero past the clamp at ln 100).  ls_partial: fp32 [mvuld_swin_qkv_bwd_blocks(...), nH]. */
int mvuld_swin_qkv_bwd_blocks(int B, int H, int W, int nH);
int mvuld_swin_qkv_bwd(const float* dq, const float* dk, const float* dv, const void* qh, const void* kh, const float* rq,
                       const float* rk, const float* qscale, const float* logit_scale, void* dqkv, float* dlogit_scale,
                       float* ls_partial, int B, int H, int W, int C, int nH, int ws, int shift, mvuld_stream_t stream);
/* Training-mode plumbing: erf GELU bf16 -> bf16 (the pre-activation is kept for mvuld_gelu_bwd); the inverse of
 * mvuld_patch_merge_gather for fp32 gradients; PatchEmbed's 4x4 patches as a bf16 [tokens, 48] matrix (tap order of
 * proj.weight.view(E, 48)) so that the projection and its weight gradient run on the GEMM. */
int mvuld_gelu_fwd(const void* x, void* y, long long n, mvuld_stream_t stream);
int mvuld_patch_merge_scatter(const float* dg, float* dx, int B, int H, int W, int C, mvuld_stream_t stream);
int mvuld_patch_im2col(const float* img, void* out, int B, int Hi, int Wi, mvuld_stream_t stream);

/* Key-padded self-attention for the text encoder: q,k,v bf16 [B, nH, L, 64] (q pre-scaled), kv_len int32 [B];
 * out bf16 [B*L, nH*64].  unixcoder.py:35-36. */
int mvuld_seq_attention(const void* q, const void* k, const void* v, const int* kv_len, void* out, int B, int L,
                        int nH, int hd, mvuld_stream_t stream);
/* Training forward of mvuld_seq_attention: also writes lse (log2-domain log-sum-exp of every row, fp32 [B*nH, L]). */
int mvuld_seq_attention_train(const void* q, const void* k, const void* v, const int* kv_len, void* out, float* lse,
                              int B, int L, int nH, int hd, mvuld_stream_t stream);
/* Backward of the RoBERTa self-attention (HF RobertaSelfAttention under autograd, unixcoder.py:33-38): prep gathers dO
 * (bf16 token-major [B*L, nH*64]) head-major and packs ld = (lse, rowsum(dO o O)); bwd writes fp32 [B*nH, L, 64]
 * dq = G k, dk = G^T q_stored, dv = P^T dO (G = dL / d natural logits; rows / key tiles past kv_len are skipped: the
 * caller zero-fills); qkv_bwd turns them into d(x Wqkv^T + b), bf16 token-major [B*L, 3*nH*64] (q | k | v). */
int mvuld_seq_attention_bwd_prep(const void* dO, const void* O, const float* lse, void* dOh, void* ld, int B, int L,
                                 int nH, mvuld_stream_t stream);
int mvuld_seq_attention_bwd(const void* q, const void* k, const void* v, const void* dOh, const void* ld,
                            const int* kv_len, float* dq, float* dk, float* dv, int B, int L, int nH, int hd,
                            mvuld_stream_t stream);
int mvuld_seq_qkv_bwd(const float* dq, const float* dk, const float* dv, void* dqkv, int B, int L, int nH, int hd,
                      mvuld_stream_t stream);
/* Packed variant (several short sequences per row, block-diagonal attention): token (b, i) attends to keys
 * [seg_lo[b*L + i], seg_hi[b*L + i]) of row b; padding tokens carry their own position (lo = i, hi = i + 1) so that
 * they stay finite.  kv_len[b] = tokens in use in row b.  This is what makes the per-node line encoding of
 * mvuld/data/data_list.py:292-299 / unixcoder.py:56-68 (every line padded to 512 tokens in the reference) affordable.
 * tile_lo / tile_hi (optional, both or neither): int32 [B, ceil(L / 128)], the range of 128-key tiles that query tile
 * (b, t) has to visit -- the union of its rows' key ranges; the other key tiles of the row are skipped. */
int mvuld_seq_attention_packed(const void* q, const void* k, const void* v, const int* kv_len, const int* seg_lo,
                               const int* seg_hi, const int* tile_lo, const int* tile_hi, void* out, int B, int L,
                               int nH, int hd, mvuld_stream_t stream);
/* out[out_row[s] (or s when null), :] = mean of token rows [seg_start[s], seg_start[s] + seg_len[s]) of tok fp32
 * [T, C] (unixcoder.py:37 per sequence). */
int mvuld_seq_segment_mean(const float* tok, const int* seg_start, const int* seg_len, const int* out_row, float* out,
                           int n, int C, mvuld_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Row kernels of the image / text branches.
 * ---------------------------------------------------------------------------------------------------------- */
/* mode 0: x = LN(y); 1: x = shortcut + LN(y) (res-post-norm, swin_transformer_v2.py:301,304);
 * 2: x = LN(y + shortcut) (RoBERTa).  y bf16 [M,C]; shortcut fp32; outputs fp32 x32 and/or bf16 xb. */
int mvuld_ln_rows(const void* y, const float* shortcut, const float* gamma, const float* beta, float* x32, void* xb,
                  int M, int C, float eps, int mode, mvuld_stream_t stream);
/* Training variant of mvuld_roberta_embed: also keeps ysum = word + position + type rows (bf16 [M, C], the LayerNorm
 * input).  Backward: mvuld_ln_rows_bwd (mode 0) on ysum, then mvuld_embed_grad_rows per table. */
int mvuld_roberta_embed_train(const long long* ids, const int* pos, const float* word, const float* posemb,
                              const float* type0, const float* gamma, const float* beta, float* x32, void* xb, void* ysum,
                              int M, int C, float eps, mvuld_stream_t stream);
/* Backward of the masked mean (unixcoder.py:37): dtok[b, t] = dsent[b] / len[b] for t < len[b], else 0. */
int mvuld_masked_mean_bwd(const float* dsent, const int* len, float* dtok, int B, int L, int C, mvuld_stream_t stream);
/* nn.Embedding backward: dtab[v] += sum of the rows d[rows[k]], k in [indptr[v], indptr[v+1]) (rows grouped by index
 * with mvuld_csr_from_coo, token order kept: fixed summation order); index skip_index (padding_idx) gets nothing. */
int mvuld_embed_grad_rows(const float* d, const int* indptr, const int* rows, float* dtab, int n_index, int C,
                          int skip_index, mvuld_stream_t stream);
/* PatchEmbed: Conv2d(3,E,4,4) + LayerNorm, img fp32 NCHW -> x32 / xb [B*(H/4)*(W/4), E]. swin_transformer_v2.py:485-493 */
int mvuld_patch_embed(const float* img, const float* w, const float* bias, const float* gamma, const float* beta,
                      float* x32, void* xb, int B, int Himg, int Wimg, int E, float eps, mvuld_stream_t stream);
/* PatchMerging 2x2 gather-concat in the order (0,0),(1,0),(0,1),(1,1). swin_transformer_v2.py:352-359 */
int mvuld_patch_merge_gather(const void* xb, void* out, int B, int H, int W, int C, mvuld_stream_t stream);
/* final LayerNorm + mean over tokens -> fp32 [B, C]. swin_transformer_v2.py:632-634 */
int mvuld_ln_meanpool(const float* x, const float* gamma, const float* beta, float* out, int B, int T, int C,
                      float eps, mvuld_stream_t stream);
/* position ids (cumsum(ids != pad) * (ids != pad) + pad), valid length per sequence, and a flag cleared when a
 * non-pad token follows a pad (the kv-length attention path requires suffix padding, as unixcoder.py:150 produces). */
int mvuld_seq_positions(const long long* ids, int B, int L, int pad, int* pos, int* len, int* suffix_ok,
                        mvuld_stream_t stream);
/* word + position + token_type[0] embeddings + LayerNorm (HF RobertaEmbeddings). */
int mvuld_roberta_embed(const long long* ids, const int* pos, const float* word, const float* posemb,
                        const float* type0, const float* gamma, const float* beta, float* x32, void* xb, int M, int C,
                        float eps, mvuld_stream_t stream);
/* masked mean over tokens (unixcoder.py:37). */
int mvuld_masked_mean(const float* tok, const int* len, float* out, int B, int L, int C, mvuld_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Graph branch (CSR gather / segment reduce).
 * ---------------------------------------------------------------------------------------------------------- */
/* In-edge CSR sorted by (dst, edge id) from int64 COO (dgl.graph + dgl.batch order).  workspace == NULL queries
 * *workspace_bytes.  status[0] = 1 if an endpoint is outside [0, N). */
int mvuld_csr_from_coo(const long long* src, const long long* dst, int E, int N, void* workspace,
                       size_t* workspace_bytes, int* indptr, int* idx_src, int* eids, int* status,
                       mvuld_stream_t stream);
/* dgl.add_self_loop + dgl.batch on the device (mvuld/data/data_list.py:314, mvuld/data/bigvul_dataset.py:177-205):
 * raw per-graph edge lists with LOCAL int32 node ids concatenated graph by graph (etype_local int64 or null), edge_off /
 * node_off int64 [B + 1] exclusive prefix sums of the per-graph counts.  Output (int64, DGL order): per graph its edges
 * in input order shifted by node_off[k], then -- if add_self_loops -- its N_k loops (i, i) with zero edge data.
 * total_out = E_raw (+ N).  *status is set to 1 if a local id is outside [0, N_k). */
int mvuld_collate_edges(const int* src_local, const int* dst_local, const long long* etype_local,
                        const long long* edge_off, const long long* node_off, int B, int add_self_loops, long long* src,
                        long long* dst, long long* etype, long long total_out, int* status, mvuld_stream_t stream);
/* etype_sorted[i] = uint8(etype[eids[i]]); status[0] = 1 if any etype outside [0, n_etypes) (DGL asserts). */
int mvuld_gather_etype(const long long* etype, const int* eids, int E, int n_etypes, unsigned char* out, int* status,
                       mvuld_stream_t stream);
/* a[dst] = sum over in-edges of msgs[src, etype, :]; msgs bf16 [N, T, D], out bf16 [N, D] with row stride ldo.
 * DGL GatedGraphConv message + reduce (baselines/models/reveal/ggnn/model.py:23, devign/model.py:35). */
int mvuld_ggnn_gather_sum(const void* msgs, const int* indptr, const int* idx_src, const unsigned char* etype,
                          void* out, int ldo, int N, int T, int D, mvuld_stream_t stream);
/* GRUCell of GatedGraphConv as ONE GEMM with the gates in its epilogue: A = [a | h] bf16 [M, K = 2D] (row stride lda),
 * Wg bf16 [4D, K] with rows 4j..4j+3 = (W_ir | W_hr), (W_iz | W_hz), (W_in | 0), (0 | W_hn) of feature j, bias4 fp32
 * [4D] = (b_ir + b_hr, b_iz + b_hz, b_in, b_hn) interleaved; h32 fp32 [M, D] updated in place, the bf16 state is
 * written to hb_out (row stride ldhb), which must not be the buffer A is read from. */
int mvuld_gemm_gru(const void* A, int lda, const void* Wg, int ldw, int M, int D, int K, const float* bias4, float* h32,
                   void* hb_out, int ldhb, mvuld_stream_t stream);
/* h0 = cat(x, 0) zero padding of GatedGraphConv; hb has row stride ldb. */
int mvuld_ggnn_init(const float* x, float* h32, void* hb, int ldb, long long N, int in_dim, int D,
                    mvuld_stream_t stream);
/* per-graph segment sum (reveal/ggnn/model.py:26-28,46-56): feat fp32 [N, D], offsets int64 [B+1] -> out [B, D]. */
int mvuld_segment_sum(const float* feat, const long long* offsets, float* out, int B, int D, mvuld_stream_t stream);
/* GATConv pieces (GraphModel.py:99-105,167-170): el/er scores, then edge-softmax + weighted aggregate + bias. */
int mvuld_gat_scores(const void* z, const float* attn_l, const float* attn_r, float* el, float* er, int N, int H,
                     int F, mvuld_stream_t stream);
int mvuld_gat_aggregate(const void* z, const float* el, const float* er, const int* indptr, const int* idx_src,
                        const float* bias, void* out, int N, int H, int F, float slope, int* zero_deg_flag,
                        mvuld_stream_t stream);
/* unbatch_features pad/truncate (GraphModel.py:30-54) fused with BatchNorm1d(max_node) over the slot axis (:135,186)
 * as a per-slot affine; gather_map (int64 [B, max_node], -1 = zero row) is optional. */
int mvuld_unbatch_pad_bn(const void* feat, const long long* offsets, const float* bn_scale, const float* bn_shift,
                         void* out, long long* gather_map, int B, int max_node, int F, mvuld_stream_t stream);
/* ELU(fc_bbox(bn_bbox(pad(pos)))) into columns [col0, col0+OUT) of the concat buffers (GraphModel.py:187,189). */
int mvuld_pos_branch(const float* pos, const long long* offsets, const float* bn_scale, const float* bn_shift,
                     const float* w, const float* bias, float* z32, void* zb, int B, int max_node, int OUT, int ld,
                     int col0, mvuld_stream_t stream);
int mvuld_f32_to_bf16(const float* in, void* out, long long n, mvuld_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fusion branch.
 * ---------------------------------------------------------------------------------------------------------- */
/* Rs_GCN.py:57-66: tpg fp32 [B*n, 3C] = (theta | phi | g); y = (theta phi^T / n) g; r_out optional fp32 [B, n, n] (the
 * affinity the reference returns).  y3 = bf16x3 split operand [B*n, 3C] = (hi | lo | hi) of the fp32 y,
 * to be multiplied against a (W_hi | W_hi | W_lo) weight (mvuld_split3_bf16, w_side = 1): fp32-class products on the
 * bf16 tensor-core GEMM.  Rs_GCN has no softmax and feeds a BatchNorm: plain bf16 operands cost 1-2 % per block. */
int mvuld_rs_gcn_affinity_f32(const float* tpg, void* y3, float* r_out, int B, int n, int C, mvuld_stream_t stream);
/* x fp32 [R, C] (row stride ldx) -> out bf16 [R, 3C]: (hi | lo | hi) for w_side == 0, (hi | hi | lo) otherwise. */
int mvuld_split3_bf16(const float* x, int ldx, void* out, int R, int C, int w_side, mvuld_stream_t stream);
/* GraphModel.py:200-209: l2norm(dim=1) + mean + concat + BN(folded) + Linear -> logits fp32 [B, num_classes].
 * Same kernel for the RQ2 ablations (mvuld/models/new_model.py): mode 0 = cat(img, graph, txt) as above; mode 1 =
 * cat(img, graph), wf [num_classes, 2D] (Multi_DefectModel_noFunc, new_model.py:317-318); mode 2 = txt * graph,
 * wf [num_classes, D] (Multi_DefectModel_noGlobalImage, new_model.py:196-197).  feat_out (optional) has that width. */
int mvuld_fusion_head_mode(const float* z, const float* img, const float* txt, const float* wf, const float* bf,
                           float* logits, float* feat_out, int B, int n, int D, int num_classes, int mode,
                           mvuld_stream_t stream);

/* Small-N fp32 linear for classification heads (swin_transformer_v2.py:642; reveal/ggnn/model.py:29-30):
 * out[M,N] = x[M,K] w[N,K]^T + b; out_sigmoid (optional) = sigmoid(out). */
int mvuld_linear_small(const float* x, const float* w, const float* b, float* out, float* out_sigmoid, int M, int N,
                       int K, mvuld_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fusion ablation classes (mvuld/models/GraphModel.py:618-1274, mvuld/models/myModels.py:280-428).
 * ---------------------------------------------------------------------------------------------------------- */
/* Per-node ELU(Linear(4, OUT)(pos_emb)) into bf16 columns [col0, col0 + OUT) of a [N, ld] buffer
 * (GraphModel.py:791 _GATPOS fc_bbox, :1132 _NOGAT3, :1242 _NOGAT4).  pos fp32 [N, 4], w fp32 [OUT, 4]. */
int mvuld_node_linear4(const float* pos, const float* w, const float* bias, void* out_bf16, int N, int OUT, int ld,
                       int col0, mvuld_stream_t stream);
/* ELU(bn_gat(pad(h))) -> fp32 + bf16 [B * max_node, F] (GraphModel.py:923-928, _011: no fc_gat behind the slot BN). */
int mvuld_unbatch_pad_bn_elu(const void* feat, const long long* offsets, const float* bn_scale, const float* bn_shift,
                             float* z32, void* zb, int B, int max_node, int F, mvuld_stream_t stream);
/* nn.GRU(H, H, 1, batch_first=True) recurrence, h_0 = 0, -> last hidden state fp32 [B, H] (myModels.py:324,385-387).
 * gi fp32 [B, T, 3H] = x_t W_ih^T + b_ih (gate order r | z | n); w_hh fp32 [3H, H]; b_hh fp32 [3H];
 * workspace: mvuld_gru_sequence_workspace(B, H) bytes.  Cooperative launch (one grid barrier per step). */
long long mvuld_gru_sequence_workspace(int B, int H);
int mvuld_gru_sequence(const float* gi, const float* w_hh, const float* b_hh, float* h_out, void* workspace, int B,
                       int T, int H, mvuld_stream_t stream);
/* mode 0: out[b, col0 + c] = softmax_c(tanh(x[b, c] * h[b, c])) * h[b, c]  (myModels.py:407-413, fusion 'attention');
 * mode 1: out[b, col0 + c] = x[b, c] * h[b, c] (myModels.py:419, fusion 'dot').  fp32, C <= 1024. */
int mvuld_gate_fusion(const float* x, const float* h, float* out, int B, int C, int ld, int col0, int mode,
                      mvuld_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training step of the fusion model (mvuld/main_bigvul.py:294-342 drives GraphModel.py:150-211 in train mode with
 * CrossEntropyLoss, clip_grad_norm_(5.0) and AdamW).  Dense backward passes reuse mvuld_gemm_bf16 on transposed
 * operands; these entry points are the non-GEMM pieces.  Activation gradients bf16 / fp32, parameter gradients fp32
 * (accumulated: callers zero them once per step).
 * ---------------------------------------------------------------------------------------------------------- */
/* out[c, r] = in[r, c] (bf16, in row stride ldi), out row stride ldo >= R with zero fill: operand transposes of
 * dW = dY^T X. */
/* The same for a table of matrices in ONE launch: desc = nmat x {in, out, R, C, ldi, ldo} (six 64-bit words each, device
 * memory), tile_end[k] = exclusive end of matrix k's 32 x 32 tiles (ceil(ldo / 32) x ceil(C / 32) per matrix) in the
 * flattened grid, total_tiles = tile_end[nmat - 1]. */
int mvuld_transpose_bf16_batched(const long long* desc, const int* tile_end, int nmat, int total_tiles,
                                 mvuld_stream_t stream);
int mvuld_transpose_bf16(const void* in, int ldi, void* out, int R, int C, int ldo, mvuld_stream_t stream);
/* out[c] += sum_r x[r, c] (bias gradients); x bf16 (is_bf16 != 0) or fp32, row stride ldx >= C. */
/* Weight gradient dW[n_out, k_in] = dY^T X (fp32, row stride ld_dw, overwritten) from ROW-major bf16 dY [M, n_out] and
 * X [M, k_in]: both operands are consumed MN-major by tcgen05 (no transposed copies), the M rows are split over CTAs
 * and the partial tiles summed in split order (bit-reproducible).  partials: fp32 workspace of
 * mvuld_gemm_dw_workspace(M, n_out, k_in) floats (may be null when that is 0). */
long long mvuld_gemm_dw_workspace(int M, int n_out, int k_in);
int mvuld_gemm_dw(const void* dY, int ld_dy, const void* X, int ld_x, float* dW, int ld_dw, float* partials, int M,
                  int n_out, int k_in, mvuld_stream_t stream);
int mvuld_colsum(const void* x, int is_bf16, int ldx, float* out, float* partials, int R, int C, mvuld_stream_t stream);
/* rows of the fp32 [slabs, C] partials workspace of mvuld_colsum (null allowed when this returns 1): the row slabs are
 * summed in slab order, so the column sums are bit-reproducible. */
int mvuld_colsum_slabs(int R, int C);
/* dx = dy * ELU'(pre) through y = dropout(ELU(pre), p) with the mask regenerated from seed (F.elu + nn.Dropout,
 * GraphModel.py:171,176); is_f32 selects fp32 tensors (no dropout). */
int mvuld_elu_bwd(const void* dy, const void* y, void* dx, long long n, int is_f32, unsigned long long seed, float p,
                  mvuld_stream_t stream);
/* out = mask(seed) * x / (1 - p), bf16: nn.Dropout forward, and the backward of GATConv's feat_drop. */
int mvuld_dropout_bf16(const void* x, void* out, long long n, unsigned long long seed, float p, mvuld_stream_t stream);
/* BatchNorm1d in training mode over the rows of x fp32 [R, C] (swinbn, bn_text, final_fc_bn, Rs_GCN W[1]):
 * saves mean / rstd, updates the running statistics when given; y32 (row stride ldy) = BN(x) + res (row stride ldr,
 * optional: the Rs_GCN residual, Rs_GCN.py:70; may alias y32). */
int mvuld_bn_cols_fwd(const float* x, const float* gamma, const float* beta, float eps, const float* res, int ldr,
                      float* y32, int ldy, void* yb, float* mean, float* rstd, float* run_mean, float* run_var,
                      float momentum, int R, int C, mvuld_stream_t stream);
/* dy has row stride ldy (a column slice of the [B, 1536] feature row); x, dx32, dxb are dense [R, C]. */
int mvuld_bn_cols_bwd(const float* x, const float* dy, int ldy, const float* gamma, const float* mean,
                      const float* rstd, float* dx32, void* dxb, float* dgamma, float* dbeta, int R, int C,
                      mvuld_stream_t stream);
/* BatchNorm1d(max_node) over the node-slot axis of x bf16 [B, n, F] (bn_gat / bn_bbox, GraphModel.py:135,186). */
int mvuld_bn_slot_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                      float* rstd, float* run_mean, float* run_var, float momentum, int B, int n, int F,
                      mvuld_stream_t stream);
int mvuld_bn_slot_bwd(const void* x, const void* dy, const float* gamma, const float* mean, const float* rstd,
                      void* dx, float* dgamma, float* dbeta, int B, int n, int F, mvuld_stream_t stream);
/* Row kernels of the encoder backward (the "whole path trainable" reading of configs[4]; driven by SwinTrainer /
 * RobertaTrainer).  LayerNorm backward of the three mvuld_ln_rows forms: dout fp32 [M,C] = gradient of the LN output (mode 1:
 * the block output's gradient, which is also the shortcut's); the LN input is recomputed from y (bf16) (+ shortcut in
 * mode 2); dv = gradient of the LN input as bf16 and / or fp32; dgamma / dbeta accumulated.  swin_transformer_v2.py:301,
 * 304,362; HF RobertaSelfOutput / RobertaOutput. */
int mvuld_ln_rows_bwd(const void* y, const float* shortcut, const float* gamma, const float* dout, void* dv_bf16,
                      float* dv_f32, float* dgamma, float* dbeta, float* dbias, float* partials, int M, int C, float eps,
                      int mode, mvuld_stream_t stream);
/* dbias (may be null): += column sums of dv, i.e. the bias gradient of the dense layer whose output the LayerNorm
 * normalises (proj / fc2, RoBERTa output.dense) -- saves a separate pass over dv.
 * rows of the fp32 [rows, 3, C] partials workspace mvuld_ln_rows_bwd needs for M rows: dgamma / dbeta are summed over
 * the blocks in a fixed order (bit-reproducible gradients, no atomics). */
int mvuld_ln_rows_bwd_blocks(int M);
/* GELU backward fused with the column sums of its result (fc1's bias gradient): dpre [M, C] bf16 = dh * GELU'(pre),
 * dbias[c] += sum_rows dpre[:, c] (fixed summation order).  partials: fp32 [mvuld_colsum_slabs(M, C), C] (may be null
 * when that is 1).  C %% 8 == 0. */
int mvuld_gelu_bwd_colsum(const void* pre, const void* dh, void* dpre, float* dbias, float* partials, int M, int C,
                          mvuld_stream_t stream);
/* exact (erf) GELU backward: dpre = dh * GELU'(pre), bf16, n %% 8 == 0 (Mlp, swin_transformer_v2.py:26-32). */
int mvuld_gelu_bwd(const void* pre, const void* dh, void* dpre, long long n, mvuld_stream_t stream);
/* fp32 strided ELU backward with a bf16 result (image / text projections, GraphModel.py:153-159). */
int mvuld_elu_bwd_rows(const float* dy, int ldy, const float* y, int ldyy, void* dx, int ldx, int R, int C,
                       mvuld_stream_t stream);
/* bn_bbox (GraphModel.py:137,187) batch statistics of the padded [B, n, 4] box tensor -> the per-slot affine that
 * mvuld_pos_branch applies; and the backward of ELU(fc_bbox(bn_bbox(.))) (dpre = ELU-backpropagated bf16 columns
 * [col0, col0 + OUT) of a [B*n, ld] gradient), accumulating dW [OUT, 4], db, dgamma, dbeta. */
int mvuld_pos_slot_stats(const float* pos, const long long* offsets, const float* gamma, const float* beta, float eps,
                         float* scale, float* shift, float* mean, float* rstd, float* run_mean, float* run_var,
                         float momentum, int B, int n, mvuld_stream_t stream);
int mvuld_pos_branch_bwd(const float* pos, const long long* offsets, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, const float* w, const void* dpre, float* dw, float* db,
                         float* dgamma, float* dbeta, float* partials /* fp32 [n, 160] workspace */, int B, int n, int OUT,
                         int ld, int col0, mvuld_stream_t stream);
/* backward of unbatch_features pad / truncate (GraphModel.py:30-54). */
int mvuld_unbatch_pad_bwd(const void* dhp, const long long* offsets, void* dh, int B, int max_node, int F,
                          mvuld_stream_t stream);
/* GATConv backward: edge-softmax backward per destination (in-CSR), gather over out-edges per source (out-CSR sorted
 * by src, pos_in[k] = in-CSR position of out-edge k), attention vector gradients (accumulated).  Scratch: alpha_e,
 * ds_e fp32 [E, H]; del, der fp32 [N, H]. */
int mvuld_gat_bwd(const void* z, const void* dout, const float* el, const float* er, const int* indptr,
                  const int* idx_src, const int* out_indptr, const int* out_dst, const int* pos_in,
                  const float* attn_l, const float* attn_r, float* alpha_e, float* ds_e, float* del, float* der,
                  void* dz, float* dattn_l, float* dattn_r, int N, int H, int F, float slope, mvuld_stream_t stream);
/* Rs_GCN.py:57-66 backward: dtpg = (dtheta | dphi | dg) from dy. */
int mvuld_rs_gcn_affinity_bwd(const void* tpg, const void* dy, void* dtpg, int B, int n, int C, mvuld_stream_t stream);
/* l2norm over the node axis + node mean (GraphModel.py:200-204), forward into a strided feature row, and backward. */
int mvuld_l2norm_mean_fwd(const float* z, float* out, int ldo, float* inv_s, int B, int n, int D, mvuld_stream_t stream);
int mvuld_l2norm_mean_bwd(const float* z, const float* inv_s, const float* dm, int ldm, float* dz32, void* dzb, int B,
                          int n, int D, mvuld_stream_t stream);
/* CrossEntropyLoss (main_bigvul.py:298,332): loss_sum += scale * sum CE; dlogits = scale * (softmax - onehot). */
int mvuld_ce_loss(const float* logits, const long long* labels, float* loss_sum, float* dlogits, int B, int C,
                  float scale, mvuld_stream_t stream);
int mvuld_linear_small_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int M,
                           int N, int K, mvuld_stream_t stream);
/* out += sum x^2 (global gradient norm of clip_grad_norm_, utils_multi.py:233).  partials: scratch of 1184 floats; the
 * reduction order is fixed, so identical (all-reduced) gradients give a bit-identical norm on every rank. */
int mvuld_sumsq_f32(const float* x, long long n, float* partials, float* out, mvuld_stream_t stream);
/* clipped AdamW on a flat fp32 buffer with per-segment weight decay (optimizer.py:11-33: decay / no-decay groups). */
int mvuld_adamw(float* p, const float* g, float* m, float* v, long long n, const long long* seg_end,
                const float* seg_wd, int nseg, const float* gnorm_sq, float max_norm, float lr, float beta1,
                float beta2, float eps, int step, mvuld_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVULD_B200_H_ */
