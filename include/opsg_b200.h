/* libopsg_b200 — C ABI of the B200 (sm_100a) relation-head hot path.
 *
 * Drop-in boundary for OpenPSG's RelationTransformerHeadV4 inference path
 * (reference: kings_sgg/models/relation_heads/relation_transformer_head_v4.py:107-358).  The reference is
 * pure Python calling PyTorch / HuggingFace eager ops, so "the FFI for this path" is the set of eager
 * calls the head makes; each entry point below names the reference call sites it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - opsg_bf16 is a raw bfloat16 (uint16_t storage); matrices are row-major with explicit leading dims
 *     in ELEMENTS; bf16 matrices consumed by the tensor-core kernels need 16-byte aligned bases and
 *     leading dimensions that are multiples of 8 elements (TMA requirement);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued, nothing synchronises the host;
 *   - no allocation happens inside the library (split-K workspaces etc. are caller-provided);
 *   - return value: 0 = OPSG_OK, negative = error; opsg_last_error_string() describes the last error of
 *     the calling thread.  There is no CPU fallback: on a machine without an sm_100 device every compute
 *     entry point returns OPSG_E_NO_DEVICE.
 */
#ifndef OPSG_B200_H_
#define OPSG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t opsg_bf16;

enum {
  OPSG_OK = 0,
  OPSG_E_INVALID = -1,     /* bad argument (shape, alignment, null pointer) */
  OPSG_E_CUDA = -2,        /* CUDA runtime / driver error; see opsg_last_error_string() */
  OPSG_E_NO_DEVICE = -3,   /* no CUDA device, or device is not compute capability 10.x */
  OPSG_E_UNSUPPORTED = -4  /* valid request outside the implemented envelope (e.g. L > 256) */
};

enum { OPSG_ACT_NONE = 0, OPSG_ACT_GELU = 1, OPSG_ACT_RELU = 2 };
enum { OPSG_OUT_BF16 = 0, OPSG_OUT_F32 = 1, OPSG_OUT_F32_ATOMIC = 2 };

int opsg_version(void);
const char* opsg_last_error_string(void);
/* 0 when the current device can run the kernels (compute capability 10.x), else OPSG_E_NO_DEVICE. */
int opsg_device_check(void);
int opsg_num_sms(void);

/* ---- a3 / K2: panoptic map -> per-object token bitmasks ------------------------------------------
 * Replaces v4:416-429 (F.interpolate(nearest) -> F.pad(0) -> F.interpolate(nearest) -> `== object_id`).
 * pan: int32 [pan_h, pan_w]; obj_ids: int32 [num_objects]; bits_out: uint32 [num_objects, words],
 * bit (l % 32) of word (l / 32) is set iff token l = ty * tok_w + tx belongs to the object.
 * words >= ceil(tok_h*tok_w / 32); unused high bits are written as zero.  The N^2 `logical_or` pair masks
 * of v4:430-433 are never materialised: consumers OR two rows of bits_out.  Bit-exact. */
int opsg_pair_mask_bits(const int32_t* pan, int pan_h, int pan_w, int img_h, int img_w, int pad_h, int pad_w,
                        int tok_h, int tok_w, const int32_t* obj_ids, int num_objects, uint32_t* bits_out,
                        int words, void* stream);

/* ---- a3 / K1: PatchEmbed operand layout -----------------------------------------------------------
 * Replaces the input side of timm PatchEmbed's Conv2d(k=s=patch) at v4:410: feat fp32 [C, h, w] ->
 * bf16 [L, C*patch*patch] (L = (h/patch)*(w/patch), row-major tokens, K order (c, py, px) = the conv
 * weight's own layout), so the projection is one opsg_gemm_bf16 against weight.reshape(C_out, -1). */
int opsg_patch_im2col(const float* feat, int channels, int h, int w, int patch, opsg_bf16* out, void* stream);

/* ---- K3 / K6 / K9 / K10c: D = act(A . W^T + bias + residual) on tcgen05 tensor cores ----------------
 * Replaces every nn.Linear on the path (HF modeling_instructblip.py:499-501,504,548,598,606; v4:208,294;
 * OPT q/k/v/out/fc1/fc2/lm_head).  A: bf16 [M, K] (lda), W: bf16 [N, K] (ldw) = nn.Linear.weight,
 * D: [M, N] (ldd) bf16 or fp32 per out_mode.  bias: fp32 [N] (or [M] when bias_along_m != 0), may be NULL.
 * residual: bf16 [M, N] (ldr) added before the activation's output is stored, may be NULL.
 * k_splits > 1 requires out_mode == OPSG_OUT_F32_ATOMIC: each split atomically adds its partial product
 * into D (caller pre-initialises D, e.g. with the bias); bias/residual/act must then be NULL/NONE. */
int opsg_gemm_bf16(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, void* D, int ldd, int M, int N, int K,
                   const float* bias, int bias_along_m, const opsg_bf16* residual, int ldr, int act, int out_mode,
                   int k_splits, void* stream);

/* LayerNorm-folding variant (CTA-pair kernel) for the Q-Former's Linear -> (+residual) -> LayerNorm -> Linear chains
 * (HF :542-553, 598-610, 687-695): activations stay UN-normalised in memory together with per-row (sum, sum of
 * squares) statistics and every consumer applies the pending LayerNorm itself, so no separate LayerNorm pass runs:
 *   a_stats  != NULL : A holds un-normalised rows x with statistics a_stats[M, 2] over their K features; W must be the
 *                      consumer weight with the LayerNorm's gamma folded in (W'[n,k] = gamma[k] W[n,k]), a_colsum[n] =
 *                      sum_k W'[n,k], and bias[n] must already include sum_k beta[k] W[n,k].  The kernel computes
 *                      rstd_m (acc[m,n] - mean_m a_colsum[n]) + bias[n]  ==  LayerNorm(x_m) . W^T + bias.
 *   r_stats  != NULL : the residual tensor is un-normalised too; (residual - mean) rstd r_gamma[n] + r_beta[n] is added.
 *   stats_out != NULL: (sum, sum of squares) of every output row (fp32, before the bf16 rounding) are accumulated with
 *                      atomics into stats_out[M, 2], which the caller zeroes beforehand.
 * bf16 output only; N, lda, ldw, ldd, ldr multiples of 8; eps = the LayerNorm epsilon. */
int opsg_gemm_bf16_ln(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, opsg_bf16* D, int ldd, int M, int N, int K,
                      const float* bias, const opsg_bf16* residual, int ldr, int act, const float* a_stats,
                      const float* a_colsum, const float* r_stats, const float* r_gamma, const float* r_beta,
                      float* stats_out, float eps, void* stream);

/* Small-M variant for weight-streaming GEMMs (LLM decode: M = number of selected pairs <= 128, OPT q/k/v/out/fc1/
 * fc2 of one decode step; v4:305-312).  Same operands and epilogue as opsg_gemm_bf16 (bias along N only).  K is cut
 * into slices of <= 12 K-blocks; a CTA keeps its slice of the activations resident in tensor memory (the MMA's A
 * operand) and streams only weights, 64 rows x 64 columns per pipeline stage; the fp32 partial rows of the slices go
 * through a caller-provided workspace of opsg_gemm_streamk_workspace_bytes(N, K) bytes and a second kernel sums them in
 * a fixed order and applies bias / activation / residual (csrc/gemm_skinny.cu).  Deterministic (no atomics).
 * Layouts that kernel does not take (N, ldd or ldr not a multiple of 4) and OPSG_SKINNY=0 use the older stream-K
 * decomposition over 256-wide tiles (gemm.cu) behind the same entry.  out_mode: OPSG_OUT_BF16 or OPSG_OUT_F32. */
size_t opsg_gemm_streamk_workspace_bytes(int N, int K);
int opsg_gemm_bf16_streamk(const opsg_bf16* A, int lda, const opsg_bf16* W, int ldw, void* D, int ldd, int M, int N,
                           int K, const float* bias, const opsg_bf16* residual, int ldr, int act, int out_mode,
                           void* workspace, size_t workspace_bytes, void* stream);

/* fp32 [rows, cols] (ld_in) -> bf16 (ld_out); init_rows_f32 broadcasts a row vector (bias) into [rows, cols]. */
int opsg_cast_f32_bf16(const float* in, int ld_in, opsg_bf16* out, int ld_out, int rows, int cols, void* stream);
int opsg_init_rows_f32(float* out, int ld_out, const float* row, int rows, int cols, void* stream);

/* ---- a4 / K7: Q-Former embeddings -----------------------------------------------------------------
 * Replaces InstructBlipQFormerEmbeddings.forward (HF :753-782) for B pair queries.
 * query: fp32 [n_query, d] (= cat(rel_cls_query, relation_query), v4:155-157); input_ids int32 [B, T];
 * word_emb/pos_emb fp32 tables; h_out bf16 [B*n_query + B*T, d]: row p*n_query+q = LN(query[q]),
 * row B*n_query + p*T + t = LN(word_emb[ids[p,t]] + pos_emb[t]). */
int opsg_qformer_embed_ln(const float* query, int n_query, const int32_t* input_ids, int B, int T,
                          const float* word_emb, int vocab, const float* pos_emb, const float* gamma,
                          const float* beta, float eps, int d, opsg_bf16* h_out, void* stream);

/* y = LayerNorm(x) * gamma + beta over the last dim; x,y bf16 [rows, cols] (same ld = cols); cols % 8 == 0, <= 8192. */
int opsg_layernorm_bf16(const opsg_bf16* x, const float* gamma, const float* beta, float eps, opsg_bf16* y,
                        int rows, int cols, void* stream);

/* ---- a5 / K4: Q-Former self-attention core --------------------------------------------------------
 * Replaces the score/softmax/context part of InstructBlipQFormerMultiHeadAttention.forward (HF :504-536)
 * for B pair queries whose rows live in the split layout of opsg_qformer_embed_ln.  qkv: bf16 [R, 3*d]
 * (q | k | v, heads of head_dim inside each); text_mask int32 [B, T] (0 = padded key: the reference's
 * -10000 additive bias underflows to weight exactly 0); ctx_out bf16 [R_out, d].  Queries: the n_query
 * query rows of every pair and, if text_queries != 0, the T text rows too.  shared_query_qkv (bf16 [n_query, 3*d] or
 * NULL): q/k/v of the query rows when they are identical for every pair -- layer 0, where they are the projection of
 * LN(query tokens) and the reference recomputes them B times; the query rows of qkv are then never read. */
int opsg_self_attn_small(const opsg_bf16* qkv, const opsg_bf16* shared_query_qkv, const int32_t* text_mask, int B,
                         int n_query, int T, int num_heads, int head_dim, int text_queries, opsg_bf16* ctx_out,
                         void* stream);

/* ---- a6 / K5: pair-query x image-feature masked cross-attention (the north-star kernel) ------------
 * Replaces the score/softmax/context part of the cross-attention MHA (HF :499-536) as called with
 * encoder_hidden_states = image tokens expanded to N^2 copies and encoder_attention_mask = pair masks
 * (v4:168-170,183-184).  q: bf16 [B*n_query, d] (row p*n_query + r), k: bf16 [L, d] (ld_k),
 * vt: bf16 [d, ld_vt] = V transposed (row h*head_dim + e, col = key), both projected ONCE per image;
 * bits: uint32 [N, words] from opsg_pair_mask_bits; pair_index int32 [B] or NULL (p -> pair p);
 * pair p = (i, j) = (pair / N, pair % N) attends keys where bits[i] | bits[j]; masked keys get the
 * reference's finfo.min bias (weight exactly 0; an all-masked pair attends uniformly to all L keys).
 * ctx_out: bf16 [B*n_query, d].  Requires head_dim == 64.  L <= 256 image tokens run on the tcgen05 kernels (one 256-key score
 * tile per unit in tensor memory); longer inputs (the reference has no limit, v4:408-435) take an online-softmax mma.sync kernel
 * that reads the mask bits directly (pass bias_tiles == NULL: the operand tiles and opsg_token_order cover L <= 256 only).
 *
 * bias_tiles (optional, recommended): the pair masks pre-arranged as tensor-core operand tiles by
 * opsg_xattn_bias_tiles() into a caller-provided buffer of opsg_xattn_bias_tiles_bytes(B, n_query) bytes.
 * They depend only on (bits, pair_index, N, B, n_query, L) — one build per image serves every head of both
 * Q-Former layers — and let the kernel apply the mask as an additive bias inside the QK^T MMA.  With
 * bias_tiles == NULL the self-contained (slower) kernel that ORs the bit rows per score row is used. */
/* Key order for opsg_xattn_pairs: perm_out int32 [L] lists the image tokens sorted by owning object (the first object of
 * `bits` whose mask holds the token; unowned tokens last; stable), bits_sorted_out [num_objects, words] are the same masks
 * in that order.  Softmax attention does not depend on the order of the keys (HF:ib:499-536 as called at v4:183-184); with
 * an object's keys contiguous, most 16-key chunks are invisible to a 32-row group of pair queries and K5 skips their
 * exponentials.  Project K / V from the token rows gathered by perm_out and pass bits_sorted_out to the two calls below. */
int opsg_token_order(const uint32_t* bits, int words, int num_objects, int L, int32_t* perm_out, uint32_t* bits_sorted_out,
                     void* stream);
size_t opsg_xattn_bias_tiles_bytes(int B, int n_query);
int opsg_xattn_bias_tiles(const uint32_t* bits, int words, const int32_t* pair_index, int num_objects, int B,
                          int n_query, int L, void* tiles_out, void* stream);
int opsg_xattn_pairs(const opsg_bf16* q, const opsg_bf16* k, int ld_k, const opsg_bf16* vt, int ld_vt,
                     const uint32_t* bits, int words, const int32_t* pair_index, int num_objects, int B,
                     int n_query, int L, int num_heads, int head_dim, const void* bias_tiles, opsg_bf16* ctx_out,
                     void* stream);

/* ---- a8 / K8: relation-existence filter -----------------------------------------------------------
 * Replaces v4:206-209 (Linear(768,1) + sigmoid on out[:,0]) and v4:236-237 (topk(B).indices[:k]).
 * x: bf16 rows of length d with row stride ld_x elements (out[:,0] = every n_query-th row);
 * w fp32 [d], b fp32 [1]; logits_out/probs_out fp32 [B]; mask_out uint8 [B] = prob > threshold decided
 * in logit space; topk_out int32 [k] = indices of the k largest logits, ties -> lower index,
 * descending.  k <= B, B <= 65536. */
int opsg_exist_filter_topk(const opsg_bf16* x, int ld_x, int B, int d, const float* w, const float* b,
                           float threshold, int k, float* logits_out, float* probs_out, uint8_t* mask_out,
                           int32_t* topk_out, void* stream);

/* ---- a11 / K11: per-mask feature pooling + pair gather ---------------------------------------------
 * Replaces detectors/openseed_relation.py:441-527 (same code in mask2former_relation.py:275-295 and
 * mask2former_relation_v2.py:392-465).
 * opsg_mask_pool_labels: the reference's mask chain (:441-462: mask = pan == id -> nearest to img_shape -> zero pad to
 *   pad_shape -> nearest to the feature size) for ALL objects at once: label_out int32 [feat_h, feat_w] = index of the
 *   first listed object whose id equals the source pixel, num_objects where nobody does (padding, unlisted ids);
 *   rep_out int32 [num_objects] = first object with the same id (objects repeating an id share its mask).
 * opsg_mask_pool_pairs: obj = sum(feat * mask) / (sum(mask) + 1e-8) (:466-468), optional class embedding
 *   cls_table[cls_ids[o]] added (cls_mode 1) or concatenated (2) (:469-474), optional background feature
 *   sum(feat * (1 - mask)) / (sum(1 - mask) + 1e-8) added (:487-493); pair = cat(obj[i], obj[j]) (:502-527).
 *   feat fp32 [C, h, w] is read once; no atomics, fixed summation order (bit-reproducible).  obj_out fp32 [N, C'] with
 *   C' = C (+ cls_dim when concatenating), pair_out fp32 [N*N, 2C'] or NULL; workspace of
 *   opsg_mask_pool_workspace_bytes bytes.  num_objects <= 255. */
int opsg_mask_pool_labels(const int32_t* pan, int pan_h, int pan_w, int img_h, int img_w, int pad_h, int pad_w, int feat_h,
                          int feat_w, const int32_t* obj_ids, int num_objects, int32_t* label_out, int32_t* rep_out,
                          void* stream);
size_t opsg_mask_pool_workspace_bytes(int channels, int h, int w, int num_objects);
int opsg_mask_pool_pairs(const float* feat, int channels, int h, int w, const int32_t* label, const int32_t* rep,
                         int num_objects, const float* cls_table, const int32_t* cls_ids, int cls_dim, int cls_mode,
                         int use_background, float* workspace, size_t workspace_bytes, float* obj_out, float* pair_out,
                         void* stream);

/* ---- a9-a10 / K9-K10: LLM prefix assembly, attention, greedy step ---------------------------------
 * opsg_gather_rows_bf16: out[r] = src[idx[r]] for contiguous row blocks (pair_feature[si], v4:294).
 * opsg_embed_gather: out[r] = table[ids[r]] (+ pos_table[pos[r]] if given), bf16 out (v4:296; OPT :64-70).
 * opsg_llm_attn: causal multi-head attention over a static KV cache; q bf16 [nseq*q_len, H*hd],
 *   k/v cache bf16 [nseq, max_ctx, H*hd]; key_mask uint8 [nseq, max_ctx] (0 = padded key); query t of a
 *   sequence sits at absolute position q_pos0 + t and sees keys <= its position (HF OPT :75-101); q_pos0 + q_len <= 256,
 *   head_dim 64, 80 or 128.
 * opsg_argmax_rows: greedy token = argmax over fp32 logits rows (ties -> lower index). */
int opsg_gather_rows_bf16(const opsg_bf16* src, int row_elems, const int32_t* idx, int n_rows, opsg_bf16* out,
                          void* stream);
int opsg_embed_gather(const opsg_bf16* table, int d, const int32_t* ids, const opsg_bf16* pos_table,
                      const int32_t* pos, int n_rows, opsg_bf16* out, int ld_out, void* stream);
/* opsg_llm_build_prefix (v4:294-301 + HF OPT :56-70, :321-330): the embedded prompt of nseq selected pairs,
 *   out[s, t, :] = (t < n_prefix ? proj[s*proj_rows_per_seq + proj_row0 + t, :] : table[ids[s, t - n_prefix], :])
 *                  + pos_table[pos[s, t], :]      (pos_table may be NULL: no learned positions, e.g. RoPE models)
 *   proj = language_projection applied to the gathered Q-Former rows (33 rows per pair, row 0 = cls row is skipped
 *   with proj_row0 = 1, n_prefix = 32); ids int32 [nseq, T] left-padded prompt tokens; pos int32 [nseq, n_prefix+T];
 *   out bf16 [nseq, n_prefix + T, d] contiguous. */
int opsg_llm_build_prefix(const opsg_bf16* proj, int proj_rows_per_seq, int proj_row0, int n_prefix,
                          const opsg_bf16* table, const int32_t* ids, int T, const opsg_bf16* pos_table,
                          const int32_t* pos, int nseq, int d, opsg_bf16* out, void* stream);
int opsg_llm_attn(const opsg_bf16* q, int ld_q, const opsg_bf16* k_cache, const opsg_bf16* v_cache, int max_ctx,
                  const uint8_t* key_mask, int nseq, int q_len, int q_pos0, int num_heads, int head_dim, float scale,
                  opsg_bf16* out, int ld_out, void* stream);
/* Decode step (one new token per sequence at position q_pos0): qkv bf16 [nseq, ld_qkv] holds the fused [q | k | v]
 * projection of the new tokens; the kernel attends over cache keys < q_pos0 plus the new key and WRITES the new k / v
 * rows into the caches at position q_pos0 (opsg_kv_append + opsg_llm_attn in one launch).  q_pos0 + 1 <= 256. */
int opsg_llm_attn_append(const opsg_bf16* qkv, int ld_qkv, opsg_bf16* k_cache, opsg_bf16* v_cache, int max_ctx,
                         const uint8_t* key_mask, int nseq, int q_pos0, int num_heads, int head_dim, float scale,
                         opsg_bf16* out, int ld_out, void* stream);
int opsg_kv_append(const opsg_bf16* qkv, int ld_qkv, int nseq, int q_len, int pos0, int d_model, opsg_bf16* k_cache,
                   opsg_bf16* v_cache, int max_ctx, void* stream);
int opsg_argmax_rows(const float* logits, int ld, int rows, int cols, int32_t* out, void* stream);

/* ---- a10 / K10 for Llama-family decoders (the LLM configs/psg/baseline_v4_ov.py:60-61 names; v4:99-105) ------------
 * opsg_rmsnorm_bf16: y = weight * (x * rsqrt(mean(x^2) + eps)) over the last dim (HF modeling_llama.py:52-69, statistics
 *   in fp32); x, y bf16 [rows, cols] with leading dims ld_x / ld_y; cols % 8 == 0, cols <= 8192.
 * opsg_rope_bf16: rotary position embedding, HF rotate_half convention (modeling_llama.py:137-166), applied IN PLACE to
 *   the first n_parts column blocks of num_heads * head_dim elements of every row of x (bf16 [rows, ld]; n_parts = 2 for
 *   the q | k part of a fused q | k | v projection).  Row r uses position pos[r] (HF generate: cumsum(attention_mask) - 1);
 *   cos_table / sin_table fp32 [table_rows, head_dim / 2] hold cos / sin(pos * theta^(-2e / head_dim)) as
 *   LlamaRotaryEmbedding computes them (:120-135).  head_dim % 16 == 0.
 * opsg_swiglu_bf16: out[r, c] = silu(gate_up[r, c]) * gate_up[r, ffn + c] (modeling_llama.py:181-183 with gate_proj and
 *   up_proj fused into one [2 * ffn, d] projection); gate_up bf16 [rows, ld_gu >= 2 * ffn], out bf16 [rows, ld_out]. */
int opsg_rmsnorm_bf16(const opsg_bf16* x, int ld_x, const float* weight, float eps, opsg_bf16* y, int ld_y, int rows,
                      int cols, void* stream);
int opsg_rope_bf16(opsg_bf16* x, int ld, int rows, int n_parts, int num_heads, int head_dim, const int32_t* pos,
                   const float* cos_table, const float* sin_table, int table_rows, void* stream);
int opsg_swiglu_bf16(const opsg_bf16* gate_up, int ld_gu, int rows, int ffn, opsg_bf16* out, int ld_out, void* stream);

/* ---- a9 / a10 bookkeeping that the reference leaves to torch / HF generate ------------------------------------------
 * opsg_llm_prompt_layout: positions and key mask of the batched prompt [n_prefix projected rows ; left-padded text]
 *   (v4:294-301): with m = [1 x n_prefix ; text_mask] and c = cumsum(m), pos_out int32 [nseq, n_prefix + T] =
 *   m ? c - 1 + pos_offset : max(pos_offset - 1, 0)  (pos_offset 2 = OPT learned positions, HF opt :64-70; 0 = rotary
 *   positions as HF generate derives them, generation/utils.py:707-729); key_mask_out uint8 [nseq, n_prefix + T +
 *   max_new_tokens] (generated positions 1); last_rows_out int32 [nseq] = row of the last prompt token in the
 *   [nseq * (n_prefix + T)] row stack; dec_pos_out int32 [max_new_tokens - 1, nseq] = position of the token fed at
 *   decode step j + 1.
 * opsg_copy_bytes: dst = src, device to device, as a kernel (CUDA-graph input staging off the copy engines).
 * opsg_transpose_i32: dst[c, r] = src[r, c]. */
int opsg_llm_prompt_layout(const int32_t* text_mask, int nseq, int T, int n_prefix, int max_new_tokens, int pos_offset,
                           int32_t* pos_out, uint8_t* key_mask_out, int32_t* last_rows_out, int32_t* dec_pos_out,
                           void* stream);
int opsg_copy_bytes(void* dst, const void* src, size_t nbytes, void* stream);
/* Second half of the deterministic split-K GEMM (opsg_gemm_bf16 with out_mode OPSG_OUT_F32 and k_splits > 1 writes split s
 * to slice s of a fp32 [k_splits][M][N] buffer): out = bf16(bias + sum of the slices, in split order (+ residual, bf16
 * [rows, ld_res], may be NULL, may alias out)).  K1 PatchEmbed (timm PatchEmbed, v4:410) uses it so that two runs return the
 * same bits; the LLM's out_proj / fc2 (down_proj) Linears of a stacked decode step (a few hundred rows, N = hidden size: too
 * few output tiles for the machine; HF:opt:181,232-236 / HF llama :262,184) take it with their residual. */
int opsg_splitk_reduce_bf16(const float* partials, int splits, int rows, int cols, const float* bias, const opsg_bf16* residual,
                            int ld_res, opsg_bf16* out, int ld_out, void* stream);
int opsg_transpose_i32(const int32_t* src, int rows, int cols, int32_t* dst, void* stream);

/* ---- f1 / f2: the integer passes either side of the head in the reference's inference loop --------------------------
 * opsg_pan_relabel: detectors/openseed_relation_v2.py:112-128 on the device: pan_out[p] = new_ids[s] for the LAST listed
 *   segment s with seg_ids[s] == pan_in[p], 0 if none (the reference: D2H, one np.where pass per segment, H2D).
 * opsg_pan_colorize: tools/infer.py:149-169: out_rgb uint8 [n_pixels, 3] = sum over the listed objects whose id equals the
 *   pixel of that object's colour (uint8 wrap-around), 0 elsewhere: the RGB-encoded panoptic map of the submission PNG. */
int opsg_pan_relabel(const int32_t* pan_in, long long n_pixels, const int32_t* seg_ids, const int32_t* new_ids, int n_segments,
                     int32_t* pan_out, void* stream);
int opsg_pan_colorize(const int32_t* pan, long long n_pixels, const int32_t* obj_ids, const uint8_t* colors, int n_objects,
                      uint8_t* out_rgb, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPSG_B200_H_ */
