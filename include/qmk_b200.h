/*
 * qmk_b200.h — C ABI of the B200 (sm_100a) decode engine for the Qwen3-TTS talker and code predictor.
 *
 * This is the drop-in boundary for ONE path of jayanth-kumar-morem/qwen-megakernel-tts: the fused
 * decode step behind torch.ops.qwen_megakernel_C.decode.  Plain pointers and sizes only (no torch
 * types).  All pointers are DEVICE pointers unless stated; `stream` is a cudaStream_t passed as void*.
 *
 * Upstream interfaces replaced (paths relative to the upstream repo root):
 *   launch_ldg_decode_direct      csrc/kernel.cu:1485-1513   (declared csrc/torch_bindings.cpp:45-53)
 *   struct LDGLayerWeights        csrc/kernel.cu:78-90       (mirrored csrc/torch_bindings.cpp:21-33)
 *   op decode(...)                csrc/torch_bindings.cpp:55-81, schema :130-141
 *   CodePredictorKernel.predict   qwen_megakernel/model_tts.py:729-773  -> qmk_cp_predict (one launch)
 *
 * Return convention: 0 = success, negative = error (text via qmk_last_error()).  The upstream C entry
 * returns void and checks nothing; the drop-in symbol keeps the void signature and records its status
 * for qmk_legacy_status().
 */
#ifndef QMK_B200_H
#define QMK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMK_ABI_VERSION 3

/* Model constants (upstream kernel.cu:21-28, model_tts.py:19-34); compile-time in the kernel. */
#define QMK_HIDDEN 1024
#define QMK_INTER 3072
#define QMK_Q_SIZE 2048
#define QMK_KV_SIZE 1024
#define QMK_HEAD_DIM 128
#define QMK_NUM_Q_HEADS 16
#define QMK_NUM_KV_HEADS 8
#define QMK_CP_GROUPS 15

/* Error codes */
#define QMK_OK 0
#define QMK_ERR_ARG (-1)      /* invalid argument (null pointer, position out of range, ...) */
#define QMK_ERR_CUDA (-2)     /* CUDA runtime error */
#define QMK_ERR_KERNEL (-3)   /* device-side watchdog fired (a wait inside the persistent kernel timed out) */
#define QMK_ERR_UNSUPPORTED (-4)

/* Per-layer weight pointers, bf16 row-major [out, in]; same field order as upstream LDGLayerWeights
 * (kernel.cu:78-90) so the caller's packed blob (model_tts.py:182-193) is consumed unchanged. */
typedef struct LDGLayerWeights {
  const void* input_layernorm_weight;     /* [1024]        */
  const void* q_proj_weight;              /* [2048, 1024]  */
  const void* k_proj_weight;              /* [1024, 1024]  */
  const void* v_proj_weight;              /* [1024, 1024]  */
  const void* q_norm_weight;              /* [128]         */
  const void* k_norm_weight;              /* [128]         */
  const void* o_proj_weight;              /* [1024, 2048]  */
  const void* post_attn_layernorm_weight; /* [1024]        */
  const void* gate_proj_weight;           /* [3072, 1024]  */
  const void* up_proj_weight;             /* [3072, 1024]  */
  const void* down_proj_weight;           /* [1024, 3072]  */
} LDGLayerWeights;

/* An engine owns one set of inter-SM exchange buffers, so all launches on it (every model created on it) execute in
 * submission order: a launch on a different stream than the engine's previous launch first waits (event) for everything
 * submitted to that stream.  Use one engine per GPU, or one per stream if launches are meant to overlap. */
typedef struct qmk_engine qmk_engine; /* per-device context: exchange buffers, watchdog word, epoch */
typedef struct qmk_model qmk_model;   /* a layer stack re-packed into per-SM weight streams */

int qmk_abi_version(void);
const char* qmk_last_error(void);
/* Hash of the sources this library was compiled from (csrc + include); the Python loader compares it with the tree. */
const char* qmk_source_hash(void);

/* ---- engine ------------------------------------------------------------------------------------ */
/* num_ctas = 0 -> the default kernel generation for `device`: the group kernel (8 kv-head groups x 16 = 128 persistent CTAs,
 * csrc/qmk_device2.cuh) when the device has >= 128 SMs, else the row-split kernel with one persistent CTA per SM (csrc/
 * qmk_device.cuh).  An explicit num_ctas other than 128, or QMK_ENGINE=1, selects the row-split kernel. */
int qmk_engine_create(int device, int num_ctas, qmk_engine** out);
void qmk_engine_destroy(qmk_engine* e);
int qmk_engine_num_ctas(const qmk_engine* e);
/* Synchronise `stream` and return the device-side status of all launches since the last call
 * (QMK_OK or QMK_ERR_KERNEL); detail[0..3] (optional, host) receives the watchdog record. */
int qmk_engine_sync_status(qmk_engine* e, void* stream, int32_t* detail);

/* Optional phase trace (debug/profiling): with stride > 0 every CTA records clock64() at the start of each
 * phase of the next launches into a [num_ctas][stride] device array; stride 0 disables it.
 * qmk_engine_trace_read synchronises `stream`, copies the array to host_out and returns the stride. */
int qmk_engine_trace_enable(qmk_engine* e, int stride);
int qmk_engine_trace_read(qmk_engine* e, void* stream, long long* host_out, int64_t max_elems);

/* Exchange diagnostics: per CTA [3][8] int32 = {poll delay in cycles after the CTA's own publish, number of
 * gathers in which a poll had to be repeated, cycles/16 spent waiting for weights} for the 8 exchange kinds
 * (qkv-in, attention-in, o-in, gate/up-in, down-in, head-in, argmax-in, token-in).  Synchronises `stream`;
 * returns the row length (24). */
int qmk_engine_poll_stats(qmk_engine* e, void* stream, int32_t* host_out, int64_t max_elems);

/* ---- model (weight re-packing; replaces upstream _pack_layer_weights as the packing layer) ------ */
/* `layers` is the caller's DEVICE blob of num_layers LDGLayerWeights structs.  The weights are copied
 * into the engine's own per-CTA stream layout; the originals are not referenced afterwards.
 * residual_fp32: 1 = fp32 residual stream across layers (upstream talker PyTorch path,
 * validate_kernel.py:123-188), 0 = bf16 residual (upstream CodePredictor, model_tts.py:567-619). */
int qmk_model_create(qmk_engine* e, const LDGLayerWeights* layers, int num_layers,
                     const void* final_norm_weight, int residual_fp32, void* stream, qmk_model** out);
/* Register an LM head [rows, 1024] bf16 (talker codec head: 3072 rows; code-predictor group heads:
 * 2048 rows).  Returns the head index (>= 0) or a negative error. */
int qmk_model_add_head(qmk_model* m, const void* lm_head_weight, int rows, void* stream);
/* Register the code-predictor embedding table of group g (bf16 [2048,1024]); used by qmk_cp_predict. */
int qmk_model_set_group_embedding(qmk_model* m, int group, const void* embedding_weight);
/* Multimodal RoPE (the talker of Qwen3-TTS is trained with mrope_section = [24, 20, 20]; upstream implements standard RoPE and
 * documents the gap, README.md:208): rotary frequency i takes the cos/sin table row of the position of ITS axis
 * (temporal / height / width).  section[3] must add up to 64; interleaved = 0: consecutive sections (Qwen2-VL
 * apply_multimodal_rotary_pos_emb), 1: axes interleaved t,h,w,t,h,w.. (Qwen3-VL apply_interleaved_mrope).  section = NULL
 * restores standard RoPE (the default).  The per-axis positions of a step are given to qmk_decode_step_mrope. */
int qmk_model_set_mrope(qmk_model* m, const int32_t* section, int interleaved);
void qmk_model_destroy(qmk_model* m);
int64_t qmk_model_packed_bytes(const qmk_model* m);

/* ---- one decode step (replaces launch_ldg_decode_direct) ----------------------------------------- */
/* Semantics of upstream decode (torch_bindings.cpp:55-81 / kernel.cu:1317-1432):
 *   input_token_id >= 0 : layer-0 input = embed_weight[input_token_id]        (kernel.cu:1364-1367)
 *   input_token_id <  0 : layer-0 input = hidden_buffer (bf16[1024], caller-written sentinel path)
 *   KV row `position` of every layer is written into k_cache/v_cache ([L][8][max_seq_len][128] bf16)
 *   normalized_out (f32[1024])  = post-final-RMSNorm hidden (bf16-rounded values)
 *   hidden_buffer  (bf16[1024]) = last layer output (clobbered)
 *   out_token (int32[1])        = argmax of head `head_index` (lowest index on ties); head_index < 0
 *                                 skips the LM head (code-predictor steps).
 * mode: 0 = fused persistent kernel (one cooperative launch), 1 = staged (one launch per phase; a
 * debugging/bisect mode of the row-split kernel that runs the same device code without inter-CTA waits; the group
 * kernel ignores it and runs fused).
 * Asynchronous on `stream`. */
int qmk_decode_step(qmk_model* m, int head_index, int input_token_id, const void* embed_weight,
                    const void* cos_table, const void* sin_table, void* k_cache, void* v_cache,
                    void* hidden_buffer, float* normalized_out, int32_t* out_token, int position,
                    int max_seq_len, float attn_scale, int mode, void* stream);

/* Same step (fused mode) with separate RoPE positions: `position` is the KV row that is written / the last row attention reads,
 * rope_pos[3] (HOST array) the table rows of the three M-RoPE axes (see qmk_model_set_mrope). */
int qmk_decode_step_mrope(qmk_model* m, int head_index, int input_token_id, const void* embed_weight,
                          const void* cos_table, const void* sin_table, void* k_cache, void* v_cache,
                          void* hidden_buffer, float* normalized_out, int32_t* out_token, int position,
                          const int32_t* rope_pos, int max_seq_len, float attn_scale, void* stream);

/* Same step with the frame loop's embedding sum fused in front of it (upstream tts_engine.py:319-335):
 *   input = talker_embed[codes[0]] + sum_{g<15} group_embedding_tables[g][codes[g+1]] + extra_embed   (bf16 adds,
 *   upstream order), codes = the int64[16] device tensor predict() returned, extra_embed = the trailing-text or
 *   tts_pad embedding (bf16[1024], device).  group_embedding_tables is a HOST array of 15 device pointers; talker_vocab /
 *   group_vocab are the row counts of the tables (codes are clamped to them on the device: an aborted predict writes
 *   negative sentinels).
 * Replaces 32 torch launches per frame; hidden_buffer is only written (last-layer output). */
int qmk_decode_step_codes(qmk_model* m, int head_index, const int64_t* codes, const void* talker_embed_weight,
                          int talker_vocab, const void* const* group_embedding_tables, int group_vocab,
                          const void* extra_embed_bf16, const void* cos_table, const void* sin_table, void* k_cache,
                          void* v_cache, void* hidden_buffer, float* normalized_out, int32_t* out_token, int position,
                          int max_seq_len, float attn_scale, void* stream);

/* ---- one code-predictor frame in ONE launch (replaces the 16-step loop of
 *      CodePredictorKernel.predict, model_tts.py:729-773) ------------------------------------------ */
/* talker_hidden: f32[1024]; talker_embed_weight: bf16[3072,1024]; out_codes: int64[16] =
 * [first_token, g0..g14].  do_sample=0 -> argmax on bf16 logits; else temperature / top-k (ties kept)
 * / softmax / inverse-CDF draw from a counter-based generator keyed by (seed, frame_counter).
 * k_cache/v_cache: [5][8][max_seq_len][128] bf16 scratch owned by the caller (max_seq_len >= 16).
 * logits_out (optional, f32[15*2048]) and hidden_out (optional, f32[15*1024]) receive the per-group
 * bf16 logits / post-norm hidden for teacher-forced parity checks; forced_tokens (optional, int32[15],
 * device) overrides the fed-back tokens. */
int qmk_cp_predict(qmk_model* m, const float* talker_hidden, int first_codebook_token,
                   const void* talker_embed_weight, const void* cos_table, const void* sin_table,
                   void* k_cache, void* v_cache, int max_seq_len, int do_sample, float temperature,
                   int top_k, uint64_t seed, uint64_t frame_counter, const int32_t* forced_tokens,
                   int64_t* out_codes, float* logits_out, float* hidden_out, void* stream);

/* Same frame with the talker's token taken from DEVICE memory (the int32 out_token of the preceding talker step on the
 * same stream), so a frame loop can queue predict(f+1) behind step(f) without a host round trip; the token is clamped
 * to [0, talker_vocab).  The host reads tokens / codes asynchronously (EOS is then detected one frame late). */
int qmk_cp_predict_dev(qmk_model* m, const float* talker_hidden, const int32_t* first_token_dev, int talker_vocab,
                       const void* talker_embed_weight, const void* cos_table, const void* sin_table, void* k_cache,
                       void* v_cache, int max_seq_len, int do_sample, float temperature, int top_k, uint64_t seed,
                       uint64_t frame_counter, int64_t* out_codes, void* stream);

/* ---- N codec frames without the host ------------------------------------------------
 * Upstream analogue: launch_ldg_generate_nosync + ldg_update_step (kernel.cu:1555-1613, 1437-1448; op generate_nosync,
 * torch_bindings.cpp:93-127) -- N steps queued with no host sync.  Here one persistent launch runs the whole upstream frame
 * loop (tts_engine.py:301-335) n_frames times:
 *     if token == eos_token: stop                      (device-side EOS flag, gen_state[1])
 *     codes[f] = code_predictor.predict(hidden, token)  16 five-layer steps, 15 heads, greedy or top-k sampling
 *     e = talker_embed[codes[0]] + sum_g group_table[g][codes[g+1]] + (trailing_text[offset + f] or pad_embed)
 *     token, hidden = talker.step_with_embed(e)         at KV row position + f
 * talker_token / talker_hidden are IN/OUT device buffers: the (int32 token, f32[1024] post-norm hidden) of the talker step
 * that precedes the first frame, overwritten with the last frame's.  codes_out (int64[n_frames][16]), tokens_out
 * (int32[n_frames], optional: the talker token produced by each frame) and gen_state (int32[4]: frames completed, EOS seen,
 * last token, reserved) may be device memory or mapped pinned host memory (a host thread can then consume frames while the
 * kernel is still generating: gen_state[0] is written after a frame's codes, behind a system-scope fence).
 * n_frames = 1 is "one launch per frame".  Long runs are split into chained launches of a few hundred frames (16-bit epochs);
 * once EOS was seen the remaining launches exit immediately.  reset_state = 1 zeroes gen_state first; with 0 a further call
 * continues the same utterance (trailing-text index = trailing_offset + gen_state[0] + f). */
typedef struct qmk_generate_args {
  qmk_model* talker;
  int32_t talker_head, talker_vocab;
  const void* talker_embed_weight;   /* bf16 [talker_vocab, 1024] */
  const void* talker_cos;
  const void* talker_sin;
  void* talker_k_cache;
  void* talker_v_cache;
  int32_t talker_max_seq, position;  /* KV row of the first frame's talker step */
  const int32_t* rope_pos;           /* optional HOST int32[3]: M-RoPE positions of that step (advance by 1 per frame) */
  void* hidden_buffer;               /* bf16[1024] out: last-layer output of the last talker step */
  float* talker_hidden;              /* f32[1024] in/out */
  int32_t* talker_token;             /* int32[1] in/out */
  qmk_model* cp;                     /* code predictor: 15 heads + 14 group embeddings registered */
  const void* cp_cos;
  const void* cp_sin;
  void* cp_k_cache;
  void* cp_v_cache;
  int32_t cp_max_seq, cp_vocab;
  const void* const* group_embedding_tables;   /* HOST array of 15 device pointers, bf16 [cp_vocab, 1024] each */
  int32_t n_frames, eos_token;       /* eos_token < 0: never stop early */
  const void* trailing_text;         /* bf16 [n_trailing, 1024] or NULL */
  int32_t n_trailing, trailing_offset;
  const void* pad_embed;             /* bf16[1024] */
  int32_t do_sample, top_k;
  float temperature;
  int32_t reset_state;
  uint64_t seed, frame_counter;
  int64_t* codes_out;
  int32_t* tokens_out;
  int32_t* gen_state;
} qmk_generate_args;
int qmk_generate_nosync(const qmk_generate_args* args, void* stream);
int qmk_generate_args_size(void);   /* sizeof(qmk_generate_args) of this build (binding self-check) */

/* ---- batched multi-stream decode (no upstream counterpart: upstream is strictly B = 1) ----
 * B = 16 .. 64 concurrent utterances, each numerically the B = 1 step (own position, own KV cache).  The projections
 * run on tcgen05 / TMEM tensor-core tiles (csrc/qmk_bgemm.cuh), weights are read in place in the upstream [out, in]
 * layout.  `layers_host` is a HOST array of num_layers LDGLayerWeights structs holding device pointers. */
typedef struct qmk_batched qmk_batched;
int qmk_batched_create(int device, const LDGLayerWeights* layers_host, int num_layers, const void* final_norm_weight,
                       const void* lm_head_weight, int lm_head_rows, const void* embed_weight, const void* cos_table,
                       const void* sin_table, int residual_fp32, int batch, int max_seq_len, qmk_batched** out);
void qmk_batched_destroy(qmk_batched* h);
/* One decode step for all B streams (asynchronous on `stream`, no host sync):
 *   token_ids  int32[B] device or NULL; an entry < 0 (or NULL) takes row b of `embeds` (bf16[B][1024]) as the input
 *   positions  int32[B] device, per-stream KV row to write; advanced by one by the call
 *   k_cache / v_cache  bf16 [B][L][8][max_seq_len][128]
 *   hidden_out f32[B][1024] post-final-norm hidden (optional), tokens_out int32[B] argmax (lowest index on ties) */
int qmk_batched_step(qmk_batched* h, const int32_t* token_ids, const void* embeds, int32_t* positions, void* k_cache,
                     void* v_cache, float* hidden_out, int32_t* tokens_out, void* stream);
const char* qmk_batched_last_error(void);

/* ---- batched code-predictor frame / batched frame loop (BASELINE.json configs[3], [4] in codec frames/s) -------------------
 * A code-predictor stack is a qmk_batched with num_layers = 5, residual_fp32 = 0, its group heads registered with
 * qmk_batched_add_head and its steps issued with qmk_batched_step_ex (input table per step, fp32 input for step 0, head
 * index or none, greedy or temperature / top-k / multinomial selection with the B = 1 engine's sampler, int64 code outputs). */
typedef struct qmk_batched_step_args {
  const int32_t* token_ids;     /* int32[B] or NULL; entry < 0 takes the stream's row of embeds_bf16 */
  const void* token_table;      /* bf16 [table_rows, 1024] the ids index; NULL = the create-time embedding table */
  int32_t table_rows;
  const void* embeds_bf16;      /* bf16[B][1024] or NULL */
  const float* embeds_f32;      /* f32[B][1024] or NULL: rounded to bf16 on load (takes precedence) */
  int32_t* positions;           /* int32[B], advanced by one */
  void* k_cache;
  void* v_cache;
  float* hidden_out;            /* f32[B][1024] or NULL */
  int32_t head;                 /* -1: no LM head; 0: the create-time head; n >= 1: n-th qmk_batched_add_head */
  int32_t do_sample, top_k, group;
  float temperature;
  uint64_t seed, counter;       /* stream b draws like a B = 1 engine seeded seed + b * 0x632BE59BD9B4E019 at frame `counter` */
  int32_t* tokens_out;          /* int32[B] (head >= 0) */
  int64_t* codes_out;           /* optional: codes_out[b * codes_stride + codes_col] = token */
  int32_t codes_stride, codes_col;
  const uint64_t* counter_ptr;  /* optional (ABI 3): device word added to `counter` when the kernel runs -- a CUDA graph captured
                                   from these calls draws fresh numbers on every replay (qmk_batched_counter_add advances it) */
  int32_t depth_hint;           /* optional (ABI 3): an upper bound of the streams' positions known to the host (positions live on
                                   the device); >= 256 at B = 16 / 32 (>= 1024 at B >= 48) splits every context over 4 / 2 (2) CTAs.  0 = unknown / shallow */
} qmk_batched_step_args;
int qmk_batched_add_head(qmk_batched* h, const void* lm_head_weight, int rows);
int qmk_batched_step_ex(qmk_batched* h, const qmk_batched_step_args* args, void* stream);
/* out[b] = talker_embed[codes[b][0]] + sum_g group_tables[g][codes[b][g+1]] + extra[b * extra_stride]  (bf16 adds, upstream order:
 * tts_engine.py:319-333 for B streams).  codes: int64[B][16]; group_tables: HOST array of 15 device pointers. */
int qmk_batched_embed_sum(int batch, const int64_t* codes, const void* talker_embed, int talker_rows,
                          const void* const* group_tables, int group_rows, const void* extra_bf16, int extra_stride,
                          void* out_bf16, void* stream);
/* *counter += inc on `stream` (one thread; the frame counter of a captured frame loop).
 * CUDA graphs: qmk_batched_step / _step_ex / _embed_sum / _counter_add only enqueue kernels on `stream` (no allocation, no
 * synchronisation, no host-side state that a replay would miss: positions, tokens and the counter live in device memory), so a
 * step or a whole frame may be captured with cudaStreamBeginCapture and replayed; the programmatic-launch edges are kept. */
int qmk_batched_counter_add(uint64_t* counter, uint64_t inc, void* stream);
/* csrc/qmk_bstep.cuh holds a PERSISTENT form of the step: one cooperative launch runs all layers, the LM head and the argmax
 * (grid barriers between phases, weights re-packed k-block-major and prefetched through a shared-memory ring by a producer
 * warp, tcgen05 / TMEM projections).  It serves the one-pass prefill below.  For decode steps it is opt-in
 * (QMK_BATCHED_PERSISTENT=1 at create time): measured on B200 it is slower than the chain of per-projection launches
 * (1.4 / 2.1 ms vs 1.0 / 1.1 ms at B = 16 / 64), see DESIGN.md.  Return value: bit 0 = kernel available, bit 1 = used for decode. */
int qmk_batched_is_persistent(const qmk_batched* h);
/* Synchronise `stream`; QMK_ERR_KERNEL if a wait inside the persistent kernel timed out since the last call (tokens_out then
 * holds -1001). */
int qmk_batched_sync_status(qmk_batched* h, void* stream);
/* Debug (QMK_BATCHED_TRACE=1): clock64 of CTA 0 entering / leaving every grid barrier of the latest persistent step. */
int qmk_batched_trace_read(qmk_batched* h, void* stream, long long* host_out, int max_elems);
/* Debug: timeline of the launch chain (current device).  enable != 0 arms %globaltimer stamps in the first and last CTA of every
 * chain kernel; with host_out the records {tag = kernel id * 16 + event * 2 + (last CTA), ns} since the previous call are copied
 * (max_records pairs of 64-bit words) and their number is returned.  enable == 0 disarms. */
int qmk_batched_chain_trace(int enable, void* stream, unsigned long long* host_out, int max_records);
/* Prefill of ONE utterance as a single batched pass (upstream runs the 8 prefill embeddings
 * through 8 sequential decode steps, tts_engine.py:281-282): lane i of `embeds` (bf16[n][1024], n <= batch) is position
 * position0 + i; all lanes share the caller's B = 1 cache [L][8][max_seq_len][128] and attend causally.  Writes the same KV
 * rows as n sequential steps; hidden_out_last (f32[1024]) / token_out_last (int32[1]) receive the last position's
 * post-norm hidden state and argmax token.  Default: the launch chain with lane = position (9 kernels per layer: the QKV
 * epilogue is split into "all lanes' K / V rows" and "causal attention"); QMK_PREFILL_PERSISTENT=1 at create time selects the
 * persistent kernel's prefill mode (needs >= 144 SMs; slower: 1.44 vs 0.78 ms for 8 positions on B200). */
int qmk_batched_prefill(qmk_batched* h, const void* embeds, int n, int position0, void* k_cache, void* v_cache,
                        float* hidden_out_last, int32_t* token_out_last, void* stream);

/* ---- text side of the prefill ------------------------------------------------------------
 * Replaces upstream TextProjection.embed_text_ids (model_tts.py:348-374; called once per utterance on the content tokens,
 * tts_engine.py:262-263): out = fc2( silu( fc1( text_embedding[ids] ) ) ) with upstream's rounding points (bf16 after every
 * operator, fp32 accumulation, bias added in fp32).  text_embedding bf16[vocab_rows][2048], fc1_weight bf16[2048][2048],
 * fc1_bias bf16[2048], fc2_weight bf16[1024][2048], fc2_bias bf16[1024]: device pointers in the upstream [out, in] layout,
 * read in place (the caller keeps them alive), 16-byte aligned.  Both projections run on tcgen05 (csrc/qmk_bgemm.cuh): one chain
 * of five launches per 512 tokens, blocks of 64 tokens along gridDim.z (csrc/qmk_text.cuh). */
typedef struct qmk_text_proj qmk_text_proj;
int qmk_text_proj_create(int device, const void* text_embedding, int vocab_rows, const void* fc1_weight, const void* fc1_bias,
                         const void* fc2_weight, const void* fc2_bias, qmk_text_proj** out);
/* ids: int64[n_ids] in DEVICE memory (values outside [0, vocab_rows) are clamped; upstream's F.embedding would trap);
 * out_bf16: bf16[n_ids][1024] in device memory.  n_ids == 0 is a no-op.  Asynchronous on `stream`, only enqueues kernels
 * (CUDA-graph capturable).  Calls on one handle share its staging buffers: a call on another stream than the previous call
 * first waits for that stream (event); calls recorded into a graph are ordered by whoever replays it. */
int qmk_text_proj_embed(qmk_text_proj* h, const int64_t* ids, int n_ids, void* out_bf16, void* stream);
void qmk_text_proj_destroy(qmk_text_proj* h);

/* ---- upstream-compatible entry point (same symbol, same argument list) ---------------------------- */
/* The scratch arguments (g_activations .. g_mlp_intermediate, block_max_*) are accepted and ignored:
 * the engine keeps its own exchange buffers.  The first call for a given `layer_weights` blob re-packs
 * the weights (one-time, synchronous); qmk_legacy_configure() sets how that blob is interpreted. */
void launch_ldg_decode_direct(int input_token_id, int* output_token_id, const void* embed_weight,
                              const LDGLayerWeights* layer_weights, const void* final_norm_weight,
                              const void* lm_head_weight, const void* cos_table, const void* sin_table,
                              void* k_cache, void* v_cache, void* hidden_buffer, void* g_activations,
                              void* g_residual, void* g_q, void* g_k, void* g_v, void* g_attn_out,
                              void* g_mlp_intermediate, void* g_normalized, void* block_max_vals,
                              void* block_max_idxs, int num_layers, int position, int max_seq_len,
                              float attn_scale, void* stream);
/* How a blob is interpreted WITHOUT any extra call (so that upstream's classes run unchanged):
 *   residual stream: num_layers == 5 -> bf16 (upstream PyTorch CodePredictor, model_tts.py:567-619), otherwise fp32
 *                    (upstream PyTorchTalkerReference, validate_kernel.py:123-188);
 *   LM head:         3072 rows (upstream compile-time LDG_VOCAB_SIZE, build_tts.py:47), unless the table is all zero -- the
 *                    dummy the upstream CodePredictorKernel passes (model_tts.py:657-659) -- then the head is skipped and
 *                    output_token is 0 (the argmax of all-zero logits).  Checked once per table (one reduction + sync).
 * qmk_legacy_configure overrides either (-1 = keep the rule above; lm_head_rows 0 = never run a head).
 * Caching: the re-packed copy is keyed by (device, blob address, num_layers) and the final-norm address.  Upstream reads the
 * weights live on every call; if they change in place, or the blob's address is re-used, call qmk_legacy_invalidate (the
 * Python op does it automatically when it sees a different blob tensor / version at a cached address). */
int qmk_legacy_configure(const LDGLayerWeights* layer_weights, int residual_fp32, int lm_head_rows);
int qmk_legacy_status(void);        /* host-side status of the last launch_ldg_decode_direct call */
int qmk_legacy_sync_status(void* stream);   /* synchronise + return and clear the DEVICE-side status (watchdog) of the current device's legacy engine */
void qmk_legacy_invalidate(const LDGLayerWeights* layer_weights);   /* drop the cached model of a blob (NULL: all) */
int qmk_legacy_check_blob(const LDGLayerWeights* layer_weights, int num_layers, const uint64_t* host_blob);   /* 1 = cached copy was stale and has been dropped */
void qmk_legacy_release(void);      /* drop cached engines / re-packed models */

#ifdef __cplusplus
}
#endif
#endif /* QMK_B200_H */
