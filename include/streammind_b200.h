/* streammind_b200 -- C ABI of the B200-native StreamMind hot path.
 *
 * The reference (xinding-sys/StreamMind) has NO plugin / operator / FFI interface: its seams are
 * Python nn.Module method boundaries.  Each entry point below therefore cites
 * the reference *Python* interface it replaces; the Python host in streammind_b200/ binds these with
 * ctypes and re-exposes the reference's own method names and argument meaning.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  Unless stated otherwise pointers are
 *     DEVICE pointers on the handle's device (tensor.data_ptr()); `stream` is a cudaStream_t passed
 *     as void* (NULL = legacy default stream).  Calls are asynchronous on `stream`.
 *   - every function returns 0 on success, non-zero on error; sm_last_error() gives the message
 *     (the Python wrapper raises RuntimeError, matching the reference's exception convention).
 *   - a handle is bound to one device and one video stream's state (Mamba conv/ssm state, KV cache,
 *     frame count); it is not thread-safe but may be called from any host thread (the reference's
 *     serving worker calls generate() from a non-main thread: streammind/serve/model_worker.py:271).
 *   - model dtype: one of fp16 / bf16 for every sub-model of a handle (the reference loads fp16:
 *     streammind/model/builder.py:54,201; bf16 in eval: eval/inference_video_ego4d_stream_parallel_new.py:160).
 */
#ifndef STREAMMIND_B200_H_
#define STREAMMIND_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sm_handle sm_handle;

enum { SM_DTYPE_F16 = 0, SM_DTYPE_BF16 = 1, SM_DTYPE_F32 = 2 };

typedef struct sm_config {
    int dtype;              /* SM_DTYPE_F16 | SM_DTYPE_BF16 */
    int max_frames;         /* frames per sm_vit_encode / sm_frame_step call (chunk size), >= 1 */
    /* CLIP vision tower (hf CLIPVisionConfig; clip_encoder.py).  vit_layers = 0 disables the tower. */
    int vit_image, vit_patch, vit_hidden, vit_layers /* layers actually executed = index of
        hidden_states[select_layer], 23 for CLIP-L with select_layer = -2 */, vit_heads, vit_ffn;
    float vit_eps;
    /* projector: PreNet -> Mamba-1 block -> PostNet (multimodal_projector/builder.py:390-414). 0 = off */
    int proj_d_model, proj_d_state, proj_d_conv, proj_expand;
    float proj_eps;
    /* gate: ClsNet = Mistral decoder at L = 1 (multimodal_projector/builder.py:370-385). layers 0 = off */
    int gate_layers, gate_heads, gate_kv_heads, gate_head_dim, gate_ffn;
    float gate_eps;
    /* LLM: MistralForCausalLM (language_model/videollama2_mistral.py:146). layers 0 = off */
    int llm_hidden, llm_layers, llm_heads, llm_kv_heads, llm_head_dim, llm_ffn, llm_vocab, llm_max_ctx;
    float llm_eps, llm_rope_theta;
    int use_graphs;         /* capture the per-frame step into CUDA graphs */
    int n_streams;          /* video streams held by this handle (0 / 1 = one).  Each stream has its own Mamba conv / ssm state and
                               KV cache; all share the weights, and sm_llm_decode_multi decodes several of them per pass over the
                               weights (the reference keeps this state on the model object: videollama2_mistral.py:159-162) */
} sm_config;

/* Lifetime.  Replaces model construction in load_pretrained_model (streammind/model/builder.py:30). */
int sm_create(sm_handle** out, int device, const sm_config* cfg);
void sm_destroy(sm_handle* h);
const char* sm_last_error(const sm_handle* h);   /* h may be NULL: error of the last failed sm_create */

/* Weight upload by the reference model's own state_dict key
 * (e.g. "model.vision_tower.vision_tower.vision_model.encoder.layers.0.self_attn.q_proj.weight",
 *  "model.mm_projector.mamba_model.ssms.0.mixer.in_proj.weight", "model.layers.3.mlp.up_proj.weight").
 * Data is copied into the kernel layout (q/k/v and gate/up are packed, the patch conv is re-pitched);
 * the caller may free `data` afterwards.  Keys that do not influence the path (q_proj/k_proj of the
 * gate, ViT layers beyond vit_layers, post_layernorm, position_ids, gate embed_tokens) return 0 and
 * are skipped.  Unknown keys are an error.  dtype must equal cfg.dtype.
 * Replaces from_pretrained / load_state_dict (streammind/model/builder.py:141-203). */
int sm_load_weight(sm_handle* h, const char* name, const void* data, int data_on_host, int dtype, int ndim,
                   const int64_t* shape);
/* Verifies that every weight the enabled sub-models need was loaded; returns non-zero and lists the
 * missing keys in sm_last_error otherwise. */
int sm_finalize_weights(sm_handle* h);

/* Per-stream state of the SELECTED stream: zero the Mamba conv/ssm state and the KV length, ordered on `stream` (frames
 * still in flight in the pipelined path are joined first).  The reference keeps this state on the model object and
 * never resets it (videollama2_mistral.py:159-165). */
int sm_stream_reset(sm_handle* h, void* stream);

/* Multi-stream handles (sm_config.n_streams > 1): the stream every single-stream entry point below acts on
 * (sm_frame_step / submit, sm_projector_step, sm_llm_prefill, sm_llm_decode, sm_kv_len, sm_kv_set_len, sm_stream_reset).
 * Default 0.  An open batch of pipelined tickets is closed first, so a batch never mixes streams. */
int sm_stream_select(sm_handle* h, int stream_id);
int sm_num_streams(const sm_handle* h);

/* Frame preprocessing: mm_utils.process_video / process_image with aspect_ratio 'pad'
 * (streammind/mm_utils.py:446-464) = expand2square(frame, background) (mm_utils.py:257-268) ->
 * CLIPImageProcessor.preprocess (transformers 4.44.2: bicubic resize through PIL.Image.resize of Pillow 9.4.0 ->
 * centre crop -> rescale 1/255 -> normalize -> channels first) -> cast to the model dtype (.half() in
 * video_score_stream_demo.py:285-287).  frames [n, H, W, 3] uint8 RGB (host memory unless frames_on_device) ->
 * pixels_out [n, 3, vit_image, vit_image] model dtype on the device: the input of sm_vit_encode / sm_frame_submit.
 * mean / std: the processor's image_mean / image_std as float32; background: tuple(int(x*255) for x in image_mean).
 * Bit-exact with the reference's PIL + numpy arithmetic (tests/test_preprocess_gpu.py).  Calls on one handle share
 * staging buffers: issue them on one stream (or synchronise between streams). */
int sm_preprocess_frames(sm_handle* h, const unsigned char* frames, int n, int H, int W, int frames_on_device,
                         const float* mean /*[3]*/, const float* std /*[3]*/, const int* background /*[3]*/,
                         void* pixels_out, void* stream);

/* Host-only (no GPU needed): the fixed-point bicubic tap table sm_preprocess_frames uses for an axis of in_size ->
 * out_size pixels = Pillow's precompute_coeffs + normalize_coeffs_8bpc (Resample.c).  ksize_out: taps per output pixel;
 * bounds_out [out_size, 2] (first source pixel, tap count) and kk_out [out_size, ksize] may be NULL; returns 2 when
 * kk_capacity (ints) is too small.  Exported so the table can be checked against the oracle on a CPU-only box. */
int sm_resample_table(int in_size, int out_size, int* ksize_out, int* bounds_out, int* kk_out, long long kk_capacity);

/* CLIPVisionTower.forward + feature_select (multimodal_encoder/clip_encoder.py:41-53,31-39).
 * pixels [B,3,H,W] model dtype, contiguous NCHW -> feats_out [B, P, C] (may be NULL) and
 * pooled_out [B, C] = mean over patches (multimodal_projector/builder.py:405; may be NULL). */
int sm_vit_encode(sm_handle* h, const void* pixels, int B, void* feats_out, void* pooled_out, void* stream);

/* Mean over the patch axis of externally supplied features (builder.py:405): feats [n, P, C] -> [n, C]. */
int sm_pool_features(sm_handle* h, const void* feats, int n, void* pooled_out, void* stream);

/* One projector step per frame: PreNet -> LayerNorm -> Mamba.step -> +residual -> LayerNorm -> PostNet
 * (Video_Mamba_seq.forward core, multimodal_projector/builder.py:403-414, evaluated incrementally;
 * equals row t of the reference's full-sequence call).  pooled [n, C] -> tok_out [n, d_model];
 * advances the stream's Mamba state by n frames. */
int sm_projector_step(sm_handle* h, const void* pooled, int n, void* tok_out, void* stream);

/* Gate: cls_demo branch (multimodal_projector/builder.py:547-562): tok [d_model] -> logits_out [2] fp32
 * (index 0 = silence, 1 = respond; videollama2_arch.py:944-948). */
int sm_gate_score(sm_handle* h, const void* tok, float* logits_out, void* stream);

/* Fused per-frame step = sm_vit_encode + sm_projector_step + sm_gate_score for each of the B frames
 * (encode_images_or_videos_score_cls_inference_allframe_demo, videollama2_arch.py:173-203, in
 * incremental form).  `pixels` may be a pinned HOST pointer (pixels_on_host = 1; the H2D copy is then
 * part of the call) or a device pointer.  Outputs (each may be NULL): feats_out [B,P,C],
 * toks_out [B, d_model], logits_out [B, 2] fp32 (device), and logits_host [B, 2] fp32 (pinned host,
 * written by an async D2H copy on `stream`; the caller synchronises the stream before reading). */
int sm_frame_step(sm_handle* h, const void* pixels, int pixels_on_host, int B, void* feats_out, void* toks_out,
                  float* logits_out, float* logits_host, void* stream);

/* sm_frame_step for n DIFFERENT streams at once (multi-stream batching): frame i belongs to stream slot
 * first_stream + i.  One tower batch, then every pass over the projector weights serves <= 4 streams (each with its own
 * Mamba conv / ssm state) and every pass over the gate weights serves all n frames.  n <= max_frames.  Outputs as in
 * sm_frame_step (row i = stream first_stream + i); results equal n separate sm_frame_step calls on the selected streams. */
int sm_frame_step_multi(sm_handle* h, const void* pixels, int pixels_on_host, int n, int first_stream, void* toks_out,
                        float* logits_out, float* logits_host, void* stream);

/* Pipelined form of sm_frame_step for throughput streams.  The vision tower of the call runs on an internal
 * high-priority stream; projector + gate run on a second internal low-priority stream, so the gate of frame t
 * overlaps the tower of frame t+1 (the tower is tensor/latency-bound and leaves HBM mostly idle, the gate is
 * HBM-bound).  `pixels` must be ready on
 * `stream` at call time.  Outputs are NOT ordered on `stream`: *ticket identifies the call, and
 * sm_frame_wait(ticket, stream, block_host) makes `stream` wait for (stream != NULL or block_host == 0) and/or
 * blocks the host until (block_host != 0) that call's outputs -- including the pinned-host logits -- are
 * complete.  At most 16 tickets are in flight (the 17th submit blocks the host on the oldest).  On a streaming handle
 * (max_frames == 1) the towers of up to 8 consecutive tickets run as ONE chunk (pixels are staged at submit time) and
 * projector + gate share each pass over their weights between 4 frames; the batch is closed by its 8th ticket or by
 * sm_frame_wait / a serial entry point touching one of its tickets.  Otherwise every ticket runs its own tower on one
 * of 8 streams ("lanes").  Results are identical to sm_frame_step (same kernels and arithmetic; stream state advances in
 * ticket order). */
int sm_frame_submit(sm_handle* h, const void* pixels, int pixels_on_host, int B, void* feats_out, void* toks_out,
                    float* logits_out, float* logits_host, void* stream, long long* ticket);
int sm_frame_wait(sm_handle* h, long long ticket, void* stream, int block_host);

/* Cognition sampling of a frame-token segment before the LLM (videollama2_arch.py:595-611, used by the eval
 * forward at :676-681).  toks [n, d] model dtype (device) -> out [k, d], the kept rows in their original order, and
 * (optional) indices_out [k] int32 (device).  k = sm_cognition_count(n, percentage, mode).
 *   mode 0 "log" = exponential_sampling as the reference ships it: torch.linspace(0, n - 1, k).int() rows (float32
 *           linspace exactly as ATen's CPU kernel evaluates it on this host), k = int(percentage n) or 1;
 *   mode 1 "similarity" = similarity_sampling: the k = max(int(percentage n), 1) rows most cosine-similar to the LAST
 *           row (torch cosine_similarity arithmetic in the model dtype; ties -> lower index). */
int sm_cognition_sample(sm_handle* h, const void* toks, int n, int d, int mode, double percentage, void* out, int32_t* indices_out,
                        void* stream);
int sm_cognition_count(int n, double percentage, int mode);
/* Host-only: torch.linspace(0, n - 1, steps).int() -> out [steps] as ATen's CPU kernel evaluates it (the reference calls
 * torch.linspace without a device, i.e. on the host): fused = 1 one FMA per element (ATen's AVX2 / AVX512 builds), 0 product
 * and sum rounded separately (its DEFAULT build), -1 = what this host's torch would do (the rule sm_cognition_sample uses).
 * Checked against torch under each ATEN_CPU_CAPABILITY in tests/test_cognition_cpu.py. */
int sm_linspace_indices(int n, int steps, int fused, int* out);

/* embed_tokens (videollama2_arch.py:967,977): ids [n] int32 device -> out [n, hidden]. */
int sm_embed_tokens(sm_handle* h, const int32_t* ids, int n, void* out, void* stream);

/* Append P positions to the KV cache (MistralForCausalLM.forward on inputs_embeds with
 * past_key_values, hf modeling_mistral.py:402-472, as driven by generate(): videollama2_mistral.py:426-431).
 * embeds [P, hidden].  If last_logits != NULL it receives the fp32 logits [vocab] of the last position. */
int sm_llm_prefill(sm_handle* h, const void* embeds, int P, float* last_logits, void* stream);

/* Greedy decode loop on the device (GenerationMixin.generate(do_sample=False) + the id rule of
 * KeywordsStoppingCriteria, mm_utils.py:631-636): must follow sm_llm_prefill.  Produces up to max_new
 * tokens, stops after emitting any of stop_ids (host array).  ids_out_host [max_new] and n_out_host
 * are HOST pointers filled on return (the call synchronises `stream`).  The KV cache afterwards holds
 * every produced token except the last one (which was never fed back), as in HF.  A stream whose cache cannot hold
 * max_new more tokens produces as many as fit (n_out_host < max_new) instead of failing.
 * One token = ONE launch of the persistent weight-streaming kernel (csrc/decode_stream.cuh). */
int sm_llm_decode(sm_handle* h, int max_new, const int32_t* stop_ids, int n_stop, int32_t* ids_out_host,
                  int32_t* n_out_host, void* stream);

/* The same loop for n <= 4 streams of the handle at once (multi-stream batching): every pass over the
 * 14.2 GB of LLM weights serves one new token of each listed stream (each with its own KV cache, position and stop
 * state; a finished stream idles).  Every stream must have been prefilled (sm_stream_select + sm_llm_prefill).
 * max_new [n]; ids_out_host [n][out_stride]; n_out_host [n].  Token ids are identical to n separate sm_llm_decode calls. */
int sm_llm_decode_multi(sm_handle* h, int n, const int* stream_ids, const int* max_new, const int32_t* stop_ids, int n_stop,
                        int32_t* ids_out_host, int out_stride, int32_t* n_out_host, void* stream);

/* Device time of the decode steps (CUDA events around the token launches of every sm_llm_decode* call, on the call's
 * stream) since the last reset: milliseconds, kernel launches (steps), tokens produced by those steps (sum over
 * streams) and the sum over those tokens of the KV length each one attended to.  bench.py derives the achieved
 * HBM GB/s of the decode kernel from it.  Blocks until the pending calls' events have completed. */
int sm_decode_stats(sm_handle* h, double* ms, long long* steps, long long* tokens, long long* ctx_sum, int reset);

/* Debug / measurement: while device_buf != NULL, CTA 0 of the decode kernel adds the nanoseconds (globaltimer) it spends in
 * each phase to device_buf[0..5]: 0 vector staging (prologue), 1 weight ring, 2 epilogue, 3 grid barriers, 4 attention,
 * 5 token selection (tools/decode_probe.py). */
int sm_debug_decode_phases(sm_handle* h, long long* device_buf);

/* Debug / test: fp32 logits [vocab] of the LAST decode step of lane `lane` (0 for sm_llm_decode) -> logits_out (device). */
int sm_debug_decode_logits(sm_handle* h, int lane, float* logits_out, void* stream);

/* KV length bookkeeping (prefix reuse across fires; the reference re-prefills from scratch:
 * videollama2_mistral.py:413 past_key_values=None). */
int sm_kv_len(const sm_handle* h);
int sm_kv_set_len(sm_handle* h, int len);   /* truncate to a common prefix; len <= current length */

/* Generic building blocks exported for unit tests (tests/test_kernels_gpu.py) */
int sm_test_gemm(sm_handle* h, const void* x /*[M,K]*/, const void* w /*[N,K]*/, const void* bias /*[N]|NULL*/,
                 void* out /*[M,N]*/, int M, int N, int K, int epi, int force_swap /*-1 auto*/, int force_bn /*0 auto*/,
                 void* stream);
/* Debug: while device_buf != NULL every GEMM CTA writes 8 phase timestamps (globaltimer, ns); the tcgen05 attention
 * kernel writes the clock64 trace of one mid-grid CTA (8 rows of 64 slots, tools/attn_trace.py) into the same buffer. */
int sm_test_gemm_trace(sm_handle* h, long long* device_buf);
int sm_test_attention(sm_handle* h, const void* qkv /*[B*S, 3*H*D]*/, void* out /*[B*S, H*D]*/, int B, int S, int H,
                      int D, void* stream);
/* The LLM prefill attention alone (causal, GQA, head_dim 128; hf MistralForCausalLM SDPA as reached from
 * videollama2_mistral.py:234-243): P query rows at positions pos0 .. pos0 + P - 1 (row i of q: Hq heads x 128 at column
 * h * 128, row pitch q_pitch elements, already rotated) against a cache [Hk][max_ctx][128] holding pos0 + P positions;
 * out [P, Hq * 128].  n_splits: -1 = the mma.sync kernel, 0 = tcgen05 kernel with the planned key-range split, > 0 forced. */
int sm_test_kv_attention(sm_handle* h, const void* q, int q_pitch, const void* kcache, const void* vcache, int max_ctx, void* out,
                         int P, int pos0, int Hq, int Hk, int n_splits, void* stream);

/* Per-kernel-class CUDA-event timing (used by bench.py's roofline pass; adds two event records per
 * launch, so keep it off on the timed path; ignored while capturing / replaying graphs).
 * sm_profile_read synchronises the device, fills accumulated milliseconds and launch counts per class
 * (index = class id, name via sm_profile_class_name), clears the accumulators and returns the number
 * of classes. */
int sm_profile_enable(sm_handle* h, int on);
int sm_profile_read(sm_handle* h, int max_classes, double* ms_by_class, long long* launches_by_class);
const char* sm_profile_class_name(int cls);

/* Debug / measurement: launch only the kernel classes whose bit is set (class ids as in
 * sm_profile_class_name); outputs are then meaningless, timings of the remaining kernels are exact.
 * bench.py uses it to time one kernel class at a time inside the same captured step. */
int sm_debug_kernel_filter(sm_handle* h, unsigned mask);

/* Debug / test: which attention kernel serves d = 64 non-causal attention (the vision tower).  -1 = default
 * (tcgen05 kernel csrc/attention_tc.cuh when the launch has >= 148 CTAs of 128 query rows, else the mma.sync
 * kernel csrc/attention.cuh; SMB_ATTN_TC=0/2 in the environment overrides), 0 = mma.sync kernel always,
 * 2 = tcgen05 kernel wherever its layout conditions hold. */
int sm_debug_attention_mode(sm_handle* h, int mode);

/* Launch accounting: number of this library's kernel launches (graph-replayed kernels included) since
 * the last call with reset != 0. */
long long sm_launch_count(sm_handle* h, int reset);

#ifdef __cplusplus
}
#endif
#endif /* STREAMMIND_B200_H_ */
