/* parrot_b200 -- C ABI of the Blackwell-native Parrot (Char2Wav reader) hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI:
 * its hot path is the Theano function compiled from model.py:Parrot.compute_cost
 * (model.py:552-824) / sample_model_fun (model.py:827-1059) and stepped by
 * blocks GradientDescent (train.py:100-108).  The entry points below are what a
 * ctypes binding replaces those compiled functions with; INTEGRATION.md shows
 * the binding.  Conventions:
 *   - plain C, no C++/torch types; every function returns 0 on success, non-zero
 *     on error, and parrot_last_error() describes the last failure of the thread;
 *   - all pointers named d_* are DEVICE pointers owned by the caller;
 *   - no hidden device allocation: the caller provides one workspace whose size
 *     parrot_workspace_bytes() reports; no hidden synchronisation: everything is
 *     enqueued on the caller's cudaStream_t (passed as void*);
 *   - a handle is built for one (batch, frames, text length) triple, like the
 *     reference bakes batch_size into its graph (train.py:87).
 * Layouts follow the reference: frame tensors time-major (T+1 or T, B, .),
 * text tensors batch-major (B, U); parameters in Blocks orientation (in, out).
 */
#ifndef PARROT_B200_H
#define PARROT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirror of the Parrot constructor keywords (model.py:251-277) + problem sizes. */
typedef struct parrot_config {
  int32_t input_dim;          /* model.py:253 */
  int32_t output_dim;         /* model.py:254 */
  int32_t rnn_h_dim;          /* model.py:255 */
  int32_t readouts_dim;       /* model.py:256 */
  int32_t weak_feedback;      /* model.py:257 */
  int32_t full_feedback;      /* model.py:258 */
  int32_t layer_norm;         /* model.py:260 */
  int32_t use_speaker;        /* model.py:261 */
  int32_t num_speakers;       /* model.py:262 */
  int32_t speaker_dim;        /* model.py:263 */
  int32_t which_cost;         /* model.py:264  0 = MSE, 1 = GMM */
  int32_t k_gmm;              /* model.py:265 */
  int32_t num_characters;     /* model.py:268 */
  int32_t attention_type;     /* model.py:269  0 = graves, 1 = softmax */
  int32_t attention_size;     /* model.py:270 */
  int32_t encoder_type;       /* model.py:274  0 = None, 1 = bidirectional */
  int32_t encoder_dim;        /* model.py:275 */
  int32_t encoder_time_axis;  /* 0 = literal reference behaviour (scan over axis 0), 1 = text axis */
  float sampling_bias;        /* model.py:266 */
  float epsilon;              /* model.py:267 */
  float attention_alignment;  /* model.py:271 */
  float sharpening_coeff;     /* model.py:272 */
  float timing_coeff;         /* model.py:273 */
  int32_t batch_size;         /* B on this device */
  int32_t seq_len;            /* T: frames per segment after the one-frame shift (model.py:559-560) */
  int32_t text_len;           /* U */
  int32_t gemm_impl;          /* 0 = tcgen05 tensor cores, 1 = SIMT verification twin */
  int32_t sampling;           /* 0 = training handle (compute_cost), 1 = sampling handle (sample_model) */
} parrot_config;

typedef struct parrot_model parrot_model;

const char* parrot_last_error(void);
int parrot_abi_version(void);

/* ---- parameter inventory (Blocks brick paths, e.g. "/parrot/rnn1.state_to_gates") ---- */
/* number of parameter tensors / total floats of the flat buffer for this configuration */
int parrot_param_count(const parrot_config* cfg, int32_t* count, int64_t* total_floats);
/* i-th tensor: name (copied, NUL-terminated), offset in floats, shape (rows = in, cols = out; cols = 0 for vectors) */
int parrot_param_info(const parrot_config* cfg, int32_t i, char* name, int32_t name_cap, int64_t* offset,
                      int32_t* rows, int32_t* cols);

/* ---- lifetime ---- */
int parrot_workspace_bytes(const parrot_config* cfg, size_t* bytes);
/* d_params / d_grads: flat fp32 buffers (grads holds total_floats + 1 floats: the last one receives
 * sum(mask), so that data-parallel ranks can allreduce gradients and the mask count together). */
int parrot_create(const parrot_config* cfg, float* d_params, float* d_grads, void* d_workspace,
                  size_t workspace_bytes, void* stream, parrot_model** out);
int parrot_destroy(parrot_model* m);
/* named internal buffers (for attention_vars, carried state, tests): byte offset into the workspace */
int parrot_buffer_info(parrot_model* m, const char* name, int64_t* byte_offset, int64_t* numel);

/* Re-derive the bf16 hi/lo operand planes from the fp32 parameters (call after every optimizer step;
 * parrot_compute_cost does it automatically when parrot_mark_params_dirty was called). */
int parrot_pack_weights(parrot_model* m, void* stream);
int parrot_mark_params_dirty(parrot_model* m);
/* per-launch CUDA-event timing of the engine / attention launches, aggregated by table name
 * ("fwdA", "fwdB", "bwd1", "bwd2", "readout", "output", "dread", "dh_readout", "wgrad", "attn_fwd", "attn_bwd",
 * "gru_bwd_pre", "sec_*" for whole sections).  parrot_get_profile synchronises the device. */
int parrot_set_profiling(parrot_model* m, int enable);
/* debug: the persistent forward scan writes [cta][bars][2] globaltimer stamps (barrier passed, arrival) */
int parrot_debug_set_stamps(parrot_model* m, unsigned long long* d_stamps, int bars);
/* debug / measurement: average time of one engine table launched back to back (optional per-CTA timeline) */
int parrot_debug_time_table(parrot_model* m, const char* name, int tick, int reverse, int reps, float* avg_ms,
                            unsigned long long* d_timeline, void* stream);
int parrot_get_profile(parrot_model* m, const char* key, double* total_ms, int64_t* launches);

/* ---- the hot path, fine grained (SURVEY 8b "minimum surface") ---- */
/* replaces Encoder.apply (model.py:233-247) + the context mask (model.py:645-646) */
int parrot_encoder_fwd(parrot_model* m, const int32_t* d_labels, const float* d_labels_mask, void* stream);
int parrot_encoder_bwd(parrot_model* m, void* stream);
/* replaces theano.scan(step) (model.py:726-737): T decoder steps, teacher forced */
int parrot_decoder_scan_fwd(parrot_model* m, const float* d_features, const float* d_feedback_noise,
                            float noise_level, float start_flag, void* stream);
int parrot_decoder_scan_bwd(parrot_model* m, void* stream);
/* replaces the readout / emitter / cost tail (model.py:739-784); d_cost receives
 * [cost, sum(cost*mask), sum(mask), 1/(sum(mask)+1e-5)] */
int parrot_readout_emit_fwd(parrot_model* m, const float* d_features, const float* d_features_mask,
                            float* d_cost, void* stream);
int parrot_readout_emit_bwd(parrot_model* m, int unnormalised, void* stream);
/* stand-alone attention window step (model.py:664-690), also the attn-step latency metric.
 * h1 [B][H], k_prev [B][A], ctx [B][U][C] -> k_out [B][A], w_out [B][C], phi_out [B][U], ab_out [B][2A] */
int parrot_attention_step(const parrot_config* cfg, const float* d_h1, const float* d_wT, const float* d_batt,
                          const float* d_ctx, const float* d_k_prev, float* d_k_out, float* d_w_out,
                          float* d_phi_out, float* d_ab_out, float* d_e_out, int training, void* stream);

/* ---- the hot path, coarse grained: what Parrot.compute_cost + theano.grad run ---- */
/* forward: encoder, scan, emitter, cost; applies the carried-state updates (model.py:786-791) */
int parrot_compute_cost(parrot_model* m, const float* d_features, const float* d_features_mask,
                        const int32_t* d_labels, const float* d_labels_mask, const int32_t* d_speaker,
                        float start_flag, const float* d_feedback_noise, float noise_level,
                        const float* d_gmm_unis, const float* d_gmm_normals, float* d_cost, void* stream);
/* backward of the last parrot_compute_cost into d_grads (overwrites).  unnormalised != 0 leaves out the
 * 1/(sum(mask)+1e-5) factor (data-parallel: divide after the allreduce). */
int parrot_backward(parrot_model* m, int unnormalised, void* stream);

/* free-running generation (model.py:1061-1083): num_steps = cfg.seq_len, num_samples = cfg.batch_size.
 * Noise: pass d_unis [T][B] / d_normals [T][B][D] for injected draws, or null + seed for Philox. */
int parrot_sample_scan(parrot_model* m, const int32_t* d_labels, const float* d_labels_mask,
                       const int32_t* d_speaker, const float* d_unis, const float* d_normals, uint64_t seed,
                       void* stream);

/* ---- optimizer: StepClipping(threshold) + Adam (train.py:100-108) on flat buffers ---- */
/* Every gradient is first multiplied by grad_scale and, when d_mask_sum is not null, by
 * 1/(d_mask_sum[0] + 1e-5) read on the device (data-parallel: the un-normalised gradients and sum(mask) are
 * all-reduced together, model.py:784 is a masked mean over the global batch) -- no host synchronisation.
 * d_stats receives [global grad norm, clip multiplier, total multiplier]; d_scratch needs 1024 doubles. */
int parrot_adam_clip_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n,
                          float grad_scale, const float* d_mask_sum, float threshold, float learning_rate,
                          float beta1, float beta2, float epsilon, int64_t time_step, float* d_stats,
                          double* d_scratch, void* stream);

/* ---- data-parallel collective: ONE ncclAllReduce(SUM, fp32) per optimizer step over the
 * flat [un-normalised gradients || sum(mask)] buffer, enqueued on the caller's stream right behind the last weight
 * gradient kernel.  The reference is single-process (no collective anywhere): this replaces nothing in it,
 * it is what train.py:100-108's GradientDescent step becomes when the minibatch is sharded over GPUs.
 * NCCL is resolved at run time (dlopen of libnccl.so.2, the copy the process already loaded if any); a 1-rank
 * communicator needs neither NCCL nor a GPU and its allreduce is the identity.
 *   rank 0:      parrot_comm_unique_id(id)            -> 128 bytes, broadcast by the launcher (any side channel)
 *   every rank:  parrot_comm_init(nranks, rank, id, &comm)   (cudaSetDevice must already be done)
 *   per step:    parrot_comm_allreduce(comm, d_flat, count, stream)      in place, fp32 SUM
 */
typedef struct parrot_comm parrot_comm;
#define PARROT_COMM_ID_BYTES 128
int parrot_comm_unique_id(void* id128);
int parrot_comm_init(int32_t nranks, int32_t rank, const void* id128, parrot_comm** out);
int parrot_comm_allreduce(parrot_comm* comm, float* d_buf, int64_t count, void* stream);
int parrot_comm_info(parrot_comm* comm, int32_t* nranks, int32_t* rank, int32_t* nccl_version);
int parrot_comm_destroy(parrot_comm* comm);

/* generic bf16x3 tensor-core GEMM used by the tests:  C[M][N] = A[M][K] * B[N][K]^T  (fp32 in/out) */
int parrot_gemm_nt(const float* d_A, const float* d_B, float* d_C, int32_t M, int32_t N, int32_t K,
                   int32_t impl, void* d_workspace, size_t workspace_bytes, void* stream);
size_t parrot_gemm_nt_workspace_bytes(int32_t M, int32_t N, int32_t K);

/* number of kernels launched by this library in this process (bench.py "gpu_launches") */
int64_t parrot_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PARROT_B200_H */
