/* l3b200.h -- C ABI of the B200-native L3-Net AVC hot path (libl3b200.so).
 *
 * The reference (marl/l3embedding) has no FFI: its seam is the Keras Model API used by
 *   l3embedding/train.py:267,282,408-414   (MODELS[...](), compile, fit_generator -> train_on_batch)
 *   l3embedding/model.py:85-181            (load_model, load_embedding)
 *   data/usc/features.py:304               (model.predict on (n,1,48000) frames)
 * Each entry point below names the reference call it replaces.  Conventions:
 *   - every function returns 0 on success, <0 on error; l3_last_error() gives the message (thread local);
 *   - the caller owns every buffer: device pointers unless the name says _host;
 *   - all work is asynchronous on the ctx stream unless the doc says it synchronises;
 *   - activations NHWC, conv kernels HWIO, dense kernels (in,out) -- Keras array layouts;
 *   - no torch / C++ types in any signature.
 */
#ifndef L3B200_H
#define L3B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define L3_VERSION 1

/* l3embedding/model.py:307-313 MODELS keys (tiny_L3 is non-functional upstream and not provided) */
enum { L3_MODEL_ORIG = 0, L3_MODEL_KAPREDBINPUTBN = 1, L3_MODEL_MELSPEC1 = 2, L3_MODEL_MELSPEC2 = 3 };
/* activation storage / conv operand precision:
 *   F32   = parity mode, fp32 storage, SIMT fp32 convolutions (run-to-run deterministic);
 *   BF16  = throughput mode, bf16 storage, tcgen05 convolutions with fp32 accumulation;
 *   F32TC = parity mode ON TENSOR CORES: fp32 storage, the Cin%64==0 convolutions run on tcgen05 with every fp32 operand
 *           split into two 16-bit parts concatenated along K (forward: fp16 parts, 22 significant bits, weights pre-scaled
 *           by 2^10; backward: bf16 parts, fp32's range) and fp32 accumulation -- meets the 1e-3 embedding bar of the
 *           reference comparison at several times the SIMT mode's speed (activations must stay below 65504) */
enum { L3_DTYPE_F32 = 0, L3_DTYPE_BF16 = 1, L3_DTYPE_F32TC = 2 };
/* input formats */
enum { L3_VIDEO_U8 = 0, L3_VIDEO_F32 = 1 };   /* u8 raw frames (scaled on device, train.py:186) | f32 in [-1,1] */
enum { L3_AUDIO_I16 = 0, L3_AUDIO_F32 = 1 };  /* int16 PCM (pcm2float on device, audio.py:21-31) | f32 in [-1,1) */
/* workspace flags */
enum { L3_WS_TRAINING = 1, L3_WS_VISION = 2, L3_WS_AUDIO = 4, L3_WS_HOST_STAGING = 8 };
/* embedding pooling (audio_model.py:461-478) */
enum { L3_POOL_ORIGINAL = 0, L3_POOL_SHORT = 1 };

typedef struct l3_ctx l3_ctx;

int l3_version(void);
const char* l3_last_error(void);

/* ---- model inventory (replaces keras Model.get_weights()/weights introspection, model.py:77) ------------ */
int64_t l3_param_count(int model_type);   /* trainable fp32 scalars in the flat arena (9 508 746 for melspec2) */
int64_t l3_l2_count(int model_type);      /* leading scalars that are conv/dense kernels (l2 1e-5 regularised) */
int64_t l3_state_count(int model_type);   /* BN moving mean/variance scalars */
int l3_num_tensors(int model_type);
/* tensor i: name (Keras-like 'audio/conv1a/kernel'), arena (0 = params, 1 = bn state), offset (floats), dims */
int l3_tensor_info(int model_type, int i, char* name, int name_cap, int* arena, int64_t* offset, int* ndim,
                   int64_t dims[4]);
/* front-end output geometry: (n_freq_or_mels, n_frames) */
int l3_frontend_shape(int model_type, int* n_out, int* n_frames);
/* audio tower embedding-map geometry (H, W) of the raw conv4b output */
int l3_embedding_map_shape(int model_type, int* h, int* w);

/* ---- context -------------------------------------------------------------------------------------------- */
int64_t l3_workspace_bytes(int model_type, int max_batch, int dtype, int flags);
/* params/grads/adam_m/adam_v: l3_param_count floats each (grads/adam may be NULL for inference);
 * bn_state: l3_state_count floats; workspace: l3_workspace_bytes bytes, 256-byte aligned;
 * stream: a cudaStream_t (NULL = legacy default stream).                                                   */
l3_ctx* l3_ctx_create(int model_type, int max_batch, int dtype, int flags, float* params, float* grads,
                      float* adam_m, float* adam_v, float* bn_state, void* workspace, int64_t workspace_bytes,
                      void* stream);
void l3_ctx_destroy(l3_ctx* ctx);
/* force the SIMT convolution path even in bf16 mode (debug / A-B checks) */
int l3_ctx_set_use_tensor_cores(l3_ctx* ctx, int enable);
int l3_ctx_uses_tensor_cores(l3_ctx* ctx);
/* inference on the tensor-core path folds BatchNorm (moving statistics) + ReLU into the convolution epilogue and
 * writes the next layer's input directly (default on); 0 keeps the layer-by-layer path (A/B checks, tests) */
int l3_ctx_set_fused_inference(l3_ctx* ctx, int enable);

/* ---- hot path -------------------------------------------------------------------------------------------- */
/* async H2D of one batch from (pinned) host memory into one of the ctx's TWO staging slots, on the ctx's own copy
 * stream (needs L3_WS_HOST_STAGING).  Staged batches are consumed first-in-first-out by the calls that take NULL
 * inputs; at most two may be pending.  This is the one entry point that may be called from a second host thread while
 * another thread drives the step: a prefetch thread uploads batch k+1 while step k runs (the reference's generator,
 * train.py:142-195, is serial with the step).  Replaces the feed_dict copy of keras train_on_batch. */
int l3_upload_batch_host(l3_ctx* ctx, const void* video_host, int video_fmt, const void* audio_host, int audio_fmt,
                         const float* labels_host, int batch);
/* forward + backward of one AVC batch = the device part of keras train_on_batch (train.py:408-414).
 * video/audio/labels are device pointers, or all NULL to use the staged batch.  Gradients of the mean loss
 * over `global_batch` samples are left in the grads arena (sum over ranks == global gradient).              */
int l3_forward_backward(l3_ctx* ctx, const void* video, int video_fmt, const void* audio, int audio_fmt,
                        const float* labels, int batch, int global_batch);
/* Keras-2.0.9 Adam (train.py:282) incl. the l2(1e-5) regulariser gradient; increments the step counter. */
int l3_adam_step(l3_ctx* ctx, float lr);
int l3_adam_set_t(l3_ctx* ctx, int64_t t);
int64_t l3_adam_get_t(l3_ctx* ctx);        /* steps taken so far (checkpointed with the moments; <0 on error) */
/* out[4] = {sum of per-sample cross-entropy, #correct, l2 penalty (1e-5*sum w^2), batch}; synchronises. */
int l3_get_metrics(l3_ctx* ctx, float out[4]);
/* one keras train_on_batch on the oldest staged batch: forward_backward + metrics + adam; synchronises once. */
int l3_train_step_staged(l3_ctx* ctx, int batch, float lr, float out_metrics[4]);
/* single-GPU convenience: l3_upload_batch_host + l3_train_step_staged in one call. */
int l3_train_step_host(l3_ctx* ctx, const void* video_host, int video_fmt, const void* audio_host, int audio_fmt,
                       const float* labels_host, int batch, float lr, float out_metrics[4]);
/* inference-mode forward (keras predict / evaluate, BN moving statistics): probs (batch,2) device floats;
 * labels may be NULL; with labels, metrics are accumulated for l3_get_metrics.                              */
int l3_predict(l3_ctx* ctx, const void* video, int video_fmt, const void* audio, int audio_fmt, const float* labels,
               int batch, float* probs_out, float* logits_out);
/* load_embedding(...,'audio', pooling) + model.predict (model.py:131-181, features.py:304):
 * audio (n,1,48000) device -> out (n, 6144|512) device floats, n <= max_batch per call.                      */
int l3_embed_audio(l3_ctx* ctx, const void* audio, int audio_fmt, int n, int pooling, float* out);
/* get_l3_frames_uniform (data/usc/features.py:256-306) without materialising the framed copy: `signal` is ONE device
 * array of n_samples; frame i is the 1 s window starting at i*hop (hop in samples; hop*sizeof(sample) must be a
 * multiple of 16).  out (n_frames, 6144|512).  Runs in chunks of the context's max_batch. */
int l3_embed_audio_frames(l3_ctx* ctx, const void* signal, int audio_fmt, int64_t n_samples, int hop, int n_frames,
                          int pooling, float* out);
/* load_embedding(...,'vision',...) (vision_model.py:198-218): video (n,224,224,3) -> (n, 8192) */
int l3_embed_vision(l3_ctx* ctx, const void* video, int video_fmt, int n, float* out);

/* ---- data parallelism inside the library ------------------------------------------------------------------------
 * Replaces l3embedding/training_utils.py:21-170 (multi_gpu_model: batch slices, one replica per GPU, gradients summed
 * into shared variables).  One process / context per GPU; the host slices the batch (training_utils.py:121-133) and
 * every rank runs the same step on its slice with its own BN batch statistics (reference semantics: no sync-BN).
 * After l3_dp_init, l3_forward_backward sums the gradient arena over the ranks ITSELF: NCCL all-reduces on the
 * context's communication stream, in buckets that leave while the backward pass is still running (conv4b / conv4a /
 * conv3 of each tower as soon as their weight gradients are enqueued; a last grouped launch for the small rest).
 * l3_adam_step and l3_get_metrics wait for them; l3_get_metrics then reports the sums over the GLOBAL batch.
 * NCCL (libnccl.so.2) is bound at run time; bootstrap: rank 0 calls l3_dp_unique_id and hands the 128 bytes to the
 * other ranks by any means (the Python host uses torch.distributed / a file; a C host may use MPI or a socket). */
int l3_dp_unique_id(char out[128]);
int l3_dp_init(l3_ctx* ctx, const char id[128], int rank, int nranks);   /* collective: every rank must call it */
int l3_dp_info(l3_ctx* ctx, int* rank, int* nranks);                     /* returns 1 when initialised, else 0 */
int l3_dp_nccl_version(void);                                            /* e.g. 22809; 0 when NCCL is not loadable */
/* one keras train_on_batch of the GLOBAL batch on this rank's staged slice: forward_backward (+ overlapped gradient
 * exchange) + metrics + adam; synchronises once.  out_metrics = sums over the global batch. */
int l3_dp_train_step_staged(l3_ctx* ctx, int batch, int global_batch, float lr, float out_metrics[4]);
/* BN moving statistics are per replica during training (reference: one momentum update per replica); replaces them
 * by their mean over the ranks, so that every rank evaluates -- and rank 0 saves -- the same model. */
int l3_dp_average_bn_state(l3_ctx* ctx);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------------- */
/* kernel launches issued by this library in the calling process since it was loaded */
uint64_t l3_launch_count(void);
/* The two towers are independent until the head; by default the audio tower's kernels run on a second internal
 * stream (forked from / joined to the context stream with events) so HBM-bound kernels of one tower overlap
 * tensor-core kernels of the other.  0 serialises everything on the context stream (used for per-kernel timing). */
int l3_ctx_set_two_streams(l3_ctx* ctx, int enable);
/* optional CUDA-event timing of the convolution / front-end launches on the ctx stream.  l3_ctx_profile_read
 * synchronises and returns the milliseconds and launch counts accumulated since the previous read, per class:
 * [0] conv forward, [1] conv dgrad, [2] conv wgrad, [3] audio front-end. */
int l3_ctx_profile_enable(l3_ctx* ctx, int enable);
int l3_ctx_profile_read(l3_ctx* ctx, float ms_out[4], int launches_out[4]);

/* ---- single ops (unit tests / parity bisecting) ------------------------------------------------------------ */
/* kapre Spectrogram/Melspectrogram (+pcm2float): audio (n,48000) -> out (n, n_out, n_frames) float */
int l3_frontend_fwd(l3_ctx* ctx, const void* audio, int audio_fmt, int n, float* out);
/* Conv2D 3x3 same: in zero-haloed padded (B,H+2,W+2,Cin), w HWIO fp32, bias fp32 (may be NULL), out unpadded
 * (B,H,W,Cout); dtype of in/out per `dtype`; use_tc=1 runs the tcgen05 kernel (bf16 only, Cin%64==0, Cout%64==0),
 * scratch then holds the packed weights (>= 2*9*Cin*Cout bytes). */
int l3_conv3x3_fwd(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int Cin,
                   int Cout, int dtype, int use_tc, void* scratch, void* stream);
/* the tensor-core forward with its fused BatchNorm batch statistics (what the training step runs): as l3_conv3x3_fwd
 * with dtype bf16 / use_tc 1 (Cin 1|3 with Cout 64, or Cin%64==0 and Cout%64==0), plus stats: device double[2*Cout]
 * <- per-channel sum and sum of squares of the stored bf16 output (of relu(output) if relu_stats; Cin%64 layers only). */
int l3_conv3x3_fwd_stats(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int Cin,
                         int Cout, void* scratch, double* stats, int relu_stats, void* stream);
/* data gradient of the same conv: dz zero-haloed padded (B,H+2,W+2,Cout) -> da unpadded (B,H,W,Cin);
 * scratch >= 9*Cin*Cout floats. */
int l3_conv3x3_dgrad(const void* dz, const float* w, void* da, int B, int H, int W, int Cin, int Cout, int dtype,
                     int use_tc, void* scratch, void* stream);
/* the tensor-core data gradient with pass 1 of the BN/ReLU backward of the layer below fused into its epilogue (what
 * the training step runs for un-pooled Conv -> BN -> ReLU layers of 128 channels and more): as l3_conv3x3_dgrad with dtype bf16 / use_tc 1, plus
 * z_below (B,H,W,Cin) bf16, scale / shift (Cin floats: bn(z) = scale*z + shift) and
 * sums: device double[2*Cin] <- sum(dy), sum(dy*z_below) per channel, dy = da where bn(z_below) > 0, else 0. */
int l3_conv3x3_dgrad_stats(const void* dz, const float* w, void* da, int B, int H, int W, int Cin, int Cout, void* scratch,
                           const void* z_below, const float* scale, const float* shift, double* sums, void* stream);
/* weight/bias gradient: a zero-haloed padded (B,H+2,W+2,Cin), dz zero-haloed padded (B,H+2,W+2,Cout);
 * dw (3,3,Cin,Cout) and db (Cout, may be NULL) are overwritten. */
int l3_conv3x3_wgrad(const void* a, const void* dz, float* dw, float* db, int B, int H, int W, int Cin, int Cout,
                     int dtype, int use_tc, void* stream);
/* ---- element-wise layer ops of the training step, stand-alone (unit tests).  dtype selects float / bf16 storage; the
 * ops are the kernels the step itself launches.  Reference semantics: keras BatchNormalization (axis -1) ->
 * Activation('relu') -> MaxPooling2D(2) as instantiated at audio_model.py:376-437, vision_model.py:130-190; relu_first
 * is the Conv -> ReLU -> BN order of vision_model.py:135-139. ------------------------------------------------------ */
/* a = pool2x2?( relu_first ? scale*relu(z)+shift : relu(scale*z+shift) ): z unpadded (B,H,W,C), a zero-haloed padded
 * (B,OH+2,OW+2,C) with OH = pool ? H/2 : H ('valid' pooling drops odd rows/columns).  The halo of `a` is not written.
 * zsel / sel (optional, pool only): the winning pre-activation (B,OH,OW,C) and one byte per pooled element
 * (window position 0..3 | 4*(max > 0)) -- the record the backward pass of the step consumes.                       */
int l3_act_fwd(const void* z, void* a, int B, int H, int W, int C, const float* scale, const float* shift, int pool,
               int relu_first, int dtype, void* zsel, uint8_t* sel, void* stream);
/* Backward of the same block in training mode (gradient through the batch statistics included):
 * da unpadded (B,OH,OW,C), z unpadded (B,H,W,C) -> dz zero-haloed padded (B,H+2,W+2,C) incl. its halo;
 * bn4 = float[4*C] {scale = gamma*invstd, shift = beta - mean*scale, mean, invstd} of the batch statistics of z
 * (of relu(z) if relu_first); d_gamma / d_beta: float[C] outputs.  zsel / sel: the forward record or NULL (the
 * routing is then re-derived from z).                                                                              */
int l3_bn_act_bwd(const void* da, const void* z, void* dz, int B, int H, int W, int C, const float* bn4, int pool,
                  int relu_first, int dtype, const void* zsel, const uint8_t* sel, float* d_gamma, float* d_beta,
                  void* stream);
/* MaxPooling2D over the whole map of relu(scale*z+shift) (audio_model.py:436, vision_model.py:189):
 * out (B,C) float, argmax (B,C) int = first maximum in row-major order.                                            */
int l3_gmaxpool_fwd(const void* z, int B, int H, int W, int C, const float* scale, const float* shift, int dtype,
                    float* out, int* argmax, void* stream);
/* its backward fused with the BN backward of the last layer: dpool (B,C) float -> dz zero-haloed padded (B,H+2,W+2,C). */
int l3_gmaxpool_bwd(const float* dpool, const int* argmax, const void* z, void* dz, int B, int H, int W, int C,
                    const float* bn4, int dtype, float* d_gamma, float* d_beta, void* stream);

/* device-side peek at internal activations for tests: which = "audio/z3", "vision/a1", "audio/x0", "concat" ...
 * copies up to `cap` floats (converted to f32) into out_host; returns element count or <0; synchronises. */
int64_t l3_debug_read(l3_ctx* ctx, const char* which, int batch, float* out_host, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* L3B200_H */
