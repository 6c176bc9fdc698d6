// C ABI (include/l3b200.h) and step orchestration of the B200-native L3-Net AVC path.
//
// What the reference does with one keras `train_on_batch` / `predict` (l3embedding/train.py:408-414,
// data/usc/features.py:304) over the graph built by l3embedding/model.py:198-284 is done here as an explicit
// sequence of kernel launches on one CUDA stream over caller-owned arenas:
//   params / grads / adam_m / adam_v : flat fp32 arenas, [conv+dense kernels (l2-regularised) | biases, BN gamma/beta]
//   bn_state                         : BN moving mean / variance
//   workspace                        : activations (NHWC; conv inputs in a zero-haloed (B,H+2,W+2,C) layout so a
//                                      3x3 tap is a constant row shift of the flattened pixel index), BN scratch.
// Data-parallel training: every rank calls l3_forward_backward on its slice with global_batch = N*batch, sums the
// grads arena over ranks (NCCL all-reduce, done by the host), then calls l3_adam_step; BN statistics stay per
// replica as under the reference's multi_gpu_model (l3embedding/training_utils.py:141-162).
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/l3b200.h"
#include "kernels.h"

namespace l3 {

unsigned long long g_launch_count = 0;
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------------------
// model specifications (restated from l3embedding/audio_model.py, vision_model.py, model.py)
// ---------------------------------------------------------------------------------------------------------
struct ModelSpec {
  int mel;          // kapre Melspectrogram (1) or Spectrogram (0)
  int n_dft;        // audio_model.py:26,138,245,355
  int n_mels;       // :248,358
  int same_pad;     // padding='same' (mel models) / 'valid' (kapre default)
  int decibel;      // return_decibel_* ; 0 -> log(max(x,1e-12))/5 (audio_model.py:43)
  int audio_bn0;    // input BatchNormalization (audio_model.py:151,260,370)
  int vision_bn0;   // vision_model.py:124
  int embed_pool[2][2];  // audio_model.py:461-478: [original|short][h,w]
};
static const ModelSpec kSpecs[4] = {
    /* orig           */ {0, 512, 0, 0, 0, 0, 0, {{8, 8}, {32, 24}}},
    /* kapredbinputbn */ {0, 512, 0, 0, 1, 1, 1, {{8, 8}, {32, 24}}},
    /* melspec1       */ {1, 2048, 128, 1, 1, 1, 1, {{4, 8}, {16, 24}}},
    /* melspec2       */ {1, 2048, 256, 1, 1, 1, 1, {{8, 8}, {32, 24}}},
};
static const int kConvCh[9] = {0, 64, 64, 128, 128, 256, 256, 512, 512};  // Cout of conv l = kConvCh[l+1]
static const char* kConvNames[8] = {"1a", "1b", "2a", "2b", "3a", "3b", "4a", "4b"};
static const int kSR = 48000, kHop = 242;
static const int kVisionEmbedPool = 7;  // vision_model.py:212

static void audio_geometry(const ModelSpec& sp, int* n_out, int* n_frames, int* left_pad) {
  *n_out = sp.mel ? sp.n_mels : sp.n_dft / 2 + 1;
  if (sp.same_pad) {  // TF SAME: ceil(L/hop) frames, total pad split floor/ceil
    *n_frames = (kSR + kHop - 1) / kHop;
    int total = (*n_frames - 1) * kHop + sp.n_dft - kSR;
    if (total < 0) total = 0;
    *left_pad = total / 2;
  } else {
    *n_frames = (kSR - sp.n_dft) / kHop + 1;
    *left_pad = 0;
  }
}

struct TensorInfo {
  std::string name;
  int arena;  // 0 params, 1 bn state
  long long offset;
  int ndim;
  long long dims[4];
  long long size() const {
    long long n = 1;
    for (int i = 0; i < ndim; ++i) n *= dims[i];
    return n;
  }
};
struct Layout {
  std::vector<TensorInfo> t;  // keras layer order (vision tower, audio tower, dense_1, dense_2)
  long long n_params = 0, n_l2 = 0, n_state = 0;
  int find(const std::string& nm) const {
    for (size_t i = 0; i < t.size(); ++i)
      if (t[i].name == nm) return (int)i;
    return -1;
  }
};

static Layout build_layout(int model_type) {
  const ModelSpec& sp = kSpecs[model_type];
  Layout L;
  // pass 0 assigns the l2-regularised kernels (front of the params arena), pass 1 the rest
  std::vector<TensorInfo> all;
  auto add = [&](const std::string& nm, int arena, int nd, long long d0, long long d1, long long d2, long long d3) {
    TensorInfo ti;
    ti.name = nm;
    ti.arena = arena;
    ti.offset = -1;
    ti.ndim = nd;
    ti.dims[0] = d0; ti.dims[1] = d1; ti.dims[2] = d2; ti.dims[3] = d3;
    all.push_back(ti);
  };
  auto add_bn = [&](const std::string& pre, int c) {
    add(pre + "/gamma", 0, 1, c, 1, 1, 1);
    add(pre + "/beta", 0, 1, c, 1, 1, 1);
    add(pre + "/moving_mean", 1, 1, c, 1, 1, 1);
    add(pre + "/moving_variance", 1, 1, c, 1, 1, 1);
  };
  for (int tw = 0; tw < 2; ++tw) {
    std::string T = tw == 0 ? "vision" : "audio";
    int c0 = tw == 0 ? 3 : 1;
    if (tw == 0 ? sp.vision_bn0 : sp.audio_bn0) add_bn(T + "/bn0", c0);
    for (int l = 0; l < 8; ++l) {
      int ci = l == 0 ? c0 : kConvCh[l], co = kConvCh[l + 1];
      add(T + "/conv" + kConvNames[l] + "/kernel", 0, 4, 3, 3, ci, co);
      add(T + "/conv" + kConvNames[l] + "/bias", 0, 1, co, 1, 1, 1);
      add_bn(T + "/bn" + kConvNames[l], co);
    }
  }
  add("dense_1/kernel", 0, 2, 1024, 128, 1, 1);
  add("dense_1/bias", 0, 1, 128, 1, 1, 1);
  add("dense_2/kernel", 0, 2, 128, 2, 1, 1);
  add("dense_2/bias", 0, 1, 2, 1, 1, 1);
  auto is_kernel = [](const std::string& n) { return n.size() > 7 && n.compare(n.size() - 7, 7, "/kernel") == 0; };
  long long po = 0, so = 0;
  for (auto& ti : all)
    if (ti.arena == 0 && is_kernel(ti.name)) { ti.offset = po; po += ti.size(); }
  L.n_l2 = po;
  for (auto& ti : all) {
    if (ti.arena == 0 && !is_kernel(ti.name)) { ti.offset = po; po += ti.size(); }
    if (ti.arena == 1) { ti.offset = so; so += ti.size(); }
  }
  L.n_params = po;
  L.n_state = so;
  L.t = all;
  return L;
}

// ---------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------
struct ConvLayer {
  int H, W, Cin, Cout;      // conv resolution
  int pool;                 // 2x2 max-pool after the activation
  int relu_first;           // vision conv1b: Conv -> ReLU -> BN (vision_model.py:39-43,135-139)
  const float *w, *b;       // params
  float *dw, *db;           // grads
  BnRef bn;
  void* in;                 // padded input  (B,H+2,W+2,Cin)
  void* z;                  // raw conv output, unpadded (B,H,W,Cout)
  void* a;                  // padded activation output at post-pool resolution (null for the last layer)
  void* zsel;               // pooled layers, training: winning pre-activation per pooled element (B,H/2,W/2,Cout)
  uint8_t* sel;             //   and window position | 4*(max > 0), one byte per pooled element
  float* w_t;               // flipped/transposed fp32 kernel for dgrad-as-conv (SIMT path)
  bf16* w_pk;               // tcgen05 operand packs (forward, dgrad)
  bf16* wt_pk;
  int tc;                   // this layer runs on the tcgen05 path (bf16 throughput mode)
  int tcs;                  // this layer runs on the tcgen05 path with SPLIT 16-bit operands and fp32 storage (parity mode
                            // L3_DTYPE_F32TC): w_pk = fp16 parts [hi ; hi ; lo] of 2^10 * w, wt_pk = bf16 parts (dgrad)
};
struct Tower {
  int present;
  int C0, H0, W0;
  int has_bn0;
  BnRef bn0;
  float* x0;    // (B,H0,W0,C0) float: scaled video / front-end output
  void* xin;    // padded T
  ConvLayer L[8];
  int* argmax;  // (B,512)
  unsigned long long* pool_scratch;  // (B,512) packed (value, ~index) keys of the global max-pool
  float* d1;    // 9*64 floats + 1 int flag: first-layer "ones" weight gradient for the input-BN backward
  void* sp_a;     // L3_DTYPE_F32TC: split operand buffers, (B,H+2,W+2,3*C) 16-bit parts of a convolution input / of dz
  void* sp_z;
  float* dw4;     //   and the 2Cin x 2Cout weight-gradient scratch of launch_wgrad3x3_tc_split
  double* wg64;   // parity mode (f32), training: fp64 merge scratch of the weight-gradient kernels (one launch at a time
  double* wg64_0; //   per stream: the side stream's layers share wg64, the first layer on the main stream has wg64_0)
  int concat_off;
  void *g0, *g1;         // backward buffers of this tower: padded dz / unpadded da
  void* g0b;             // second dz buffer: layer l's dz lives in {g0, g0b}[l & 1] so that its weight gradient can run
                         // on the side stream while the layers below already write the other buffer
  cudaStream_t stream;   // stream this tower's kernels are launched on (set by the caller of tower_*)
  cudaStream_t wstream;  // low-priority side stream for the weight-gradient kernels (null: same stream)
  cudaEvent_t ev_dz[2], ev_wg[2];   // dz(l) ready / wgrad(l) done, per dz buffer
  int wg_pending[2];
};

}  // namespace l3

using namespace l3;

struct l3_ctx {
  int model_type, max_batch, dtype, flags;
  ModelSpec spec;
  Layout layout;
  float *params, *grads, *adam_m, *adam_v, *bn_state;
  char* ws;
  long long ws_bytes;
  cudaStream_t stream;
  FrontendPlan fe;
  void* fe_tables;
  int* clip_max;
  Tower vision, audio;
  HeadRef head;
  double* l2_out;
  float* unit_scale;   // 512 x 1.0f / 512 x 0.0f: identity BN coefficients for the pool-only pass of the fused inference path
  float* zero_shift;
  // the two towers are independent until the head: the audio tower runs on a second stream so its HBM-bound
  // kernels overlap the vision tower's tensor-core kernels (and vice versa)
  cudaStream_t stream2;
  cudaEvent_t ev_fork, ev_join, ev_pack;
  int two_streams;
  int pack_event_pending;   // this step's operand packs were enqueued on the context stream: towers wait for ev_pack
  // Inside a tower the backward chain is dgrad -> BN/ReLU backward (HBM-bound) -> dgrad ...; the weight gradients hang
  // off it as leaves.  With wgrad_streams they run on a low-priority side stream per tower and fill the tensor pipe
  // while the chain is in its HBM-bound kernels (needs the double dz buffer and the internal tower streams).
  cudaStream_t stream_v;        // vision tower's own (high-priority) stream when two_streams: the caller's stream has
  cudaEvent_t ev_join_v;        // the default (lowest) priority and would not win against the side streams
  int wgrad_streams;
  // host staging: two device slots filled by l3_upload_batch_host on a copy stream (FIFO), so that the upload of batch
  // k+1 -- issued by a prefetch thread -- overlaps the step of batch k.  ev_staged[s]: upload into slot s complete;
  // ev_consumed[s]: the step that read slot s has finished reading it (the slot may be overwritten).
  int device;
  void *st_video[2], *st_audio[2];
  float* st_labels[2];
  int st_video_fmt[2], st_audio_fmt[2], st_batch[2];
  int st_head, st_tail, st_count;
  bool st_consumed_valid[2];
  cudaStream_t copy_stream;
  cudaEvent_t ev_staged[2], ev_consumed[2];
  std::mutex st_mu;
  float* metrics_host;   // pinned: {ce sum, #correct} + l2 (double) of the last enqueue_metrics
  // data parallelism (l3_dp_init): NCCL communicator + communication stream; the gradient buckets of a step are
  // all-reduced while the backward pass is still running (tower_backward_layer, dp_fire_tail)
  DpState* dp;
  int last_global_batch;
  long long adam_t;
  int use_tc;
  int fuse_inference;    // fused BN + ReLU conv epilogue on the inference path (l3_ctx_set_fused_inference; default on)
  int last_batch;
  // optional per-kernel-class device timing (CUDA events on the ctx stream)
  int prof_on;
  std::vector<cudaEvent_t> prof_ev;   // pairs (start, stop)
  std::vector<int> prof_cls;
  size_t prof_used;
};

namespace l3 {

// classes reported by l3_ctx_profile_read
enum { PROF_CONV_FWD = 0, PROF_CONV_DGRAD = 1, PROF_CONV_WGRAD = 2, PROF_FRONTEND = 3, PROF_N = 4 };
struct ProfScope {
  l3_ctx* c;
  cudaEvent_t stop;
  cudaStream_t st;
  ProfScope(l3_ctx* ctx, int cls, cudaStream_t stream) : c(ctx), stop(nullptr), st(stream) {
    if (!c->prof_on) return;
    if (c->prof_used * 2 + 2 > c->prof_ev.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      c->prof_ev.push_back(a);
      c->prof_ev.push_back(b);
      c->prof_cls.push_back(cls);
    }
    c->prof_cls[c->prof_used] = cls;
    cudaEventRecord(c->prof_ev[c->prof_used * 2], st);
    stop = c->prof_ev[c->prof_used * 2 + 1];
    c->prof_used++;
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};

struct Bump {
  char* base;
  long long off;
  void* take(long long bytes) {
    long long o = off;
    off += (bytes + 255) / 256 * 256;
    return base ? (void*)(base + o) : nullptr;
  }
};

static size_t esize(int dtype) { return dtype == L3_DTYPE_BF16 ? 2 : 4; }

static void carve_bn(Bump& bp, BnRef& bn, int C) {
  bn.C = C;
  bn.sum = (double*)bp.take(sizeof(double) * 2 * C);
  bn.mean = (float*)bp.take(4 * C);
  bn.invstd = (float*)bp.take(4 * C);
  bn.scale = (float*)bp.take(4 * C);
  bn.shift = (float*)bp.take(4 * C);
  bn.c1 = (float*)bp.take(4 * C);
  bn.c2 = (float*)bp.take(4 * C);
}


// Walks every workspace allocation; with ctx->ws == nullptr it only measures.
static long long carve(l3_ctx* c) {
  Bump bp{c->ws, 0};
  const long long B = c->max_batch;
  const size_t es = esize(c->dtype);
  const bool training = c->flags & L3_WS_TRAINING;
  int n_out, n_frames, left;
  audio_geometry(c->spec, &n_out, &n_frames, &left);
  c->fe_tables = bp.take(frontend_table_bytes(c->spec.n_dft, c->spec.mel ? c->spec.n_mels : 1));
  c->clip_max = (int*)bp.take(4 * B);
  for (int sl = 0; sl < 2; ++sl) {
    const bool on = (c->flags & L3_WS_HOST_STAGING) != 0;
    c->st_video[sl] = on ? bp.take(B * 224 * 224 * 3 * 4) : nullptr;
    c->st_audio[sl] = on ? bp.take(B * kSR * 4) : nullptr;
    c->st_labels[sl] = on ? (float*)bp.take(B * 2 * 4) : nullptr;
  }
  for (int t = 0; t < 2; ++t) {
    long long g0_max = 0, g1_max = 0, spa_max = 0, spz_max = 0, dw4_max = 0;
    Tower& tw = t == 0 ? c->vision : c->audio;
    tw.present = (c->flags & (t == 0 ? L3_WS_VISION : L3_WS_AUDIO)) ? 1 : 0;
    tw.C0 = t == 0 ? 3 : 1;
    tw.H0 = t == 0 ? 224 : n_out;
    tw.W0 = t == 0 ? 224 : n_frames;
    tw.has_bn0 = t == 0 ? c->spec.vision_bn0 : c->spec.audio_bn0;
    tw.concat_off = t == 0 ? 0 : 512;  // model.py:25 concatenate([vision, audio])
    if (!tw.present) continue;
    tw.x0 = (float*)bp.take(4 * B * tw.H0 * tw.W0 * tw.C0);
    tw.xin = bp.take(es * B * (tw.H0 + 2) * (tw.W0 + 2) * tw.C0);
    if (tw.has_bn0) carve_bn(bp, tw.bn0, tw.C0);
    int H = tw.H0, W = tw.W0;
    void* prev = tw.xin;
    for (int l = 0; l < 8; ++l) {
      ConvLayer& L = tw.L[l];
      L.H = H; L.W = W;
      L.Cin = l == 0 ? tw.C0 : kConvCh[l];
      L.Cout = kConvCh[l + 1];
      L.pool = (l == 1 || l == 3 || l == 5) ? 1 : 0;
      L.relu_first = (t == 0 && l == 1) ? 1 : 0;
      L.in = prev;
      L.z = bp.take(es * B * H * W * L.Cout);
      carve_bn(bp, L.bn, L.Cout);
      int OH = L.pool ? H / 2 : H, OW = L.pool ? W / 2 : W;  // valid pooling floors; vision sizes are even
      L.a = l < 7 ? bp.take(es * B * (OH + 2) * (OW + 2) * L.Cout) : nullptr;
      const bool rec = training && L.pool;
      L.zsel = rec ? bp.take(es * B * OH * OW * L.Cout) : nullptr;
      L.sel = rec ? (uint8_t*)bp.take(B * OH * OW * L.Cout) : nullptr;
      L.w_t = training ? (float*)bp.take(4LL * 9 * L.Cin * L.Cout) : nullptr;
      L.tc = (c->dtype == L3_DTYPE_BF16 && L.Cin % 64 == 0 && L.Cout % 64 == 0) ? 1 : 0;
      L.tcs = (c->dtype == L3_DTYPE_F32TC && L.Cin % 64 == 0 && L.Cout % 64 == 0) ? 1 : 0;
      const long long pk = 2LL * 9 * L.Cin * L.Cout * (L.tcs ? 3 : 1);
      L.w_pk = (L.tc || L.tcs) ? (bf16*)bp.take(pk) : nullptr;
      L.wt_pk = ((L.tc || L.tcs) && training) ? (bf16*)bp.take(pk) : nullptr;
      long long dz = B * (H + 2) * (W + 2) * L.Cout, da = B * H * W * L.Cin;
      if (dz > g0_max) g0_max = dz;
      if (da > g1_max) g1_max = da;
      if (L.tcs) {
        const long long a3 = B * (H + 2) * (W + 2) * 3 * L.Cin, w4 = 9LL * 4 * L.Cin * L.Cout;
        if (a3 > spa_max) spa_max = a3;
        if (3 * dz > spz_max) spz_max = 3 * dz;
        if (w4 > dw4_max) dw4_max = w4;
      }
      prev = L.a;
      H = OH; W = OW;
    }
    tw.argmax = (int*)bp.take(4 * B * 512);
    tw.pool_scratch = (unsigned long long*)bp.take(8 * B * 512);
    tw.d1 = (float*)bp.take(4 * (9 * 64 + 4));
    tw.sp_a = spa_max ? bp.take(2 * spa_max) : nullptr;
    tw.sp_z = (spz_max && training) ? bp.take(2 * spz_max) : nullptr;
    tw.dw4 = (dw4_max && training) ? (float*)bp.take(4 * dw4_max) : nullptr;
    const bool f64_merge = training && c->dtype != L3_DTYPE_BF16;
    tw.wg64 = f64_merge ? (double*)bp.take(8LL * (9 * 512 * 512 + 512)) : nullptr;
    tw.wg64_0 = f64_merge ? (double*)bp.take(8LL * (9 * 3 * 64 + 10 * 64)) : nullptr;
    tw.g0 = training ? bp.take(es * g0_max) : nullptr;
    tw.g0b = training ? bp.take(es * g0_max) : nullptr;
    tw.g1 = training ? bp.take(es * g1_max) : nullptr;
  }
  HeadRef& h = c->head;
  h.concat = (float*)bp.take(4 * B * 1024);
  h.hidden = (float*)bp.take(4 * B * 128);
  h.probs = (float*)bp.take(4 * B * 2);
  h.logits = (float*)bp.take(4 * B * 2);
  h.dlogits = (float*)bp.take(4 * B * 2);
  h.dhidden = (float*)bp.take(4 * B * 128);
  h.dconcat = (float*)bp.take(4 * B * 1024);
  h.metrics = (float*)bp.take(256);
  c->l2_out = (double*)bp.take(256);
  c->unit_scale = (float*)bp.take(4 * 512);
  c->zero_shift = (float*)bp.take(4 * 512);
  return bp.off;
}

static float* P(l3_ctx* c, const char* nm) { int i = c->layout.find(nm); return i < 0 ? nullptr : c->params + c->layout.t[i].offset; }
static float* G(l3_ctx* c, const char* nm) { int i = c->layout.find(nm); return (i < 0 || !c->grads) ? nullptr : c->grads + c->layout.t[i].offset; }
static float* S(l3_ctx* c, const char* nm) { int i = c->layout.find(nm); return i < 0 ? nullptr : c->bn_state + c->layout.t[i].offset; }

static void bind_bn(l3_ctx* c, BnRef& bn, const std::string& pre) {
  bn.gamma = P(c, (pre + "/gamma").c_str());
  bn.beta = P(c, (pre + "/beta").c_str());
  bn.d_gamma = G(c, (pre + "/gamma").c_str());
  bn.d_beta = G(c, (pre + "/beta").c_str());
  bn.moving_mean = S(c, (pre + "/moving_mean").c_str());
  bn.moving_var = S(c, (pre + "/moving_variance").c_str());
}
static void bind_params(l3_ctx* c) {
  for (int t = 0; t < 2; ++t) {
    Tower& tw = t == 0 ? c->vision : c->audio;
    if (!tw.present) continue;
    std::string T = t == 0 ? "vision" : "audio";
    if (tw.has_bn0) bind_bn(c, tw.bn0, T + "/bn0");
    for (int l = 0; l < 8; ++l) {
      ConvLayer& L = tw.L[l];
      std::string cn = T + "/conv" + kConvNames[l];
      L.w = P(c, (cn + "/kernel").c_str());
      L.b = P(c, (cn + "/bias").c_str());
      L.dw = G(c, (cn + "/kernel").c_str());
      L.db = G(c, (cn + "/bias").c_str());
      bind_bn(c, L.bn, T + "/bn" + kConvNames[l]);
    }
  }
  HeadRef& h = c->head;
  h.w1 = P(c, "dense_1/kernel"); h.b1 = P(c, "dense_1/bias");
  h.w2 = P(c, "dense_2/kernel"); h.b2 = P(c, "dense_2/bias");
  h.dw1 = G(c, "dense_1/kernel"); h.db1 = G(c, "dense_1/bias");
  h.dw2 = G(c, "dense_2/kernel"); h.db2 = G(c, "dense_2/bias");
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
static const float kBnMomentum = 0.99f, kBnEps = 1e-3f;  // keras BatchNormalization defaults
static const int kBnUnbiasedMoving = 1;                  // TF fused batch norm feeds the Bessel-corrected variance


// want_stats: training-mode BN statistics of the output; *stats_done tells the caller they were fused
static const float kSplitWeightScale = 1024.f, kSplitWeightScaleInv = 1.f / 1024.f;   // exact powers of two

template <typename T>
static int conv_forward(l3_ctx* c, ConvLayer& L, int B, bool want_stats, bool* stats_done, cudaStream_t s, void* tw_sp_a) {
  ProfScope ps(c, PROF_CONV_FWD, s);
  *stats_done = false;
  if (L.tc && c->use_tc) {
    const bool fuse = want_stats && conv_tc_fuses_stats();
    *stats_done = fuse;
    return launch_conv3x3_tc((const bf16*)L.in, L.w_pk, L.b, (bf16*)L.z, B, L.H, L.W, L.Cin, L.Cout,
                             fuse ? L.bn.sum : nullptr, L.relu_first, s);
  }
  if (L.tcs && c->use_tc) {
    // parity mode on tensor cores: fp16 parts of the fp32 input, weights pre-scaled by 2^10 (kept inside fp16's normal
    // range), fp32 accumulate and output; the BN statistics are taken by the caller from the fp32 tensor
    const long long rows_p = (long long)B * (L.H + 2) * (L.W + 2);
    if (launch_split16((const float*)L.in, tw_sp_a, rows_p, L.Cin, 1, s)) return -1;
    return launch_conv3x3_tc_split(tw_sp_a, L.w_pk, L.b, (float*)L.z, B, L.H, L.W, L.Cin, L.Cout, 1, kSplitWeightScaleInv, s);
  }
  if (L.Cin <= 3 && L.Cout == 64 && c->use_tc && c->dtype == L3_DTYPE_BF16) {
    *stats_done = want_stats;
    return launch_first_conv_tc((const bf16*)L.in, L.w, L.b, (bf16*)L.z, B, L.H, L.W, L.Cin, L.Cout,
                                want_stats ? L.bn.sum : nullptr, s);
  }
  if (L.Cin <= 3 && L.Cout == 64) {
    *stats_done = want_stats;
    return launch_first_conv<T>((const T*)L.in, L.w, L.b, (T*)L.z, B, L.H, L.W, L.Cin, L.Cout,
                                want_stats ? L.bn.sum : nullptr, s);
  }
  return launch_conv3x3_simt<T>((const T*)L.in, L.w, L.b, (T*)L.z, B, L.H, L.W, L.Cin, L.Cout, s);
}

// input: raw (device) video or audio -> x0 -> (input BN) -> xin
template <typename T>
static int tower_input(l3_ctx* c, Tower& tw, bool is_audio, const void* src, int fmt, int B, bool training) {
  cudaStream_t s = tw.stream;
  // mode of the fused input pass (launch_input_stage): 0 = x0 is final, 1 = u8 video, 2 = dB map awaiting its reference
  int mode = 0;
  const uint8_t* u8 = nullptr;
  if (is_audio) {
    ProfScope ps(c, PROF_FRONTEND, s);
    if (launch_frontend(c->fe, src, fmt == L3_AUDIO_I16, B, tw.x0, c->clip_max, s, /*finish=*/0)) return -1;
    if (c->fe.decibel) mode = 2;
  } else {
    long long n = (long long)B * 224 * 224 * 3;
    if (fmt == L3_VIDEO_U8) {
      mode = 1;
      u8 = (const uint8_t*)src;
    } else {
      L3_CHECK_CUDA(cudaMemcpyAsync(tw.x0, src, n * 4, cudaMemcpyDeviceToDevice, s));
    }
  }
  const float *sc = nullptr, *sh = nullptr;
  if (tw.has_bn0) {
    long long rows = (long long)B * tw.H0 * tw.W0;
    if (training) {
      // batch statistics first: pass 1 produces x0 and its sums, pass 2 (below) normalises
      if (launch_input_stage<T>(mode, u8, tw.x0, (T*)nullptr, B, tw.H0, tw.W0, tw.C0, nullptr, nullptr, c->clip_max,
                                tw.bn0.sum, s))
        return -1;
      mode = 0;
    }
    if (launch_bn_finalize(tw.bn0, rows, training, kBnMomentum, kBnEps, kBnUnbiasedMoving, s)) return -1;
    sc = tw.bn0.scale;
    sh = tw.bn0.shift;
  }
  return launch_input_stage<T>(mode, u8, tw.x0, (T*)tw.xin, B, tw.H0, tw.W0, tw.C0, sc, sh, c->clip_max, nullptr, s);
}

// n_layers_act: layers [0, 7) always activate into the next input; the last conv's z is the embedding tap
template <typename T>
static int tower_forward(l3_ctx* c, Tower& tw, int B, bool training, bool embed_only) {
  cudaStream_t s = tw.stream;
  for (int l = 0; l < 8; ++l) {
    ConvLayer& L = tw.L[l];
    bool stats_done = false;
    if (l == 1 && c->pack_event_pending && s != c->stream) L3_CHECK_CUDA(cudaStreamWaitEvent(s, c->ev_pack, 0));
    if (!training && l < 7 && L.tc && c->use_tc && c->fuse_inference && !(L.pool && L.relu_first)) {
      // inference on the tensor-core path: BN uses the moving statistics, so scale / shift are known BEFORE the
      // convolution and BN + ReLU ride in its epilogue on the fp32 accumulator.  Un-pooled layers store straight into
      // the next layer's padded input (no z tensor, no activation pass); pooled layers store the activation un-padded
      // and a pool-only pass (identity coefficients; values are >= 0) writes the next input.  The Conv -> ReLU -> BN
      // -> pool layer (vision conv1b: BN output may be negative) keeps the un-fused path.
      long long rows = (long long)B * L.H * L.W;
      if (launch_bn_finalize(L.bn, rows, 0, kBnMomentum, kBnEps, kBnUnbiasedMoving, s)) return -1;
      {
        ProfScope ps(c, PROF_CONV_FWD, s);
        if (launch_conv3x3_tc_act((const bf16*)L.in, L.w_pk, L.b, (bf16*)(L.pool ? L.z : L.a), B, L.H, L.W, L.Cin, L.Cout,
                                  L.bn.scale, L.bn.shift, L.relu_first, L.pool ? 0 : 1, s))
          return -1;
      }
      if (L.pool && launch_act_fwd<T>((const T*)L.z, (T*)L.a, B, L.H, L.W, L.Cout, c->unit_scale, c->zero_shift, 1, 0, s))
        return -1;
      continue;
    }
    if (!training && l == 0 && L.Cin <= 3 && L.Cout == 64 && c->use_tc && c->dtype == L3_DTYPE_BF16 && c->fuse_inference) {
      // the same for the first layer's own kernel (Cin 1 / 3; Conv -> BN -> ReLU, not pooled)
      long long rows = (long long)B * L.H * L.W;
      if (launch_bn_finalize(L.bn, rows, 0, kBnMomentum, kBnEps, kBnUnbiasedMoving, s)) return -1;
      ProfScope ps(c, PROF_CONV_FWD, s);
      if (launch_first_conv_tc((const bf16*)L.in, L.w, L.b, (bf16*)L.a, B, L.H, L.W, L.Cin, L.Cout, nullptr, s, L.bn.scale,
                               L.bn.shift, 1))
        return -1;
      continue;
    }
    if (conv_forward<T>(c, L, B, training && !(l == 7 && embed_only), &stats_done, s, tw.sp_a)) return -1;
    if (l == 7 && embed_only) return 0;  // raw conv4b output incl. bias, before BN/ReLU (audio_model.py:482)
    long long rows = (long long)B * L.H * L.W;
    if (training && !stats_done && launch_channel_stats<T>((const T*)L.z, rows, L.Cout, L.relu_first, L.bn.sum, s)) return -1;
    if (launch_bn_finalize(L.bn, rows, training, kBnMomentum, kBnEps, kBnUnbiasedMoving, s)) return -1;
    if (l < 7) {
      if (launch_act_fwd<T>((const T*)L.z, (T*)L.a, B, L.H, L.W, L.Cout, L.bn.scale, L.bn.shift, L.pool, L.relu_first, s,
                            training ? (T*)L.zsel : nullptr, training ? L.sel : nullptr))
        return -1;
    } else {
      // MaxPooling2D over the whole final map + Flatten (audio_model.py:436-437, vision_model.py:189-190)
      if (launch_gmaxpool_fwd<T>((const T*)L.z, B, L.H * L.W, L.Cout, L.bn.scale, L.bn.shift,
                                 c->head.concat + tw.concat_off, 1024, tw.argmax, tw.pool_scratch, s))
        return -1;
    }
  }
  return 0;
}


// Backward of one tower, cut into begin / per-layer / end so that the host can interleave the two towers layer by
// layer (their kernels run concurrently on the towers' streams) and fire the data-parallel gradient buckets as soon as a
// layer's weight gradient is enqueued.
// dz(l) lives in dzbuf[l & 1]; with `defer` the weight gradient of layer l runs on the tower's side stream, ordered by
// two events per buffer: ev_dz (dz(l) complete -> wgrad(l) may read it) and ev_wg (wgrad(l) done -> the BN/ReLU
// backward of layer l-2 may overwrite the buffer)
static bool tower_defers_wgrad(l3_ctx* c, Tower& tw) {
  // (the split-operand parity mode shares one pair of operand buffers per tower: its weight gradients stay in stream order)
  return tw.wstream != nullptr && c->wgrad_streams && c->two_streams && tw.stream != c->stream && c->dtype != L3_DTYPE_F32TC;
}

template <typename T>
static int tower_backward_begin(l3_ctx* c, Tower& tw, int B) {
  cudaStream_t s = tw.stream;
  tw.wg_pending[0] = tw.wg_pending[1] = 0;
  ConvLayer& L = tw.L[7];
  return launch_gmaxpool_bwd<T>(c->head.dconcat + tw.concat_off, 1024, tw.argmax, (const T*)L.z, (T*)tw.g0b, L.bn, B, L.H, L.W,
                                L.Cout, s);
}

// data-parallel bucket of this tower's conv-kernel gradients [l_lo, l_hi] (contiguous in the arena), ordered after
// everything enqueued so far on `producer`
static int dp_fire_kernels(l3_ctx* c, Tower& tw, int l_lo, int l_hi, cudaStream_t producer) {
  if (!c->dp) return 0;
  long long n = 0;
  for (int l = l_lo; l <= l_hi; ++l) n += 9LL * tw.L[l].Cin * tw.L[l].Cout;
  void* p = tw.L[l_lo].dw;
  return dp_allreduce_ranges(c->dp, producer, &p, &n, 1, 0);
}

template <typename T>
static int tower_backward_layer(l3_ctx* c, Tower& tw, int B, int l) {
  cudaStream_t s = tw.stream;
  T* dzbuf[2] = {(T*)tw.g0, (T*)tw.g0b};
  const bool defer = tower_defers_wgrad(c, tw);
  cudaStream_t ws = defer ? tw.wstream : s;
  T* da = (T*)tw.g1;
  ConvLayer& L = tw.L[l];
  T* dz = dzbuf[l & 1];
  long long rows = (long long)B * L.H * L.W;
  if (l == 7) {
    // dz holds the scattered dy of the global max-pool and bn.sum its sums: finish BN backward in place
    if (launch_bn_bwd_finalize(L.bn, rows, 0, s)) return -1;
    if (launch_bn_bwd_apply<T>(dz, (const T*)L.z, B, L.H, L.W, L.Cout, L.bn, L.relu_first, s)) return -1;
  }
  // weight / bias gradient (layer 0 stays on the main stream: the input-BN gradient is derived from its result)
  {
    const bool side = defer && l > 0;
    cudaStream_t sw = side ? ws : s;
    if (side) {
      L3_CHECK_CUDA(cudaEventRecord(tw.ev_dz[l & 1], s));
      L3_CHECK_CUDA(cudaStreamWaitEvent(ws, tw.ev_dz[l & 1], 0));
    }
    {
      ProfScope ps(c, PROF_CONV_WGRAD, sw);
      if (L.tcs && c->use_tc) {
        // parity mode on tensor cores: bf16 parts of both operands (gradients need fp32's range), all four part products
        const long long rows_p = (long long)B * (L.H + 2) * (L.W + 2);
        if (launch_split16((const float*)dz, tw.sp_z, rows_p, L.Cout, 0, sw)) return -1;   // also the dgrad operand below
        if (launch_split16((const float*)L.in, tw.sp_a, rows_p, L.Cin, 0, sw)) return -1;
        if (launch_wgrad3x3_tc_split(tw.sp_a, tw.sp_z, L.dw, tw.dw4, B, L.H, L.W, L.Cin, L.Cout, sw)) return -1;
        if (L.relu_first) {   // the one Conv -> ReLU -> BN layer: its bias gradient is not identically zero
          if (launch_channel_stats<float>((const float*)dz, rows_p, L.Cout, 0, tw.wg64, sw)) return -1;
          if (launch_f64_to_f32(tw.wg64, L.db, L.Cout, sw)) return -1;
        }
      } else if (L.tc && c->use_tc) {
        // Conv -> BN layers: sum_pixels(dz) == 0 identically (BN backward removes the mean), so the bias gradient
        // stays at the zero the grads arena was cleared to; only the ReLU-before-BN layer needs the reduction.
        // (that reduction is HBM-bound: on the tower's own stream, not queued behind the side stream's weight gradients,
        // where it used to finish last of the whole step)
        if (launch_wgrad3x3_tc((const bf16*)L.in, (const bf16*)dz, L.dw, nullptr, B, L.H, L.W, L.Cin, L.Cout, sw)) return -1;
        if (L.relu_first && launch_bias_grad_tc((const bf16*)dz, L.db, B, L.H, L.W, L.Cout, s)) return -1;
      } else if (l == 0 && c->use_tc && c->dtype == L3_DTYPE_BF16 && L.Cout == 64) {
        // (the first layer is Conv -> BN in every model type: its bias gradient is identically zero as well -- summing
        // the stored dz would only add up its rounding errors)
        if (launch_first_wgrad_tc((const bf16*)L.in, (const bf16*)dz, L.dw, nullptr, tw.has_bn0 ? tw.d1 : nullptr, B, L.H, L.W,
                                  L.Cin, L.Cout, sw))
          return -1;
      } else if (l == 0) {
        // parity mode derives the input-BN gradient directly (below), not from d1
        if (launch_first_wgrad<T>((const T*)L.in, (const T*)dz, L.dw, nullptr, (tw.has_bn0 && sizeof(T) != 4) ? tw.d1 : nullptr, B,
                                  L.H, L.W, L.Cin, L.Cout, sw, tw.wg64_0))
          return -1;
      } else {
        if (launch_wgrad3x3_simt<T>((const T*)L.in, (const T*)dz, L.dw, L.db, B, L.H, L.W, L.Cin, L.Cout, sw, tw.wg64)) return -1;
      }
    }
    if (side) {
      L3_CHECK_CUDA(cudaEventRecord(tw.ev_wg[l & 1], ws));
      tw.wg_pending[l & 1] = 1;
    }
    // gradient buckets of the big layers leave as soon as their weight gradient is enqueued: conv4b, conv4a, conv3a+3b
    // hold 74 % + 18 % of the bytes and are complete after ~27 % / ~45 % of the backward FLOPs, conv2a + 2b
    // follow; only conv1a + 1b of each tower (0.15 MB) wait for the final grouped launch (dp_fire_tail), whose latency
    // is the one part of the exchange the backward pass cannot hide.  Layers 7..1 share the stream `sw`.
    if (l == 7 || l == 6) { if (dp_fire_kernels(c, tw, l, l, sw)) return -1; }
    else if (l == 4) { if (dp_fire_kernels(c, tw, 4, 5, sw)) return -1; }
    else if (l == 2) { if (dp_fire_kernels(c, tw, 2, 3, sw)) return -1; }
  }
  if (l == 0) {
    if (tw.has_bn0) {
      // input BN: d_gamma / d_beta need only sum(da), sum(da*xhat) -- computed without storing da
      ProfScope ps(c, PROF_CONV_DGRAD, s);
      if (sizeof(T) == 4) {
        // parity mode: sum(da), sum(da*xhat) straight from dz with fp64 accumulation.  (The algebraic route below
        // forms them as a residual of cancelling sums of the weight gradient: fine at bf16 accuracy, but it amplified
        // the weight gradient's rounding noise past the 1e-2 parity bar.)
        if (launch_first_dgrad_bnstats<T>((const T*)dz, L.w, tw.x0, tw.bn0, B, L.H, L.W, L.Cin, L.Cout, nullptr, s)) return -1;
      } else {
        int* fallback = reinterpret_cast<int*>(tw.d1 + 9 * 64);
        if (launch_bn0_from_dw(L.w, L.dw, tw.d1, tw.bn0, L.Cin, fallback, s)) return -1;
        if (launch_first_dgrad_bnstats<T>((const T*)dz, L.w, tw.x0, tw.bn0, B, L.H, L.W, L.Cin, L.Cout, fallback, s)) return -1;
      }
      if (launch_bn_bwd_finalize(tw.bn0, rows, 0, s)) return -1;
    }
    return 0;
  }
  // data gradient: da = conv(dz, flip/transpose(w))
  ConvLayer& Lp = tw.L[l - 1];
  // Pass 1 of the BN/ReLU backward of the layer below (sum dy, sum dy*z) rides in the data gradient's epilogue where that
  // pays: measured at B = 64 (tools/run_op.py dgrad vs dgrad_stats) the fused launch costs +14 / +19 / +55 us at 512 / 256 /
  // 128 channels against 38 / 80 / 110 us for the stand-alone pass it replaces (which re-reads da and z from HBM); at 64
  // channels (N = 64 tiles, shared-memory bound) it is neutral, and pooled layers take their sums from the recorded routing.
  const bool fuse_stats = L.tc && c->use_tc && conv_tc_fuses_bwd_stats(Lp.Cout) && !Lp.pool && !Lp.relu_first;
  {
    ProfScope ps(c, PROF_CONV_DGRAD, s);
    if (L.tcs && c->use_tc) {
      // sp_z holds the bf16 parts of dz (split for the weight gradient above, same stream)
      if (launch_conv3x3_tc_split(tw.sp_z, L.wt_pk, nullptr, (float*)da, B, L.H, L.W, L.Cout, L.Cin, 0, 1.f, s)) return -1;
    } else if (fuse_stats) {
      if (launch_dgrad3x3_tc_bwdstats((const bf16*)dz, L.wt_pk, (bf16*)da, B, L.H, L.W, L.Cout, L.Cin, (const bf16*)Lp.z,
                                      Lp.bn.scale, Lp.bn.shift, Lp.bn.sum, s))
        return -1;
    } else if (L.tc && c->use_tc) {
      if (launch_conv3x3_tc((const bf16*)dz, L.wt_pk, nullptr, (bf16*)da, B, L.H, L.W, L.Cout, L.Cin, nullptr, 0, s)) return -1;
    } else {
      if (launch_flip_transpose(L.w, L.w_t, L.Cin, L.Cout, s)) return -1;
      if (launch_conv3x3_simt<T>((const T*)dz, L.w_t, nullptr, da, B, L.H, L.W, L.Cout, L.Cin, s)) return -1;
    }
  }
  // activation + BN backward of layer l-1: da (at its pooled resolution) -> dz(l-1) (padded, full resolution)
  long long rows_p = (long long)B * Lp.H * Lp.W;
  if (!fuse_stats && launch_bwd_stats<T>(da, (const T*)Lp.z, B, Lp.H, Lp.W, Lp.Cout, Lp.bn, Lp.pool, Lp.relu_first, s,
                                         (const T*)Lp.zsel, Lp.sel))
    return -1;
  if (launch_bn_bwd_finalize(Lp.bn, rows_p, sizeof(T) == 4 ? 2 : 1, s)) return -1;
  // dz(l-1) goes into the buffer the weight gradient of layer l+1 read
  const int nb = (l - 1) & 1;
  if (tw.wg_pending[nb]) {
    L3_CHECK_CUDA(cudaStreamWaitEvent(s, tw.ev_wg[nb], 0));
    tw.wg_pending[nb] = 0;
  }
  return launch_bwd_apply<T>(da, (const T*)Lp.z, dzbuf[nb], B, Lp.H, Lp.W, Lp.Cout, Lp.bn, Lp.pool, Lp.relu_first, s, Lp.sel);
}

// the tower is complete only when its side stream is
static int tower_backward_end(l3_ctx* c, Tower& tw) {
  for (int b2 = 0; b2 < 2; ++b2)
    if (tw.wg_pending[b2]) {
      L3_CHECK_CUDA(cudaStreamWaitEvent(tw.stream, tw.ev_wg[b2], 0));
      tw.wg_pending[b2] = 0;
    }
  return 0;
}

// both towers, interleaved layer by layer on the host (on the device they overlap on their own streams)
template <typename T>
static int towers_backward(l3_ctx* c, int B) {
  Tower* tws[2] = {&c->vision, &c->audio};
  for (Tower* tw : tws)
    if (tower_backward_begin<T>(c, *tw, B)) return -1;
  for (int l = 7; l >= 0; --l)
    for (Tower* tw : tws)
      if (tower_backward_layer<T>(c, *tw, B, l)) return -1;
  for (Tower* tw : tws)
    if (tower_backward_end(c, *tw)) return -1;
  return 0;
}

// loss sum and #correct of the GLOBAL batch: the two scalars are copied aside (head.metrics[8..9]) and summed over the
// ranks right after the forward pass; l3_get_metrics reports them when data parallelism is on
static int dp_fire_metrics(l3_ctx* c) {
  float* g = c->head.metrics + 8;
  L3_CHECK_CUDA(cudaMemcpyAsync(g, c->head.metrics, 8, cudaMemcpyDeviceToDevice, c->stream));
  void* p = g;
  long long n = 2;
  return dp_allreduce_ranges(c->dp, c->stream, &p, &n, 1, 0);
}

// The dense kernels' gradients are complete as soon as the head's backward is (before the towers' backward starts)
static int dp_fire_dense(l3_ctx* c) {
  void* p = c->head.dw1;                                  // dense_1 and dense_2 kernels are adjacent (build_layout)
  long long n = 1024LL * 128 + 128 * 2;
  return dp_allreduce_ranges(c->dp, c->stream, &p, &n, 1, 0);
}

// The rest of the gradient arena in ONE grouped launch once both towers have joined the context stream: conv1a + conv1b
// of each tower and everything that is not a kernel (biases, BN gamma / beta) -- 0.35 MB of 38 MB.
static int dp_fire_tail(l3_ctx* c) {
  void* ptrs[3];
  long long counts[3];
  int n = 0;
  for (Tower* tw : {&c->vision, &c->audio}) {
    long long k = 0;
    for (int l = 0; l <= 1; ++l) k += 9LL * tw->L[l].Cin * tw->L[l].Cout;
    ptrs[n] = tw->L[0].dw;
    counts[n++] = k;
  }
  ptrs[n] = c->grads + c->layout.n_l2;
  counts[n++] = c->layout.n_params - c->layout.n_l2;
  return dp_allreduce_ranges(c->dp, c->stream, ptrs, counts, n, 0);
}

// bf16 tensor-core operand packs of every layer (forward, and dgrad when training), one launch
static int pack_all_weights(l3_ctx* c, bool vision, bool audio, bool with_dgrad, cudaStream_t stream = nullptr) {
  if (!c->use_tc) return 0;
  PackBatch pb;
  pb.n = 0;
  for (int t = 0; t < 2; ++t) {
    Tower& tw = t == 0 ? c->vision : c->audio;
    if (!tw.present || !(t == 0 ? vision : audio)) continue;
    for (int l = 0; l < 8; ++l) {
      ConvLayer& L = tw.L[l];
      if (L.tcs) {
        PackJob f{L.w, L.w_pk, L.Cin, L.Cout, 0};
        f.split = 1; f.fp16 = 1; f.scale = kSplitWeightScale;
        pb.job[pb.n++] = f;
        if (with_dgrad && L.wt_pk) {
          PackJob d{L.w, L.wt_pk, L.Cin, L.Cout, 1};
          d.split = 1;
          pb.job[pb.n++] = d;
        }
        continue;
      }
      if (!L.tc) continue;
      pb.job[pb.n++] = PackJob{L.w, L.w_pk, L.Cin, L.Cout, 0};
      if (with_dgrad && L.wt_pk) pb.job[pb.n++] = PackJob{L.w, L.wt_pk, L.Cin, L.Cout, 1};
    }
  }
  return launch_pack_weights_batch(pb, stream ? stream : c->stream);
}

// towers on their own streams between fork() and join(); everything else on the caller's stream
static int fork_streams(l3_ctx* c) {
  c->vision.stream = (c->two_streams && c->stream_v) ? c->stream_v : c->stream;
  c->audio.stream = c->two_streams ? c->stream2 : c->stream;
  if (!c->two_streams) return 0;
  L3_CHECK_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  L3_CHECK_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  if (c->stream_v) L3_CHECK_CUDA(cudaStreamWaitEvent(c->stream_v, c->ev_fork, 0));
  return 0;
}
static int join_streams(l3_ctx* c) {
  if (!c->two_streams) return 0;
  L3_CHECK_CUDA(cudaEventRecord(c->ev_join, c->stream2));
  L3_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  if (c->stream_v) {
    L3_CHECK_CUDA(cudaEventRecord(c->ev_join_v, c->stream_v));
    L3_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join_v, 0));
  }
  return 0;
}

template <typename T>
static int forward_all(l3_ctx* c, const void* video, int vfmt, const void* audio, int afmt, const float* labels, int B,
                       bool training, float grad_scale) {
  if (fork_streams(c)) return -1;
  // The bf16 operand packs are first needed by the SECOND conv layer (the first-layer kernels read the fp32 weights
  // themselves): they are built on the context stream, beside the towers' input stages, and each tower waits for them
  // right before its first tensor-core layer (tower_forward).  On the vision stream they used to delay the whole tower.
  c->pack_event_pending = 0;
  if (c->vision.stream != c->stream) {
    if (pack_all_weights(c, true, true, training, c->stream)) return -1;
    L3_CHECK_CUDA(cudaEventRecord(c->ev_pack, c->stream));
    c->pack_event_pending = 1;
  } else if (pack_all_weights(c, true, true, training, c->stream)) {
    return -1;
  }
  // interleave the two towers' launches so neither stream starves while the host is still enqueuing the other
  if (tower_input<T>(c, c->vision, false, video, vfmt, B, training)) return -1;
  if (tower_input<T>(c, c->audio, true, audio, afmt, B, training)) return -1;
  if (tower_forward<T>(c, c->vision, B, training, false)) return -1;
  if (tower_forward<T>(c, c->audio, B, training, false)) return -1;
  if (join_streams(c)) return -1;
  return launch_head_fwd(c->head, labels, B, grad_scale, c->stream);
}

static int check_batch(l3_ctx* c, int batch) {
  L3_REQUIRE(c != nullptr, "null ctx");
  L3_REQUIRE(batch >= 1 && batch <= c->max_batch, "batch %d outside [1, %d]", batch, c->max_batch);
  return 0;
}

// ---- activation peek for parity bisecting -----------------------------------------------------------------
__device__ __forceinline__ float to_f(uint8_t x) { return (float)x; }
__device__ __forceinline__ float to_f(int x) { return (float)x; }
template <typename T>
__global__ void k_gather_unpad(const T* __restrict__ src, float* __restrict__ dst, long long n, int H, int W, int C,
                               int padded) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!padded) {
      dst[i] = to_f(src[i]);
    } else {
      int ch = (int)(i % C);
      long long p = i / C;
      int x = (int)(p % W);
      int y = (int)((p / W) % H);
      long long b = p / ((long long)W * H);
      dst[i] = to_f(src[pad_off(b, y, x, H, W, C) + ch]);
    }
  }
}

}  // namespace l3

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" {

int l3_version(void) { return L3_VERSION; }
const char* l3_last_error(void) { return g_err; }

static int valid_model(int m) {
  if (m < 0 || m > 3) {
    set_error("invalid model type %d", m);
    return 0;
  }
  return 1;
}
int64_t l3_param_count(int m) { return valid_model(m) ? build_layout(m).n_params : -1; }
int64_t l3_l2_count(int m) { return valid_model(m) ? build_layout(m).n_l2 : -1; }
int64_t l3_state_count(int m) { return valid_model(m) ? build_layout(m).n_state : -1; }
int l3_num_tensors(int m) { return valid_model(m) ? (int)build_layout(m).t.size() : -1; }
int l3_tensor_info(int m, int i, char* name, int name_cap, int* arena, int64_t* offset, int* ndim, int64_t dims[4]) {
  if (!valid_model(m)) return -2;
  Layout L = build_layout(m);
  L3_REQUIRE(i >= 0 && i < (int)L.t.size(), "tensor index %d out of range", i);
  const TensorInfo& t = L.t[i];
  if (name && name_cap > 0) {
    strncpy(name, t.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (arena) *arena = t.arena;
  if (offset) *offset = t.offset;
  if (ndim) *ndim = t.ndim;
  if (dims)
    for (int k = 0; k < 4; ++k) dims[k] = t.dims[k];
  return 0;
}
int l3_frontend_shape(int m, int* n_out, int* n_frames) {
  if (!valid_model(m)) return -2;
  int a, b, lp;
  audio_geometry(kSpecs[m], &a, &b, &lp);
  if (n_out) *n_out = a;
  if (n_frames) *n_frames = b;
  return 0;
}
int l3_embedding_map_shape(int m, int* h, int* w) {
  if (!valid_model(m)) return -2;
  int H, W, lp;
  audio_geometry(kSpecs[m], &H, &W, &lp);
  for (int i = 0; i < 3; ++i) { H /= 2; W /= 2; }
  if (h) *h = H;
  if (w) *w = W;
  return 0;
}

int64_t l3_workspace_bytes(int model_type, int max_batch, int dtype, int flags) {
  if (!valid_model(model_type)) return -2;
  if (max_batch < 1 || (dtype != L3_DTYPE_F32 && dtype != L3_DTYPE_BF16 && dtype != L3_DTYPE_F32TC)) {
    set_error("bad max_batch %d / dtype %d", max_batch, dtype);
    return -2;
  }
  l3_ctx tmp{};
  tmp.model_type = model_type;
  tmp.max_batch = max_batch;
  tmp.dtype = dtype;
  tmp.flags = flags;
  tmp.spec = kSpecs[model_type];
  tmp.ws = nullptr;
  return carve(&tmp);
}

// the towers' own streams (highest priority: the caller's stream has the default, lowest one) and the low-priority side
// streams of the weight gradients, with the events that order them
static int create_streams(l3_ctx* c) {
  int least = 0, greatest = 0;
  L3_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  L3_CHECK_CUDA(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, greatest));
  L3_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  L3_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  L3_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming));
  if (c->wgrad_streams) {
    L3_CHECK_CUDA(cudaStreamCreateWithPriority(&c->stream_v, cudaStreamNonBlocking, greatest));
    L3_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_join_v, cudaEventDisableTiming));
    for (Tower* tw : {&c->vision, &c->audio}) {
      L3_CHECK_CUDA(cudaStreamCreateWithPriority(&tw->wstream, cudaStreamNonBlocking, least));
      for (int i = 0; i < 2; ++i) {
        L3_CHECK_CUDA(cudaEventCreateWithFlags(&tw->ev_dz[i], cudaEventDisableTiming));
        L3_CHECK_CUDA(cudaEventCreateWithFlags(&tw->ev_wg[i], cudaEventDisableTiming));
      }
    }
  }
  return 0;
}

l3_ctx* l3_ctx_create(int model_type, int max_batch, int dtype, int flags, float* params, float* grads, float* adam_m,
                      float* adam_v, float* bn_state, void* workspace, int64_t workspace_bytes, void* stream) {
  int64_t need = l3_workspace_bytes(model_type, max_batch, dtype, flags);
  if (need < 0) return nullptr;
  if (!params || !bn_state || !workspace) {
    set_error("params, bn_state and workspace must be non-null");
    return nullptr;
  }
  if ((flags & L3_WS_TRAINING) && (!grads || !adam_m || !adam_v)) {
    set_error("training context needs grads / adam_m / adam_v arenas");
    return nullptr;
  }
  if (workspace_bytes < need) {
    set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
    return nullptr;
  }
  if (((uintptr_t)workspace & 255) != 0) {
    set_error("workspace must be 256-byte aligned");
    return nullptr;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: libl3b200 has no CPU fallback");
    return nullptr;
  }
  l3_ctx* c = new l3_ctx();
  c->model_type = model_type;
  c->max_batch = max_batch;
  c->dtype = dtype;
  c->flags = flags;
  c->spec = kSpecs[model_type];
  c->layout = build_layout(model_type);
  c->params = params; c->grads = grads; c->adam_m = adam_m; c->adam_v = adam_v; c->bn_state = bn_state;
  c->ws = (char*)workspace;
  c->ws_bytes = workspace_bytes;
  c->stream = (cudaStream_t)stream;
  c->adam_t = 0;
  c->use_tc = (dtype != L3_DTYPE_F32) && conv_tc_supported();
  if (dtype == L3_DTYPE_F32TC && !c->use_tc) {
    set_error("L3_DTYPE_F32TC needs the tcgen05 path (sm_100)");
    delete c;
    return nullptr;
  }
  c->fuse_inference = 1;
  c->pack_event_pending = 0;
  c->device = 0;
  cudaGetDevice(&c->device);
  c->st_head = c->st_tail = c->st_count = 0;
  c->st_batch[0] = c->st_batch[1] = 0;
  c->st_consumed_valid[0] = c->st_consumed_valid[1] = false;
  c->copy_stream = nullptr;
  c->metrics_host = nullptr;
  c->dp = nullptr;
  c->last_global_batch = 0;
  if (cudaMallocHost((void**)&c->metrics_host, 64) != cudaSuccess) {
    set_error("cudaMallocHost failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  if (flags & L3_WS_HOST_STAGING) {
    bool ok = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
      ok = cudaEventCreateWithFlags(&c->ev_staged[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      set_error("staging stream / events: %s", cudaGetErrorString(cudaGetLastError()));
      delete c;
      return nullptr;
    }
  }
  c->last_batch = 0;
  c->prof_on = 0;
  c->prof_used = 0;
  c->vision.stream = c->audio.stream = c->stream;
  c->stream2 = nullptr;
  c->stream_v = nullptr;
  c->vision.wstream = c->audio.wstream = nullptr;
  {
    c->two_streams = 1;     // l3_ctx_set_two_streams(ctx, 0) serialises the towers (per-kernel timing)
    c->wgrad_streams = 1;
    if (c->two_streams && create_streams(c)) {
      delete c;
      return nullptr;
    }
  }
  carve(c);
  bind_params(c);
  // zero the workspace once: the halos of every fixed-geometry padded activation buffer stay zero forever
  if (cudaMemsetAsync(workspace, 0, need, c->stream) != cudaSuccess) {
    set_error("workspace memset failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  {
    std::vector<float> ones(512, 1.0f);
    if (cudaMemcpyAsync(c->unit_scale, ones.data(), 4 * 512, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
      set_error("workspace init failed: %s", cudaGetErrorString(cudaGetLastError()));
      delete c;
      return nullptr;
    }
  }
  int n_out, n_frames, left;
  audio_geometry(c->spec, &n_out, &n_frames, &left);
  FrontendPlan& fe = c->fe;
  fe.n_dft = c->spec.n_dft; fe.n_hop = kHop; fe.n_frames = n_frames; fe.left_pad = left; fe.n_out = n_out;
  fe.mel = c->spec.mel; fe.decibel = c->spec.decibel; fe.n_samples = kSR; fe.clip_stride = kSR;
  if (frontend_build_tables(&fe, kSR, c->spec.n_mels, c->fe_tables, c->stream)) {
    delete c;
    return nullptr;
  }
  return c;
}

void l3_ctx_destroy(l3_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_staged[i]); cudaEventDestroy(ctx->ev_consumed[i]); }
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->metrics_host) cudaFreeHost(ctx->metrics_host);
  if (ctx->dp) dp_destroy(ctx->dp);
  if (ctx->stream2) {
    cudaStreamSynchronize(ctx->stream2);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    cudaEventDestroy(ctx->ev_pack);
    cudaStreamDestroy(ctx->stream2);
  }
  if (ctx->stream_v) {
    cudaStreamSynchronize(ctx->stream_v);
    cudaEventDestroy(ctx->ev_join_v);
    cudaStreamDestroy(ctx->stream_v);
  }
  for (Tower* tw : {&ctx->vision, &ctx->audio})
    if (tw->wstream) {
      cudaStreamSynchronize(tw->wstream);
      for (int i = 0; i < 2; ++i) { cudaEventDestroy(tw->ev_dz[i]); cudaEventDestroy(tw->ev_wg[i]); }
      cudaStreamDestroy(tw->wstream);
    }
  delete ctx;
}

uint64_t l3_launch_count(void) { return g_launch_count; }

int l3_ctx_set_two_streams(l3_ctx* c, int enable) {
  L3_REQUIRE(c != nullptr, "null ctx");
  if (enable && !c->stream2 && create_streams(c)) return -1;
  L3_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  c->two_streams = enable ? 1 : 0;
  return 0;
}

int l3_ctx_profile_enable(l3_ctx* c, int enable) {
  L3_REQUIRE(c != nullptr, "null ctx");
  c->prof_on = enable ? 1 : 0;
  c->prof_used = 0;
  return 0;
}
int l3_ctx_profile_read(l3_ctx* c, float ms_out[4], int launches_out[4]) {
  L3_REQUIRE(c != nullptr && ms_out != nullptr, "null argument");
  L3_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < PROF_N; ++i) {
    ms_out[i] = 0.f;
    if (launches_out) launches_out[i] = 0;
  }
  for (size_t i = 0; i < c->prof_used; ++i) {
    float ms = 0.f;
    L3_CHECK_CUDA(cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
    ms_out[c->prof_cls[i]] += ms;
    if (launches_out) launches_out[c->prof_cls[i]]++;
  }
  c->prof_used = 0;
  return 0;
}

int l3_ctx_set_use_tensor_cores(l3_ctx* c, int enable) {
  L3_REQUIRE(c != nullptr, "null ctx");
  L3_REQUIRE(!enable || (c->dtype != L3_DTYPE_F32 && conv_tc_supported()), "tensor-core path needs a bf16 / f32tc context on sm_100");
  L3_REQUIRE(enable || c->dtype != L3_DTYPE_F32TC, "an f32tc context has no SIMT fallback for its split layers: use L3_DTYPE_F32");
  c->use_tc = enable ? 1 : 0;
  return 0;
}
int l3_ctx_uses_tensor_cores(l3_ctx* c) { return c ? c->use_tc : 0; }
int l3_ctx_set_fused_inference(l3_ctx* c, int enable) {
  L3_REQUIRE(c != nullptr, "null ctx");
  c->fuse_inference = enable ? 1 : 0;
  return 0;
}

int l3_upload_batch_host(l3_ctx* c, const void* video_host, int video_fmt, const void* audio_host, int audio_fmt,
                         const float* labels_host, int batch) {
  if (check_batch(c, batch)) return -2;
  L3_REQUIRE(c->flags & L3_WS_HOST_STAGING, "context was created without L3_WS_HOST_STAGING");
  L3_REQUIRE(video_host && audio_host, "video and audio host buffers are required");
  L3_CHECK_CUDA(cudaSetDevice(c->device));   // may be called from a prefetch thread
  int slot;
  {
    std::lock_guard<std::mutex> lk(c->st_mu);
    L3_REQUIRE(c->st_count < 2, "both staging slots hold batches that no step has consumed yet");
    slot = c->st_head;
  }
  cudaStream_t cs = c->copy_stream;
  if (c->st_consumed_valid[slot]) L3_CHECK_CUDA(cudaStreamWaitEvent(cs, c->ev_consumed[slot], 0));
  {
    size_t n = (size_t)batch * 224 * 224 * 3 * (video_fmt == L3_VIDEO_U8 ? 1 : 4);
    L3_CHECK_CUDA(cudaMemcpyAsync(c->st_video[slot], video_host, n, cudaMemcpyHostToDevice, cs));
  }
  {
    size_t n = (size_t)batch * kSR * (audio_fmt == L3_AUDIO_I16 ? 2 : 4);
    L3_CHECK_CUDA(cudaMemcpyAsync(c->st_audio[slot], audio_host, n, cudaMemcpyHostToDevice, cs));
  }
  if (labels_host) L3_CHECK_CUDA(cudaMemcpyAsync(c->st_labels[slot], labels_host, (size_t)batch * 8, cudaMemcpyHostToDevice, cs));
  L3_CHECK_CUDA(cudaEventRecord(c->ev_staged[slot], cs));
  std::lock_guard<std::mutex> lk(c->st_mu);
  c->st_video_fmt[slot] = video_fmt;
  c->st_audio_fmt[slot] = audio_fmt;
  c->st_batch[slot] = batch;
  c->st_head ^= 1;
  c->st_count++;
  return 0;
}

// NULL inputs = the oldest staged batch: *slot receives its index (the caller releases it with release_slot once the
// kernels reading it are enqueued); explicit device pointers leave *slot = -1
static int resolve_inputs(l3_ctx* c, const void*& video, int& vfmt, const void*& audio, int& afmt, const float*& labels,
                          int batch, bool need_labels, int* slot) {
  *slot = -1;
  if (!video && !audio) {
    std::lock_guard<std::mutex> lk(c->st_mu);
    L3_REQUIRE(c->st_count > 0, "no staged batch: call l3_upload_batch_host first");
    const int sl = c->st_tail;
    L3_REQUIRE(c->st_batch[sl] == batch, "the staged batch has %d samples, not %d", c->st_batch[sl], batch);
    L3_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_staged[sl], 0));
    video = c->st_video[sl]; vfmt = c->st_video_fmt[sl];
    audio = c->st_audio[sl]; afmt = c->st_audio_fmt[sl];
    if (!labels) labels = c->st_labels[sl];
    *slot = sl;
  }
  L3_REQUIRE(video && audio, "video and audio must both be given (or both NULL for the staged batch)");
  L3_REQUIRE(!need_labels || labels, "labels required");
  L3_REQUIRE(c->vision.present && c->audio.present, "context lacks a tower (L3_WS_VISION | L3_WS_AUDIO)");
  return 0;
}
// the forward pass (both towers joined, head launched) has been enqueued on c->stream: everything that reads the slot
// precedes this point in stream order
static int release_slot(l3_ctx* c, int slot) {
  if (slot < 0) return 0;
  L3_CHECK_CUDA(cudaEventRecord(c->ev_consumed[slot], c->stream));
  std::lock_guard<std::mutex> lk(c->st_mu);
  c->st_consumed_valid[slot] = true;
  c->st_tail ^= 1;
  c->st_count--;
  return 0;
}

int l3_forward_backward(l3_ctx* c, const void* video, int video_fmt, const void* audio, int audio_fmt,
                        const float* labels, int batch, int global_batch) {
  if (check_batch(c, batch)) return -2;
  L3_REQUIRE(c->flags & L3_WS_TRAINING, "context was created without L3_WS_TRAINING");
  L3_REQUIRE(global_batch >= batch, "global_batch %d < batch %d", global_batch, batch);
  L3_CHECK_CUDA(cudaSetDevice(c->device));
  int slot;
  if (resolve_inputs(c, video, video_fmt, audio, audio_fmt, labels, batch, true, &slot)) return -2;
  const float gs = 1.0f / (float)global_batch;
  if (c->dp) {
    // the previous step's collectives (and its Adam update, which waited for them) precede this memset in stream order
    // only if the caller ran l3_adam_step; make the dependency explicit either way
    if (dp_join(c->dp, c->stream)) return -1;
    dp_begin_step(c->dp);
  }
  L3_CHECK_CUDA(cudaMemsetAsync(c->grads, 0, sizeof(float) * c->layout.n_params, c->stream));
  int rc;
  if (c->dtype == L3_DTYPE_BF16) {
    rc = forward_all<bf16>(c, video, video_fmt, audio, audio_fmt, labels, batch, true, gs);
    if (!rc) rc = release_slot(c, slot);
    if (!rc && c->dp) rc = dp_fire_metrics(c);
    if (!rc) rc = launch_head_bwd(c->head, batch, c->stream);
    if (!rc && c->dp) rc = dp_fire_dense(c);
    if (!rc) rc = fork_streams(c);
    if (!rc) rc = towers_backward<bf16>(c, batch);
    if (!rc) rc = join_streams(c);
  } else {
    rc = forward_all<float>(c, video, video_fmt, audio, audio_fmt, labels, batch, true, gs);
    if (!rc) rc = release_slot(c, slot);
    if (!rc && c->dp) rc = dp_fire_metrics(c);
    if (!rc) rc = launch_head_bwd(c->head, batch, c->stream);
    if (!rc && c->dp) rc = dp_fire_dense(c);
    if (!rc) rc = fork_streams(c);
    if (!rc) rc = towers_backward<float>(c, batch);
    if (!rc) rc = join_streams(c);
  }
  if (!rc && c->dp) rc = dp_fire_tail(c);
  c->last_batch = batch;
  c->last_global_batch = global_batch;
  return rc;
}

int l3_adam_step(l3_ctx* c, float lr) {
  L3_REQUIRE(c != nullptr, "null ctx");
  L3_REQUIRE(c->flags & L3_WS_TRAINING, "context was created without L3_WS_TRAINING");
  L3_CHECK_CUDA(cudaSetDevice(c->device));
  if (c->dp && dp_join(c->dp, c->stream)) return -1;   // the summed gradients must have landed
  // keras 2.0.9 Adam.get_updates: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); p -= lr_t * m / (sqrt(v) + eps)
  c->adam_t += 1;
  const double b1 = 0.9, b2 = 0.999;
  const double t = (double)c->adam_t;
  const double lr_t = (double)lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t));
  return launch_adam(c->params, c->grads, c->adam_m, c->adam_v, c->layout.n_params, c->layout.n_l2, (float)lr_t,
                     (float)b1, (float)b2, 1e-8f, 1e-5f, c->stream);
}
int l3_adam_set_t(l3_ctx* c, int64_t t) {
  L3_REQUIRE(c != nullptr && t >= 0, "bad adam step");
  c->adam_t = t;
  return 0;
}

int64_t l3_adam_get_t(l3_ctx* c) {
  L3_REQUIRE(c != nullptr, "null ctx");
  return c->adam_t;
}

// l2 penalty kernel + asynchronous read-back of {ce sum, #correct, l2} into the pinned metrics buffer
static int enqueue_metrics(l3_ctx* c) {
  if (launch_l2_penalty(c->params, c->layout.n_l2, c->l2_out, c->stream)) return -1;
  const float* src = c->head.metrics;
  if (c->dp && c->last_global_batch > 0) {
    if (dp_join(c->dp, c->stream)) return -1;
    src = c->head.metrics + 8;     // sums over the global batch (dp_fire_metrics)
  }
  L3_CHECK_CUDA(cudaMemcpyAsync(c->metrics_host, src, 8, cudaMemcpyDeviceToHost, c->stream));
  L3_CHECK_CUDA(cudaMemcpyAsync(c->metrics_host + 2, c->l2_out, 8, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}
static void read_metrics(l3_ctx* c, float out[4]) {
  double l2;
  memcpy(&l2, c->metrics_host + 2, 8);
  out[0] = c->metrics_host[0];
  out[1] = c->metrics_host[1];
  out[2] = (float)(1e-5 * l2);
  out[3] = (float)((c->dp && c->last_global_batch > 0) ? c->last_global_batch : c->last_batch);
}

int l3_get_metrics(l3_ctx* c, float out[4]) {
  L3_REQUIRE(c != nullptr && out != nullptr, "null argument");
  L3_CHECK_CUDA(cudaSetDevice(c->device));
  if (enqueue_metrics(c)) return -1;
  L3_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  read_metrics(c, out);
  return 0;
}

int l3_train_step_staged(l3_ctx* c, int batch, float lr, float out_metrics[4]) {
  int rc = l3_forward_backward(c, nullptr, 0, nullptr, 0, nullptr, batch, batch);
  // loss terms of the weights the batch was run with (as keras reports them): read back asynchronously BEFORE the
  // update is enqueued, one synchronisation at the end of the step
  if (!rc && out_metrics) rc = enqueue_metrics(c);
  if (!rc) rc = l3_adam_step(c, lr);
  if (!rc) L3_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  if (!rc && out_metrics) read_metrics(c, out_metrics);
  return rc;
}

int l3_dp_train_step_staged(l3_ctx* c, int batch, int global_batch, float lr, float out_metrics[4]) {
  L3_REQUIRE(c != nullptr && c->dp != nullptr, "data parallelism is not initialised (l3_dp_init)");
  int rc = l3_forward_backward(c, nullptr, 0, nullptr, 0, nullptr, batch, global_batch);
  if (!rc && out_metrics) rc = enqueue_metrics(c);
  if (!rc) rc = l3_adam_step(c, lr);
  if (!rc) L3_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  if (!rc && out_metrics) read_metrics(c, out_metrics);
  return rc;
}

// ---- data parallelism -------------------------------------------------------------------------------------------
int l3_dp_unique_id(char out[128]) {
  L3_REQUIRE(out != nullptr, "null argument");
  return dp_unique_id(out);
}
int l3_dp_nccl_version(void) { return dp_nccl_version(); }

int l3_dp_init(l3_ctx* c, const char id[128], int rank, int nranks) {
  L3_REQUIRE(c != nullptr && id != nullptr, "null argument");
  L3_REQUIRE(c->flags & L3_WS_TRAINING, "data parallelism needs a training context");
  L3_REQUIRE(c->dp == nullptr, "data parallelism is already initialised for this context");
  L3_CHECK_CUDA(cudaSetDevice(c->device));
  c->dp = dp_create(id, rank, nranks);
  return c->dp ? 0 : -1;
}

int l3_dp_info(l3_ctx* c, int* rank, int* nranks) {
  L3_REQUIRE(c != nullptr, "null ctx");
  if (rank) *rank = dp_rank(c->dp);
  if (nranks) *nranks = dp_nranks(c->dp);
  return c->dp ? 1 : 0;
}

__global__ void k_scale(float* __restrict__ p, long long n, float f) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] *= f;
}
int l3_dp_average_bn_state(l3_ctx* c) {
  L3_REQUIRE(c != nullptr && c->dp != nullptr, "data parallelism is not initialised (l3_dp_init)");
  L3_CHECK_CUDA(cudaSetDevice(c->device));
  void* p = c->bn_state;
  long long n = c->layout.n_state;
  if (dp_allreduce_ranges(c->dp, c->stream, &p, &n, 1, 0)) return -1;
  if (dp_join(c->dp, c->stream)) return -1;
  k_scale<<<ceil_div(n, 256), 256, 0, c->stream>>>(c->bn_state, n, 1.0f / (float)dp_nranks(c->dp));
  L3_CHECK_LAUNCH();
  return 0;
}

int l3_train_step_host(l3_ctx* c, const void* video_host, int video_fmt, const void* audio_host, int audio_fmt,
                       const float* labels_host, int batch, float lr, float out_metrics[4]) {
  int rc = l3_upload_batch_host(c, video_host, video_fmt, audio_host, audio_fmt, labels_host, batch);
  if (!rc) rc = l3_train_step_staged(c, batch, lr, out_metrics);
  return rc;
}

int l3_predict(l3_ctx* c, const void* video, int video_fmt, const void* audio, int audio_fmt, const float* labels,
               int batch, float* probs_out, float* logits_out) {
  if (check_batch(c, batch)) return -2;
  L3_CHECK_CUDA(cudaSetDevice(c->device));
  int slot;
  if (resolve_inputs(c, video, video_fmt, audio, audio_fmt, labels, batch, false, &slot)) return -2;
  int rc = c->dtype == L3_DTYPE_BF16
               ? forward_all<bf16>(c, video, video_fmt, audio, audio_fmt, labels, batch, false, 1.0f / batch)
               : forward_all<float>(c, video, video_fmt, audio, audio_fmt, labels, batch, false, 1.0f / batch);
  if (!rc) rc = release_slot(c, slot);
  if (rc) return rc;
  if (probs_out) L3_CHECK_CUDA(cudaMemcpyAsync(probs_out, c->head.probs, (size_t)batch * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (logits_out) L3_CHECK_CUDA(cudaMemcpyAsync(logits_out, c->head.logits, (size_t)batch * 8, cudaMemcpyDeviceToDevice, c->stream));
  c->last_batch = batch;
  c->last_global_batch = 0;      // evaluation metrics are local: the host sums them over the ranks it sharded over
  return 0;
}

int l3_embed_audio(l3_ctx* c, const void* audio, int audio_fmt, int n, int pooling, float* out) {
  if (check_batch(c, n)) return -2;
  L3_REQUIRE(c->audio.present, "context lacks the audio tower");
  L3_REQUIRE(pooling == L3_POOL_ORIGINAL || pooling == L3_POOL_SHORT, "bad pooling %d", pooling);
  L3_REQUIRE(audio && out, "null argument");
  Tower& tw = c->audio;
  tw.stream = c->stream;
  const int ph = c->spec.embed_pool[pooling][0], pw = c->spec.embed_pool[pooling][1];
  ConvLayer& L = tw.L[7];
  if (pack_all_weights(c, false, true, false)) return -1;
  if (c->dtype == L3_DTYPE_BF16) {
    if (tower_input<bf16>(c, tw, true, audio, audio_fmt, n, false)) return -1;
    if (tower_forward<bf16>(c, tw, n, false, true)) return -1;
    return launch_embed_pool<bf16>((const bf16*)L.z, n, L.H, L.W, L.Cout, ph, pw, out, c->stream);
  }
  if (tower_input<float>(c, tw, true, audio, audio_fmt, n, false)) return -1;
  if (tower_forward<float>(c, tw, n, false, true)) return -1;
  return launch_embed_pool<float>((const float*)L.z, n, L.H, L.W, L.Cout, ph, pw, out, c->stream);
}

int l3_embed_audio_frames(l3_ctx* c, const void* signal, int audio_fmt, int64_t n_samples, int hop, int n_frames,
                          int pooling, float* out) {
  L3_REQUIRE(c != nullptr && signal && out, "null argument");
  L3_REQUIRE(hop >= 1 && n_frames >= 1, "bad hop %d / n_frames %d", hop, n_frames);
  L3_REQUIRE((int64_t)(n_frames - 1) * hop + kSR <= n_samples, "signal of %lld samples is too short for %d frames at hop %d",
             (long long)n_samples, n_frames, hop);
  const int es = audio_fmt == L3_AUDIO_I16 ? 2 : 4;
  L3_REQUIRE(((int64_t)hop * es) % 16 == 0, "hop %d breaks the 16-byte alignment of the frame starts (frame on the host)", hop);
  L3_REQUIRE(pooling == L3_POOL_ORIGINAL || pooling == L3_POOL_SHORT, "bad pooling %d", pooling);
  int eh, ew;
  l3_embedding_map_shape(c->model_type, &eh, &ew);
  const int dim = (eh / c->spec.embed_pool[pooling][0]) * (ew / c->spec.embed_pool[pooling][1]) * 512;
  int rc = 0;
  c->fe.clip_stride = hop;   // overlapping 1 s windows read straight from the signal: no 10x framed copy
  for (int f0 = 0; f0 < n_frames && !rc; f0 += c->max_batch) {
    const int nb = n_frames - f0 < c->max_batch ? n_frames - f0 : c->max_batch;
    rc = l3_embed_audio(c, (const char*)signal + (int64_t)f0 * hop * es, audio_fmt, nb, pooling, out + (int64_t)f0 * dim);
  }
  c->fe.clip_stride = kSR;
  return rc;
}

int l3_embed_vision(l3_ctx* c, const void* video, int video_fmt, int n, float* out) {
  if (check_batch(c, n)) return -2;
  L3_REQUIRE(c->vision.present, "context lacks the vision tower");
  L3_REQUIRE(video && out, "null argument");
  Tower& tw = c->vision;
  tw.stream = c->stream;
  ConvLayer& L = tw.L[7];
  if (pack_all_weights(c, true, false, false)) return -1;
  if (c->dtype == L3_DTYPE_BF16) {
    if (tower_input<bf16>(c, tw, false, video, video_fmt, n, false)) return -1;
    if (tower_forward<bf16>(c, tw, n, false, true)) return -1;
    return launch_embed_pool<bf16>((const bf16*)L.z, n, L.H, L.W, L.Cout, kVisionEmbedPool, kVisionEmbedPool, out, c->stream);
  }
  if (tower_input<float>(c, tw, false, video, video_fmt, n, false)) return -1;
  if (tower_forward<float>(c, tw, n, false, true)) return -1;
  return launch_embed_pool<float>((const float*)L.z, n, L.H, L.W, L.Cout, kVisionEmbedPool, kVisionEmbedPool, out, c->stream);
}

int l3_frontend_fwd(l3_ctx* c, const void* audio, int audio_fmt, int n, float* out) {
  if (check_batch(c, n)) return -2;
  L3_REQUIRE(audio && out, "null argument");
  return launch_frontend(c->fe, audio, audio_fmt == L3_AUDIO_I16, n, out, c->clip_max, c->stream);
}

// ---- stand-alone ops (unit tests) -------------------------------------------------------------------------
namespace {
// stream-ordered temporary
struct TempBuf {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  int alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    L3_CHECK_CUDA(cudaMallocAsync(&p, bytes, s));
    return 0;
  }
  ~TempBuf() {
    if (p) cudaFreeAsync(p, s);
  }
};
// L3_DTYPE_F32TC forward (flip = 0: fp16 parts, weights x 2^10) / data gradient (flip = 1: bf16 parts) of one layer
int conv_split_op(const float* in, const float* w, const float* bias, float* out, int B, int H, int W, int Cin, int Cout,
                  int flip, cudaStream_t s) {
  L3_REQUIRE(conv_tc_supported(), "tcgen05 path unavailable on this device");
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "split conv: C%%64");
  const int Ci = flip ? Cout : Cin, Co = flip ? Cin : Cout;   // channels of the convolution that actually runs
  const long long rows_p = (long long)B * (H + 2) * (W + 2);
  TempBuf sp, pk;
  if (sp.alloc((size_t)rows_p * 3 * Ci * 2, s) || pk.alloc((size_t)9 * 3 * Ci * Co * 2, s)) return -1;
  PackBatch pb;
  pb.n = 1;
  pb.job[0] = PackJob{w, (bf16*)pk.p, Cin, Cout, flip};
  pb.job[0].split = 1;
  pb.job[0].fp16 = flip ? 0 : 1;
  pb.job[0].scale = flip ? 1.f : kSplitWeightScale;
  if (launch_pack_weights_batch(pb, s)) return -1;
  if (launch_split16(in, sp.p, rows_p, Ci, flip ? 0 : 1, s)) return -1;
  return launch_conv3x3_tc_split(sp.p, pk.p, bias, out, B, H, W, Ci, Co, flip ? 0 : 1, flip ? 1.f : kSplitWeightScaleInv, s);
}
}  // namespace

int l3_conv3x3_fwd(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int Cin, int Cout,
                   int dtype, int use_tc, void* scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == L3_DTYPE_F32TC) {
    L3_REQUIRE(use_tc, "L3_DTYPE_F32TC is a tensor-core mode");
    return conv_split_op((const float*)in, w, bias, (float*)out, B, H, W, Cin, Cout, 0, s);
  }
  if (use_tc && (Cin == 1 || Cin == 3)) {
    L3_REQUIRE(dtype == L3_DTYPE_BF16 && Cout == 64, "tc first-layer conv: bf16, Cout 64");
    L3_REQUIRE(conv_tc_supported(), "tcgen05 path unavailable on this device");
    return launch_first_conv_tc((const bf16*)in, w, bias, (bf16*)out, B, H, W, Cin, Cout, nullptr, s);
  }
  if (use_tc) {
    L3_REQUIRE(dtype == L3_DTYPE_BF16 && Cin % 64 == 0 && Cout % 64 == 0 && scratch, "tc conv: bf16, C%%64, scratch");
    L3_REQUIRE(conv_tc_supported(), "tcgen05 path unavailable on this device");
    if (launch_pack_weights_tc(w, (bf16*)scratch, Cin, Cout, 0, s)) return -1;
    return launch_conv3x3_tc((const bf16*)in, (const bf16*)scratch, bias, (bf16*)out, B, H, W, Cin, Cout, nullptr, 0, s);
  }
  if (dtype == L3_DTYPE_BF16) return launch_conv3x3_simt<bf16>((const bf16*)in, w, bias, (bf16*)out, B, H, W, Cin, Cout, s);
  return launch_conv3x3_simt<float>((const float*)in, w, bias, (float*)out, B, H, W, Cin, Cout, s);
}
int l3_conv3x3_fwd_stats(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int Cin,
                         int Cout, void* scratch, double* stats, int relu_stats, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  L3_REQUIRE(conv_tc_supported(), "tcgen05 path unavailable on this device");
  L3_REQUIRE(stats != nullptr, "fwd_stats: stats is NULL");
  if (Cin == 1 || Cin == 3) {
    L3_REQUIRE(Cout == 64 && !relu_stats, "tc first-layer conv: Cout 64, no relu statistics");
    return launch_first_conv_tc((const bf16*)in, w, bias, (bf16*)out, B, H, W, Cin, Cout, stats, s);
  }
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0 && scratch, "tc conv: C%%64, scratch");
  L3_REQUIRE(conv_tc_fuses_stats(), "the selected L3_CONV_TC_VARIANT has no fused statistics");
  if (launch_pack_weights_tc(w, (bf16*)scratch, Cin, Cout, 0, s)) return -1;
  return launch_conv3x3_tc((const bf16*)in, (const bf16*)scratch, bias, (bf16*)out, B, H, W, Cin, Cout, stats, relu_stats, s);
}
int l3_conv3x3_dgrad(const void* dz, const float* w, void* da, int B, int H, int W, int Cin, int Cout, int dtype,
                     int use_tc, void* scratch, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == L3_DTYPE_F32TC) {
    L3_REQUIRE(use_tc, "L3_DTYPE_F32TC is a tensor-core mode");
    return conv_split_op((const float*)dz, w, nullptr, (float*)da, B, H, W, Cin, Cout, 1, s);
  }
  L3_REQUIRE(scratch, "dgrad needs scratch of 9*Cin*Cout floats");
  if (use_tc) {
    L3_REQUIRE(dtype == L3_DTYPE_BF16 && Cin % 64 == 0 && Cout % 64 == 0, "tc dgrad: bf16, C%%64");
    L3_REQUIRE(conv_tc_supported(), "tcgen05 path unavailable on this device");
    if (launch_pack_weights_tc(w, (bf16*)scratch, Cin, Cout, 1, s)) return -1;
    return launch_conv3x3_tc((const bf16*)dz, (const bf16*)scratch, nullptr, (bf16*)da, B, H, W, Cout, Cin, nullptr, 0, s);
  }
  if (launch_flip_transpose(w, (float*)scratch, Cin, Cout, s)) return -1;
  if (dtype == L3_DTYPE_BF16)
    return launch_conv3x3_simt<bf16>((const bf16*)dz, (const float*)scratch, nullptr, (bf16*)da, B, H, W, Cout, Cin, s);
  return launch_conv3x3_simt<float>((const float*)dz, (const float*)scratch, nullptr, (float*)da, B, H, W, Cout, Cin, s);
}
int l3_conv3x3_dgrad_stats(const void* dz, const float* w, void* da, int B, int H, int W, int Cin, int Cout, void* scratch,
                           const void* z_below, const float* scale, const float* shift, double* sums, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  L3_REQUIRE(scratch && z_below && scale && shift && sums, "dgrad_stats: NULL argument");
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc dgrad: C%%64");
  L3_REQUIRE(conv_tc_supported(), "tcgen05 path unavailable on this device");
  if (launch_pack_weights_tc(w, (bf16*)scratch, Cin, Cout, 1, s)) return -1;
  return launch_dgrad3x3_tc_bwdstats((const bf16*)dz, (const bf16*)scratch, (bf16*)da, B, H, W, Cout, Cin,
                                     (const bf16*)z_below, scale, shift, sums, s);
}
int l3_conv3x3_wgrad(const void* a, const void* dz, float* dw, float* db, int B, int H, int W, int Cin, int Cout,
                     int dtype, int use_tc, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == L3_DTYPE_F32TC) {
    L3_REQUIRE(use_tc && conv_tc_supported(), "L3_DTYPE_F32TC is a tensor-core mode (sm_100)");
    L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "split wgrad: C%%64");
    const long long rows_p = (long long)B * (H + 2) * (W + 2);
    TempBuf spa, spz, w4, s64;
    if (spa.alloc((size_t)rows_p * 3 * Cin * 2, s) || spz.alloc((size_t)rows_p * 3 * Cout * 2, s) ||
        w4.alloc(sizeof(float) * 9 * 4 * (size_t)Cin * Cout, s) || s64.alloc(sizeof(double) * 2 * Cout, s))
      return -1;
    if (launch_split16((const float*)a, spa.p, rows_p, Cin, 0, s)) return -1;
    if (launch_split16((const float*)dz, spz.p, rows_p, Cout, 0, s)) return -1;
    if (launch_wgrad3x3_tc_split(spa.p, spz.p, dw, (float*)w4.p, B, H, W, Cin, Cout, s)) return -1;
    if (db) {
      if (launch_channel_stats<float>((const float*)dz, rows_p, Cout, 0, (double*)s64.p, s)) return -1;
      if (launch_f64_to_f32((const double*)s64.p, db, Cout, s)) return -1;
    }
    return 0;
  }
  L3_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 9 * Cin * Cout, s));
  if (db) L3_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cout, s));
  if (use_tc && (Cin == 1 || Cin == 3)) {
    L3_REQUIRE(dtype == L3_DTYPE_BF16 && Cout == 64, "tc first-layer wgrad: bf16, Cout 64");
    L3_REQUIRE(conv_tc_supported(), "tcgen05 path unavailable on this device");
    return launch_first_wgrad_tc((const bf16*)a, (const bf16*)dz, dw, db, nullptr, B, H, W, Cin, Cout, s);
  }
  if (use_tc) {
    L3_REQUIRE(dtype == L3_DTYPE_BF16 && Cin % 64 == 0 && Cout % 64 == 0, "tc wgrad: bf16, C%%64");
    L3_REQUIRE(conv_tc_supported(), "tcgen05 path unavailable on this device");
    return launch_wgrad3x3_tc((const bf16*)a, (const bf16*)dz, dw, db, B, H, W, Cin, Cout, s);
  }
  if (dtype == L3_DTYPE_BF16) return launch_wgrad3x3_simt<bf16>((const bf16*)a, (const bf16*)dz, dw, db, B, H, W, Cin, Cout, s);
  return launch_wgrad3x3_simt<float>((const float*)a, (const float*)dz, dw, db, B, H, W, Cin, Cout, s);
}

// ---- stand-alone element-wise layer ops (unit tests) ----------------------------------------------------------
namespace {
// a BnRef over a stream-ordered temporary, filled from the caller's {scale, shift, mean, invstd}
struct TempBn {
  BnRef bn;
  char* mem = nullptr;
  cudaStream_t s = nullptr;
  int init(int C, const float* bn4, float* d_gamma, float* d_beta, cudaStream_t stream) {
    s = stream;
    const size_t bytes = sizeof(double) * 2 * C + sizeof(float) * 2 * C;
    L3_CHECK_CUDA(cudaMallocAsync((void**)&mem, bytes, s));
    bn = BnRef{};
    bn.C = C;
    bn.sum = (double*)mem;
    bn.c1 = (float*)(mem + sizeof(double) * 2 * C);
    bn.c2 = bn.c1 + C;
    bn.scale = const_cast<float*>(bn4);
    bn.shift = const_cast<float*>(bn4) + C;
    bn.mean = const_cast<float*>(bn4) + 2 * C;
    bn.invstd = const_cast<float*>(bn4) + 3 * C;
    bn.d_gamma = d_gamma;
    bn.d_beta = d_beta;
    return 0;
  }
  ~TempBn() {
    if (mem) cudaFreeAsync(mem, s);
  }
};
}  // namespace

int l3_act_fwd(const void* z, void* a, int B, int H, int W, int C, const float* scale, const float* shift, int pool,
               int relu_first, int dtype, void* zsel, uint8_t* sel, void* stream) {
  L3_REQUIRE(z && a && scale && shift, "act_fwd: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == L3_DTYPE_BF16)
    return launch_act_fwd<bf16>((const bf16*)z, (bf16*)a, B, H, W, C, scale, shift, pool, relu_first, s, (bf16*)zsel, sel);
  return launch_act_fwd<float>((const float*)z, (float*)a, B, H, W, C, scale, shift, pool, relu_first, s, (float*)zsel, sel);
}

int l3_bn_act_bwd(const void* da, const void* z, void* dz, int B, int H, int W, int C, const float* bn4, int pool,
                  int relu_first, int dtype, const void* zsel, const uint8_t* sel, float* d_gamma, float* d_beta,
                  void* stream) {
  L3_REQUIRE(da && z && dz && bn4 && d_gamma && d_beta, "bn_act_bwd: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  TempBn t;
  if (t.init(C, bn4, d_gamma, d_beta, s)) return -1;
  const long long rows = (long long)B * H * W;
  if (dtype == L3_DTYPE_BF16) {
    if (launch_bwd_stats<bf16>((const bf16*)da, (const bf16*)z, B, H, W, C, t.bn, pool, relu_first, s, (const bf16*)zsel, sel)) return -1;
    if (launch_bn_bwd_finalize(t.bn, rows, 1, s)) return -1;
    return launch_bwd_apply<bf16>((const bf16*)da, (const bf16*)z, (bf16*)dz, B, H, W, C, t.bn, pool, relu_first, s, sel);
  }
  if (launch_bwd_stats<float>((const float*)da, (const float*)z, B, H, W, C, t.bn, pool, relu_first, s, (const float*)zsel, sel)) return -1;
  if (launch_bn_bwd_finalize(t.bn, rows, 2, s)) return -1;
  return launch_bwd_apply<float>((const float*)da, (const float*)z, (float*)dz, B, H, W, C, t.bn, pool, relu_first, s, sel);
}

int l3_gmaxpool_fwd(const void* z, int B, int H, int W, int C, const float* scale, const float* shift, int dtype,
                    float* out, int* argmax, void* stream) {
  L3_REQUIRE(z && scale && shift && out && argmax, "gmaxpool_fwd: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* scratch = nullptr;
  L3_CHECK_CUDA(cudaMallocAsync((void**)&scratch, sizeof(unsigned long long) * (size_t)B * C, s));
  int rc = dtype == L3_DTYPE_BF16
               ? launch_gmaxpool_fwd<bf16>((const bf16*)z, B, H * W, C, scale, shift, out, C, argmax, scratch, s)
               : launch_gmaxpool_fwd<float>((const float*)z, B, H * W, C, scale, shift, out, C, argmax, scratch, s);
  cudaFreeAsync(scratch, s);
  return rc;
}

int l3_gmaxpool_bwd(const float* dpool, const int* argmax, const void* z, void* dz, int B, int H, int W, int C,
                    const float* bn4, int dtype, float* d_gamma, float* d_beta, void* stream) {
  L3_REQUIRE(dpool && argmax && z && dz && bn4 && d_gamma && d_beta, "gmaxpool_bwd: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  TempBn t;
  if (t.init(C, bn4, d_gamma, d_beta, s)) return -1;
  const long long rows = (long long)B * H * W;
  if (dtype == L3_DTYPE_BF16) {
    if (launch_gmaxpool_bwd<bf16>(dpool, C, argmax, (const bf16*)z, (bf16*)dz, t.bn, B, H, W, C, s)) return -1;
    if (launch_bn_bwd_finalize(t.bn, rows, 0, s)) return -1;
    return launch_bn_bwd_apply<bf16>((bf16*)dz, (const bf16*)z, B, H, W, C, t.bn, 0, s);
  }
  if (launch_gmaxpool_bwd<float>(dpool, C, argmax, (const float*)z, (float*)dz, t.bn, B, H, W, C, s)) return -1;
  if (launch_bn_bwd_finalize(t.bn, rows, 0, s)) return -1;
  return launch_bn_bwd_apply<float>((float*)dz, (const float*)z, B, H, W, C, t.bn, 0, s);
}

int64_t l3_debug_read(l3_ctx* c, const char* which, int batch, float* out_host, int64_t cap) {
  if (check_batch(c, batch)) return -2;
  L3_REQUIRE(which && out_host, "null argument");
  std::string w(which);
  const void* src = nullptr;
  long long n = 0;
  int H = 1, W = 1, C = 1, padded = 0, is_float = 0, kind = 0;   // kind 1: uint8 record, 2: int32
  auto tower_of = [&](const std::string& t) -> Tower* { return t == "vision" ? &c->vision : t == "audio" ? &c->audio : nullptr; };
  size_t slash = w.find('/');
  if (w == "concat") { src = c->head.concat; n = (long long)batch * 1024; is_float = 1; }
  else if (w == "hidden") { src = c->head.hidden; n = (long long)batch * 128; is_float = 1; }
  else if (w == "logits") { src = c->head.logits; n = (long long)batch * 2; is_float = 1; }
  else if (w == "probs") { src = c->head.probs; n = (long long)batch * 2; is_float = 1; }
  else if (w == "dconcat") { src = c->head.dconcat; n = (long long)batch * 1024; is_float = 1; }
  else if (slash != std::string::npos) {
    Tower* tw = tower_of(w.substr(0, slash));
    std::string k = w.substr(slash + 1);
    L3_REQUIRE(tw && tw->present, "unknown tower in '%s'", which);
    if (k == "x0") { src = tw->x0; n = (long long)batch * tw->H0 * tw->W0 * tw->C0; is_float = 1; }
    else if (k == "xin") { src = tw->xin; H = tw->H0; W = tw->W0; C = tw->C0; padded = 1; n = (long long)batch * H * W * C; }
    else if (k.size() == 2 && (k[0] == 'z' || k[0] == 'a') && k[1] >= '0' && k[1] <= '7') {
      ConvLayer& L = tw->L[k[1] - '0'];
      if (k[0] == 'z') { src = L.z; H = L.H; W = L.W; C = L.Cout; }
      else {
        L3_REQUIRE(L.a, "layer 7 has no activation buffer");
        src = L.a; H = L.pool ? L.H / 2 : L.H; W = L.pool ? L.W / 2 : L.W; C = L.Cout; padded = 1;
      }
      n = (long long)batch * H * W * C;
    }
    else if (k.size() == 4 && k.compare(0, 3, "sel") == 0 && k[3] >= '0' && k[3] <= '7') {
      // recorded max-pool routing of a pooled layer (training forward): window position | 4 * (max > 0)
      ConvLayer& L = tw->L[k[3] - '0'];
      L3_REQUIRE(L.sel, "layer %c has no routing record (not pooled, or not a training context)", k[3]);
      src = L.sel; n = (long long)batch * (L.H / 2) * (L.W / 2) * L.Cout; kind = 1;
    }
    else if (k == "argmax") { src = tw->argmax; n = (long long)batch * 512; kind = 2; }   // global max-pool routing
  }
  L3_REQUIRE(src, "unknown buffer '%s'", which);
  L3_REQUIRE(n <= cap, "buffer '%s' has %lld elements, capacity %lld", which, n, (long long)cap);
  float* tmp = nullptr;
  L3_CHECK_CUDA(cudaMalloc(&tmp, n * 4));
  int blocks = (int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
  if (kind == 1) k_gather_unpad<uint8_t><<<blocks, 256, 0, c->stream>>>((const uint8_t*)src, tmp, n, H, W, C, 0);
  else if (kind == 2) k_gather_unpad<int><<<blocks, 256, 0, c->stream>>>((const int*)src, tmp, n, H, W, C, 0);
  else if (is_float) k_gather_unpad<float><<<blocks, 256, 0, c->stream>>>((const float*)src, tmp, n, H, W, C, 0);
  else if (c->dtype == L3_DTYPE_BF16) k_gather_unpad<bf16><<<blocks, 256, 0, c->stream>>>((const bf16*)src, tmp, n, H, W, C, padded);
  else k_gather_unpad<float><<<blocks, 256, 0, c->stream>>>((const float*)src, tmp, n, H, W, C, padded);
  cudaError_t e = cudaMemcpyAsync(out_host, tmp, n * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  L3_CHECK_CUDA(e);
  return n;
}

}  // extern "C"
