// Internal launcher declarations (host side). Every launcher is asynchronous on `s`.
#pragma once
#include "common.cuh"

namespace l3 {

// ---- front-end (frontend.cu) ------------------------------------------------------------
struct FrontendPlan {
  int n_dft;        // 512 | 2048
  int n_hop;        // 242
  int n_frames;     // 197 | 199
  int left_pad;     // 0 | 982
  int n_out;        // 257 (linear) | n_mels
  int mel;          // 1 -> mel filterbank
  int decibel;      // 1 -> 10*log10, per-clip max, clip -80 ; 0 -> log(max(x,1e-12))/5
  int n_samples;    // 48000
  long long clip_stride;  // samples between the starts of consecutive clips (n_samples; the hop when framing on device)
  // device constants (built once per ctx by frontend_build_tables)
  const float2* tw1;        // FFT step-1 twiddles [16][n_dft/16]   (frontend_fft.cuh)
  const float2* tw2;        // FFT step-2 twiddles [16][n_dft/256]
  const float* window;      // n_dft periodic hann
  const float* window_i16;  // the same times 2^-15 (exact): int16 samples are windowed and scaled by one multiply
  const int* mel_start;     // [n_mels]
  const int* mel_count;     // [n_mels]
  const int* mel_offset;    // [n_mels] into mel_weight
  const float* mel_weight;  // packed non-zeros
  int mel_nnz;              // number of packed non-zeros
};
// bytes of device memory needed for the tables
size_t frontend_table_bytes(int n_dft, int n_mels);
// fills `dev_mem` (device) with tables and completes `plan` pointers. Synchronous H2D copies on `s`.
int frontend_build_tables(FrontendPlan* plan, int sr, int n_mels, void* dev_mem, cudaStream_t s);
// audio: int16 (is_i16=1) or float (B, n_samples). raw: (B, n_out, n_frames) float scratch; clip_max: int[B].
// out: (B, n_out, n_frames) float final.
// finish = 0 (decibel plans): stop after the per-clip maximum; `out` then holds the un-referenced dB map and the caller
// applies max(x - clip max, -80) itself (launch_input_stage mode 2 fuses it with the input BatchNorm's passes)
int launch_frontend(const FrontendPlan& p, const void* audio, int is_i16, int B, float* out, int* clip_max,
                    cudaStream_t s, int finish = 1);

// ---- elementwise / reductions (elementwise.cu) ------------------------------------------
template <typename T>
int launch_channel_stats(const T* x, long long rows, int C, int relu, double* sum2C, cudaStream_t s);
int launch_bn_finalize(const BnRef& bn, long long count, int training, float momentum, float eps, int unbiased,
                       cudaStream_t s);
// input stage of a tower in one pass (C = 1 | 3).  mode 0: x0 holds the float input; 1: u8 video -> 2*(x/255)-1 -> x0;
// 2: dB finish of the front-end's raw map in place (x0 = max(x0 - clip max, -80)).  sum != null: per-channel sum / sum of
// squares (fp64, zeroed here).  xin != null: (scale ? v*scale+shift : v) -> zero-haloed padded (B,H+2,W+2,C).
template <typename T>
int launch_input_stage(int mode, const uint8_t* u8, float* x0, T* xin, int B, int H, int W, int C, const float* scale,
                       const float* shift, const int* clip_max, double* sum, cudaStream_t s);
template <typename T>
int launch_zero_halo(T* buf, int B, int H, int W, int C, cudaStream_t s);
// zsel / sel (optional, pooled layers in training): the winning pre-activation (B,H/2,W/2,C) and one byte per pooled
// element (window position | 4 * (max > 0)) for the backward kernels
template <typename T>
int launch_act_fwd(const T* z, T* a, int B, int H, int W, int C, const float* scale, const float* shift, int pool,
                   int relu_first, cudaStream_t s, T* zsel = nullptr, uint8_t* sel = nullptr);
template <typename T>
int launch_gmaxpool_fwd(const T* z, int B, int HW, int C, const float* scale, const float* shift, float* out,
                        int out_stride, int* argmax, unsigned long long* scratch /* B*C */, cudaStream_t s);
template <typename T>
int launch_gmaxpool_bwd(const float* dpool, int dpool_stride, const int* argmax, const T* z, T* dy, const BnRef& bn,
                        int B, int H, int W, int C, cudaStream_t s);   // dy: padded
// activation + BN backward in two passes over (da, z): pass 1 -> bn.sum = {sum dy, sum dy*xhat}; pass 2 -> dz (padded)
template <typename T>
int launch_bwd_stats(const T* da, const T* z, int B, int H, int W, int C, const BnRef& bn, int pool, int relu_first,
                     cudaStream_t s, const T* zsel = nullptr, const uint8_t* sel = nullptr);
template <typename T>
int launch_bwd_apply(const T* da, const T* z, T* dz, int B, int H, int W, int C, const BnRef& bn, int pool,
                     int relu_first, cudaStream_t s, const uint8_t* sel = nullptr);
// raw_sums: bn.sum[C..2C) holds 0: sum(dy*xhat); 1: sum(dy*xin) (launch_bwd_stats<bf16>); 2: sum(dy*(xin-mean)) (<float>)
int launch_bn_bwd_finalize(const BnRef& bn, long long count, int raw_sums, cudaStream_t s);
template <typename T>
int launch_bn_bwd_apply(T* dy_inout, const T* z, int B, int H, int W, int C, const BnRef& bn, int relu_first,
                        cudaStream_t s);   // dy: padded, z: unpadded
// embedding head: MaxPooling2D(pool,'same') over raw z (B,H,W,C) -> (B, OH*OW*C) float, flatten (h,w,c)
template <typename T>
int launch_embed_pool(const T* z, int B, int H, int W, int C, int ph, int pw, float* out, cudaStream_t s);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, long long n_l2, float lr_t, float b1,
                float b2, float eps, float l2, cudaStream_t s);
int launch_l2_penalty(const float* p, long long n_l2, double* out, cudaStream_t s);
int launch_zero(void* p, size_t bytes, cudaStream_t s);
// Parity mode on tensor cores: fp32 (rows, C) -> 16-bit parts [hi | lo | hi] (rows, 3*C); fp16 = 1: fp16 parts of scale*x
// (forward operands), 0: bf16 parts (backward operands).  rows counts PADDED pixel rows (halo rows split to zeros).
int launch_split16(const float* src, void* dst, long long rows, int C, int fp16, cudaStream_t s);
// dst[i] = (float)src[i]
int launch_f64_to_f32(const double* src, float* dst, long long n, cudaStream_t s);

// ---- SIMT fp32-accumulate convolutions (conv_simt.cu) --------------------------------------
// out[b,y,x,co] = bias[co] + sum_{ky,kx,ci} in[b,y+ky-1,x+kx-1,ci] * w[(ky*3+kx)*Cin*Cout + ci*Cout + co]
template <typename T>
int launch_conv3x3_simt(const T* in, const float* w, const float* bias, T* out, int B, int H, int W, int Cin,
                        int Cout, cudaStream_t s);
// dw[(ky*3+kx)*Cin*Cout + ci*Cout + co] += sum_{b,y,x} a[b,y+ky-1,x+kx-1,ci] * dz[b,y,x,co]   (dw pre-zeroed)
// db[co] += sum dz
// T = float (parity mode): the split-K partials are merged in fp64 and rounded once -- dw / db are OVERWRITTEN and the
// result is independent of the merge order; scratch64 = 9*Cin*Cout + Cout doubles (null: stream-ordered allocation)
template <typename T>
int launch_wgrad3x3_simt(const T* a, const T* dz, float* dw, float* db, int B, int H, int W, int Cin, int Cout,
                         cudaStream_t s, double* scratch64 = nullptr);
// first-layer (Cin = 1|3, Cout = 64) backward: weight/bias gradient, and the input-BN backward sums computed
// straight from dz without materialising the data gradient (bn.sum <- sum(da), sum(da*xhat))
// stats (optional): double[2*64] per-channel sum / sum of squares of the stored output (BN batch statistics)
template <typename T>
int launch_first_conv(const T* in, const float* w, const float* bias, T* out, int B, int H, int W, int C0, int Cout,
                      double* stats, cudaStream_t s);
// d1 (optional, 9*64 floats): weight gradient w.r.t. an all-ones input plane, consumed by launch_bn0_from_dw
template <typename T>
int launch_first_wgrad(const T* a, const T* dz, float* dw, float* db, float* d1, int B, int H, int W, int C0, int Cout,
                       cudaStream_t s, double* scratch64 = nullptr);
// input-BN gradients (bn.sum = {sum da, sum da*xhat}) from dw and d1 alone; *fallback = 1 if some |gamma| ~ 0
int launch_bn0_from_dw(const float* w, const float* dw, const float* d1, const BnRef& bn, int C0, int* fallback,
                       cudaStream_t s);
// direct computation of the same sums from dz; with only_if != null the kernel is a no-op unless *only_if != 0
template <typename T>
int launch_first_dgrad_bnstats(const T* dz, const float* w, const float* x0, const BnRef& bn, int B, int H, int W, int C0,
                               int Cout, const int* only_if, cudaStream_t s);
// w_t[(ky*3+kx)*Cout*Cin + co*Cin + ci] = w[((2-ky)*3+(2-kx))*Cin*Cout + ci*Cout + co]  (for dgrad-as-conv)
int launch_flip_transpose(const float* w, float* w_t, int Cin, int Cout, cudaStream_t s);

// ---- head (head.cu) ----------------------------------------------------------------------------
struct HeadRef {
  const float *w1, *b1, *w2, *b2;   // params: (1024,128),(128),(128,2),(2)
  float *dw1, *db1, *dw2, *db2;     // grads
  float* concat;                    // (B,1024) [vision | audio]
  float* hidden;                    // (B,128)
  float* probs;                     // (B,2)
  float* logits;                    // (B,2)
  float* dlogits;                   // (B,2)
  float* dhidden;                   // (B,128)
  float* dconcat;                   // (B,1024)
  float* metrics;                   // [0]=sum ce, [1]=#correct   (device, accumulated with atomics; zero first)
};
int launch_head_fwd(const HeadRef& h, const float* labels, int B, float grad_scale, cudaStream_t s);
int launch_head_bwd(const HeadRef& h, int B, cudaStream_t s);

// ---- tcgen05 bf16 implicit-GEMM convolutions (conv_tc.cu) ---------------------------------------
// 1 when the running device is sm_100 and the driver exposes cuTensorMapEncodeTiled.
int conv_tc_supported();
int conv_tc_fuses_stats();   // the active forward variant computes BN statistics in its epilogue
// Weight pre-pack: fp32 HWIO (3,3,Cin,Cout) -> bf16 K-major rows [(tap*Cin/64 + kc)*Cout + co][64 ci] (one TMA box
// row = 128 B).  flip_transpose=1 packs the dgrad operand: taps flipped and Cin/Cout swapped, i.e. the result is the
// forward pack of a conv with Cin' = Cout, Cout' = Cin.
int launch_pack_weights_tc(const float* w, bf16* packed, int Cin, int Cout, int flip_transpose, cudaStream_t s);
// the same for every tensor-core layer of a step in ONE launch (weights change once per step, in k_adam)
static const int kMaxPackJobs = 32;
struct PackJob {
  const float* w;
  bf16* out;
  int Cin, Cout, flip;
  int split = 0;       // parity mode on tensor cores: [hi ; hi ; lo] parts of scale * w over 3x the input channels
  int fp16 = 0;        //   parts are fp16 (forward) / bf16 (backward)
  float scale = 1.f;
};
struct PackBatch {
  PackJob job[kMaxPackJobs];
  int n;
};
int launch_pack_weights_batch(const PackBatch& pb, cudaStream_t s);
// Forward / dgrad conv on tensor cores. in: zero-haloed padded bf16 (B,H+2,W+2,Cin), Cin%64==0, Cout%64==0;
// out: unpadded bf16 (B,H,W,Cout); bias may be null.  stats (optional): double[2*Cout] receives the per-channel sum and
// sum of squares of the stored output (of relu(output) if relu_stats) -- the BatchNorm batch statistics, fused into
// the epilogue so the tensor is not read again.
int launch_conv3x3_tc(const bf16* in, const bf16* packed_w, const float* bias, bf16* out, int B, int H, int W, int Cin,
                      int Cout, double* stats, int relu_stats, cudaStream_t s);
// Inference forward with BatchNorm (moving statistics folded into scale / shift) + ReLU fused into the epilogue, stored
// un-padded (out_padded = 0) or straight into the next layer's zero-haloed padded input (out_padded = 1)
int launch_conv3x3_tc_act(const bf16* in, const bf16* packed_w, const float* bias, bf16* out, int B, int H, int W, int Cin,
                          int Cout, const float* scale, const float* shift, int relu_first, int out_padded, cudaStream_t s);
// Data gradient of a conv (Cin -> Cout) on tensor cores with pass 1 of the BN/ReLU backward of the layer BELOW fused
// into the epilogue: da (B,H,W,Cin) = conv(dz, packed_wt); sums[0..Cin) = sum(dy), sums[Cin..2Cin) = sum(dy * z_below)
// with dy = da where scale*z_below + shift > 0 (the layer below is Conv -> BN -> ReLU, not pooled); sums are zeroed here.
// Parity mode on tensor cores (split 16-bit operands, fp32 accumulate and output): see conv_tc.cu
int launch_conv3x3_tc_split(const void* in_split, const void* packed_w_split, const float* bias, float* out, int B, int H, int W,
                            int Cin, int Cout, int fp16, float out_scale, cudaStream_t s);
int conv_tc_fuses_bwd_stats(int channels_below);   // 1 where the fused epilogue beats the stand-alone statistics pass
int launch_dgrad3x3_tc_bwdstats(const bf16* dz, const bf16* packed_wt, bf16* da, int B, int H, int W, int Cout, int Cin,
                                const bf16* z_below, const float* scale_below, const float* shift_below, double* sums,
                                cudaStream_t s);
// Weight gradient on tensor cores: a padded (B,H+2,W+2,Cin), dz padded (B,H+2,W+2,Cout) with zero halos;
// dw (3,3,Cin,Cout) fp32 and db (Cout) are accumulated into (pre-zeroed by the caller).
int launch_wgrad3x3_tc(const bf16* a, const bf16* dz, float* dw, float* db, int B, int H, int W, int Cin, int Cout,
                       cudaStream_t s);
int launch_bias_grad_tc(const bf16* dz, float* db, int B, int H, int W, int Cout, cudaStream_t s);
// Parity mode on tensor cores: bf16 split operands [hi | lo | hi]; dw4 = fp32 scratch of 9 * 2*Cin * 2*Cout floats
int launch_wgrad3x3_tc_split(const void* a_split, const void* dz_split, float* dw, float* dw4, int B, int H, int W, int Cin,
                             int Cout, cudaStream_t s);
// First-layer (Cin = 1|3, Cout = 64) weight gradient on tensor cores: xin padded (B,H+2,W+2,C0), dz padded
// (B,H+2,W+2,64); accumulates dw (9*C0*64), db (64) and d1 (9*64, optional; zeroed here) -- see launch_first_wgrad.
int launch_first_wgrad_tc(const bf16* xin, const bf16* dz, float* dw, float* db, float* d1, int B, int H, int W, int C0,
                          int Cout, cudaStream_t s);

// First-layer (Cin = 1|3, Cout = 64) forward on tensor cores: xin padded bf16 (B,H+2,W+2,C0), w fp32 HWIO, out unpadded
// bf16 (B,H,W,64); stats as in launch_conv3x3_tc (zeroed here).
// act_scale / act_shift (inference, optional): BN + ReLU fused into the epilogue; with out_padded the result is stored
// into the next layer's zero-haloed padded input.
int launch_first_conv_tc(const bf16* xin, const float* w, const float* bias, bf16* out, int B, int H, int W, int C0,
                         int Cout, double* stats, cudaStream_t s, const float* act_scale = nullptr,
                         const float* act_shift = nullptr, int out_padded = 0);

// ---- data-parallel gradient exchange (dp.cu): NCCL bound at run time ---------------------------------------------
static const int kDpMaxBuckets = 24;
struct DpState;
int dp_unique_id(char out[128]);
int dp_nccl_version();
DpState* dp_create(const char id_bytes[128], int rank, int nranks);   // ncclCommInitRank on the current device
void dp_destroy(DpState* d);
int dp_rank(const DpState* d);
int dp_nranks(const DpState* d);
void dp_begin_step(DpState* d);
// in-place sum over the ranks of disjoint ranges (one grouped NCCL launch on the communication stream), ordered after
// everything enqueued so far on `producer`
int dp_allreduce_ranges(DpState* d, cudaStream_t producer, void* const* ptrs, const long long* counts, int n_ranges,
                        int is_f64);
// `consumer` waits for every collective issued so far
int dp_join(DpState* d, cudaStream_t consumer);

}  // namespace l3
