// HBM-bound kernels of the L3 AVC step: input scaling, BatchNorm statistics / finalize / backward,
// activation (+2x2 max-pool) forward and backward, global max-pool, embedding pool, Adam.
// Reference semantics: keras BatchNormalization / Activation('relu') / MaxPooling2D as instantiated in
// l3embedding/audio_model.py:370-437 and l3embedding/vision_model.py:124-190; Adam at l3embedding/train.py:282.
// All activations are NHWC; per-thread work is 8 channels (one 16-byte bf16 vector) with coalesced access.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "kernels.h"

namespace l3 {

static const int kThreads = 256;
static const int kMaxBlocks = 148 * 8;  // grid-stride kernels: a multiple of the SM count
// Grid caps of the big grid-stride kernels below (148 x 8 / 148 x 16 blocks): measured in round 2 with everything from one
// to six blocks per SM -- the training step does not move (10.7 .. 11.1 ms, inside the run-to-run spread): under the
// board's power cap the step is bound by the work, not by how the element-wise blocks interleave with the convolutions.
static inline int ew_cap(int dflt) { return dflt; }

// Merge type of the per-channel reductions.  bf16 (throughput) mode: fp32 partials merged with fp32 shared-memory
// atomics, one fp64 global atomic per block and channel.  fp32 (parity) mode: every partial is widened to fp64 BEFORE
// the first merge (warp shuffles, shared and global atomics all fp64), so the result depends on the run-dependent merge
// order only at the 1e-16 level -- the BN-backward sums are residuals of cancelling terms, and an fp32 merge in arrival
// order moved the gradients of identical runs by more than the parity bar.
template <typename T> struct MergeT { typedef float type; };
template <> struct MergeT<float> { typedef double type; };

// --------------------------------------------------------------------------------------------
// per-channel sum / sum of squares over [rows][C]
// --------------------------------------------------------------------------------------------
template <typename T, bool RELU>
__global__ void k_channel_stats_vec(const T* __restrict__ x, long long rows, int C, double* __restrict__ sum) {
  typedef typename MergeT<T>::type ACC;
  extern __shared__ __align__(8) unsigned char sh_raw[];
  ACC* sh = reinterpret_cast<ACC*>(sh_raw);  // 2*C
  const int groups = C >> 3;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = blockDim.x / groups;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = (ACC)0;
  __syncthreads();
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  for (long long r = (long long)blockIdx.x * lanes + lane; r < rows; r += (long long)gridDim.x * lanes) {
    float v[8];
    load8(x + r * C + g * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = RELU ? fmaxf(v[i], 0.f) : v[i];
      s1[i] += t;
      s2[i] += t * t;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    atomicAdd(&sh[g * 8 + i], (ACC)s1[i]);
    atomicAdd(&sh[C + g * 8 + i], (ACC)s2[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sum[i], (double)sh[i]);
}

template <typename T>
int launch_channel_stats(const T* x, long long rows, int C, int relu, double* sum, cudaStream_t s) {
  // (the 1- and 3-channel input tensors take their statistics inside launch_input_stage)
  L3_REQUIRE(C % 8 == 0 && kThreads % (C / 8) == 0, "channel_stats: unsupported C=%d", C);
  L3_CHECK_CUDA(cudaMemsetAsync(sum, 0, sizeof(double) * 2 * C, s));
  int lanes = kThreads / (C / 8);
  long long want = (rows + (long long)lanes * 16 - 1) / ((long long)lanes * 16);
  int blocks = (int)(want > kMaxBlocks ? kMaxBlocks : (want < 1 ? 1 : want));
  const size_t shb = 2 * C * sizeof(typename MergeT<T>::type);
  if (relu)
    k_channel_stats_vec<T, true><<<blocks, kThreads, shb, s>>>(x, rows, C, sum);
  else
    k_channel_stats_vec<T, false><<<blocks, kThreads, shb, s>>>(x, rows, C, sum);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_channel_stats<float>(const float*, long long, int, int, double*, cudaStream_t);
template int launch_channel_stats<bf16>(const bf16*, long long, int, int, double*, cudaStream_t);

// --------------------------------------------------------------------------------------------
// BN finalize: mean / biased var -> invstd, scale, shift ; moving-stat update (momentum 0.99, eps 1e-3)
// --------------------------------------------------------------------------------------------
__global__ void k_bn_finalize(BnRef bn, double count, int training, float momentum, float eps, int unbiased) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= bn.C) return;
  float mean, var;
  if (training) {
    double m = bn.sum[c] / count;
    double v = bn.sum[bn.C + c] / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    var = (float)v;
    double vu = (unbiased && count > 1.0) ? v * (count / (count - 1.0)) : v;
    bn.moving_mean[c] = bn.moving_mean[c] * momentum + mean * (1.f - momentum);
    bn.moving_var[c] = bn.moving_var[c] * momentum + (float)vu * (1.f - momentum);
  } else {
    mean = bn.moving_mean[c];
    var = bn.moving_var[c];
  }
  float inv = rsqrtf(var + eps);
  inv = inv * (1.5f - 0.5f * (var + eps) * inv * inv);  // one Newton step: full fp32 accuracy
  float sc = bn.gamma[c] * inv;
  bn.mean[c] = mean;
  bn.invstd[c] = inv;
  bn.scale[c] = sc;
  bn.shift[c] = bn.beta[c] - mean * sc;
}
int launch_bn_finalize(const BnRef& bn, long long count, int training, float momentum, float eps, int unbiased,
                       cudaStream_t s) {
  k_bn_finalize<<<ceil_div(bn.C, 128), 128, 0, s>>>(bn, (double)count, training, momentum, eps, unbiased);
  L3_CHECK_LAUNCH();
  return 0;
}

// --------------------------------------------------------------------------------------------
// Input stage of a tower (C = 1 audio map, C = 3 video frame), ONE pass over the elements:
//   MODE 0  v = x0[i]                                     (float input already in place)
//   MODE 1  v = 2 * (u8 / 255) - 1  -> x0[i]              (train.py:186; divide in fp64, fp32 affine, as skimage does)
//   MODE 2  v = max(x0[i] - clip max, -80) -> x0[i]       (kapre amplitude_to_decibel: per-clip maximum, 80 dB floor)
//   sum  != null: per-channel sum / sum of squares of v   (training-mode input BatchNorm), fp64 accumulators
//   xin  != null: v * scale[c] + shift[c] (or v) -> T in the zero-haloed layout (B,H+2,W+2,C) every convolution reads
// Round 2: replaces four kernels (u8 -> float, dB finish, statistics, affine) that re-read the tensor three times and
// decomposed every flat index with 64-bit divisions (the affine alone took 62 us per 64 frames: 15x its HBM time).
// --------------------------------------------------------------------------------------------
static const int kInThreads = 192;   // a multiple of 3: a thread's grid-stride elements all belong to ONE channel
template <typename T, int C, int MODE>
__global__ void __launch_bounds__(kInThreads)
k_input_stage(const uint8_t* __restrict__ u8, float* __restrict__ x0, T* __restrict__ xin, unsigned n, int H, int W,
              FastDiv d_row, FastDiv d_h, FastDiv d_clip, const float* __restrict__ scale, const float* __restrict__ shift,
              const int* __restrict__ clip_max, double* __restrict__ sum) {
  __shared__ double sh[2 * C];
  if (threadIdx.x < 2 * C) sh[threadIdx.x] = 0.0;
  __syncthreads();
  const int c = (C == 1) ? 0 : (int)(threadIdx.x % C);
  float sc = 1.f, sf = 0.f;
  if (xin != nullptr && scale != nullptr) { sc = scale[c]; sf = shift[c]; }
  double s1 = 0.0, s2 = 0.0;
  const unsigned stride = gridDim.x * kInThreads;   // multiple of C
  constexpr int U = 4;                              // independent loads in flight per thread
  for (unsigned i0 = blockIdx.x * kInThreads + threadIdx.x; i0 < n; i0 += U * stride) {
    float v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const unsigned i = i0 + k * stride;
      v[k] = 0.f;
      if (i < n) {
        if (MODE == 1) v[k] = (float)u8[i];
        else v[k] = x0[i];
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const unsigned i = i0 + k * stride;
      if (i >= n) continue;
      float x = v[k];
      if (MODE == 1) {
        x = 2.0f * (float)((double)x / 255.0) - 1.0f;
        x0[i] = x;
      } else if (MODE == 2) {
        x = fmaxf(x - ordered_to_float(clip_max[fdiv(i, d_clip)]), -80.0f);
        x0[i] = x;
      }
      if (sum != nullptr) {
        s1 += (double)x;
        s2 += (double)x * (double)x;
      }
      if (xin != nullptr) {
        const unsigned row = fdiv(i, d_row), e = i - row * d_row.d;   // row = b * H + y, e = x * C + c
        const unsigned b = fdiv(row, d_h), y = row - b * d_h.d;
        const float o = (scale != nullptr) ? x * sc + sf : x;
        xin[pad_off(b, (int)y, 0, H, W, C) + e] = from_f<T>(o);
      }
    }
  }
  if (sum != nullptr) {
    atomicAdd(&sh[c], s1);
    atomicAdd(&sh[C + c], s2);
    __syncthreads();
    if (threadIdx.x < 2 * C) atomicAdd(&sum[threadIdx.x], sh[threadIdx.x]);
  }
}
template <typename T>
int launch_input_stage(int mode, const uint8_t* u8, float* x0, T* xin, int B, int H, int W, int C, const float* scale,
                       const float* shift, const int* clip_max, double* sum, cudaStream_t s) {
  L3_REQUIRE(C == 1 || C == 3, "input stage: C=%d", C);
  L3_REQUIRE(mode >= 0 && mode <= 2 && (mode != 1 || u8 != nullptr) && (mode != 2 || clip_max != nullptr), "input stage: mode");
  const long long n = (long long)B * H * W * C;
  L3_REQUIRE(n < 0x7fffffffLL, "input stage: too many elements");
  if (sum) L3_CHECK_CUDA(cudaMemsetAsync(sum, 0, sizeof(double) * 2 * C, s));
  long long want = (n + kInThreads * 8 - 1) / (kInThreads * 8);
  const int blocks = (int)(want > kMaxBlocks ? kMaxBlocks : (want < 1 ? 1 : want));
  const FastDiv d_row = make_fastdiv((uint32_t)(W * C)), d_h = make_fastdiv((uint32_t)H);
  const FastDiv d_clip = make_fastdiv((uint32_t)(H * W * C));
#define L3_IN(c_, m_) k_input_stage<T, c_, m_><<<blocks, kInThreads, 0, s>>>(u8, x0, xin, (unsigned)n, H, W, d_row, d_h, d_clip, \
                                                                          scale, shift, clip_max, sum)
  if (C == 1) { if (mode == 0) L3_IN(1, 0); else if (mode == 1) L3_IN(1, 1); else L3_IN(1, 2); }
  else { if (mode == 0) L3_IN(3, 0); else if (mode == 1) L3_IN(3, 1); else L3_IN(3, 2); }
#undef L3_IN
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_input_stage<float>(int, const uint8_t*, float*, float*, int, int, int, int, const float*, const float*,
                                       const int*, double*, cudaStream_t);
template int launch_input_stage<bf16>(int, const uint8_t*, float*, bf16*, int, int, int, int, const float*, const float*,
                                      const int*, double*, cudaStream_t);

// zero the one-pixel halo of a padded (B,H+2,W+2,C) buffer (buffers re-used at several geometries); one 16-byte
// store per thread when a pixel's channels are a multiple of 16 bytes (every layer with C >= 64)
template <typename T, int VEC>
__global__ void k_zero_halo(T* __restrict__ buf, int B, int H, int W, int C) {
  const int per_img = 2 * (W + 2) + 2 * H;  // halo pixels per image
  const int cv = C / VEC;
  const unsigned total = (unsigned)B * per_img * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * VEC;
    const unsigned q = i / cv;
    const int h = (int)(q % per_img);
    const long long b = q / per_img;
    int yp, xp;
    if (h < W + 2) { yp = 0; xp = h; }
    else if (h < 2 * (W + 2)) { yp = H + 1; xp = h - (W + 2); }
    else { int r = h - 2 * (W + 2); yp = 1 + (r >> 1); xp = (r & 1) ? W + 1 : 0; }
    T* dst = buf + ((b * (H + 2) + yp) * (long long)(W + 2) + xp) * C + c;
    if (VEC == 1) *dst = from_f<T>(0.f);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}
template <typename T>
int launch_zero_halo(T* buf, int B, int H, int W, int C, cudaStream_t s) {
  constexpr int V = 16 / (int)sizeof(T);
  const bool vec = (C % V == 0) && (((uintptr_t)buf & 15) == 0);
  long long total = (long long)B * (2 * (W + 2) + 2 * H) * (vec ? C / V : C);
  L3_REQUIRE(total < 0x7fffffffLL, "zero_halo: too many elements");
  long long want = (total + kThreads - 1) / kThreads;
  int blocks = (int)(want > kMaxBlocks ? kMaxBlocks : (want < 1 ? 1 : want));
  if (vec) k_zero_halo<T, V><<<blocks, kThreads, 0, s>>>(buf, B, H, W, C);
  else k_zero_halo<T, 1><<<blocks, kThreads, 0, s>>>(buf, B, H, W, C);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_zero_halo<float>(float*, int, int, int, int, cudaStream_t);
template int launch_zero_halo<bf16>(bf16*, int, int, int, int, cudaStream_t);

// --------------------------------------------------------------------------------------------
// activation forward: a = pool2x2?( relu_first ? relu(z)*s+t : relu(z*s+t) );  z unpadded, a zero-haloed padded
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void act8(const float (&z)[8], const float (&sc)[8], const float (&sh)[8], int relu_first,
                                     float (&y)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    y[i] = relu_first ? fmaxf(z[i], 0.f) * sc[i] + sh[i] : fmaxf(z[i] * sc[i] + sh[i], 0.f);
}

// Several independent 16-byte loads are issued per thread before any is consumed (U pixels, or U 2x2 windows): next to
// a resident tensor-core CTA of the other tower's stream only ONE of these blocks fits on an SM, and its bytes in
// flight -- not the block count -- then set the achieved HBM bandwidth.
// REC (training, pooled layers): also records, per pooled element, which window position won and whether the winner is
// positive (`sel` = position | 4 * (max > 0), one byte) and the winning pre-activation (`zsel`).  The backward kernels
// then neither recompute the four activations nor -- for the statistics pass -- read the full-resolution z: they were
// instruction-issue-bound (~450 instructions per 2x2 window) and moved 2.5x the bytes.
template <typename T, bool POOL, bool REC>
__global__ void __launch_bounds__(256, 3)
k_act_fwd(const T* __restrict__ z, T* __restrict__ a, int H, int W, int C, int OH, int OW, long long npix,
          const float* __restrict__ scale, const float* __restrict__ shift, int relu_first, T* __restrict__ zsel,
          uint8_t* __restrict__ sel) {
  const int groups = C >> 3;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = blockDim.x / groups;
  float sc[8], sh[8];
  load8(scale + g * 8, sc);
  load8(shift + g * 8, sh);
  // fp32 (parity mode): half as many, the raw loads are twice as wide; the recording variant keeps one window in flight
  // (its argmax / winner registers would spill otherwise)
  constexpr int U = REC ? 1 : (POOL ? 2 : 4) / (sizeof(T) == 4 ? 2 : 1);
  const long long stride = (long long)gridDim.x * lanes;
  for (long long p0 = (long long)blockIdx.x * (lanes * U) + lane; p0 < npix; p0 += stride * U) {   // block = U*lanes consecutive pixels
    Raw8<T> v[U][POOL ? 4 : 1];
    long long dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * lanes;
      dst[u] = -1;
      if (p < npix) {
        // 32-bit index math (npix < 2^31): 64-bit divisions cost ~100 instructions each in these issue-bound kernels
        const unsigned pr = (unsigned)p / (unsigned)OW;
        const int ox = (int)((unsigned)p - pr * (unsigned)OW);
        const long long b = pr / (unsigned)OH;
        const int oy = (int)(pr - (unsigned)b * (unsigned)OH);
        dst[u] = pad_off(b, oy, ox, OH, OW, C) + g * 8;
        if (!POOL) {
          ldraw(z + p * C + g * 8, v[u][0]);
        } else {
          const T* z00 = z + ((b * H + 2 * oy) * W + 2 * ox) * C + g * 8;
          ldraw(z00, v[u][0]);
          ldraw(z00 + C, v[u][POOL ? 1 : 0]);
          ldraw(z00 + (long long)W * C, v[u][POOL ? 2 : 0]);
          ldraw(z00 + (long long)W * C + C, v[u][POOL ? 3 : 0]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (dst[u] < 0) continue;
      float y[8], x[8];
      unraw(v[u][0], x);
      act8(x, sc, sh, relu_first, y);
      if (POOL && REC) {
        // first maximum in window order (0,0),(0,1),(1,0),(1,1) -- the rule the backward pass used to re-derive
        float zs[8];
        int arg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { zs[i] = x[i]; arg[i] = 0; }
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          float t[8];
          unraw(v[u][POOL ? k : 0], x);
          act8(x, sc, sh, relu_first, t);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (t[i] > y[i]) { y[i] = t[i]; zs[i] = x[i]; arg[i] = k; }
        }
        uint32_t s01 = 0, s23 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s01 |= (uint32_t)(arg[i] | (y[i] > 0.f ? 4 : 0)) << (8 * i);
          s23 |= (uint32_t)(arg[i + 4] | (y[i + 4] > 0.f ? 4 : 0)) << (8 * i);
        }
        const long long po = (p0 + u * lanes) * C + g * 8;   // un-padded pooled element offset
        store8(zsel + po, zs);
        *reinterpret_cast<uint2*>(sel + po) = make_uint2(s01, s23);
      } else if (POOL) {
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          float t[8];
          unraw(v[u][POOL ? k : 0], x);
          act8(x, sc, sh, relu_first, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = fmaxf(y[i], t[i]);
        }
      }
      store8(a + dst[u], y);
    }
  }
}
template <typename T>
int launch_act_fwd(const T* z, T* a, int B, int H, int W, int C, const float* scale, const float* shift, int pool,
                   int relu_first, cudaStream_t s, T* zsel, uint8_t* sel) {
  L3_REQUIRE(C % 8 == 0 && kThreads % (C / 8) == 0, "act_fwd: C=%d", C);
  int OH = pool ? H / 2 : H, OW = pool ? W / 2 : W;
  long long npix = (long long)B * OH * OW;
  L3_REQUIRE(npix < 0x7fffffffLL, "act_fwd: too many pixels");
  int lanes = kThreads / (C / 8);
  long long want = (npix + (long long)lanes * 4 - 1) / ((long long)lanes * 4);
  const int cap = ew_cap(148 * 16);
  int blocks = (int)(want > cap ? cap : (want < 1 ? 1 : want));
  if (pool && zsel && sel)
    k_act_fwd<T, true, true><<<blocks, kThreads, 0, s>>>(z, a, H, W, C, OH, OW, npix, scale, shift, relu_first, zsel, sel);
  else if (pool)
    k_act_fwd<T, true, false><<<blocks, kThreads, 0, s>>>(z, a, H, W, C, OH, OW, npix, scale, shift, relu_first, nullptr, nullptr);
  else
    k_act_fwd<T, false, false><<<blocks, kThreads, 0, s>>>(z, a, H, W, C, OH, OW, npix, scale, shift, relu_first, nullptr, nullptr);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_act_fwd<float>(const float*, float*, int, int, int, int, const float*, const float*, int, int, cudaStream_t,
                                   float*, uint8_t*);
template int launch_act_fwd<bf16>(const bf16*, bf16*, int, int, int, int, const float*, const float*, int, int, cudaStream_t,
                                  bf16*, uint8_t*);

// --------------------------------------------------------------------------------------------
// global max-pool forward over relu(bn(z)) -> (B,C) float + argmax pixel (first max in row-major order)
// --------------------------------------------------------------------------------------------
// Stage 1: grid (B, slices); every block reduces its pixel slice per channel and merges with a 64-bit atomicMax of
// (value bits << 32 | ~pixel index): values are relu outputs (>= 0, so their bit patterns order like the floats) and
// among equal values the smallest pixel index wins (first maximum in row-major order, like keras/TF max-pool).
template <typename T>
__global__ void k_gmaxpool_fwd(const T* __restrict__ z, int HW, int C, const float* __restrict__ scale,
                               const float* __restrict__ shift, unsigned long long* __restrict__ best_out) {
  const int groups = C >> 3;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = blockDim.x / groups;
  const int b = blockIdx.x;
  const int per = (HW + gridDim.y - 1) / gridDim.y;
  const int p0 = blockIdx.y * per, p1 = min(HW, p0 + per);
  float sc[8], sh[8], best[8];
  int bi[8];
  load8(scale + g * 8, sc);
  load8(shift + g * 8, sh);
#pragma unroll
  for (int i = 0; i < 8; ++i) { best[i] = -1.f; bi[i] = 0; }
  for (int p = p0 + lane; p < p1; p += lanes) {
    float v[8], y[8];
    load8(z + ((long long)b * HW + p) * C + g * 8, v);
    act8(v, sc, sh, 0, y);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (y[i] > best[i]) { best[i] = y[i]; bi[i] = p; }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (best[i] >= 0.f) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(best[i]) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)bi[i]);
      atomicMax(&best_out[(long long)b * C + g * 8 + i], key);
    }
  }
}
__global__ void k_gmaxpool_fwd_finish(const unsigned long long* __restrict__ best, int n, int C, float* __restrict__ out,
                                      int out_stride, int* __restrict__ argmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = best[i];
  out[(long long)(i / C) * out_stride + (i % C)] = __uint_as_float((unsigned)(key >> 32));
  argmax[i] = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
}
template <typename T>
int launch_gmaxpool_fwd(const T* z, int B, int HW, int C, const float* scale, const float* shift, float* out,
                        int out_stride, int* argmax, unsigned long long* scratch, cudaStream_t s) {
  L3_REQUIRE(C % 8 == 0 && kThreads % (C / 8) == 0, "gmaxpool: C=%d", C);
  L3_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long) * (size_t)B * C, s));
  int slices = (148 * 4 + B - 1) / B;
  if (slices > 16) slices = 16;
  if (slices < 1) slices = 1;
  dim3 grid(B, slices);
  k_gmaxpool_fwd<T><<<grid, kThreads, 0, s>>>(z, HW, C, scale, shift, scratch);
  L3_CHECK_LAUNCH();
  k_gmaxpool_fwd_finish<<<ceil_div((long long)B * C, 256), 256, 0, s>>>(scratch, B * C, C, out, out_stride, argmax);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_gmaxpool_fwd<float>(const float*, int, int, int, const float*, const float*, float*, int, int*, unsigned long long*, cudaStream_t);
template int launch_gmaxpool_fwd<bf16>(const bf16*, int, int, int, const float*, const float*, float*, int, int*, unsigned long long*, cudaStream_t);

// global max-pool backward: dy = scatter(dpool) masked by relu, plus the BN-backward sums.  One thread per
// (sample, channel); the per-channel sums go through double atomics (B values per channel).
template <typename T>
__global__ void k_gmaxpool_bwd(const float* __restrict__ dpool, int dpool_stride, const int* __restrict__ argmax,
                               const T* __restrict__ z, T* __restrict__ dy, BnRef bn, int B, int H, int W, int C) {
  const int HW = H * W;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int c = idx % C, b = idx / C;
  const float sc = bn.scale[c], sh = bn.shift[c], mean = bn.mean[c], inv = bn.invstd[c];
  const int p = argmax[b * C + c];
  const float zz = to_f(z[((long long)b * HW + p) * C + c]);
  const float g = dpool[(long long)b * dpool_stride + c];
  if (zz * sc + sh > 0.f) {
    const T gt = from_f<T>(g);
    dy[pad_off(b, p / W, p % W, H, W, C) + c] = gt;
    const float gq = to_f(gt);
    atomicAdd(&bn.sum[c], (double)gq);
    atomicAdd(&bn.sum[C + c], (double)(gq * ((zz - mean) * inv)));
  }
}
template <typename T>
int launch_gmaxpool_bwd(const float* dpool, int dpool_stride, const int* argmax, const T* z, T* dy, const BnRef& bn,
                        int B, int H, int W, int C, cudaStream_t s) {
  L3_CHECK_CUDA(cudaMemsetAsync(dy, 0, sizeof(T) * (size_t)B * (H + 2) * (W + 2) * C, s));
  L3_CHECK_CUDA(cudaMemsetAsync(bn.sum, 0, sizeof(double) * 2 * C, s));
  k_gmaxpool_bwd<T><<<ceil_div((long long)B * C, 256), 256, 0, s>>>(dpool, dpool_stride, argmax, z, dy, bn, B, H, W, C);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_gmaxpool_bwd<float>(const float*, int, const int*, const float*, float*, const BnRef&, int, int, int, int, cudaStream_t);
template int launch_gmaxpool_bwd<bf16>(const float*, int, const int*, const bf16*, bf16*, const BnRef&, int, int, int, int, cudaStream_t);

// --------------------------------------------------------------------------------------------
// activation + BatchNorm backward, two passes over (da, z) with no intermediate tensor:
//   normal     : a = pool(relu(bn(z)))      dy = unpool(da) * [bn(z) > 0]      xhat from z
//   relu_first : a = pool(bn(relu(z)))      dy = unpool(da)                    xhat from relu(z)
//   pass 1 (k_bwd_stats): sum(dy), sum(dy*xhat) per channel                     -> bn.sum
//   pass 2 (k_bwd_apply): dz = scale*(dy - c1 - xhat*c2) [*(z>0) if relu_first] -> zero-haloed padded buffer
// max-pool routes to the first maximum in window order (0,0),(0,1),(1,0),(1,1); pixels dropped by 'valid' pooling of
// odd sizes have dy = 0 (they still receive the BN mean terms).  da: unpadded (B,OH,OW,C); z: unpadded (B,H,W,C).
// --------------------------------------------------------------------------------------------
// Accumulates the RAW sums S1 = sum(dy) and S2 = sum(dy * xin) (xin = z, or relu(z) when relu_first); the
// normalised sum the BN backward needs is sum(dy*xhat) = invstd * (S2 - mean * S1), formed in double precision by
// k_bn_bwd_finalize.  (Fewer instructions per element: these kernels are issue-bound, not HBM-bound.)
template <typename T, bool POOL>
__global__ void __launch_bounds__(256, 2)
k_bwd_stats(const T* __restrict__ da, const T* __restrict__ z, int H, int W, int C, int OH, int OW, long long npix,
            BnRef bn, int relu_first) {
  typedef typename MergeT<T>::type ACC;
  extern __shared__ __align__(8) unsigned char sh_raw[];
  ACC* sh = reinterpret_cast<ACC*>(sh_raw);  // 2*C
  const int groups = C >> 3;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = blockDim.x / groups;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = (ACC)0;
  __syncthreads();
  // fp32 (parity) mode subtracts the batch mean before accumulating (no cancellation in S2 - mean*S1: the
  // near-zero input-BN beta gradient is sensitive to it); bf16 mode keeps the raw products.
  constexpr bool kCentre = sizeof(T) == 4;
  float sc[8], sf[8], mu[8], s1[8], s2[8];
  load8(bn.scale + g * 8, sc);
  load8(bn.shift + g * 8, sf);
#pragma unroll
  for (int i = 0; i < 8; ++i) mu[i] = 0.f;
  if (kCentre) load8(bn.mean + g * 8, mu);
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  const long long stride = (long long)gridDim.x * lanes;
  if (!POOL) {
    constexpr int U = sizeof(T) == 4 ? 2 : 4;   // 8 independent 16-byte loads in flight per thread (see k_act_fwd)
    for (long long p0 = (long long)blockIdx.x * (lanes * U) + lane; p0 < npix; p0 += stride * U) {   // block = U*lanes consecutive pixels
      Raw8<T> rg[U], rz[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = p0 + u * lanes;
        ok[u] = p < npix;
        if (ok[u]) {
          ldraw(da + p * C + g * 8, rg[u]);
          ldraw(z + p * C + g * 8, rz[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        float g8[8], zs[8];
        unraw(rg[u], g8);
        unraw(rz[u], zs);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float zz = zs[i];
          // normal: dy = da where bn(z) > 0, xin = z ; relu_first: dy = da, xin = relu(z)
          const float d = relu_first ? g8[i] : (fmaf(zz, sc[i], sf[i]) > 0.f ? g8[i] : 0.f);
          float xin = relu_first ? fmaxf(zz, 0.f) : zz;
          if (kCentre) xin -= mu[i];
          s1[i] += d;
          s2[i] = fmaf(d, xin, s2[i]);
        }
      }
    }
  } else {
    // software-pipelined: the five raw loads of the next 2x2 window are in flight while the current one is reduced
    // (these kernels sat on their first use of the loaded data at ~23 % occupancy)
    auto fetch = [&](long long p, Raw8<T>& rg, Raw8<T> (&rz)[4]) {
      ldraw(da + p * C + g * 8, rg);
      const unsigned pr = (unsigned)p / (unsigned)OW;
      const int ox = (int)((unsigned)p - pr * (unsigned)OW);
      const long long b = pr / (unsigned)OH;
      const int oy = (int)(pr - (unsigned)b * (unsigned)OH);
      const T* z00 = z + ((b * H + 2 * oy) * W + 2 * ox) * C + g * 8;
      ldraw(z00, rz[0]);
      ldraw(z00 + C, rz[1]);
      ldraw(z00 + (long long)W * C, rz[2]);
      ldraw(z00 + (long long)W * C + C, rz[3]);
    };
    long long p = (long long)blockIdx.x * lanes + lane;
    Raw8<T> cg, cz[4], ng, nz[4];
    if (p < npix) fetch(p, cg, cz);
    while (p < npix) {
      const long long pn = p + stride;
      if (pn < npix) fetch(pn, ng, nz);
      float g8[8], zs[8], m[8], y[8], x[8];
      unraw(cg, g8);
      unraw(cz[0], zs);
      act8(zs, sc, sf, relu_first, m);
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        unraw(cz[q], x);
        act8(x, sc, sf, relu_first, y);
#pragma unroll
        for (int i = 0; i < 8; ++i) if (y[i] > m[i]) { m[i] = y[i]; zs[i] = x[i]; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = relu_first ? g8[i] : (m[i] > 0.f ? g8[i] : 0.f);
        float xin = relu_first ? fmaxf(zs[i], 0.f) : zs[i];
        if (kCentre) xin -= mu[i];
        s1[i] += d;
        s2[i] = fmaf(d, xin, s2[i]);
      }
      cg = ng;
#pragma unroll
      for (int q = 0; q < 4; ++q) cz[q] = nz[q];
      p = pn;
    }
  }
  // lanes of a warp that share a channel group (tid % groups) are `groups` apart: fold them first
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ACC a1 = (ACC)s1[i], a2 = (ACC)s2[i];
    for (int off = 16; off >= groups; off >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if ((threadIdx.x & 31) < groups || groups >= 32) {
      atomicAdd(&sh[g * 8 + i], a1);
      atomicAdd(&sh[C + g * 8 + i], a2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&bn.sum[i], (double)sh[i]);
}
// pooled layers with the forward pass's record: dy = da where the window maximum was positive (always, if relu_first),
// xin = the winning pre-activation -- two 16-byte loads and 8 selection bytes per 8 channels of a pooled pixel
template <typename T>
__global__ void __launch_bounds__(256, 2)
k_bwd_stats_sel(const T* __restrict__ da, const T* __restrict__ zsel, const uint8_t* __restrict__ sel, int C, long long npix,
                BnRef bn, int relu_first) {
  typedef typename MergeT<T>::type ACC;
  extern __shared__ __align__(8) unsigned char sh_raw[];
  ACC* sh = reinterpret_cast<ACC*>(sh_raw);  // 2*C
  const int groups = C >> 3;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = blockDim.x / groups;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = (ACC)0;
  __syncthreads();
  constexpr bool kCentre = sizeof(T) == 4;   // see k_bwd_stats
  float mu[8], s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mu[i] = s1[i] = s2[i] = 0.f;
  if (kCentre) load8(bn.mean + g * 8, mu);
  constexpr int U = sizeof(T) == 4 ? 2 : 4;
  const long long stride = (long long)gridDim.x * lanes;
  for (long long p0 = (long long)blockIdx.x * (lanes * U) + lane; p0 < npix; p0 += stride * U) {
    Raw8<T> rg[U], rz[U];
    uint2 rs[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * lanes;
      ok[u] = p < npix;
      if (ok[u]) {
        ldraw(da + p * C + g * 8, rg[u]);
        ldraw(zsel + p * C + g * 8, rz[u]);
        rs[u] = *reinterpret_cast<const uint2*>(sel + p * C + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      float g8[8], zs[8];
      unraw(rg[u], g8);
      unraw(rz[u], zs);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t sb = ((i < 4 ? rs[u].x : rs[u].y) >> (8 * (i & 3))) & 0xffu;
        const float d = (relu_first || (sb & 4u)) ? g8[i] : 0.f;
        float xin = relu_first ? fmaxf(zs[i], 0.f) : zs[i];
        if (kCentre) xin -= mu[i];
        s1[i] += d;
        s2[i] = fmaf(d, xin, s2[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ACC a1 = (ACC)s1[i], a2 = (ACC)s2[i];
    for (int off = 16; off >= groups; off >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if ((threadIdx.x & 31) < groups || groups >= 32) {
      atomicAdd(&sh[g * 8 + i], a1);
      atomicAdd(&sh[C + g * 8 + i], a2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&bn.sum[i], (double)sh[i]);
}

template <typename T>
int launch_bwd_stats(const T* da, const T* z, int B, int H, int W, int C, const BnRef& bn, int pool, int relu_first,
                     cudaStream_t s, const T* zsel, const uint8_t* sel) {
  if (pool && zsel && sel) {
    L3_REQUIRE(C % 8 == 0 && kThreads % (C / 8) == 0, "bwd_stats: C=%d", C);
    L3_CHECK_CUDA(cudaMemsetAsync(bn.sum, 0, sizeof(double) * 2 * C, s));
    const long long npix = (long long)B * (H / 2) * (W / 2);
    const int lanes = kThreads / (C / 8);
    long long want = (npix + (long long)lanes * 4 - 1) / ((long long)lanes * 4);
    const int cap = ew_cap(148 * 8);
    int blocks = (int)(want > cap ? cap : (want < 1 ? 1 : want));
    k_bwd_stats_sel<T><<<blocks, kThreads, 2 * C * sizeof(typename MergeT<T>::type), s>>>(da, zsel, sel, C, npix, bn, relu_first);
    L3_CHECK_LAUNCH();
    return 0;
  }
  // The 2x2-pool variants need ~128 registers: in 128-thread blocks (16 K registers) they still fit next to a resident
  // tensor-core CTA (38 K registers) of the other tower's stream; a 256-thread block would not.
  const int threads = pool ? 128 : kThreads;
  L3_REQUIRE(C % 8 == 0 && threads % (C / 8) == 0, "bwd_stats: C=%d", C);
  L3_CHECK_CUDA(cudaMemsetAsync(bn.sum, 0, sizeof(double) * 2 * C, s));
  int OH = pool ? H / 2 : H, OW = pool ? W / 2 : W;
  long long npix = (long long)B * OH * OW;
  int lanes = threads / (C / 8);
  long long want = (npix + (long long)lanes * 4 - 1) / ((long long)lanes * 4);
  const int cap = ew_cap(148 * 8);
  int blocks = (int)(want > cap ? cap : (want < 1 ? 1 : want));
  const size_t shb = 2 * C * sizeof(typename MergeT<T>::type);
  if (pool) k_bwd_stats<T, true><<<blocks, threads, shb, s>>>(da, z, H, W, C, OH, OW, npix, bn, relu_first);
  else k_bwd_stats<T, false><<<blocks, kThreads, shb, s>>>(da, z, H, W, C, OH, OW, npix, bn, relu_first);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_bwd_stats<float>(const float*, const float*, int, int, int, int, const BnRef&, int, int, cudaStream_t,
                                     const float*, const uint8_t*);
template int launch_bwd_stats<bf16>(const bf16*, const bf16*, int, int, int, int, const BnRef&, int, int, cudaStream_t,
                                    const bf16*, const uint8_t*);

// folded BN-backward coefficients (written by k_bn_bwd_finalize into bn.c1 / bn.c2):
//   dz = scale*(dy - mean(dy) - xhat*mean(dy*xhat)) = scale*dy + c1*xin + c2,   xin = z (or relu(z) if relu_first)
struct BnBwdCoef {
  float sc[8], cb[8], cc[8];
};
template <typename T>
__device__ __forceinline__ void bwd_emit(T* __restrict__ dst, const float (&v)[8], const float (&d)[8], const BnBwdCoef& k,
                                         int relu_first) {
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float xin = relu_first ? fmaxf(v[i], 0.f) : v[i];
    float r = fmaf(k.sc[i], d[i], fmaf(k.cb[i], xin, k.cc[i]));
    if (relu_first && !(v[i] > 0.f)) r = 0.f;
    o[i] = r;
  }
  store8(dst, o);
}

template <typename T, bool POOL, bool SEL>
__global__ void __launch_bounds__(256, 2)
k_bwd_apply(const T* __restrict__ da, const T* __restrict__ z, T* __restrict__ dz, int H, int W, int C, int OH, int OW,
            long long npix, BnRef bn, int relu_first, const uint8_t* __restrict__ sel) {
  const int groups = C >> 3;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = blockDim.x / groups;
  float sf[8];
  BnBwdCoef k;
  load8(bn.scale + g * 8, k.sc);
  if (!SEL) load8(bn.shift + g * 8, sf);   // only the re-derivation of the ReLU mask / pool routing needs the shift
  load8(bn.c1 + g * 8, k.cb);
  load8(bn.c2 + g * 8, k.cc);
  const float zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long p = (long long)blockIdx.x * lanes + lane; p < npix; p += (long long)gridDim.x * lanes) {
    // 32-bit index math (npix < 2^31): 64-bit divisions cost ~100 instructions each in these issue-bound kernels
    const unsigned pr = (unsigned)p / (unsigned)OW;
    const int ox = (int)((unsigned)p - pr * (unsigned)OW);
    const long long b = pr / (unsigned)OH;
    const int oy = (int)(pr - (unsigned)b * (unsigned)OH);
    float g8[8];
    load8(da + p * C + g * 8, g8);
    if (!POOL) {
      float v[8], y[8], d[8];
      load8(z + p * C + g * 8, v);
      act8(v, k.sc, sf, relu_first, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = relu_first ? g8[i] : (y[i] > 0.f ? g8[i] : 0.f);
      bwd_emit<T>(dz + pad_off(b, oy, ox, H, W, C) + g * 8, v, d, k, relu_first);
      continue;
    }
    float v[4][8], m[8];
    int arg[8];
    {
      const T* z00 = z + ((b * H + 2 * oy) * W + 2 * ox) * C + g * 8;
      load8(z00, v[0]);
      load8(z00 + C, v[1]);
      load8(z00 + (long long)W * C, v[2]);
      load8(z00 + (long long)W * C + C, v[3]);
      if (SEL) {
        // the forward pass recorded the winning window position and the sign of the maximum (k_act_fwd<.., REC>)
        const uint2 sb = *reinterpret_cast<const uint2*>(sel + p * C + g * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t b8 = ((i < 4 ? sb.x : sb.y) >> (8 * (i & 3))) & 0xffu;
          arg[i] = (int)(b8 & 3u);
          m[i] = (b8 & 4u) ? 1.f : 0.f;
        }
      } else {
        float y[8];
        act8(v[0], k.sc, sf, relu_first, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) arg[i] = 0;
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          act8(v[q], k.sc, sf, relu_first, y);
#pragma unroll
          for (int i = 0; i < 8; ++i) if (y[i] > m[i]) { m[i] = y[i]; arg[i] = q; }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = (arg[i] == q && (relu_first || m[i] > 0.f)) ? g8[i] : 0.f;
      bwd_emit<T>(dz + pad_off(b, 2 * oy + (q >> 1), 2 * ox + (q & 1), H, W, C) + g * 8, v[q], d, k, relu_first);
    }
    // rows / columns dropped by valid pooling of odd sizes: dy = 0
    const bool last_x = (W & 1) && ox == OW - 1, last_y = (H & 1) && oy == OH - 1;
    if (last_x) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        load8(z + ((b * H + 2 * oy + r) * W + W - 1) * C + g * 8, v[0]);
        bwd_emit<T>(dz + pad_off(b, 2 * oy + r, W - 1, H, W, C) + g * 8, v[0], zero8, k, relu_first);
      }
    }
    if (last_y) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        load8(z + ((b * H + H - 1) * W + 2 * ox + r) * C + g * 8, v[0]);
        bwd_emit<T>(dz + pad_off(b, H - 1, 2 * ox + r, H, W, C) + g * 8, v[0], zero8, k, relu_first);
      }
      if (last_x) {
        load8(z + ((b * H + H - 1) * W + W - 1) * C + g * 8, v[0]);
        bwd_emit<T>(dz + pad_off(b, H - 1, W - 1, H, W, C) + g * 8, v[0], zero8, k, relu_first);
      }
    }
  }
}
// (Measured: more loads in flight per thread -- 2 / 4 pixels per iteration, or prefetching the next 2x2 window --
// makes this write-heavy kernel SLOWER; it wants resident warps: 4 blocks / SM at 63 registers.)
// dz: zero-haloed padded (B,H+2,W+2,C); the halo is (re)zeroed here because the buffer is shared between layers
template <typename T>
int launch_bwd_apply(const T* da, const T* z, T* dz, int B, int H, int W, int C, const BnRef& bn, int pool,
                     int relu_first, cudaStream_t s, const uint8_t* sel) {
  const int threads = pool ? 128 : kThreads;   // see launch_bwd_stats
  L3_REQUIRE(C % 8 == 0 && threads % (C / 8) == 0, "bwd_apply: C=%d", C);
  if (launch_zero_halo<T>(dz, B, H, W, C, s)) return -1;
  int OH = pool ? H / 2 : H, OW = pool ? W / 2 : W;
  long long npix = (long long)B * OH * OW;
  int lanes = threads / (C / 8);
  long long want = (npix + (long long)lanes * 2 - 1) / ((long long)lanes * 2);
  const int cap = ew_cap(148 * 16);
  int blocks = (int)(want > cap ? cap : (want < 1 ? 1 : want));
  if (pool && sel) k_bwd_apply<T, true, true><<<blocks, threads, 0, s>>>(da, z, dz, H, W, C, OH, OW, npix, bn, relu_first, sel);
  else if (pool) k_bwd_apply<T, true, false><<<blocks, threads, 0, s>>>(da, z, dz, H, W, C, OH, OW, npix, bn, relu_first, nullptr);
  else k_bwd_apply<T, false, false><<<blocks, kThreads, 0, s>>>(da, z, dz, H, W, C, OH, OW, npix, bn, relu_first, nullptr);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_bwd_apply<float>(const float*, const float*, float*, int, int, int, int, const BnRef&, int, int, cudaStream_t,
                                     const uint8_t*);
template int launch_bwd_apply<bf16>(const bf16*, const bf16*, bf16*, int, int, int, int, const BnRef&, int, int, cudaStream_t,
                                    const uint8_t*);

// BN backward finalize: dgamma, dbeta and the folded coefficients of dz = scale*dy + c1*xin + c2
//   c1 = -scale*invstd*mean(dy*xhat) ; c2 = -scale*mean(dy) - c1*mean
__global__ void k_bn_bwd_finalize(BnRef bn, double count, int raw) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= bn.C) return;
  double s1 = bn.sum[c], s2 = bn.sum[bn.C + c];
  // raw 1: sum(dy*xin) ; raw 2: sum(dy*(xin-mean))   ->  sum(dy*xhat)
  if (raw == 1) s2 = (double)bn.invstd[c] * (s2 - (double)bn.mean[c] * s1);
  else if (raw == 2) s2 = (double)bn.invstd[c] * s2;
  bn.d_beta[c] = (float)s1;
  bn.d_gamma[c] = (float)s2;
  const double sc = (double)bn.scale[c];
  const double cb = -sc * (double)bn.invstd[c] * (s2 / count);
  bn.c1[c] = (float)cb;
  bn.c2[c] = (float)(-sc * (s1 / count) - cb * (double)bn.mean[c]);
}
int launch_bn_bwd_finalize(const BnRef& bn, long long count, int raw_sums, cudaStream_t s) {
  k_bn_bwd_finalize<<<ceil_div(bn.C, 128), 128, 0, s>>>(bn, (double)count, raw_sums);
  L3_CHECK_LAUNCH();
  return 0;
}

// BN backward apply (in place on the padded dy buffer; z unpadded):
//   dz = scale*dy + c1*xin + c2 (folded coefficients)  [ * (z>0) and xin = relu(z) when relu_first ]
template <typename T>
__global__ void k_bn_bwd_apply(T* __restrict__ dy, const T* __restrict__ z, long long total, int H, int W, int C,
                               BnRef bn, int relu_first) {
  const int groups = C >> 3;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int g = (int)(idx % groups);
  long long p = idx / groups;
  const int xx = (int)(p % W);
  const int yy = (int)((p / W) % H);
  const long long b = p / ((long long)W * H);
  T* dptr = dy + pad_off(b, yy, xx, H, W, C) + g * 8;
  float sc[8], cb[8], cc[8], d[8], v[8], o[8];
  load8(bn.scale + g * 8, sc);
  load8(bn.c1 + g * 8, cb);
  load8(bn.c2 + g * 8, cc);
  load8(dptr, d);
  load8(z + p * C + g * 8, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float xin = relu_first ? fmaxf(v[i], 0.f) : v[i];
    float r = fmaf(sc[i], d[i], fmaf(cb[i], xin, cc[i]));
    if (relu_first && !(v[i] > 0.f)) r = 0.f;
    o[i] = r;
  }
  store8(dptr, o);
}
template <typename T>
int launch_bn_bwd_apply(T* dy, const T* z, int B, int H, int W, int C, const BnRef& bn, int relu_first,
                        cudaStream_t s) {
  long long total = (long long)B * H * W * (C / 8);
  k_bn_bwd_apply<T><<<ceil_div(total, kThreads), kThreads, 0, s>>>(dy, z, total, H, W, C, bn, relu_first);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_bn_bwd_apply<float>(float*, const float*, int, int, int, int, const BnRef&, int, cudaStream_t);
template int launch_bn_bwd_apply<bf16>(bf16*, const bf16*, int, int, int, int, const BnRef&, int, cudaStream_t);

// --------------------------------------------------------------------------------------------
// embedding head: MaxPooling2D(pool, padding='same') over the raw conv4b map, Flatten (h,w,c)
// (audio_model.py:480-484, vision_model.py:212-215).  Windows divide evenly for every model type.
// --------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_embed_pool(const T* __restrict__ z, int H, int W, int C, int ph, int pw, int OH, int OW,
                             long long total, float* __restrict__ out) {
  const int groups = C >> 3;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int g = (int)(idx % groups);
  long long p = idx / groups;
  int ox = (int)(p % OW);
  int oy = (int)((p / OW) % OH);
  long long b = p / ((long long)OW * OH);
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  for (int dy = 0; dy < ph; ++dy)
    for (int dx = 0; dx < pw; ++dx) {
      int y = oy * ph + dy, x = ox * pw + dx;
      if (y >= H || x >= W) continue;
      float v[8];
      load8(z + ((b * H + y) * W + x) * C + g * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
    }
  store8(out + p * C + g * 8, m);
}
template <typename T>
int launch_embed_pool(const T* z, int B, int H, int W, int C, int ph, int pw, float* out, cudaStream_t s) {
  L3_REQUIRE(H % ph == 0 && W % pw == 0, "embed_pool: pool must divide the map (%dx%d by %dx%d)", H, W, ph, pw);
  int OH = H / ph, OW = W / pw;
  long long total = (long long)B * OH * OW * (C / 8);
  k_embed_pool<T><<<ceil_div(total, kThreads), kThreads, 0, s>>>(z, H, W, C, ph, pw, OH, OW, total, out);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_embed_pool<float>(const float*, int, int, int, int, int, int, float*, cudaStream_t);
template int launch_embed_pool<bf16>(const bf16*, int, int, int, int, int, int, float*, cudaStream_t);

// --------------------------------------------------------------------------------------------
// Keras-2.0.9 Adam over the flat trainable arena; the first n_l2 scalars are conv/dense kernels and
// receive the l2(1e-5) regulariser gradient 2*l2*w (model.py:23-31, audio_model.py:376-432).
// --------------------------------------------------------------------------------------------
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, long long n_l2, float lr_t, float b1, float b2, float eps,
                       float l2) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float w = p[i];
    float gi = g[i] + (i < n_l2 ? 2.f * l2 * w : 0.f);
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = w - lr_t * mi / (sqrtf(vi) + eps);
  }
}
int launch_adam(float* p, const float* g, float* m, float* v, long long n, long long n_l2, float lr_t, float b1,
                float b2, float eps, float l2, cudaStream_t s) {
  k_adam<<<kMaxBlocks, kThreads, 0, s>>>(p, g, m, v, n, n_l2, lr_t, b1, b2, eps, l2);
  L3_CHECK_LAUNCH();
  return 0;
}

__global__ void k_l2_penalty(const float* __restrict__ p, long long n, double* __restrict__ out) {
  __shared__ double sh[8];
  double acc = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    acc += (double)w * w;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    atomicAdd(out, t);
  }
}
int launch_l2_penalty(const float* p, long long n_l2, double* out, cudaStream_t s) {
  L3_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double), s));
  k_l2_penalty<<<148 * 2, kThreads, 0, s>>>(p, n_l2, out);
  L3_CHECK_LAUNCH();
  return 0;
}

// --------------------------------------------------------------------------------------------
// parity mode on tensor cores: split an fp32 tensor into 16-bit parts, x ~ hi + lo, laid out [hi | lo | hi] along the
// channel axis -- against weights packed [hi ; hi ; lo] one K-concatenated MMA chain accumulates
// a_hi w_hi + a_lo w_hi + a_hi w_lo in fp32.  fp16 parts carry 2 x 11 significant bits (forward: activations are O(1),
// well inside fp16's range), bf16 parts 2 x 8 bits with fp32's range (backward: gradients can be tiny).
// --------------------------------------------------------------------------------------------
template <bool FP16>
__global__ void k_split16(const float* __restrict__ src, unsigned short* __restrict__ dst, long long rows, int C) {
  const int groups = C >> 3;
  const long long total = rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int g = (int)(i - r * groups);
    float v[8];
    load8(src + r * C + g * 8, v);
    unsigned short hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (FP16) {
        const __half h = __float2half_rn(v[k]);
        const __half l = __float2half_rn(v[k] - __half2float(h));
        hi[k] = *reinterpret_cast<const unsigned short*>(&h);
        lo[k] = *reinterpret_cast<const unsigned short*>(&l);
      } else {
        const bf16 h = __float2bfloat16_rn(v[k]);
        const bf16 l = __float2bfloat16_rn(v[k] - __bfloat162float(h));
        hi[k] = *reinterpret_cast<const unsigned short*>(&h);
        lo[k] = *reinterpret_cast<const unsigned short*>(&l);
      }
    }
    uint4 uh, ul;
    uh.x = hi[0] | ((uint32_t)hi[1] << 16); uh.y = hi[2] | ((uint32_t)hi[3] << 16);
    uh.z = hi[4] | ((uint32_t)hi[5] << 16); uh.w = hi[6] | ((uint32_t)hi[7] << 16);
    ul.x = lo[0] | ((uint32_t)lo[1] << 16); ul.y = lo[2] | ((uint32_t)lo[3] << 16);
    ul.z = lo[4] | ((uint32_t)lo[5] << 16); ul.w = lo[6] | ((uint32_t)lo[7] << 16);
    unsigned short* d = dst + r * 3 * C + g * 8;
    *reinterpret_cast<uint4*>(d) = uh;
    *reinterpret_cast<uint4*>(d + C) = ul;
    *reinterpret_cast<uint4*>(d + 2 * C) = uh;
  }
}
int launch_split16(const float* src, void* dst, long long rows, int C, int fp16, cudaStream_t s) {
  L3_REQUIRE(C % 8 == 0, "split16: C=%d", C);
  const long long total = rows * (C / 8);
  long long want = (total + kThreads - 1) / kThreads;
  int blocks = (int)(want > 148 * 16 ? 148 * 16 : (want < 1 ? 1 : want));
  if (fp16) k_split16<true><<<blocks, kThreads, 0, s>>>(src, (unsigned short*)dst, rows, C);
  else k_split16<false><<<blocks, kThreads, 0, s>>>(src, (unsigned short*)dst, rows, C);
  L3_CHECK_LAUNCH();
  return 0;
}

__global__ void k_f64_to_f32_ew(const double* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}
int launch_f64_to_f32(const double* src, float* dst, long long n, cudaStream_t s) {
  int blocks = (int)((n + 255) / 256 > kMaxBlocks ? kMaxBlocks : (n + 255) / 256);
  k_f64_to_f32_ew<<<blocks < 1 ? 1 : blocks, 256, 0, s>>>(src, dst, n);
  L3_CHECK_LAUNCH();
  return 0;
}

int launch_zero(void* p, size_t bytes, cudaStream_t s) {
  L3_CHECK_CUDA(cudaMemsetAsync(p, 0, bytes, s));
  return 0;
}

}  // namespace l3
