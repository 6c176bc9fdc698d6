// On-device audio front-end: pcm2float -> zero-padded framing -> Hann -> FFT -> |.|^2 -> (mel) -> sqrt -> dB.
// Replaces the kapre Spectrogram / Melspectrogram layers instantiated at
//   l3embedding/audio_model.py:39-43   (orig:          n_dft 512, valid, magnitude, log(max(x,1e-12))/5)
//   l3embedding/audio_model.py:149-150 (kapredbinputbn: n_dft 512, valid, magnitude, dB)
//   l3embedding/audio_model.py:257-259 (melspec1:      n_dft 2048, same, 128 mels htk, dB)
//   l3embedding/audio_model.py:367-369 (melspec2:      n_dft 2048, same, 256 mels htk, dB)
// and pcm2float (l3embedding/audio.py:21-31).  kapre evaluates the STFT as two dense strided convolutions
// (1.67 GFLOP/clip); algebraically that is |rfft(frame*hann)|^2, which is what is computed here with an
// fp32 FFT held in registers (frontend_fft.cuh; two real frames packed into one complex transform).  The mel
// projection uses the <=2-non-zeros-per-bin structure of the librosa filterbank (band lists) instead of a dense GEMM.
//
// HBM traffic per clip: 96 KB in (int16) + n_out*n_frames*4 B out; each CTA stages the sample window of its frames
// into shared memory with one 1-D TMA bulk copy (cp.async.bulk); the 8.5x frame overlap (2048/242) between CTAs is
// served by L2.
#include <math.h>
#include <vector>
#include "kernels.h"
#include "frontend_fft.cuh"

namespace l3 {

// One CTA = 128 threads = FR consecutive frames of one clip, two real frames packed into one complex transform that
// lives in REGISTERS (frontend_fft.cuh: N = 16 x 16 x R3, three shared-memory exchanges).  N = 2048: one transform at a time
// (T = 128 threads), two iterations; N = 512: four transforms side by side (T = 32), one iteration.  ~37 KB (50 KB) of
// shared memory and <= 85 registers: six (four) CTAs per SM.
static const int kFeThreads = 128;
template <int N>
struct FeCfg {
  static const int T = FftGeom<N>::T;
  static const int NC = kFeThreads / T;          // transforms in flight per CTA
  static const int FR = (N == 2048) ? 4 : 8;     // frames per CTA
  static const int ITERS = FR / (2 * NC);
  static const int CTAS = (N == 2048) ? 6 : 4;   // resident CTAs per SM aimed at (registers allow it; int16 input: 37 KB each)
};

// ---- host-side table construction (float64, cast to float32 as kapre stores them) ---------------------
static void build_mel(int sr, int n_fft, int n_mels, std::vector<int>& start, std::vector<int>& count,
                      std::vector<int>& offset, std::vector<float>& weight) {
  // librosa 0.5.1 filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=True, norm=1)
  const int nf = n_fft / 2 + 1;
  std::vector<double> mel_f(n_mels + 2);
  const double mmax = 2595.0 * log10(1.0 + (sr / 2.0) / 700.0);
  for (int i = 0; i < n_mels + 2; ++i) {
    double m = mmax * i / (n_mels + 1);
    mel_f[i] = 700.0 * (pow(10.0, m / 2595.0) - 1.0);
  }
  start.assign(n_mels, 0);
  count.assign(n_mels, 0);
  offset.assign(n_mels, 0);
  weight.clear();
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    int first = -1, last = -1;
    std::vector<float> row(nf);
    for (int k = 0; k < nf; ++k) {
      double f = (sr / 2.0) * k / (nf - 1);
      double lower = -(mel_f[i] - f) / fd0;
      double upper = (mel_f[i + 2] - f) / fd1;
      double w = fmax(0.0, fmin(lower, upper)) * enorm;
      row[k] = (float)w;
      if (row[k] != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    offset[i] = (int)weight.size();
    if (first >= 0) {
      start[i] = first;
      count[i] = last - first + 1;
      for (int k = first; k <= last; ++k) weight.push_back(row[k]);
    }
  }
}

size_t frontend_table_bytes(int n_dft, int n_mels) {
  size_t nf = n_dft / 2 + 1;
  // twiddles (step 1: 16 x n_dft/16, step 2: 16 x n_dft/256) + 2 windows + 3 int tables + weights (upper bound: 2
  // non-zeros per bin + slack)
  return (size_t)(n_dft + n_dft / 16 + 16) * sizeof(float2) + 2 * (size_t)n_dft * sizeof(float) +
         3 * (size_t)(n_mels + 1) * sizeof(int) + (4 * nf + 64) * sizeof(float) + 256;
}

int frontend_build_tables(FrontendPlan* plan, int sr, int n_mels, void* dev_mem, cudaStream_t s) {
  const int N = plan->n_dft;
  std::vector<float2> tw1(N), tw2(N / 16);
  if (N == 2048) fft_build_twiddles<2048>(tw1.data(), tw2.data());
  else fft_build_twiddles<512>(tw1.data(), tw2.data());
  std::vector<float> win(N), win16(N);
  for (int t = 0; t < N; ++t) {
    win[t] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * t / N));
    win16[t] = win[t] * (1.0f / 32768.0f);   // exact: a power of two
  }
  char* p = reinterpret_cast<char*>(dev_mem);
  auto put = [&](const void* src, size_t bytes) -> void* {
    void* dst = p;
    if (bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    p += (bytes + 15) / 16 * 16;
    return dst;
  };
  plan->tw1 = (const float2*)put(tw1.data(), tw1.size() * sizeof(float2));
  plan->tw2 = (const float2*)put(tw2.data(), tw2.size() * sizeof(float2));
  plan->window = (const float*)put(win.data(), win.size() * sizeof(float));
  plan->window_i16 = (const float*)put(win16.data(), win16.size() * sizeof(float));
  if (plan->mel) {
    std::vector<int> st, ct, of;
    std::vector<float> wt;
    build_mel(sr, N, n_mels, st, ct, of, wt);
    L3_REQUIRE(wt.size() <= (size_t)(4 * (N / 2 + 1) + 64), "mel table overflow");
    plan->mel_start = (const int*)put(st.data(), st.size() * sizeof(int));
    plan->mel_count = (const int*)put(ct.data(), ct.size() * sizeof(int));
    plan->mel_offset = (const int*)put(of.data(), of.size() * sizeof(int));
    plan->mel_weight = (const float*)put(wt.data(), wt.size() * sizeof(float));
    plan->mel_nnz = (int)wt.size();
  } else {
    plan->mel_start = plan->mel_count = plan->mel_offset = nullptr;
    plan->mel_weight = nullptr;
    plan->mel_nnz = 0;
  }
  L3_CHECK_CUDA(cudaStreamSynchronize(s));  // host vectors go out of scope
  return 0;
}

// ---- device ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct LdgTw {
  const float2* p;
  __device__ __forceinline__ float2 operator()(int j) const { return __ldg(p + j); }
};

// grid: (ceil(n_frames / FR), B).  Shared: FFT exchange buffers (re | im), power spectra, output tile, clip window
// (staged by a 1-D TMA bulk copy).
template <int N, bool I16>
__global__ void __launch_bounds__(kFeThreads, FeCfg<N>::CTAS)
k_frontend(FrontendPlan p, const void* __restrict__ audio, float* __restrict__ raw, int* __restrict__ clip_max) {
  typedef FftGeom<N> G;
  typedef FeCfg<N> F;
  constexpr int NF = N / 2 + 1, PWS = NF + 3, T = G::T, NC = F::NC, FR = F::FR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* const bre_all = reinterpret_cast<float*>(smem_raw);       // NC x BUF
  float* const bim_all = bre_all + NC * G::BUF;                    // NC x BUF
  float* const pw_all = bim_all + NC * G::BUF;                     // NC x 2 x PWS
  float* const tile = pw_all + NC * 2 * PWS;                       // n_out x FR
  unsigned char* stage = reinterpret_cast<unsigned char*>(tile + p.n_out * FR);
  stage = reinterpret_cast<unsigned char*>(((uintptr_t)stage + 15) & ~(uintptr_t)15);
  __shared__ __align__(8) unsigned long long bar;
  __shared__ float red[kFeThreads / 32];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FR;
  const int nfr = min(FR, p.n_frames - f0);
  // sample window needed by this CTA: [s_lo, s_hi) in clip coordinates (may exceed [0, n_samples))
  const int s_lo = f0 * p.n_hop - p.left_pad;
  const int s_hi = (f0 + nfr - 1) * p.n_hop - p.left_pad + N;
  const int c_lo = max(s_lo, 0), c_hi = min(s_hi, p.n_samples);
  constexpr int ES = I16 ? 2 : 4;
  // 16-byte aligned source window for the bulk copy
  const int a_lo = c_lo & ~(16 / ES - 1);
  const int a_hi = min((c_hi + 16 / ES - 1) & ~(16 / ES - 1), p.n_samples);  // n_samples*ES is a multiple of 16
  const uint32_t bytes = (uint32_t)(a_hi - a_lo) * ES;
  const unsigned char* src = reinterpret_cast<const unsigned char*>(audio) + ((size_t)b * p.clip_stride + a_lo) * ES;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(stage)),
        "l"(src), "r"(bytes), "r"(smem_u32(&bar))
        : "memory");
  }
  __syncthreads();   // the barrier is initialised for everybody
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar))
          : "memory");
    }
  }

  const int g = threadIdx.x / T, u = threadIdx.x % T;
  float* const bre = bre_all + g * G::BUF;
  float* const bim = bim_all + g * G::BUF;
  float* const pw0 = pw_all + g * 2 * PWS;
  float* const pw1 = pw0 + PWS;
  const LdgTw tw1{p.tw1}, tw2{p.tw2};             // tables through L1: shared by every CTA
  const float* __restrict__ win = I16 ? p.window_i16 : p.window;
  const float* __restrict__ melw = p.mel_weight;
  float local_max = -INFINITY;

#pragma unroll 1
  for (int it = 0; it < F::ITERS; ++it) {
    const int fa = 2 * (it * NC + g), fb = fa + 1;   // frame fa -> real part, fb -> imaginary part
    Cx v[16];
    // windowed samples x[u + T r]; zero outside the clip (TF SAME zero padding) and for frames past the clip's last one.
    // int16 samples: (float)s * (w * 2^-15) == ((float)s * 2^-15) * w bit for bit (pcm2float, audio.py:21-31)
    {
      const int ia0 = (f0 + fa) * p.n_hop - p.left_pad + u;
      if (fb < nfr && ia0 - u >= 0 && ia0 - u + p.n_hop + N <= p.n_samples) {
        // both frames of the pair lie inside the clip (all but the first and last few frames): no bounds checks
        if (I16) {
          const short* sp = reinterpret_cast<const short*>(stage) + (ia0 - a_lo);
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float w = __ldg(win + u + T * r);
            v[r] = cx((float)sp[T * r] * w, (float)sp[T * r + p.n_hop] * w);
          }
        } else {
          const float* sp = reinterpret_cast<const float*>(stage) + (ia0 - a_lo);
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float w = __ldg(win + u + T * r);
            v[r] = cx(sp[T * r] * w, sp[T * r + p.n_hop] * w);
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int ia = ia0 + T * r, ib = ia + p.n_hop;
          float xa = 0.f, xb = 0.f;
          if (fa < nfr && ia >= 0 && ia < p.n_samples)
            xa = I16 ? (float)reinterpret_cast<const short*>(stage)[ia - a_lo] : reinterpret_cast<const float*>(stage)[ia - a_lo];
          if (fb < nfr && ib >= 0 && ib < p.n_samples)
            xb = I16 ? (float)reinterpret_cast<const short*>(stage)[ib - a_lo] : reinterpret_cast<const float*>(stage)[ib - a_lo];
          const float w = __ldg(win + u + T * r);
          v[r] = cx(xa * w, xb * w);
        }
      }
    }
    fft_step1<N>(v, u, bre, bim, tw1);
    __syncthreads();
    fft_step2_load<N>(v, u, bre, bim);
    __syncthreads();   // S1 fully read before S2 overwrites the buffer
    fft_step2_store<N>(v, u, bre, bim, tw2);
    __syncthreads();
    fft_step3_load<N>(v, u, bre, bim);
    __syncthreads();
    fft_step3_store<N>(v, u, bre, bim);
    __syncthreads();
    // separate the two real spectra: XA[k] = (Z[k] + conj(Z[N-k]))/2 ; XB[k] = (Z[k] - conj(Z[N-k]))/(2i)
    for (int k = u; k < NF; k += T) {
      const int nk = (N - k) & (N - 1);
      const float zkx = bre[k], zky = bim[k], znx = bre[nk], zny = bim[nk];
      const float ar = 0.5f * (zkx + znx), ai = 0.5f * (zky - zny);
      const float br = 0.5f * (zky + zny), bi = -0.5f * (zkx - znx);
      pw0[k] = ar * ar + ai * ai;
      pw1[k] = br * br + bi * bi;
    }
    __syncthreads();
    // projection + log: one output bin (both frames of the pair) per thread-iteration
    const bool has_a = fa < nfr, has_b = fb < nfr;
    for (int m = u; m < p.n_out; m += T) {
      float va, vb;
      if (p.mel) {
        const int k0 = __ldg(p.mel_start + m), n = __ldg(p.mel_count + m);
        const float* w = melw + __ldg(p.mel_offset + m);
        float sa = 0.f, sb = 0.f;
        for (int k = 0; k < n; ++k) {
          const float wk = __ldg(w + k);
          sa = fmaf(pw0[k0 + k], wk, sa);
          sb = fmaf(pw1[k0 + k], wk, sb);
        }
        va = sqrtf(sa);
        vb = sqrtf(sb);
      } else {
        va = sqrtf(pw0[m]);
        vb = sqrtf(pw1[m]);
      }
      if (p.decibel) {
        va = 10.0f * (logf(fmaxf(va, 1e-10f)) / 2.302585092994046f);
        vb = 10.0f * (logf(fmaxf(vb, 1e-10f)) / 2.302585092994046f);
        if (has_a) local_max = fmaxf(local_max, va);
        if (has_b) local_max = fmaxf(local_max, vb);
      } else {
        va = logf(fmaxf(va, 1e-12f)) / 5.0f;
        vb = logf(fmaxf(vb, 1e-12f)) / 5.0f;
      }
      tile[m * FR + fa] = va;
      tile[m * FR + fb] = vb;
    }
    // (the next iteration's first shared-memory write -- step 1 into bre / bim -- is ordered after this iteration's last
    // read of them by the barrier above; pw is rewritten only after three more barriers)
  }
  __syncthreads();
  // write the tile: rows of nfr contiguous floats
  for (int i = threadIdx.x; i < p.n_out * FR; i += blockDim.x) {
    int m = i / FR, f = i % FR;
    if (f < nfr) raw[((size_t)b * p.n_out + m) * p.n_frames + f0 + f] = tile[i];
  }
  if (p.decibel) {
    float m = warp_max(local_max);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < kFeThreads / 32; ++i) m = fmaxf(m, red[i]);
      atomicMax(&clip_max[b], float_to_ordered(m));
    }
  }
}

__global__ void k_frontend_init_max(int* clip_max, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) clip_max[i] = float_to_ordered(-INFINITY);
}

// dB finish: subtract the per-clip max (kapre amplitude_to_decibel, per-sample axes), clip at -80
__global__ void k_frontend_finish(float* __restrict__ x, const int* __restrict__ clip_max, long long per_clip,
                                  long long total) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    float mx = ordered_to_float(clip_max[i / per_clip]);
    x[i] = fmaxf(x[i] - mx, -80.0f);
  }
}

template <int N, bool I16>
static int launch_fe(const FrontendPlan& p, const void* audio, int B, float* out, int* clip_max, cudaStream_t s) {
  typedef FftGeom<N> G;
  typedef FeCfg<N> F;
  constexpr int NF = N / 2 + 1;
  constexpr int ES = I16 ? 2 : 4;
  size_t stage_elems = (size_t)(F::FR - 1) * p.n_hop + N + 16;
  size_t smem = (size_t)F::NC * G::BUF * 8 + (size_t)F::NC * 2 * (NF + 3) * 4 + (size_t)p.n_out * F::FR * 4 + 16 + stage_elems * ES;
  static PerDeviceOnce once;
  if (once.needed()) {
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_frontend<N, I16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_frontend<N, I16>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
    once.mark();
  }
  L3_REQUIRE(smem <= 64 * 1024, "frontend smem %zu too large", smem);
  dim3 grid(ceil_div(p.n_frames, F::FR), B);
  k_frontend<N, I16><<<grid, kFeThreads, smem, s>>>(p, audio, out, clip_max);
  L3_CHECK_LAUNCH();
  return 0;
}

int launch_frontend(const FrontendPlan& p, const void* audio, int is_i16, int B, float* out, int* clip_max,
                    cudaStream_t s, int finish) {
  L3_REQUIRE(p.n_dft == 512 || p.n_dft == 2048, "frontend: n_dft %d", p.n_dft);
  L3_REQUIRE((p.clip_stride * (is_i16 ? 2 : 4)) % 16 == 0, "frontend: clip stride %lld breaks 16-byte alignment", p.clip_stride);
  if (p.decibel) {
    k_frontend_init_max<<<ceil_div(B, 128), 128, 0, s>>>(clip_max, B);
    L3_CHECK_LAUNCH();
  }
  int rc;
  if (p.n_dft == 2048)
    rc = is_i16 ? launch_fe<2048, true>(p, audio, B, out, clip_max, s) : launch_fe<2048, false>(p, audio, B, out, clip_max, s);
  else
    rc = is_i16 ? launch_fe<512, true>(p, audio, B, out, clip_max, s) : launch_fe<512, false>(p, audio, B, out, clip_max, s);
  if (rc) return rc;
  if (p.decibel && finish) {
    long long per_clip = (long long)p.n_out * p.n_frames, total = per_clip * B;
    int blocks = (int)((total + 1023) / 1024);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_frontend_finish<<<blocks, 256, 0, s>>>(out, clip_max, per_clip, total);
    L3_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace l3
