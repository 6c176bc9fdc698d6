// On-device audio front-end: pcm2float -> zero-padded framing -> Hann -> FFT -> |.|^2 -> (mel) -> sqrt -> dB.
// Replaces the kapre Spectrogram / Melspectrogram layers instantiated at
//   l3embedding/audio_model.py:39-43   (orig:          n_dft 512, valid, magnitude, log(max(x,1e-12))/5)
//   l3embedding/audio_model.py:149-150 (kapredbinputbn: n_dft 512, valid, magnitude, dB)
//   l3embedding/audio_model.py:257-259 (melspec1:      n_dft 2048, same, 128 mels htk, dB)
//   l3embedding/audio_model.py:367-369 (melspec2:      n_dft 2048, same, 256 mels htk, dB)
// and pcm2float (l3embedding/audio.py:21-31).  kapre evaluates the STFT as two dense strided convolutions
// (1.67 GFLOP/clip); algebraically that is |rfft(frame*hann)|^2, which is what is computed here with an
// fp32 in-shared-memory FFT (two real frames packed into one complex transform).  The mel projection uses the
// <=2-non-zeros-per-bin structure of the librosa filterbank (band lists) instead of a dense GEMM.
//
// HBM traffic per clip: 96 KB in (int16) + n_out*n_frames*4 B out; the whole clip is staged into shared memory
// once with a 1-D TMA bulk copy (cp.async.bulk) so the 8.5x frame overlap (2048/242) never re-reads HBM.
#include <math.h>
#include <vector>
#include "kernels.h"

namespace l3 {

// 4 frames = 2 packed complex FFTs per CTA, transformed CONCURRENTLY (one barrier per pass).  Round 2: was 8 frames with
// the Hann window and the mel weights staged in shared memory (137 KB: ONE 16-warp CTA per SM, and the kernel is bound
// by barrier / shared-memory latency, not by issue slots); with 4 frames and those two tables read through L1
// (__ldg: they are shared by every CTA) a CTA needs 57 KB and <= 40 registers, so THREE CTAs = 48 warps share an SM.
static const int kFramesPerCta = 4;
static const int kFePairs = kFramesPerCta / 2;
static const int kFeThreads = 512;
static const int kFeCtasPerSm = 3;

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
  // twiddle + window + 3 int tables + weights (upper bound: 2 non-zeros per bin + slack)
  return (size_t)(n_dft / 2) * sizeof(float2) + (size_t)n_dft * sizeof(float) + 3 * (size_t)(n_mels + 1) * sizeof(int) +
         (4 * nf + 64) * sizeof(float) + 256;
}

int frontend_build_tables(FrontendPlan* plan, int sr, int n_mels, void* dev_mem, cudaStream_t s) {
  const int N = plan->n_dft;
  std::vector<float2> tw(N / 2);
  for (int j = 0; j < N / 2; ++j) {
    double a = -2.0 * M_PI * j / N;
    tw[j] = make_float2((float)cos(a), (float)sin(a));
  }
  std::vector<float> win(N);
  for (int t = 0; t < N; ++t) win[t] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * t / N));
  char* p = reinterpret_cast<char*>(dev_mem);
  auto put = [&](const void* src, size_t bytes) -> void* {
    void* dst = p;
    if (bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    p += (bytes + 15) / 16 * 16;
    return dst;
  };
  plan->twiddle = (const float2*)put(tw.data(), tw.size() * sizeof(float2));
  plan->window = (const float*)put(win.data(), win.size() * sizeof(float));
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

// In-place DIF FFT of G independent length-N sequences z[g*N .. g*N+N) (shared memory), radix-4 passes (two radix-2
// stages fused in registers) plus one final radix-2 pass (log2 N is odd for both 512 and 2048); output bit-reversed.
// All G transforms advance together, so a pass costs ONE block barrier: the previous version ran 11 barriers per
// 2048-point transform with four butterflies per thread between them and was barrier-latency bound.
template <int N, int G>
__device__ __forceinline__ void fft_dif_batched(float2* z, const float2* __restrict__ tw) {
  static_assert(N == 512 || N == 2048, "log2(N) must be odd");
  constexpr int TOTAL = G * (N / 4), IT = (TOTAL + kFeThreads - 1) / kFeThreads;   // quad-butterflies per thread and pass
#pragma unroll 1
  for (int h = N / 2; h >= 2; h >>= 2) {          // fused stages with halves h and h/2
    const int ts1 = (N / 2) / h, hq = h >> 1;
    // two phases -- all loads, then all math and stores (the quad-butterflies of a pass touch disjoint elements, which
    // the compiler cannot know): with 16 warps per SM the passes were waiting on one LDS round trip per butterfly
    float2 x[IT][4];
    float2* zp[IT];
    int jj[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = threadIdx.x + it * kFeThreads;
      const int q = i & (N / 4 - 1);
      const int j = q & (hq - 1);
      jj[it] = j;
      zp[it] = z + (i / (N / 4)) * N + ((q - j) << 2) + j;
      if (i < TOTAL) {
        x[it][0] = zp[it][0]; x[it][1] = zp[it][hq]; x[it][2] = zp[it][h]; x[it][3] = zp[it][h + hq];
      }
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      if (threadIdx.x + it * kFeThreads >= TOTAL) continue;
      const int j = jj[it];
      const float2 x0 = x[it][0], x1 = x[it][1], x2 = x[it][2], x3 = x[it][3];
      const float2 wa = tw[j * ts1], wb = tw[(j + hq) * ts1], wc = tw[j * 2 * ts1];
      const float2 y0 = make_float2(x0.x + x2.x, x0.y + x2.y), d0 = make_float2(x0.x - x2.x, x0.y - x2.y);
      const float2 y1 = make_float2(x1.x + x3.x, x1.y + x3.y), d1 = make_float2(x1.x - x3.x, x1.y - x3.y);
      const float2 y2 = make_float2(d0.x * wa.x - d0.y * wa.y, d0.x * wa.y + d0.y * wa.x);
      const float2 y3 = make_float2(d1.x * wb.x - d1.y * wb.y, d1.x * wb.y + d1.y * wb.x);
      const float2 e0 = make_float2(y0.x - y1.x, y0.y - y1.y), e1 = make_float2(y2.x - y3.x, y2.y - y3.y);
      zp[it][0] = make_float2(y0.x + y1.x, y0.y + y1.y);
      zp[it][hq] = make_float2(e0.x * wc.x - e0.y * wc.y, e0.x * wc.y + e0.y * wc.x);
      zp[it][h] = make_float2(y2.x + y3.x, y2.y + y3.y);
      zp[it][h + hq] = make_float2(e1.x * wc.x - e1.y * wc.y, e1.x * wc.y + e1.y * wc.x);
    }
    __syncthreads();
  }
  // last stage (half = 1): twiddle 1
#pragma unroll
  for (int i = threadIdx.x; i < G * (N / 2); i += kFeThreads) {
    const float2 a = z[2 * i], b = z[2 * i + 1];
    z[2 * i] = make_float2(a.x + b.x, a.y + b.y);
    z[2 * i + 1] = make_float2(a.x - b.x, a.y - b.y);
  }
  __syncthreads();
}

// grid: (ceil(n_frames / kFramesPerCta), B).  Shared: clip window (staged by TMA bulk copy), FFT buffer, twiddles,
// power spectra, output tile.
template <int N, bool I16>
__global__ void __launch_bounds__(kFeThreads, kFeCtasPerSm)
k_frontend(FrontendPlan p, const void* __restrict__ audio, float* __restrict__ raw, int* __restrict__ clip_max) {
  constexpr int NF = N / 2 + 1;
  constexpr int LOGN = (N == 2048) ? 11 : 9;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int G = kFePairs, PWS = NF + 3;
  float2* zbuf = reinterpret_cast<float2*>(smem_raw);                       // G x N complex (frame pairs A + iB)
  float2* tw = zbuf + G * N;                                                // N/2 complex
  float* pw = reinterpret_cast<float*>(tw + N / 2);                         // kFramesPerCta power spectra of PWS floats
  float* tile = pw + kFramesPerCta * PWS;                                   // n_out * kFramesPerCta
  const float* __restrict__ win = p.window;                                 // through L1 (shared by all CTAs)
  const float* __restrict__ melw = p.mel_weight;
  unsigned char* stage = reinterpret_cast<unsigned char*>(tile + p.n_out * kFramesPerCta);
  stage = reinterpret_cast<unsigned char*>(((uintptr_t)stage + 15) & ~(uintptr_t)15);
  __shared__ __align__(8) unsigned long long bar;
  __shared__ float red[kFeThreads / 32];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kFramesPerCta;
  const int nfr = min(kFramesPerCta, p.n_frames - f0);
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
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(stage)),
        "l"(src), "r"(bytes), "r"(smem_u32(&bar))
        : "memory");
  }
  // overlap: constant tables -> shared
  for (int i = threadIdx.x; i < N / 2; i += blockDim.x) tw[i] = p.twiddle[i];
  // wait for the clip window
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
  __syncthreads();

  float local_max = -INFINITY;
  // pack frame 2g (real) and frame 2g+1 (imag) of every pair, windowed; zero outside the clip (TF SAME zero padding)
  // and for frames past the end of the clip's frame range
  for (int i = threadIdx.x; i < G * N; i += blockDim.x) {
    const int g = i / N, t = i & (N - 1);
    const int fa = 2 * g, fb = 2 * g + 1;
    const int ia = (f0 + fa) * p.n_hop - p.left_pad + t, ib = ia + p.n_hop;
    float xa = 0.f, xb = 0.f;
    if (fa < nfr && ia >= 0 && ia < p.n_samples)
      xa = I16 ? (float)reinterpret_cast<const short*>(stage)[ia - a_lo] * (1.0f / 32768.0f)
               : reinterpret_cast<const float*>(stage)[ia - a_lo];
    if (fb < nfr && ib >= 0 && ib < p.n_samples)
      xb = I16 ? (float)reinterpret_cast<const short*>(stage)[ib - a_lo] * (1.0f / 32768.0f)
               : reinterpret_cast<const float*>(stage)[ib - a_lo];
    const float w = __ldg(win + t);
    zbuf[i] = make_float2(xa * w, xb * w);
  }
  __syncthreads();
  fft_dif_batched<N, G>(zbuf, tw);
  // separate the two real spectra: XA[k] = (Z[k] + conj(Z[N-k]))/2 ; XB[k] = (Z[k] - conj(Z[N-k]))/(2i)
  for (int i = threadIdx.x; i < G * NF; i += blockDim.x) {
    const int g = i / NF, k = i - g * NF;
    const int rk = __brev((unsigned)k) >> (32 - LOGN);
    const int rnk = __brev((unsigned)((N - k) & (N - 1))) >> (32 - LOGN);
    const float2 zk = zbuf[g * N + rk], zn = zbuf[g * N + rnk];
    const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
    const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
    pw[(2 * g) * PWS + k] = ar * ar + ai * ai;
    pw[(2 * g + 1) * PWS + k] = br * br + bi * bi;
  }
  __syncthreads();
  // projection + log: one (frame pair, output bin) item per thread-iteration
  for (int i = threadIdx.x; i < G * p.n_out; i += blockDim.x) {
    const int g = i / p.n_out, m = i - g * p.n_out;
    const float* pw0 = pw + (2 * g) * PWS;
    const float* pw1 = pw0 + PWS;
    const bool has_a = 2 * g < nfr, has_b = 2 * g + 1 < nfr;
    float va, vb;
    if (p.mel) {
      const int k0 = p.mel_start[m], n = p.mel_count[m];
      const float* w = melw + p.mel_offset[m];
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
    tile[m * kFramesPerCta + 2 * g] = va;
    tile[m * kFramesPerCta + 2 * g + 1] = vb;
  }
  __syncthreads();
  // write the tile: rows of nfr contiguous floats
  for (int i = threadIdx.x; i < p.n_out * kFramesPerCta; i += blockDim.x) {
    int m = i / kFramesPerCta, f = i % kFramesPerCta;
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
  constexpr int NF = N / 2 + 1;
  constexpr int ES = I16 ? 2 : 4;
  size_t stage_elems = (size_t)(kFramesPerCta - 1) * p.n_hop + N + 16;
  size_t smem = (size_t)kFePairs * N * 8 + (size_t)(N / 2) * 8 + kFramesPerCta * (size_t)(NF + 3) * 4 +
                (size_t)p.n_out * kFramesPerCta * 4 + 16 + stage_elems * ES;
  static PerDeviceOnce once;
  if (once.needed()) {
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_frontend<N, I16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    once.mark();
  }
  L3_REQUIRE(smem <= 100 * 1024, "frontend smem %zu too large", smem);
  dim3 grid(ceil_div(p.n_frames, kFramesPerCta), B);
  k_frontend<N, I16><<<grid, kFeThreads, smem, s>>>(p, audio, out, clip_max);
  L3_CHECK_LAUNCH();
  return 0;
}

int launch_frontend(const FrontendPlan& p, const void* audio, int is_i16, int B, float* out, int* clip_max,
                    cudaStream_t s) {
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
  if (p.decibel) {
    long long per_clip = (long long)p.n_out * p.n_frames, total = per_clip * B;
    int blocks = (int)((total + 1023) / 1024);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_frontend_finish<<<blocks, 256, 0, s>>>(out, clip_max, per_clip, total);
    L3_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace l3
