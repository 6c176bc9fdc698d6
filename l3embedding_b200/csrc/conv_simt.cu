// fp32-accumulate SIMT 3x3 'same' convolutions (implicit GEMM, shared-memory tiled, 4x4 register blocking).
// This is the parity-mode (fp32 storage) path and the K=9 / K=27 first-layer path of the throughput mode;
// the bf16 tensor-core path is conv_tc.cu.  Semantics: keras Conv2D(3x3, padding='same', bias) with HWIO kernels
// (l3embedding/audio_model.py:376-432, l3embedding/vision_model.py:130-186), NHWC activations.
#include "kernels.h"

namespace l3 {

static const int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
k_conv3x3_simt(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
               T* __restrict__ out, int B, int H, int W, int Cin, int Cout) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const long long M = (long long)B * H * W;
  const int K = 9 * Cin;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int t = threadIdx.x;
  // A-load assignment: k_local = t % 16, pixels m_local = t/16 + 16*i
  const int ak = t & 15;
  long long pbase[4];
  bool pvalid[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + (t >> 4) + 16 * i;
    pvalid[i] = m < M;
    long long mm = pvalid[i] ? m : 0;
    int x = (int)(mm % W);
    int y = (int)((mm / W) % H);
    long long b = mm / ((long long)W * H);
    pbase[i] = pad_off(b, y - 1, x - 1, H, W, Cin);  // top-left tap of the 3x3 window in the padded input
  }
  const int bn = t & 63, bk = t >> 6;
  const int tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      int k = k0 + ak;
      bool kv = k < K;
      int tap = kv ? k / Cin : 0;
      int ci = k - tap * Cin;
      long long doff = ((long long)(tap / 3) * (W + 2) + tap % 3) * Cin + ci;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (kv && pvalid[i]) v = to_f(in[pbase[i] + doff]);
        As[ak][(t >> 4) + 16 * i] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int kk = bk + 4 * i;
        int kg = k0 + kk;
        float v = 0.f;
        if (kg < K && n0 + bn < Cout) v = w[(long long)kg * Cout + n0 + bn];
        Bs[kk][bn] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < Cout) out[m * Cout + n] = from_f<T>(acc[i][j] + (bias ? bias[n] : 0.f));
    }
  }
}

template <typename T>
int launch_conv3x3_simt(const T* in, const float* w, const float* bias, T* out, int B, int H, int W, int Cin, int Cout,
                        cudaStream_t s) {
  long long M = (long long)B * H * W;
  dim3 grid(ceil_div(M, BM), ceil_div(Cout, BN));
  k_conv3x3_simt<T><<<grid, 256, 0, s>>>(in, w, bias, out, B, H, W, Cin, Cout);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_conv3x3_simt<float>(const float*, const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
template int launch_conv3x3_simt<bf16>(const bf16*, const float*, const float*, bf16*, int, int, int, int, int, cudaStream_t);

// ---- weight gradient ----------------------------------------------------------------------------------
// a: padded (B,H+2,W+2,Cin); dz: padded (B,H+2,W+2,Cout)
// grid.x = ci tiles * co tiles * 9 taps ; grid.y = split-K slices over the B*H*W pixels.
// The slice partials are merged with atomics.  ACC = double (parity mode): the merge target is an fp64 scratch
// (9*Cin*Cout + Cout doubles, zeroed by the launcher) that k_f64_to_f32 rounds once at the end, so the result does not
// depend on the order in which the slices arrive (an fp32 merge in run-dependent order made the gradients -- and the
// input-BN gradient derived from them -- differ from run to run by more than the 1e-2 parity bar at small batch).
template <typename T, typename ACC>
__global__ void __launch_bounds__(256)
k_wgrad3x3_simt(const T* __restrict__ a, const T* __restrict__ dz, ACC* __restrict__ dw, ACC* __restrict__ db,
                int B, int H, int W, int Cin, int Cout, int ci_tiles, int co_tiles, long long m_per_slice) {
  __shared__ __align__(16) float As[BK][BM];  // [pixel][ci]
  __shared__ __align__(16) float Bs[BK][BN];  // [pixel][co]
  const long long M = (long long)B * H * W;
  int bid = blockIdx.x;
  const int tap = bid % 9;
  bid /= 9;
  const int co0 = (bid % co_tiles) * BN;
  const int ci0 = (bid / co_tiles) * BM;
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  const long long mbeg = (long long)blockIdx.y * m_per_slice;
  const long long mend = min(M, mbeg + m_per_slice);
  const int t = threadIdx.x;
  const int lc = t & 63, lk = t >> 6;
  const int tx = t & 15, ty = t >> 4;
  const bool do_bias = (tap == 4) && (ci0 == 0) && (db != nullptr);
  float bsum = 0.f;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long mc = mbeg; mc < mend; mc += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk = lk + 4 * i;
      long long m = mc + kk;
      float va = 0.f, vb = 0.f;
      if (m < mend) {
        int x = (int)(m % W);
        int y = (int)((m / W) % H);
        long long b = m / ((long long)W * H);
        if (ci0 + lc < Cin) va = to_f(a[pad_off(b, y + dy, x + dx, H, W, Cin) + ci0 + lc]);
        if (co0 + lc < Cout) vb = to_f(dz[pad_off(b, y, x, H, W, Cout) + co0 + lc]);
      }
      As[kk][lc] = va;
      Bs[kk][lc] = vb;
    }
    __syncthreads();
    if (do_bias && t < 64) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) bsum += Bs[kk][t];
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 av4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bv4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {av4.x, av4.y, av4.z, av4.w}, bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ci = ci0 + ty * 4 + i;
    if (ci >= Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = co0 + tx * 4 + j;
      if (co < Cout) atomicAdd(&dw[((long long)tap * Cin + ci) * Cout + co], (ACC)acc[i][j]);
    }
  }
  if (do_bias && t < 64 && co0 + t < Cout) atomicAdd(&db[co0 + t], (ACC)bsum);
}

// dst[i] = (float)src[i]: the single rounding of an fp64 merge target
__global__ void k_f64_to_f32(const double* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}
static int f64_to_f32(const double* src, float* dst, long long n, cudaStream_t s) {
  int blocks = (int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
  k_f64_to_f32<<<blocks < 1 ? 1 : blocks, 256, 0, s>>>(src, dst, n);
  L3_CHECK_LAUNCH();
  return 0;
}
// fp64 merge scratch: the caller's (carved from the context workspace) or a stream-ordered allocation
struct Scratch64 {
  double* p;
  bool owned;
  cudaStream_t s;
  int init(double* given, long long n, cudaStream_t stream) {
    s = stream;
    p = given;
    owned = false;
    if (!p) {
      L3_CHECK_CUDA(cudaMallocAsync((void**)&p, sizeof(double) * n, s));
      owned = true;
    }
    L3_CHECK_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * n, s));
    return 0;
  }
  ~Scratch64() {
    if (owned && p) cudaFreeAsync(p, s);
  }
};

template <typename T>
int launch_wgrad3x3_simt(const T* a, const T* dz, float* dw, float* db, int B, int H, int W, int Cin, int Cout,
                         cudaStream_t s, double* scratch64) {
  long long M = (long long)B * H * W;
  int ci_tiles = ceil_div(Cin, BM), co_tiles = ceil_div(Cout, BN);
  int tiles = ci_tiles * co_tiles * 9;
  long long slices = (148LL * 8 + tiles - 1) / tiles;
  long long max_slices = (M + 255) / 256;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  long long m_per_slice = ((M + slices - 1) / slices + BK - 1) / BK * BK;
  slices = (M + m_per_slice - 1) / m_per_slice;
  dim3 grid(tiles, (unsigned)slices);
  if (sizeof(T) == 4) {
    // parity mode: fp64 merge of the slice partials, rounded once (dw / db are overwritten)
    const long long nw = 9LL * Cin * Cout;
    Scratch64 sc;
    if (sc.init(scratch64, nw + Cout, s)) return -1;
    k_wgrad3x3_simt<T, double><<<grid, 256, 0, s>>>(a, dz, sc.p, db ? sc.p + nw : nullptr, B, H, W, Cin, Cout, ci_tiles,
                                                    co_tiles, m_per_slice);
    L3_CHECK_LAUNCH();
    if (f64_to_f32(sc.p, dw, nw, s)) return -1;
    if (db && f64_to_f32(sc.p + nw, db, Cout, s)) return -1;
    return 0;
  }
  k_wgrad3x3_simt<T, float><<<grid, 256, 0, s>>>(a, dz, dw, db, B, H, W, Cin, Cout, ci_tiles, co_tiles, m_per_slice);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_wgrad3x3_simt<float>(const float*, const float*, float*, float*, int, int, int, int, int, cudaStream_t, double*);
template int launch_wgrad3x3_simt<bf16>(const bf16*, const bf16*, float*, float*, int, int, int, int, int, cudaStream_t, double*);

__global__ void k_flip_transpose(const float* __restrict__ w, float* __restrict__ wt, int Cin, int Cout) {
  long long n = 9LL * Cin * Cout;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // i indexes wt[tap][co][ci]
  int ci = (int)(i % Cin);
  int co = (int)((i / Cin) % Cout);
  int tap = (int)(i / ((long long)Cin * Cout));
  wt[i] = w[((long long)(8 - tap) * Cin + ci) * Cout + co];
}
int launch_flip_transpose(const float* w, float* w_t, int Cin, int Cout, cudaStream_t s) {
  long long n = 9LL * Cin * Cout;
  k_flip_transpose<<<ceil_div(n, 256), 256, 0, s>>>(w, w_t, Cin, Cout);
  L3_CHECK_LAUNCH();
  return 0;
}

}  // namespace l3

// ==========================================================================================================
// First-layer kernels (Cin = 1 audio / 3 vision, Cout = 64): K = 9 / 27 is far too small for a GEMM tile; all three
// products are HBM/LSU-bound, so they get dedicated direct kernels with a sliding 3x3 input window in registers.
// ==========================================================================================================
namespace l3 {

static const int kFirstSeg = 16;   // x-segments per image row = pixel lanes per block
template <typename T> struct FirstMergeT { typedef float type; };
template <> struct FirstMergeT<float> { typedef double type; };

// forward: out[b,y,x,co] = bias[co] + sum_{ky,kx,c} in[b,y+ky,x+kx,c] (padded coords) * w[ky][kx][c][co]
// block = 16 channel groups (4 output channels each, weights in registers) x 16 pixel lanes (contiguous x segments)
template <typename T, int C0>
__global__ void __launch_bounds__(256)
k_first_conv(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
             int B, int H, int W, int rows_per_block, double* __restrict__ stats) {
  constexpr int CO = 64, K = 9 * C0;
  float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};   // BN statistics of the stored values
  const int cg = threadIdx.x & 15, pl = threadIdx.x >> 4;
  float wr[K][4], b4[4];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float4 t = *reinterpret_cast<const float4*>(w + k * CO + cg * 4);
    wr[k][0] = t.x; wr[k][1] = t.y; wr[k][2] = t.z; wr[k][3] = t.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) b4[i] = bias ? bias[cg * 4 + i] : 0.f;
  const int seg = (W + kFirstSeg - 1) / kFirstSeg;
  const int x0 = pl * seg, x1 = min(W, x0 + seg);
  const int Wp = W + 2;
  const long long n_rows = (long long)B * H;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(n_rows, r0 + rows_per_block);
  for (long long r = r0; r < r1; ++r) {
    if (x0 >= x1) break;
    const long long b = r / H;
    const int y = (int)(r - b * H);
    const T* arow = in + pad_off(b, y - 1, -1, H, W, C0);   // padded row y (top tap), padded column 0
    float win[3][3][C0];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 1; kx < 3; ++kx)
#pragma unroll
        for (int c = 0; c < C0; ++c) win[ky][kx][c] = to_f(arow[((long long)ky * Wp + x0 + kx - 1) * C0 + c]);
    T* orow = out + ((b * H + y) * (long long)W) * CO + cg * 4;
    for (int x = x0; x < x1; ++x) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int c = 0; c < C0; ++c) {
          win[ky][0][c] = win[ky][1][c];
          win[ky][1][c] = win[ky][2][c];
          win[ky][2][c] = to_f(arow[((long long)ky * Wp + x + 2) * C0 + c]);
        }
      float acc[4] = {b4[0], b4[1], b4[2], b4[3]};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int c = 0; c < C0; ++c) {
            const float v = win[ky][kx][c];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(v, wr[(ky * 3 + kx) * C0 + c][i], acc[i]);
          }
      T* o = orow + (long long)x * CO;
      if (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      } else {
        uint2 u;
        u.x = pack_bf16x2(acc[0], acc[1]);
        u.y = pack_bf16x2(acc[2], acc[3]);
        *reinterpret_cast<uint2*>(o) = u;
        acc[0] = __uint_as_float(u.x << 16); acc[1] = __uint_as_float(u.x & 0xffff0000u);
        acc[2] = __uint_as_float(u.y << 16); acc[3] = __uint_as_float(u.y & 0xffff0000u);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { ssum[i] += acc[i]; ssq[i] = fmaf(acc[i], acc[i], ssq[i]); }
    }
  }
  if (stats != nullptr) {
    // merge type: fp64 from the first merge on in parity mode (T = float), fp32 shared atomics otherwise
    typedef typename FirstMergeT<T>::type ACC;
    __shared__ ACC red[2][CO];
    if (threadIdx.x < 2 * CO) (&red[0][0])[threadIdx.x] = (ACC)0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // lanes l and l+16 of a warp hold the same channel group (cg = tid & 15)
      ACC a = (ACC)ssum[i], b2 = (ACC)ssq[i];
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      b2 += __shfl_xor_sync(0xffffffffu, b2, 16);
      if ((threadIdx.x & 31) < 16) { atomicAdd(&red[0][cg * 4 + i], a); atomicAdd(&red[1][cg * 4 + i], b2); }
    }
    __syncthreads();
    if (threadIdx.x < 2 * CO) atomicAdd(&stats[threadIdx.x], (double)(&red[0][0])[threadIdx.x]);
  }
}

template <typename T>
int launch_first_conv(const T* in, const float* w, const float* bias, T* out, int B, int H, int W, int C0, int Cout,
                      double* stats, cudaStream_t s) {
  L3_REQUIRE(Cout == 64 && (C0 == 1 || C0 == 3), "first_conv: C0=%d Cout=%d", C0, Cout);
  if (stats) L3_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * Cout, s));
  long long n_rows = (long long)B * H;
  int rpb = (int)((n_rows + 148 * 16 - 1) / (148 * 16));
  if (rpb < 1) rpb = 1;
  int blocks = (int)((n_rows + rpb - 1) / rpb);
  if (C0 == 1) k_first_conv<T, 1><<<blocks, 256, 0, s>>>(in, w, bias, out, B, H, W, rpb, stats);
  else k_first_conv<T, 3><<<blocks, 256, 0, s>>>(in, w, bias, out, B, H, W, rpb, stats);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_first_conv<float>(const float*, const float*, const float*, float*, int, int, int, int, int, double*, cudaStream_t);
template int launch_first_conv<bf16>(const bf16*, const float*, const float*, bf16*, int, int, int, int, int, double*, cudaStream_t);

// dw[tap][c][co] += sum_px a[px+tap][c] * dz[px][co] ; db[co] += sum_px dz[px][co]
// d1[tap][co]    += sum_px inside(px+tap) * dz[px][co]      (weight gradient w.r.t. an all-ones input plane; used by
//                                                            k_bn0_from_dw for the input-BN gradients)
// block = 64 output channels x 4 pixel lanes (contiguous x segments); 4 pixels per iteration with all loads issued
// first (input taps are warp-broadcast loads); every thread keeps its 9*C0 (+9) partial sums in registers.
// ACC = double (parity mode): block partials are merged into an fp64 scratch and rounded once (see k_wgrad3x3_simt).
template <typename T, int C0, typename ACC>
__global__ void __launch_bounds__(256)
k_first_wgrad(const T* __restrict__ a, const T* __restrict__ dz, ACC* __restrict__ dw, ACC* __restrict__ db,
              ACC* __restrict__ d1, int B, int H, int W, int rows_per_block) {
  constexpr int CO = 64, K = 9 * C0;
  const int co = threadIdx.x & 63, lane = threadIdx.x >> 6;
  float acc[K], ones[9];
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) ones[i] = 0.f;
  float bsum = 0.f;
  const long long n_rows = (long long)B * H;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(n_rows, r0 + rows_per_block);
  const int Wp = W + 2;
  const int seg = (W + 3) / 4;
  const int x0 = lane * seg, x1 = min(W, x0 + seg);
  for (long long r = r0; r < r1; ++r) {
    const long long b = r / H;
    const int y = (int)(r - b * H);
    const T* dzrow = dz + pad_off(b, y, 0, H, W, CO) + co;
    const T* arow = a + pad_off(b, y - 1, -1, H, W, C0);   // padded row y (top tap), padded column 0
    const float fy0 = y > 0 ? 1.f : 0.f, fy2 = y < H - 1 ? 1.f : 0.f;
    for (int x = x0; x < x1; x += 4) {
      float cols[3][6][C0], g[4];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int xc = min(x + j, W + 1);
#pragma unroll
          for (int c = 0; c < C0; ++c) cols[ky][j][c] = to_f(arow[((long long)ky * Wp + xc) * C0 + c]);
        }
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = (x + i < x1) ? to_f(dzrow[(long long)(x + i) * CO]) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum += g[i];
        const float gx0 = (x + i > 0) ? g[i] : 0.f, gx2 = (x + i < W - 1) ? g[i] : 0.f;
        ones[0] = fmaf(fy0, gx0, ones[0]); ones[1] = fmaf(fy0, g[i], ones[1]); ones[2] = fmaf(fy0, gx2, ones[2]);
        ones[3] += gx0;                    ones[4] += g[i];                    ones[5] += gx2;
        ones[6] = fmaf(fy2, gx0, ones[6]); ones[7] = fmaf(fy2, g[i], ones[7]); ones[8] = fmaf(fy2, gx2, ones[8]);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int c = 0; c < C0; ++c)
              acc[(ky * 3 + kx) * C0 + c] = fmaf(cols[ky][i + kx][c], g[i], acc[(ky * 3 + kx) * C0 + c]);
      }
    }
  }
  __shared__ float red[4][K + 10][CO];
#pragma unroll
  for (int i = 0; i < K; ++i) red[lane][i][co] = acc[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) red[lane][K + i][co] = ones[i];
  red[lane][K + 9][co] = bsum;
  __syncthreads();
  for (int i = threadIdx.x; i < (K + 10) * CO; i += blockDim.x) {
    const int k = i / CO, c = i % CO;
    const ACC v = (ACC)red[0][k][c] + (ACC)red[1][k][c] + (ACC)red[2][k][c] + (ACC)red[3][k][c];
    if (k < K) atomicAdd(&dw[k * CO + c], v);
    else if (k < K + 9) { if (d1) atomicAdd(&d1[(k - K) * CO + c], v); }
    else if (db) atomicAdd(&db[c], v);
  }
}

template <typename T>
int launch_first_wgrad(const T* a, const T* dz, float* dw, float* db, float* d1, int B, int H, int W, int C0, int Cout,
                       cudaStream_t s, double* scratch64) {
  L3_REQUIRE(Cout == 64 && (C0 == 1 || C0 == 3), "first_wgrad: C0=%d Cout=%d", C0, Cout);
  long long n_rows = (long long)B * H;
  int rpb = (int)((n_rows + 148 * 8 - 1) / (148 * 8));
  if (rpb < 1) rpb = 1;
  int blocks = (int)((n_rows + rpb - 1) / rpb);
  if (sizeof(T) == 4) {
    // parity mode: fp64 merge, rounded once; dw / db / d1 are overwritten.  scratch = [dw 9*C0*64 | d1 9*64 | db 64]
    const int nw = 9 * C0 * 64;
    Scratch64 sc;
    if (sc.init(scratch64, nw + 10 * 64, s)) return -1;
    double *w64 = sc.p, *o64 = sc.p + nw, *b64 = sc.p + nw + 9 * 64;
    if (C0 == 1) k_first_wgrad<T, 1, double><<<blocks, 256, 0, s>>>(a, dz, w64, db ? b64 : nullptr, d1 ? o64 : nullptr, B, H, W, rpb);
    else k_first_wgrad<T, 3, double><<<blocks, 256, 0, s>>>(a, dz, w64, db ? b64 : nullptr, d1 ? o64 : nullptr, B, H, W, rpb);
    L3_CHECK_LAUNCH();
    if (f64_to_f32(w64, dw, nw, s)) return -1;
    if (d1 && f64_to_f32(o64, d1, 9 * 64, s)) return -1;
    if (db && f64_to_f32(b64, db, 64, s)) return -1;
    return 0;
  }
  if (d1) L3_CHECK_CUDA(cudaMemsetAsync(d1, 0, sizeof(float) * 9 * 64, s));
  if (C0 == 1) k_first_wgrad<T, 1, float><<<blocks, 256, 0, s>>>(a, dz, dw, db, d1, B, H, W, rpb);
  else k_first_wgrad<T, 3, float><<<blocks, 256, 0, s>>>(a, dz, dw, db, d1, B, H, W, rpb);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_first_wgrad<float>(const float*, const float*, float*, float*, float*, int, int, int, int, int, cudaStream_t, double*);
template int launch_first_wgrad<bf16>(const bf16*, const bf16*, float*, float*, float*, int, int, int, int, int, cudaStream_t, double*);

// Input-BN gradients from the first layer's weight gradient (no pass over the image at all).  With
// da = dgrad(dz) and xin = gamma*xhat + beta the zero-padded conv input:
//   sum_p da[p,c]          = sum_{tap,co} w[tap][c][co] * d1[tap][co]          (d1 = weight gradient of an all-ones plane)
//   sum_p da[p,c]*xin[p,c] = sum_{tap,co} w[tap][c][co] * dw[tap][c][co]
//   sum_p da*xhat          = (sum da*xin - beta * sum da) / gamma
// Writes bn.sum = {sum da, sum da*xhat}.  A channel with |gamma| ~ 0 cannot be recovered this way: *fallback is set
// and k_first_dgrad_bnstats (which computes the sums directly) runs instead.
__global__ void k_bn0_from_dw(const float* __restrict__ w, const float* __restrict__ dw, const float* __restrict__ d1,
                              BnRef bn, int C0, int* __restrict__ fallback) {
  __shared__ float r1[3][128], r2[3][128];
  const int t = threadIdx.x;  // 128 threads
  for (int c = 0; c < C0; ++c) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = t; i < 9 * 64; i += 128) {
      const int tap = i >> 6, co = i & 63;
      const float wv = w[(tap * C0 + c) * 64 + co];
      s1 = fmaf(wv, d1[i], s1);
      s2 = fmaf(wv, dw[(tap * C0 + c) * 64 + co], s2);
    }
    r1[c][t] = s1;
    r2[c][t] = s2;
  }
  __syncthreads();
  if (t == 0) {
    int fb = 0;
    for (int c = 0; c < C0; ++c) {
      double s1 = 0, s2 = 0;
      for (int i = 0; i < 128; ++i) { s1 += r1[c][i]; s2 += r2[c][i]; }
      const double gm = bn.gamma[c];
      if (fabs(gm) < 1e-12) fb = 1;
      bn.sum[c] = s1;
      bn.sum[C0 + c] = fabs(gm) < 1e-12 ? 0.0 : (s2 - (double)bn.beta[c] * s1) / gm;
    }
    if (fb)
      for (int c = 0; c < 2 * C0; ++c) bn.sum[c] = 0.0;   // the fallback kernel accumulates from zero
    *fallback = fb;
  }
}
int launch_bn0_from_dw(const float* w, const float* dw, const float* d1, const BnRef& bn, int C0, int* fallback,
                       cudaStream_t s) {
  L3_REQUIRE(C0 >= 1 && C0 <= 3, "bn0_from_dw: C0=%d", C0);
  k_bn0_from_dw<<<1, 128, 0, s>>>(w, dw, d1, bn, C0, fallback);
  L3_CHECK_LAUNCH();
  return 0;
}

// Input-BN backward without materialising the data gradient.  8 threads per input pixel (8 of the 64 dz channels
// each, so a warp reads 4 pixels x 128 B fully coalesced) compute
//   da[c] = sum_{ky,kx,co} dz[y+1-ky, x+1-kx, co] * w[ky][kx][c][co]
// reduce over the 8 lanes by shuffles, and the block accumulates sum(da) and sum(da * xhat) per input channel (xhat
// from the float front-end / video tensor x0) into bn.sum (double[2*C0]); k_bn_bwd_finalize then yields d_gamma /
// d_beta.
template <typename T, int C0>
__global__ void __launch_bounds__(256)
k_first_dgrad_bnstats(const T* __restrict__ dz, const float* __restrict__ w, const float* __restrict__ x0, BnRef bn,
                      int B, int H, int W, const int* __restrict__ only_if) {
  constexpr int CO = 64;
  if (only_if != nullptr && *only_if == 0) return;   // the algebraic path (k_bn0_from_dw) already produced bn.sum
  __shared__ __align__(16) float ws[9 * C0 * CO];
  __shared__ double red[2 * C0];
  for (int i = threadIdx.x; i < 9 * C0 * CO; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < 2 * C0) red[threadIdx.x] = 0.0;
  __syncthreads();
  const int g = threadIdx.x & 7;
  const long long npix = (long long)B * H * W;
  // the two sums are residuals of large cancelling terms (sum(da) is ~0 by construction of the BN backward above):
  // accumulated in fp64 from the first addition on, so that only the per-pixel fp32 products carry rounding
  double s1[C0], s2[C0];
  float mean[C0], inv[C0];
#pragma unroll
  for (int c = 0; c < C0; ++c) { s1[c] = s2[c] = 0.0; mean[c] = bn.mean[c]; inv[c] = bn.invstd[c]; }
  const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
  const long long p_end = (npix + stride - 1) / stride * stride;   // all lanes of a warp iterate together (shuffles)
  for (long long p = blockIdx.x * (long long)(blockDim.x >> 3) + (threadIdx.x >> 3); p < p_end; p += stride) {
    const bool live = p < npix;
    float da[C0];
#pragma unroll
    for (int c = 0; c < C0; ++c) da[c] = 0.f;
    if (live) {
      const int x = (int)(p % W);
      const int y = (int)((p / W) % H);
      const long long b = p / ((long long)W * H);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        float v[8];
        load8(dz + pad_off(b, y + 1 - ky, x + 1 - kx, H, W, CO) + g * 8, v);
#pragma unroll
        for (int c = 0; c < C0; ++c) {
          const float* wt = ws + (tap * C0 + c) * CO + g * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wt), w1 = *reinterpret_cast<const float4*>(wt + 4);
          da[c] = fmaf(v[0], w0.x, fmaf(v[1], w0.y, fmaf(v[2], w0.z, fmaf(v[3], w0.w, da[c]))));
          da[c] = fmaf(v[4], w1.x, fmaf(v[5], w1.y, fmaf(v[6], w1.z, fmaf(v[7], w1.w, da[c]))));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C0; ++c) {
      float t = da[c];
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      if (live && g == 0) {
        const double xh = ((double)x0[p * C0 + c] - (double)mean[c]) * (double)inv[c];
        s1[c] += (double)t;
        s2[c] += (double)t * xh;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C0; ++c) {
    double a1 = s1[c], a2 = s2[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&red[c], a1); atomicAdd(&red[C0 + c], a2); }
  }
  __syncthreads();
  if (threadIdx.x < 2 * C0) atomicAdd(&bn.sum[threadIdx.x], red[threadIdx.x]);
}

template <typename T>
int launch_first_dgrad_bnstats(const T* dz, const float* w, const float* x0, const BnRef& bn, int B, int H, int W, int C0,
                               int Cout, const int* only_if, cudaStream_t s) {
  L3_REQUIRE(Cout == 64 && (C0 == 1 || C0 == 3), "first_dgrad: C0=%d Cout=%d", C0, Cout);
  if (!only_if) L3_CHECK_CUDA(cudaMemsetAsync(bn.sum, 0, sizeof(double) * 2 * C0, s));
  long long npix = (long long)B * H * W;
  long long want = (npix + 31) / 32;
  int blocks = (int)(want > 148 * 16 ? 148 * 16 : want);
  if (C0 == 1) k_first_dgrad_bnstats<T, 1><<<blocks, 256, 0, s>>>(dz, w, x0, bn, B, H, W, only_if);
  else k_first_dgrad_bnstats<T, 3><<<blocks, 256, 0, s>>>(dz, w, x0, bn, B, H, W, only_if);
  L3_CHECK_LAUNCH();
  return 0;
}
template int launch_first_dgrad_bnstats<float>(const float*, const float*, const float*, const BnRef&, int, int, int, int, int, const int*, cudaStream_t);
template int launch_first_dgrad_bnstats<bf16>(const bf16*, const float*, const float*, const BnRef&, int, int, int, int, int, const int*, cudaStream_t);

}  // namespace l3
