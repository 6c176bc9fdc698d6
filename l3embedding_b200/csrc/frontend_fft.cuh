// Register-resident FFT of the audio front-end (frontend.cu): a length-N complex transform (N = 2048 or 512) is cut as
// N = 16 x 16 x R3 (R3 = 8 or 2) and run by T = N / 16 threads, every thread holding 16 complex points in registers:
//   step 1   thread t:          v[r] = x[t + T r];          16-point DFT over r;  v[q]  *= W_N^(t q);      -> S1[q][t]
//   step 2   thread (q, t2):    v[r2] = S1[q][t2 + R3 r2];  16-point DFT over r2; v[q2] *= W_N^(16 t2 q2); -> S2[q2*16 + q][t2]
//   step 3   thread, 16/R3 x:   R3-point DFT over t2 of S2[p][.], p = q2*16 + q                            -> Z[p + 256 k3]
// (decimation in frequency: X[q + 16 q2 + 256 k3]), so the spectrum lands in NATURAL order and only three shared-memory
// exchanges separate the 11 radix-2 stages.  The version this replaces ran 5 radix-4 passes + 1 radix-2 pass in shared
// memory with a block barrier each; its late passes (strides of 1..8 complex numbers) were 4..8-way bank conflicted.
//
// Everything here is __host__ __device__ so that tests/fe_fft_host.cu can run the very same index arithmetic on the CPU,
// thread by thread, against a float64 DFT (no GPU in the build container).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace l3 {

#define L3_HD __host__ __device__ __forceinline__

struct Cx {
  float x, y;
};
L3_HD Cx cx(float a, float b) { Cx c; c.x = a; c.y = b; return c; }
L3_HD Cx operator+(Cx a, Cx b) { return cx(a.x + b.x, a.y + b.y); }
L3_HD Cx operator-(Cx a, Cx b) { return cx(a.x - b.x, a.y - b.y); }
L3_HD Cx cmul(Cx a, Cx w) { return cx(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
L3_HD Cx mul_neg_i(Cx a) { return cx(a.y, -a.x); }   // a * (-i)

// y[c] = sum_a x[a] W4^(a c), W4 = -i; in place on four named values
L3_HD void dft4(Cx& x0, Cx& x1, Cx& x2, Cx& x3) {
  const Cx t0 = x0 + x2, t1 = x0 - x2, t2 = x1 + x3, t3 = mul_neg_i(x1 - x3);
  x0 = t0 + t2;
  x2 = t0 - t2;
  x1 = t1 + t3;
  x3 = t1 - t3;
}

// v[q] <- sum_r v[r] W16^(r q): r = 4a + b, q = c + 4d:  W16^(rq) = W4^(ac) W16^(bc) W4^(bd)
L3_HD void dft16(Cx (&v)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
  // over a (stride 4), for each b
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);   // v[4c + b] = Y_b[c]
  // twiddles W16^(b c), exp(-2 pi i k / 16): k = 1, 2, 3, 2, 4, 6, 3, 6, 9
  v[4 * 1 + 1] = cmul(v[4 * 1 + 1], cx(c1, -s1));
  v[4 * 1 + 2] = cmul(v[4 * 1 + 2], cx(h, -h));
  v[4 * 1 + 3] = cmul(v[4 * 1 + 3], cx(s1, -c1));
  v[4 * 2 + 1] = cmul(v[4 * 2 + 1], cx(h, -h));
  v[4 * 2 + 2] = mul_neg_i(v[4 * 2 + 2]);
  v[4 * 2 + 3] = cmul(v[4 * 2 + 3], cx(-h, -h));
  v[4 * 3 + 1] = cmul(v[4 * 3 + 1], cx(s1, -c1));
  v[4 * 3 + 2] = cmul(v[4 * 3 + 2], cx(-h, -h));
  v[4 * 3 + 3] = cmul(v[4 * 3 + 3], cx(-c1, s1));
  // over b, for each c: X[c + 4 d]
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);   // v[4c + d] = X[c + 4d]
  // to natural order: X[q] sits at v[4 (q & 3) + (q >> 2)] -- a transposition of the 4 x 4 register tile
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) {
      const Cx t = v[4 * c + d];
      v[4 * c + d] = v[4 * d + c];
      v[4 * d + c] = t;
    }
}

// v[k] <- sum_t v[t] W8^(t k): t = 2a + b, k = c + 4d:  W8^(tk) = W4^(ac) W8^(bc) W2^(bd)
L3_HD void dft8(Cx (&v)[8]) {
  const float h = 0.70710678118654752f;
  dft4(v[0], v[2], v[4], v[6]);   // v[2c]     = Y_0[c]
  dft4(v[1], v[3], v[5], v[7]);   // v[2c + 1] = Y_1[c]
  v[3] = cmul(v[3], cx(h, -h));
  v[5] = mul_neg_i(v[5]);
  v[7] = cmul(v[7], cx(-h, -h));
  Cx o[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    o[c] = v[2 * c] + v[2 * c + 1];
    o[c + 4] = v[2 * c] - v[2 * c + 1];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = o[k];
}

template <int N>
struct FftGeom {
  static const int T = N / 16;           // threads per transform
  static const int R3 = T / 16;          // radix of the last step (8 | 2)
  static const int LD1 = T + R3;         // pitch of S1 rows (bank-conflict-free reads in step 2)
  static const int LD2 = R3 + 1;         // pitch of S2 rows (odd: conflict-free reads in step 3)
  static const int BUF = (16 * LD1 > 256 * LD2) ? (16 * LD1 > N ? 16 * LD1 : N) : (256 * LD2 > N ? 256 * LD2 : N);
};

// Twiddle tables, laid out the way the threads read them (consecutive threads -> consecutive entries, no index arithmetic):
//   tw1[q * T + t]    = W_N^(t q)         q in [0,16), t  in [0,T)      (step 1)
//   tw2[q2 * R3 + t2] = W_N^(16 t2 q2)    q2 in [0,16), t2 in [0,R3)    (step 2)
// `Tw` is a functor int -> float2 over such a table (device: __ldg; host test: plain load).

// step 1 (after the caller filled v[r] = x[t + T r]): DFT, twiddle, store
template <int N, typename Tw>
L3_HD void fft_step1(Cx (&v)[16], int t, float* re, float* im, const Tw& tw1) {
  typedef FftGeom<N> G;
  dft16(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    Cx y = v[0];
    if (q != 0) {
      const float2 w = tw1(q * G::T + t);
      y = cmul(v[q], cx(w.x, w.y));
    }
    re[q * G::LD1 + t] = y.x;
    im[q * G::LD1 + t] = y.y;
  }
}
// step 2, load half: thread u = q * R3 + t2
template <int N>
L3_HD void fft_step2_load(Cx (&v)[16], int u, const float* re, const float* im) {
  typedef FftGeom<N> G;
  const int q = u / G::R3, t2 = u % G::R3;
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = cx(re[q * G::LD1 + t2 + G::R3 * r], im[q * G::LD1 + t2 + G::R3 * r]);
}
template <int N, typename Tw>
L3_HD void fft_step2_store(Cx (&v)[16], int u, float* re, float* im, const Tw& tw2) {
  typedef FftGeom<N> G;
  const int q = u / G::R3, t2 = u % G::R3;
  dft16(v);
#pragma unroll
  for (int q2 = 0; q2 < 16; ++q2) {
    Cx y = v[0];
    if (q2 != 0) {
      const float2 w = tw2(q2 * G::R3 + t2);
      y = cmul(v[q2], cx(w.x, w.y));
    }
    re[(q2 * 16 + q) * G::LD2 + t2] = y.x;
    im[(q2 * 16 + q) * G::LD2 + t2] = y.y;
  }
}
// step 3, load half: thread u owns the 16 / R3 short transforms p = u + T j; v[j * R3 + t2]
template <int N>
L3_HD void fft_step3_load(Cx (&v)[16], int u, const float* re, const float* im) {
  typedef FftGeom<N> G;
#pragma unroll
  for (int j = 0; j < 16 / G::R3; ++j)
#pragma unroll
    for (int t2 = 0; t2 < G::R3; ++t2) {
      const int p = u + G::T * j;
      v[j * G::R3 + t2] = cx(re[p * G::LD2 + t2], im[p * G::LD2 + t2]);
    }
}
// step 3, store half: Z[p + 256 k3] in natural order
template <int N>
L3_HD void fft_step3_store(Cx (&v)[16], int u, float* re, float* im) {
  typedef FftGeom<N> G;
  if (G::R3 == 8) {
    Cx a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = v[i]; b[i] = v[8 + i]; }
    dft8(a);
    dft8(b);
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) {
      re[u + 256 * k3] = a[k3].x;           im[u + 256 * k3] = a[k3].y;
      re[u + G::T + 256 * k3] = b[k3].x;    im[u + G::T + 256 * k3] = b[k3].y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const Cx s = v[2 * j] + v[2 * j + 1], d = v[2 * j] - v[2 * j + 1];
      const int p = u + G::T * j;
      re[p] = s.x;         im[p] = s.y;
      re[p + 256] = d.x;   im[p + 256] = d.y;
    }
  }
}

// host: fill the two tables (float64 angles, rounded once)
template <int N>
inline void fft_build_twiddles(float2* tw1, float2* tw2) {
  typedef FftGeom<N> G;
  const double kTwoPi = 6.283185307179586476925286766559;
  for (int q = 0; q < 16; ++q)
    for (int t = 0; t < G::T; ++t) {
      const double a = -kTwoPi * (double)((q * t) % N) / N;
      tw1[q * G::T + t] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int q2 = 0; q2 < 16; ++q2)
    for (int t2 = 0; t2 < G::R3; ++t2) {
      const double a = -kTwoPi * (double)((16 * t2 * q2) % N) / N;
      tw2[q2 * G::R3 + t2] = make_float2((float)cos(a), (float)sin(a));
    }
}

}  // namespace l3
