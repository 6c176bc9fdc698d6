// Host execution of the front-end's register FFT (l3embedding_b200/csrc/frontend_fft.cuh), thread by thread and phase by
// phase exactly as k_frontend runs it, against a float64 DFT.  Built and run by tests/test_host.py (no GPU needed).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../l3embedding_b200/csrc/frontend_fft.cuh"
using namespace l3;

struct HostTw {
  const float2* p;
  L3_HD float2 operator()(int j) const { return p[j]; }
};

template <int N>
static double run(unsigned seed) {
  typedef FftGeom<N> G;
  std::vector<float2> tw1(16 * G::T), tw2(16 * G::R3);
  fft_build_twiddles<N>(tw1.data(), tw2.data());
  const HostTw twl1{tw1.data()}, twl2{tw2.data()};
  std::vector<double> xr(N), xi(N);
  srand(seed);
  for (int n = 0; n < N; ++n) { xr[n] = rand() / (double)RAND_MAX - 0.5; xi[n] = rand() / (double)RAND_MAX - 0.5; }
  std::vector<float> re(G::BUF), im(G::BUF);
  std::vector<Cx> regs((size_t)G::T * 16);
  auto R = [&](int t) -> Cx(&)[16] { return *reinterpret_cast<Cx(*)[16]>(&regs[(size_t)t * 16]); };
  for (int t = 0; t < G::T; ++t) {
    Cx(&v)[16] = R(t);
    for (int r = 0; r < 16; ++r) v[r] = cx((float)xr[t + G::T * r], (float)xi[t + G::T * r]);
    fft_step1<N>(v, t, re.data(), im.data(), twl1);
  }
  for (int u = 0; u < G::T; ++u) fft_step2_load<N>(R(u), u, re.data(), im.data());
  for (int u = 0; u < G::T; ++u) fft_step2_store<N>(R(u), u, re.data(), im.data(), twl2);
  for (int u = 0; u < G::T; ++u) fft_step3_load<N>(R(u), u, re.data(), im.data());
  for (int u = 0; u < G::T; ++u) fft_step3_store<N>(R(u), u, re.data(), im.data());
  double worst = 0, scale = 0;
  for (int k = 0; k < N; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      double a = -2.0 * M_PI * (double)((long long)k * n % N) / N;
      sr += (double)(float)xr[n] * cos(a) - (double)(float)xi[n] * sin(a);
      si += (double)(float)xr[n] * sin(a) + (double)(float)xi[n] * cos(a);
    }
    scale = fmax(scale, hypot(sr, si));
    worst = fmax(worst, hypot(sr - re[k], si - im[k]));
  }
  return worst / scale;
}

int main() {
  double e1 = run<2048>(1), e2 = run<512>(2), e3 = run<2048>(3);
  printf("rel_err_2048 %.3e\nrel_err_512 %.3e\nrel_err_2048b %.3e\n", e1, e2, e3);
  return (e1 < 2e-6 && e2 < 2e-6 && e3 < 2e-6) ? 0 : 1;
}
