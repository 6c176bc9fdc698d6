// Shared helpers for the l3embedding_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace l3 {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (no exceptions cross the C ABI) ------------------------------------
void set_error(const char* fmt, ...);
#define L3_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      l3::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                         \
    }                                                                                    \
  } while (0)
// every kernel launch in the library is followed by L3_CHECK_LAUNCH(): it also feeds l3_launch_count()
extern unsigned long long g_launch_count;
#define L3_CHECK_LAUNCH()             \
  do {                                \
    ++l3::g_launch_count;             \
    L3_CHECK_CUDA(cudaGetLastError()); \
  } while (0)
#define L3_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      l3::set_error(__VA_ARGS__);                                                        \
      return -2;                                                                         \
    }                                                                                    \
  } while (0)

// ---- 8-wide vector access on T in {float, bf16} ----------------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
// raw 8-element loads: kept as loaded (4 registers for bf16) until unpacked at the point of use, so that kernels can
// hold many independent loads in flight without the register cost of the converted floats
template <typename T> struct Raw8;
template <> struct Raw8<bf16> { uint4 u; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void ldraw(const bf16* p, Raw8<bf16>& r) { r.u = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void ldraw(const float* p, Raw8<float>& r) {
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void unraw(const Raw8<bf16>& r, float (&v)[8]) {
  const uint32_t w[4] = {r.u.x, r.u.y, r.u.z, r.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void unraw(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// monotone float<->int map so atomicMax(int) orders floats
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// element offset of pixel (b,y,x) in a zero-haloed padded NHWC buffer (B,H+2,W+2,C)
__host__ __device__ __forceinline__ long long pad_off(long long b, int y, int x, int H, int W, int C) {
  return ((b * (H + 2) + (y + 1)) * (long long)(W + 2) + (x + 1)) * C;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// division of n < 2^31 by a launch-time constant d >= 2 (Granlund-Montgomery): q = umulhi(n, mul) >> shift.  The conv
// epilogues turn a flattened padded pixel index into (image, row, column) once per tile row, the input stage once per
// element; hardware-less 32-bit divisions cost ~30 dependent instructions each, 64-bit ones over a hundred.
struct FastDiv {
  uint32_t mul, shift, d;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;               // ceil(log2 d), d >= 2 -> l >= 1
  f.mul = (uint32_t)(((1ull << (31 + l)) / d) + 1);
  f.shift = l - 1;
  f.d = d;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return __umulhi(n, f.mul) >> f.shift; }

// cudaFuncSetAttribute (opt-in to > 48 KB of dynamic shared memory) is per DEVICE: one flag per launcher and device,
// so that a second context on another GPU of the same process gets its own opt-in.  Setting the attribute twice is
// harmless, so the flag is only marked after the calls succeeded (two racing threads both set it).
struct PerDeviceOnce {
  unsigned long long done = 0;   // bit d: configured on device d
  bool needed() const {
    int d = 0;
    cudaGetDevice(&d);
    return !((done >> (d & 63)) & 1ull);
  }
  void mark() {
    int d = 0;
    cudaGetDevice(&d);
    __atomic_fetch_or(&done, 1ull << (d & 63), __ATOMIC_RELAXED);
  }
};

// ---- per-BN-layer device record ---------------------------------------------------------
// All pointers are device pointers into caller-owned buffers.
struct BnRef {
  int C;
  const float* gamma;     // params arena
  const float* beta;
  float* moving_mean;     // bn_state arena
  float* moving_var;
  float* d_gamma;         // grads arena
  float* d_beta;
  // workspace (per layer)
  double* sum;            // [2*C]: sum, sumsq   (fwd)  /  sum_dy, sum_dy_xhat (bwd)
  float* mean;            // [C] batch mean (or moving mean at inference)
  float* invstd;          // [C]
  float* scale;           // gamma*invstd
  float* shift;           // beta - mean*scale
  float* c1;              // bwd: mean(dy)
  float* c2;              // bwd: mean(dy*xhat)
};

}  // namespace l3
