// AVC head: concat(vision 512, audio 512) -> Dense(128, relu) -> Dense(2, softmax) with keras
// categorical_crossentropy (clip 1e-7) and categorical accuracy, forward and backward.
// Reference: l3embedding/model.py:24-31, l3embedding/train.py:270-284.  0.26 MFLOP per pair: latency only,
// so plain deterministic SIMT kernels (no atomics on gradients).
#include "kernels.h"

namespace l3 {

static const int kIn = 1024, kHid = 128;

// one CTA per sample; the 1024-long dot products of the hidden layer are split over 4 x 128 threads (a single thread per
// hidden unit walked 1024 dependent global loads: 67 us on the critical path between the towers' forward and backward)
static const int kHeadSplit = 4;
__global__ void __launch_bounds__(128 * kHeadSplit)
k_head_fwd(HeadRef h, const float* __restrict__ labels, int B, float grad_scale) {
  __shared__ float xs[kIn];
  __shared__ float part[kHeadSplit][kHid];
  __shared__ float red[2][4];
  const int b = blockIdx.x, j = threadIdx.x & (kHid - 1), sp = threadIdx.x / kHid;
  for (int i = threadIdx.x; i < kIn; i += blockDim.x) xs[i] = h.concat[(long long)b * kIn + i];
  __syncthreads();
  {
    constexpr int seg = kIn / kHeadSplit;
    float a = 0.f;
#pragma unroll 8
    for (int i = sp * seg; i < (sp + 1) * seg; ++i) a = fmaf(xs[i], h.w1[i * kHid + j], a);
    part[sp][j] = a;
  }
  __syncthreads();
  if (threadIdx.x >= kHid) return;
  float acc = h.b1[j];
#pragma unroll
  for (int q = 0; q < kHeadSplit; ++q) acc += part[q][j];
  float hv = fmaxf(acc, 0.f);
  h.hidden[(long long)b * kHid + j] = hv;
  float l0 = warp_sum(hv * h.w2[j * 2 + 0]);
  float l1 = warp_sum(hv * h.w2[j * 2 + 1]);
  if ((j & 31) == 0) { red[0][j >> 5] = l0; red[1][j >> 5] = l1; }
  asm volatile("bar.sync 1, 128;" ::: "memory");   // the first 128 threads only (the others have left)
  if (j == 0) {
    float z0 = red[0][0] + red[0][1] + red[0][2] + red[0][3] + h.b2[0];
    float z1 = red[1][0] + red[1][1] + red[1][2] + red[1][3] + h.b2[1];
    float mx = fmaxf(z0, z1);
    float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
    float inv = 1.f / (e0 + e1);
    float p0 = e0 * inv, p1 = e1 * inv;
    h.logits[b * 2] = z0; h.logits[b * 2 + 1] = z1;
    h.probs[b * 2] = p0; h.probs[b * 2 + 1] = p1;
    if (labels) {
      float y0 = labels[b * 2], y1 = labels[b * 2 + 1];
      // keras: p /= sum(p); p = clip(p, 1e-7, 1-1e-7); loss = -sum(y*log(p))
      float s = p0 + p1;
      float r0 = p0 / s, r1 = p1 / s;
      const float lo = 1e-7f, hi = 1.f - 1e-7f;
      float q0 = fminf(fmaxf(r0, lo), hi), q1 = fminf(fmaxf(r1, lo), hi);
      float ce = -(y0 * logf(q0) + y1 * logf(q1));
      atomicAdd(&h.metrics[0], ce);
      int am_p = p1 > p0 ? 1 : 0, am_y = y1 > y0 ? 1 : 0;
      if (am_p == am_y) atomicAdd(&h.metrics[1], 1.f);
      // backward through clip (zero outside), renormalisation and softmax
      float g0 = (r0 > lo && r0 < hi) ? -y0 / q0 : 0.f;
      float g1 = (r1 > lo && r1 < hi) ? -y1 / q1 : 0.f;
      float dot = g0 * r0 + g1 * r1;
      float gp0 = (g0 - dot) / s, gp1 = (g1 - dot) / s;
      float dsm = gp0 * p0 + gp1 * p1;
      h.dlogits[b * 2] = p0 * (gp0 - dsm) * grad_scale;
      h.dlogits[b * 2 + 1] = p1 * (gp1 - dsm) * grad_scale;
    }
  }
}

int launch_head_fwd(const HeadRef& h, const float* labels, int B, float grad_scale, cudaStream_t s) {
  L3_CHECK_CUDA(cudaMemsetAsync(h.metrics, 0, 2 * sizeof(float), s));
  k_head_fwd<<<B, 128 * kHeadSplit, 0, s>>>(h, labels, B, grad_scale);
  L3_CHECK_LAUNCH();
  return 0;
}

// dhidden[b][j] = relu'(h) * sum_k dlogits[b][k] * w2[j][k] ; dw2, db2
__global__ void k_head_bwd_hidden(HeadRef h, int B) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < B * kHid) {
    int b = idx / kHid, j = idx % kHid;
    float hv = h.hidden[idx];
    float d = h.dlogits[b * 2] * h.w2[j * 2] + h.dlogits[b * 2 + 1] * h.w2[j * 2 + 1];
    h.dhidden[idx] = hv > 0.f ? d : 0.f;
  }
  if (idx < kHid * 2) {
    int j = idx / 2, k = idx % 2;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(h.hidden[b * kHid + j], h.dlogits[b * 2 + k], acc);
    h.dw2[idx] = acc;
  }
  if (idx < 2) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += h.dlogits[b * 2 + idx];
    h.db2[idx] = acc;
  }
}
// dw1[i][j] = sum_b x[b][i]*dh[b][j] ; db1[j] ; dconcat[b][i] = sum_j dh[b][j]*w1[i][j]
__global__ void k_head_bwd_w1(HeadRef h, int B) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < kIn * kHid) {
    int i = idx / kHid, j = idx % kHid;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(h.concat[(long long)b * kIn + i], h.dhidden[b * kHid + j], acc);
    h.dw1[idx] = acc;
  }
  if (idx < kHid) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += h.dhidden[b * kHid + idx];
    h.db1[idx] = acc;
  }
}
__global__ void k_head_bwd_x(HeadRef h, int B) {
  __shared__ float dh[kHid];
  const int b = blockIdx.x;
  if (threadIdx.x < kHid) dh[threadIdx.x] = h.dhidden[b * kHid + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.x; i < kIn; i += blockDim.x) {
    float acc = 0.f;
    const float4* wrow = reinterpret_cast<const float4*>(h.w1 + (long long)i * kHid);
#pragma unroll 8
    for (int j = 0; j < kHid / 4; ++j) {
      float4 w = wrow[j];
      acc += dh[4 * j] * w.x + dh[4 * j + 1] * w.y + dh[4 * j + 2] * w.z + dh[4 * j + 3] * w.w;
    }
    h.dconcat[(long long)b * kIn + i] = acc;
  }
}

int launch_head_bwd(const HeadRef& h, int B, cudaStream_t s) {
  int n1 = B * kHid > 256 ? B * kHid : 256;
  k_head_bwd_hidden<<<ceil_div(n1, 256), 256, 0, s>>>(h, B);
  L3_CHECK_LAUNCH();
  k_head_bwd_w1<<<ceil_div(kIn * kHid, 256), 256, 0, s>>>(h, B);
  L3_CHECK_LAUNCH();
  k_head_bwd_x<<<B, 256, 0, s>>>(h, B);
  L3_CHECK_LAUNCH();
  return 0;
}

}  // namespace l3
