// tcgen05 (5th-gen tensor core) bf16 implicit-GEMM 3x3 convolutions for sm_100a.  Replaces what cuDNN ran for the keras
// Conv2D layers of l3embedding/audio_model.py:376-432 and l3embedding/vision_model.py:130-186 (forward and backward).
//
// Kernels in this file:
//   k_conv3x3_tc3         forward / dgrad: shared-halo A regions on CTA pairs (cta_group::2), coalescing epilogue with
//                         per-configuration flavours EPI_* (BN statistics, fused BN + ReLU for inference, fp32 output)
//   k_wgrad3x3_tc2        weight gradient: shared-halo regions, tap pairing for Cin = 64
//   k_first_conv_tc       Cin = 1 / 3 forward: im2col operand built in shared memory from bulk-copy-staged input runs
//   k_first_wgrad_tc      Cin = 1 / 3 weight gradient (+ the input-BN gradient columns), same staging
//   k_pack_weights_batch  fp32 HWIO -> bf16 K-major operand rows, all layers in one launch
// (Round 1's per-tap / single-CTA versions of the forward and weight-gradient kernels were retired in round 2.)
//
// Layout trick: every convolution input lives in a zero-haloed buffer (B,H+2,W+2,C).  Flattening (b,y,x) to one
// row index m makes tap (ky,kx) of a 'same' 3x3 convolution a CONSTANT row shift (ky-1)*(W+2)+(kx-1), so the
// im2col operand of a 128-pixel tile is just a 2-D TMA box [128 rows][64 channels] at row m0+shift -- no gather,
// no boundary predicates (out-of-range rows are zero-filled by TMA, halo rows hold zeros).  Outputs are computed
// for halo rows too (1.6 % .. 15 % extra MMA work) and dropped in the epilogue.
//
// Forward GEMM  D[m][co] = sum_{tap,ci} A[m+shift(tap)][ci] * Wp[tap][ci][co]      M = pixels, N = Cout, K = 9*Cin
//   A, B both K-major in shared memory (128-byte swizzle, written by TMA), fp32 accumulators in TMEM (double
//   buffered so the epilogue of tile i overlaps the MMAs of tile i+1), persistent CTAs, warp-specialised:
//   warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM allocator), 4 (v1) / 8 (v2, v3) epilogue warps.
// Weight gradient  dW[tap][ci][co] = sum_m A[m+shift(tap)][ci] * dZ[m][co]          M = Cin, N = Cout, K = pixels
//   both operands are "MN-major" (channels contiguous), which UMMA reads directly from the same TMA boxes;
//   split-K over pixel slices with fp32 vector reductions into dW.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include "kernels.h"

namespace l3 {

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_addr(bar);
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_addr(dst)), "l"(tm), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// descriptor passed as (lo, hi) 32-bit halves: only `lo` (start address | LBO) changes between MMAs, `hi` is constant
__device__ __forceinline__ void umma_bf16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %6, 0;\n mov.b64 da, {%1, %2};\n mov.b64 db, {%3, %4};\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
// high half of a 128B-swizzle descriptor: SBO | version 1 | SWIZZLE_128B ; low half: start address | LBO
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
// one lane of the (converged) warp, the same one every time
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
// flattened padded pixel m -> output pixel index (b*H + y)*W + x of the un-padded tensor, or -1 for halo / out-of-range
__device__ __forceinline__ int out_pixel(uint32_t m, long long Mp, int H, int W, const FastDiv& dHWp, const FastDiv& dWp) {
  const uint32_t b = fdiv(m, dHWp);
  const uint32_t r = m - b * dHWp.d;
  const uint32_t yp = fdiv(r, dWp), xp = r - yp * dWp.d;
  const bool valid = ((long long)m < Mp) && (yp >= 1) && ((int)yp <= H) && (xp >= 1) && ((int)xp <= W);
  return valid ? (int)((b * (uint32_t)H + (yp - 1)) * (uint32_t)W + (xp - 1)) : -1;
}
// 16-byte vector reduction into global memory (sm_90+): one L2 atomic op for four consecutive floats
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.  Addresses / offsets in 16-byte units.
//   K-major  operand: rows of 128 B (64 bf16 along K); 8-row groups SBO = 1024 B apart; LBO unused (1)
//   MN-major operand: rows of 128 B (64 bf16 along M/N) per k; 8-k groups SBO = 1024 B apart; next 64 M/N
//                     elements LBO bytes away
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  d |= 2ull << 61;   // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B bf16, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------
// host: TMA descriptors through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    cudaGetLastError();
  }
  return fn;
}
int conv_tc_supported() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    cached = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess && major == 10 &&
        get_encode() != nullptr)
      cached = 1;
    cudaGetLastError();
  }
  return cached;
}
// 2-D bf16 tensor [rows][inner] (inner contiguous), box [box_rows][64], 128-byte swizzle, zero fill out of range
static int make_tmap(CUtensorMap* tm, const void* base, long long inner, long long rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  L3_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  L3_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld box=%d", (int)r, inner, rows, box_rows);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// weight pack: fp32 HWIO -> bf16 rows [(tap*KC + kc)*Cout + co][64 ci]  (K-major B operand, one 128-byte row each)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_pack_weights(const float* __restrict__ w, bf16* __restrict__ out, int Cin, int Cout, int flip) {
  // output conv has Ci' inputs and Co' outputs: forward (Ci',Co') = (Cin,Cout); dgrad (Ci',Co') = (Cout,Cin)
  const int Ci = flip ? Cout : Cin, Co = flip ? Cin : Cout;
  const int KC = Ci / 64;
  const long long n = 9LL * Ci * Co;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int cil = (int)(i & 63);
    long long r = i >> 6;
    int co = (int)(r % Co);
    long long r2 = r / Co;
    int kc = (int)(r2 % KC);
    int tap = (int)(r2 / KC);
    int ci = kc * 64 + cil;
    float v = flip ? w[((long long)(8 - tap) * Cin + co) * Cout + ci]   // w[8-tap][ci_orig = co'][co_orig = ci']
                   : w[((long long)tap * Cin + ci) * Cout + co];
    out[i] = __float2bfloat16_rn(v);
  }
}
// 16-bit hi / lo parts of x: x ~ hi + lo with hi = round16(x), lo = round16(x - hi)
__device__ __forceinline__ unsigned short part16(float x, int which, int fp16) {
  if (fp16) {
    const __half h = __float2half_rn(x);
    const __half r = which ? __float2half_rn(x - __half2float(h)) : h;
    return *reinterpret_cast<const unsigned short*>(&r);
  }
  const bf16 h = __float2bfloat16_rn(x);
  const bf16 r = which ? __float2bfloat16_rn(x - __bfloat162float(h)) : h;
  return *reinterpret_cast<const unsigned short*>(&r);
}
// all layers of a step in one launch: blockIdx.y = job.  split jobs (parity mode on tensor cores) write the operand of a
// convolution with 3*Ci input channels: segments [hi ; hi ; lo] of scale * w, as fp16 or bf16 parts
__global__ void k_pack_weights_batch(PackBatch pb) {
  const PackJob j = pb.job[blockIdx.y];
  const int Cin = j.Cin, Cout = j.Cout, flip = j.flip;
  const int Ci = flip ? Cout : Cin, Co = flip ? Cin : Cout;
  const int Cik = j.split ? 3 * Ci : Ci;      // channels of the kernel's K axis
  const int KC = Cik / 64;
  const long long n = 9LL * Cik * Co;
  unsigned short* const out16 = reinterpret_cast<unsigned short*>(j.out);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int cil = (int)(i & 63);
    long long r = i >> 6;
    int co = (int)(r % Co);
    long long r2 = r / Co;
    int kc = (int)(r2 % KC);
    int tap = (int)(r2 / KC);
    int cik = kc * 64 + cil;
    const int seg = cik / Ci, ci = cik - seg * Ci;
    float v = flip ? j.w[((long long)(8 - tap) * Cin + co) * Cout + ci] : j.w[((long long)tap * Cin + ci) * Cout + co];
    if (j.split) out16[i] = part16(v * j.scale, seg == 2 ? 1 : 0, j.fp16);
    else j.out[i] = __float2bfloat16_rn(v);
  }
}
int launch_pack_weights_batch(const PackBatch& pb, cudaStream_t s) {
  if (pb.n == 0) return 0;
  L3_REQUIRE(pb.n <= kMaxPackJobs, "pack batch too large");
  dim3 grid(148, pb.n);
  k_pack_weights_batch<<<grid, 256, 0, s>>>(pb);
  L3_CHECK_LAUNCH();
  return 0;
}

int launch_pack_weights_tc(const float* w, bf16* packed, int Cin, int Cout, int flip_transpose, cudaStream_t s) {
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "pack_weights: channels must be multiples of 64");
  long long n = 9LL * Cin * Cout;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_pack_weights<<<blocks, 256, 0, s>>>(w, packed, Cin, Cout, flip_transpose);
  L3_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// forward / dgrad
// ---------------------------------------------------------------------------------------------------------
static const int kBM = 128;   // pixel rows of one UMMA M tile

// ---------------------------------------------------------------------------------------------------------
// forward / dgrad, version 3: CTA pairs (tcgen05 cta_group::2), N = 256
// ---------------------------------------------------------------------------------------------------------
// Version 2 with N = 256 is bound by shared-memory bandwidth: every 128x256x16 MMA re-reads an 8 KB weight slice
// for only 128 output rows (4 + 8 KB per 128 cycles) while TMA writes the next stages into the same memory.  A CTA
// pair (two SMs of one TPC) issues ONE 256x256x16 MMA: each CTA stages its own 128 pixel rows (A) and HALF of the
// weight tile (128 of the 256 output channels); the tensor cores of both SMs read both halves, so per SM the
// operand traffic per FLOP halves.  The leader CTA (cluster rank 0) issues the MMAs; both CTAs run TMA (their
// transaction bytes land on the leader's mbarriers), and tcgen05.commit multicasts "stage free" / "accumulator
// ready" to both.
static const int kConv3Threads = 64 + 256;
static const int kMaxCoutTc = 512;   // bias staging in shared memory
// epilogue flavours of k_conv3x3_tc3 (how the per-channel BN statistics are reduced across the 32 pixel rows a warp holds)
enum { EPI_NONE = 0,        // no statistics (dgrad, inference): leanest register footprint
       EPI_BUTTERFLY = 1,   // 31-shuffle transposing butterfly per 32x32 chunk (~350 instructions)
       EPI_SMEM = 2,        // transpose through a per-warp 32x36-float scratch (needs 36 KB of shared memory: BN < 256)
       EPI_CARRY = 3,       // per-lane partial sums carried in registers over ALL tiles, one butterfly per kernel
                            // (a warp must own a single chunk: BN = 64; costs ~50 registers per thread)
       EPI_FWD2 = 4,        // statistics taken in the STORE-phase mapping (lane = 16-byte piece of 4 rows): 8 column
                            // partials per lane and chunk, carried over the tiles; works for every BN
       EPI_BWD = 5,         // dgrad: the same mapping reads the matching 16 bytes of the layer-below's z (coalesced) and
                            // carries sum(dy), sum(dy*z) -- pass 1 of the BN/ReLU backward without re-reading da
       EPI_F32 = 7,         // parity mode on tensor cores: fp32 output = out_scale * accumulator + bias (split 16-bit
                            // operands concatenated along K, see launch_conv3x3_tc_split); no statistics
       EPI_ACT = 6 };       // inference: BatchNorm (moving statistics folded to scale / shift) + ReLU applied to the fp32
                            // accumulator before the bf16 rounding; optionally stored straight into the NEXT layer's
                            // zero-haloed padded input -- no z tensor, no activation pass
template <int BN_, int MT_, int EPI_>
struct Conv3Cfg {
  static const int BN = BN_, MT = MT_;
  static const int kARows = MT * kBM + 8;          // this CTA's MT m-tiles + halo
  static const int kABoxes = (MT == 4) ? 5 : (MT == 2 ? 3 : 1);
  static const int kABoxRows = kARows / kABoxes;   // 104 / 88 / 136
  static const int kAStage = kARows * 128;
  static const int kAStages = (MT == 4) ? 2 : (MT == 2 ? 3 : 4);
  static const int kBStage = (BN / 2) * 128;       // this CTA's half of the weight tile
  static const int kBStages = 8;
  static const int EPI = EPI_;
  static const int kStatScratch = (EPI == EPI_SMEM) ? 8 * 32 * 36 * 4 : 0;   // per-epilogue-warp transposition scratch
  static const int kStoreScratch = 8 * 32 * 64;    // per-epilogue-warp store transposition scratch
  static const int kSmem = kAStages * kAStage + kBStages * kBStage + kStoreScratch + kStatScratch + 1024;
  static const int kTmemCols = 2 * MT * BN;        // 512 for (256,1), (128,2), (64,4)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose transaction bytes are credited to the mbarrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_addr(dst)), "l"(tm), "r"(smem_addr(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_lh_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %6, 0;\n mov.b64 da, {%1, %2};\n mov.b64 db, {%3, %4};\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_addr(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n .reg .b32 ra;\n mapa.shared::cluster.u32 ra, %0, %1;\n mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(
          smem_addr(bar)),
      "r"(rank)
      : "memory");
}

// Transposing butterfly over a warp: on entry lane r holds 32 per-column partial values a[0..31] (and b[0..31]); on
// exit lane i holds in a[0] (b[0]) the sum over all 32 lanes of column i.  31 shuffles per array.
__device__ __forceinline__ void col_butterfly(float (&a)[32], float (&b)[32], int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int hv = 16 >> step;              // values kept per lane after this step
    const bool upper = (lane >> (4 - step)) & 1;
#pragma unroll
    for (int j = 0; j < hv; ++j) {
      const float send_a = upper ? a[j] : a[j + hv];
      const float keep_a = upper ? a[j + hv] : a[j];
      const float send_b = upper ? b[j] : b[j + hv];
      const float keep_b = upper ? b[j + hv] : b[j];
      a[j] = keep_a + __shfl_xor_sync(0xffffffffu, send_a, 16 >> step);
      b[j] = keep_b + __shfl_xor_sync(0xffffffffu, send_b, 16 >> step);
    }
  }
}

template <int BN, int MT, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConv3Threads, 1)
k_conv3x3_tc3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const float* __restrict__ bias, bf16* __restrict__ out, int H, int W, int Cin, int Cout, long long Mp,
              int num_m_pairs, int num_n_tiles, double* __restrict__ stats, int relu_stats, FastDiv dHWp, FastDiv dWp,
              const bf16* __restrict__ zprev, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
              int out_padded, int ab_fp16, float out_scale) {
  using Cfg = Conv3Cfg<BN, MT, EPI>;
  constexpr int AST = Cfg::kAStages, BST = Cfg::kBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + AST * Cfg::kAStage;
  __shared__ __align__(8) uint64_t a_full[AST], a_empty[AST], b_full[BST], b_empty[BST], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[kMaxCoutTc];   // whole bias vector, read by the epilogue as float4
  // EPI_BWD: BN scale | shift of the layer below; EPI_ACT: of this layer
  __shared__ __align__(16) float s_aux[(EPI == EPI_BWD || EPI == EPI_ACT) ? 2 * kMaxCoutTc : 4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (bias != nullptr)
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_bias[i] = __ldg(bias + i);
  if (EPI == EPI_BWD || EPI == EPI_ACT)
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) {
      s_aux[i] = __ldg(bn_scale + i);
      s_aux[kMaxCoutTc + i] = __ldg(bn_shift + i);
    }
  if (threadIdx.x == 0) {
    for (int i = 0; i < AST; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < BST; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 16); }   // 8 warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)),
                 "n"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int KC = Cin >> 6;
  const int Wp = W + 2;
  const int num_tiles = num_m_pairs * num_n_tiles;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own 128 pixel rows + own half of the weight tile =====
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / num_n_tiles) * (2 * MT * kBM) + (int)rank * (MT * kBM);
        const int n0 = (tile % num_n_tiles) * BN + (int)rank * (BN / 2);
        for (int kc = 0; kc < KC; ++kc)
          for (int ky = 0; ky < 3; ++ky) {
            mbar_wait(&a_empty[as], aph ^ 1);
            if (rank == 0) mbar_expect_tx(&a_full[as], 2 * Cfg::kAStage);
            const int row0 = m0 + (ky - 1) * Wp - 1;
#pragma unroll
            for (int bx = 0; bx < Cfg::kABoxes; ++bx)
              tma_load_2d_pair(&tmA, &a_full[as], smem + as * Cfg::kAStage + bx * Cfg::kABoxRows * 128, kc * 64,
                               row0 + bx * Cfg::kABoxRows);
            if (++as == AST) { as = 0; aph ^= 1; }
            for (int kx = 0; kx < 3; ++kx) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (rank == 0) mbar_expect_tx(&b_full[bs], 2 * Cfg::kBStage);
              tma_load_2d_pair(&tmB, &b_full[bs], smem_b + bs * Cfg::kBStage, 0, ((ky * 3 + kx) * KC + kc) * Cout + n0);
              if (++bs == BST) { bs = 0; bph ^= 1; }
            }
          }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===== MMA issuer (leader CTA only) =====
      // A / B element format: bf16 (format code 1 in bits 7..9 / 10..12) or, for the split-fp16 operands of the parity
      // mode's forward pass, fp16 (code 0); same 16-bit layout, same MMA rate
      const uint32_t idesc = ab_fp16 ? (make_idesc(2 * kBM, BN, 0, 0) & ~((1u << 7) | (1u << 10))) : make_idesc(2 * kBM, BN, 0, 0);
      constexpr uint32_t hi = desc_hi(1024);
      const bool leader = elect_one();
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t sa_base = smem_addr(smem), sb_base = smem_addr(smem_b);
      // EPI_F32 (parity mode): the accumulator is handed to the epilogue every kFlushIts (ky, channel-chunk) iterations
      // and restarted from zero -- tcgen05 adds each MMA's products into the fp32 accumulator with truncation, a bias of
      // ~2^-25 of the running sum per MMA that grows linearly with the chain length (measured 2.9e-5 relative at
      // K = 3 * 4608); short chains summed by the epilogue in registers (round-to-nearest) keep it at ~1e-6
      constexpr int kFlushIts = (EPI == EPI_F32) ? 3 : (1 << 30);
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        uint32_t d_tmem = 0;
        uint32_t first = 1;
        for (int it = 0; it < 3 * KC; ++it) {
          if (it % kFlushIts == 0) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            d_tmem = tmem_base + acc * (MT * BN);
            first = 1;
          }
          mbar_wait(&a_full[as], aph);
          const uint32_t a_lo0 = desc_lo(sa_base + as * Cfg::kAStage, 16);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_lo0 = desc_lo(sb_base + bs * Cfg::kBStage, 16);
            if (leader) {
#pragma unroll
              for (int t = 0; t < MT; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_lh_pair(d_tmem + t * BN, a_lo0 + (uint32_t)((t * kBM + kx) * 8 + k * 2), hi,
                                    b_lo0 + (uint32_t)(k * 2), hi, idesc, (first && kx == 0 && k == 0) ? 0u : 1u);
              umma_commit_pair(&b_empty[bs]);
            }
            __syncwarp();
            if (++bs == BST) { bs = 0; bph ^= 1; }
          }
          first = 0;
          if (leader) umma_commit_pair(&a_empty[as]);
          __syncwarp();
          if (++as == AST) { as = 0; aph ^= 1; }
          if (it % kFlushIts == kFlushIts - 1 || it == 3 * KC - 1) {
            if (leader) umma_commit_pair(&tfull_bar[acc]);
            __syncwarp();
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs, own 128 rows): TMEM -> regs -> (+bias from smem) -> bf16 -> HBM, plus the BN batch
    // statistics of the stored values.  Lane = pixel row, so per-channel sums need a reduction ACROSS lanes; the
    // EPI template parameter selects how (see the enum above).  The per-item butterfly made the Cin <= 128 layers
    // epilogue-bound.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int NCH = BN / 32, NST = NCH / 2;
    static_assert(EPI != EPI_CARRY || NST == 1, "EPI_CARRY needs one chunk per warp");
    uint4* const st_scr = reinterpret_cast<uint4*>(smem_b + BST * Cfg::kBStage) + (warp - 2) * 128;   // 32 rows x 64 B
    float* const sw = reinterpret_cast<float*>(smem_b + BST * Cfg::kBStage + Cfg::kStoreScratch) + (warp - 2) * (32 * 36);
    int acc = 0;
    uint32_t acc_phase = 0;
    float st_sum[NST], st_sq[NST];
#pragma unroll
    for (int i = 0; i < NST; ++i) st_sum[i] = st_sq[i] = 0.f;
    float ca[EPI == EPI_CARRY ? 32 : 1], cb[EPI == EPI_CARRY ? 32 : 1];
#pragma unroll
    for (int j = 0; j < (EPI == EPI_CARRY ? 32 : 1); ++j) ca[j] = cb[j] = 0.f;
    // EPI_FWD2 / EPI_BWD: lane l carries, per chunk slot, the partial sums of columns (l & 3) * 8 .. + 7 over the rows
    // (l >> 2) + 8 i it stores
    constexpr bool kStorePhaseStats = (EPI == EPI_FWD2 || EPI == EPI_BWD);
    float f1[kStorePhaseStats ? NST : 1][8], f2[kStorePhaseStats ? NST : 1][8];
#pragma unroll
    for (int c = 0; c < (kStorePhaseStats ? NST : 1); ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) f1[c][j] = f2[c][j] = 0.f;
    int st_n0 = -1;
    auto flush_stats = [&]() {
      if (kStorePhaseStats) {
        if (stats != nullptr && st_n0 >= 0) {
          __syncwarp();
#pragma unroll
          for (int c = 0; c < (kStorePhaseStats ? NST : 1); ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a = f1[c][j], b = f2[c][j];
#pragma unroll
              for (int o = 4; o < 32; o <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
              }
              if (lane < 4) {
                const int col = st_n0 + (2 * c + half) * 32 + lane * 8 + j;
                atomicAdd(&stats[col], (double)a);
                atomicAdd(&stats[Cout + col], (double)b);
              }
              f1[c][j] = f2[c][j] = 0.f;
            }
        }
        return;
      }
      if (EPI != EPI_NONE && stats != nullptr && st_n0 >= 0) {
        if (EPI == EPI_CARRY) {
          __syncwarp();
          float (&fa)[32] = reinterpret_cast<float (&)[32]>(ca);
          float (&fb)[32] = reinterpret_cast<float (&)[32]>(cb);
          col_butterfly(fa, fb, lane);
          st_sum[0] = ca[0];
          st_sq[0] = cb[0];
#pragma unroll
          for (int j = 0; j < (EPI == EPI_CARRY ? 32 : 1); ++j) ca[j] = cb[j] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < NST; ++i) {
          const int col = st_n0 + (2 * i + half) * 32 + lane;
          atomicAdd(&stats[col], (double)st_sum[i]);
          atomicAdd(&stats[Cout + col], (double)st_sq[i]);
          st_sum[i] = st_sq[i] = 0.f;
        }
      }
    };
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const unsigned mbase = (unsigned)(tile / num_n_tiles) * (2 * MT * kBM) + rank * (MT * kBM) + q * 32 + lane;
      const int n0 = (tile % num_n_tiles) * BN;
      if (EPI == EPI_F32) {
        // parity mode: sum the short accumulation chains (see the MMA issuer) in registers, round-to-nearest
        constexpr int NR = (EPI == EPI_F32) ? MT * NST * 32 : 1;
        float racc[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) racc[i] = 0.f;
        const int n_groups = (3 * KC + 2) / 3;
#pragma unroll 1
        for (int grp = 0; grp < n_groups; ++grp) {
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
#pragma unroll
          for (int t = 0; t < MT; ++t)
#pragma unroll
            for (int chh = 0; chh < NST; ++chh)
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * (MT * BN) + t * BN + (2 * chh + half) * 32 + hh * 16), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) racc[(EPI == EPI_F32) ? ((t * NST + chh) * 32 + hh * 16 + j) : 0] += __uint_as_float(v[j]);
              }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        float* const outf = reinterpret_cast<float*>(out);
#pragma unroll
        for (int t = 0; t < MT; ++t) {   // unrolled: racc must be indexed statically to stay in registers
          const int opix = out_pixel(mbase + t * kBM, Mp, H, W, dHWp, dWp);
          int spix[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) spix[i] = __shfl_sync(0xffffffffu, opix, (lane >> 2) + 8 * i);
#pragma unroll
          for (int chh = 0; chh < NST; ++chh) {
            const int c0 = (2 * chh + half) * 32;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              // 16 columns (64 bytes per row) at a time through the swizzled scratch: one store instruction writes
              // 8 rows x 64 contiguous bytes
              uint4* const srow = st_scr + lane * 4;
              const int sw4 = (lane >> 1) & 3;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias != nullptr) bb = *reinterpret_cast<const float4*>(&s_bias[n0 + c0 + hh * 16 + 4 * c]);
                const int r0 = (EPI == EPI_F32) ? ((t * NST + chh) * 32 + hh * 16 + 4 * c) : 0;
                uint4 u;
                u.x = __float_as_uint(fmaf(racc[r0], out_scale, bb.x));
                u.y = __float_as_uint(fmaf(racc[(EPI == EPI_F32) ? r0 + 1 : 0], out_scale, bb.y));
                u.z = __float_as_uint(fmaf(racc[(EPI == EPI_F32) ? r0 + 2 : 0], out_scale, bb.z));
                u.w = __float_as_uint(fmaf(racc[(EPI == EPI_F32) ? r0 + 3 : 0], out_scale, bb.w));
                srow[c ^ sw4] = u;
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int R = (lane >> 2) + 8 * i;
                const uint4 val = st_scr[R * 4 + ((lane & 3) ^ ((R >> 1) & 3))];
                if (spix[i] >= 0)
                  *reinterpret_cast<uint4*>(outf + (long long)spix[i] * Cout + n0 + c0 + hh * 16 + (lane & 3) * 4) = val;
              }
              __syncwarp();   // the scratch is rewritten by the next group
            }
          }
        }
        continue;
      }
      if (n0 != st_n0) { flush_stats(); st_n0 = n0; }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < MT; ++t) {
        // output pixel of this lane's row, -1 for halo / out-of-range rows (32-bit index math: Mp < 2^31)
        const int opix = out_pixel(mbase + t * kBM, Mp, H, W, dHWp, dWp);
        const bool valid = opix >= 0;
        // destination row: the un-padded pixel index, or (EPI_ACT, out_padded) the padded index itself -- the output then IS
        // the next layer's zero-haloed input; halo rows are never written and stay zero
        const int dpix = (EPI == EPI_ACT && out_padded && valid) ? (int)(mbase + t * kBM) : opix;
        // store phase: lane l writes 16-byte piece (l & 3) of rows (l >> 2) + 8 i
        int spix[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) spix[i] = __shfl_sync(0xffffffffu, dpix, (lane >> 2) + 8 * i);
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * (MT * BN) + t * BN);
#pragma unroll
        for (int chh = 0; chh < NST; ++chh) {
          const int c0 = (2 * chh + half) * 32;
          uint32_t pk[16];
          uint4 zr[EPI == EPI_BWD ? 4 : 1];
          if (EPI == EPI_BWD) {
            // the layer-below's z for the pieces this lane will store: issued before the TMEM load / transposition
#pragma unroll
            for (int i = 0; i < (EPI == EPI_BWD ? 4 : 1); ++i)
              if (spix[i] >= 0)
                zr[i] = *reinterpret_cast<const uint4*>(zprev + (long long)spix[i] * Cout + n0 + c0 + (lane & 3) * 8);
          }
          if (EPI == EPI_CARRY || EPI == EPI_NONE || EPI == EPI_ACT || kStorePhaseStats) {
            // 16 columns at a time: half the live registers of the 32-column path
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[16];
              tmem_ld16(t_row + c0 + hh * 16, v);
              tmem_ld_wait();
              if (EPI == EPI_ACT) {
                // y = relu(scale * (acc + bias) + shift), or scale * relu(acc + bias) + shift for the Conv -> ReLU -> BN
                // layer (relu_stats doubles as that flag): the arithmetic of k_act_fwd on the un-rounded accumulator
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int col = n0 + c0 + hh * 16 + 4 * j;
                  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (bias != nullptr) bb = *reinterpret_cast<const float4*>(&s_bias[col]);
                  const float4 sc = *reinterpret_cast<const float4*>(&s_aux[col]);
                  const float4 sf = *reinterpret_cast<const float4*>(&s_aux[kMaxCoutTc + col]);
                  float y0 = __uint_as_float(v[4 * j]) + bb.x, y1 = __uint_as_float(v[4 * j + 1]) + bb.y;
                  float y2 = __uint_as_float(v[4 * j + 2]) + bb.z, y3 = __uint_as_float(v[4 * j + 3]) + bb.w;
                  if (relu_stats) {
                    y0 = fmaf(fmaxf(y0, 0.f), sc.x, sf.x); y1 = fmaf(fmaxf(y1, 0.f), sc.y, sf.y);
                    y2 = fmaf(fmaxf(y2, 0.f), sc.z, sf.z); y3 = fmaf(fmaxf(y3, 0.f), sc.w, sf.w);
                  } else {
                    y0 = fmaxf(fmaf(y0, sc.x, sf.x), 0.f); y1 = fmaxf(fmaf(y1, sc.y, sf.y), 0.f);
                    y2 = fmaxf(fmaf(y2, sc.z, sf.z), 0.f); y3 = fmaxf(fmaf(y3, sc.w, sf.w), 0.f);
                  }
                  pk[hh * 8 + 2 * j] = pack_bf16x2(y0, y1);
                  pk[hh * 8 + 2 * j + 1] = pack_bf16x2(y2, y3);
                }
              } else if (bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 bb = *reinterpret_cast<const float4*>(&s_bias[n0 + c0 + hh * 16 + 4 * j]);
                  pk[hh * 8 + 2 * j] = pack_bf16x2(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
                  pk[hh * 8 + 2 * j + 1] = pack_bf16x2(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  pk[hh * 8 + j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
              }
            }
          } else {
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);
            tmem_ld_wait();
            if (bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = *reinterpret_cast<const float4*>(&s_bias[n0 + c0 + 4 * j]);
                pk[2 * j] = pack_bf16x2(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
                pk[2 * j + 1] = pack_bf16x2(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            }
          }
          // coalesced store: lane = row would emit 32 scattered 16-byte pieces per instruction (one L2 request each --
          // measured: the request rate, not the bytes, bounded the Cout <= 128 layers).  The 32x64-byte chunk is
          // transposed through a private swizzled scratch so that one instruction writes 8 rows x 64 contiguous bytes.
          {
            uint4* const srow = st_scr + lane * 4;
            const int sw4 = (lane >> 1) & 3;
#pragma unroll
            for (int c = 0; c < 4; ++c) srow[c ^ sw4] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int R = (lane >> 2) + 8 * i;
              const uint4 val = st_scr[R * 4 + ((lane & 3) ^ ((R >> 1) & 3))];
              if (spix[i] >= 0) {
                *reinterpret_cast<uint4*>(out + (long long)spix[i] * Cout + n0 + c0 + (lane & 3) * 8) = val;
                if (kStorePhaseStats && stats != nullptr) {
                  const uint32_t w4[4] = {val.x, val.y, val.z, val.w};
                  float x[8];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    x[2 * j] = __uint_as_float(w4[j] << 16);
                    x[2 * j + 1] = __uint_as_float(w4[j] & 0xffff0000u);
                  }
                  constexpr int CS = kStorePhaseStats ? 1 : 0;
                  if (EPI == EPI_FWD2) {
                    // statistics of the values as stored (bf16-rounded)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float xv = relu_stats ? fmaxf(x[j], 0.f) : x[j];
                      f1[chh * CS][j] += xv;
                      f2[chh * CS][j] = fmaf(xv, xv, f2[chh * CS][j]);
                    }
                  } else {
                    // dy = da where bn(z) > 0 (ReLU after BN), raw sums sum(dy), sum(dy*z) -- k_bwd_stats<bf16>'s contract
                    const uint32_t z4[4] = {zr[i * CS].x, zr[i * CS].y, zr[i * CS].z, zr[i * CS].w};
                    const float* scp = &s_aux[n0 + c0 + (lane & 3) * 8];
                    const float4 sc0 = *reinterpret_cast<const float4*>(scp), sc1 = *reinterpret_cast<const float4*>(scp + 4);
                    const float4 sh0 = *reinterpret_cast<const float4*>(scp + kMaxCoutTc);
                    const float4 sh1 = *reinterpret_cast<const float4*>(scp + kMaxCoutTc + 4);
                    const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
                    const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float zz = __uint_as_float((j & 1) ? (z4[j >> 1] & 0xffff0000u) : (z4[j >> 1] << 16));
                      const float d = fmaf(zz, sc[j], sh[j]) > 0.f ? x[j] : 0.f;
                      f1[chh * CS][j] += d;
                      f2[chh * CS][j] = fmaf(d, zz, f2[chh * CS][j]);
                    }
                  }
                }
              }
            }
            __syncwarp();   // the scratch is rewritten by the next chunk
          }
          if (EPI == EPI_NONE || EPI == EPI_ACT || kStorePhaseStats || stats == nullptr) continue;
          // statistics of the values as stored (bf16-rounded); halo / out-of-range rows contribute nothing
          if (EPI == EPI_CARRY) {
            if (valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float x0 = __uint_as_float(pk[j] << 16);
                float x1 = __uint_as_float(pk[j] & 0xffff0000u);
                if (relu_stats) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                const int i0 = (EPI == EPI_CARRY) ? 2 * j : 0, i1 = (EPI == EPI_CARRY) ? 2 * j + 1 : 0;
                ca[i0] += x0; ca[i1] += x1;
                cb[i0] = fmaf(x0, x0, cb[i0]); cb[i1] = fmaf(x1, x1, cb[i1]);
              }
            }
            __syncwarp();
          } else if (EPI == EPI_SMEM) {
            float4* const wrow = reinterpret_cast<float4*>(sw + lane * 36);
            if (valid) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 x = make_float4(__uint_as_float(pk[2 * j] << 16), __uint_as_float(pk[2 * j] & 0xffff0000u),
                                       __uint_as_float(pk[2 * j + 1] << 16), __uint_as_float(pk[2 * j + 1] & 0xffff0000u));
                if (relu_stats) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                wrow[j] = x;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) wrow[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              const float x = sw[rr * 36 + lane];
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
            st_sum[chh] += s1;
            st_sq[chh] += s2;
            __syncwarp();   // the scratch is rewritten by the next chunk
          } else {
            float a[32], b2[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x0 = valid ? __uint_as_float(pk[j] << 16) : 0.f;
              float x1 = valid ? __uint_as_float(pk[j] & 0xffff0000u) : 0.f;
              if (relu_stats) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
              a[2 * j] = x0; a[2 * j + 1] = x1;
              b2[2 * j] = x0 * x0; b2[2 * j + 1] = x1 * x1;
            }
            col_butterfly(a, b2, lane);
            st_sum[chh] += a[0];
            st_sq[chh] += b2[0];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);   // the leader's MMA warp owns the accumulator hand-off
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody leaves while the peer may still signal its barriers / read its operands
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
  }
}

// EPI_BWD only: z / BN scale / BN shift of the layer whose activation gradient this dgrad produces
struct BwdFuse {
  const bf16* z;
  const float *scale, *shift;
  int out_padded = 0;   // EPI_ACT: store into the zero-haloed padded layout
  int f32_out = 0;      // EPI_F32: fp32 output = out_scale * accumulator + bias
  int ab_fp16 = 0;      //          operands are fp16 (else bf16)
  float out_scale = 1.f;
};
template <int BN, int MT, int EPI>
static int launch_conv3(const bf16* in, const bf16* packed_w, const float* bias, bf16* out, int H, int W, int Cin, int Cout,
                        long long Mp, double* stats, int relu_stats, cudaStream_t s, BwdFuse bf = BwdFuse{nullptr, nullptr, nullptr}) {
  using Cfg = Conv3Cfg<BN, MT, EPI>;
  L3_REQUIRE(Cout <= kMaxCoutTc, "conv_tc: Cout=%d exceeds the bias staging buffer", Cout);
  static PerDeviceOnce once;
  if (once.needed()) {
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_conv3x3_tc3<BN, MT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    once.mark();
  }
  CUtensorMap tmA, tmB;
  if (make_tmap(&tmA, in, Cin, Mp, Cfg::kABoxRows)) return -1;
  if (make_tmap(&tmB, packed_w, 64, 9LL * (Cin / 64) * Cout, Cfg::BN / 2)) return -1;
  const int num_mp = (int)((Mp + 2 * MT * kBM - 1) / (2 * MT * kBM)), num_n = Cout / Cfg::BN;
  long long tiles = (long long)num_mp * num_n;
  int pairs = (int)(tiles < 74 ? tiles : 74);
  if (stats) L3_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * Cout, s));
  k_conv3x3_tc3<BN, MT, EPI><<<2 * pairs, kConv3Threads, Cfg::kSmem, s>>>(tmA, tmB, bias, out, H, W, Cin, Cout, Mp, num_mp, num_n,
                                                                 stats, relu_stats,
                                                                 make_fastdiv((uint32_t)(H + 2) * (uint32_t)(W + 2)),
                                                                 make_fastdiv((uint32_t)(W + 2)), bf.z, bf.scale, bf.shift,
                                                                 bf.out_padded, bf.ab_fp16, bf.out_scale);
  L3_CHECK_LAUNCH();
  return 0;
}

static int launch_conv3_any(int BN, const bf16* in, const bf16* packed_w, const float* bias, bf16* out, int H, int W, int Cin,
                            int Cout, long long Mp, double* stats, int relu_stats, cudaStream_t s,
                            BwdFuse bf = BwdFuse{nullptr, nullptr, nullptr}) {
#define L3_GO(bn, mt, epi) return launch_conv3<bn, mt, epi>(in, packed_w, bias, out, H, W, Cin, Cout, Mp, stats, relu_stats, s, bf)
  if (bf.f32_out) {   // parity mode on tensor cores
    L3_REQUIRE(stats == nullptr, "fp32-output epilogue: no statistics");
    if (BN == 256) L3_GO(256, 1, EPI_F32);
    if (BN == 128) L3_GO(128, 2, EPI_F32);
    L3_GO(64, 4, EPI_F32);
  }
  if (bf.z == nullptr && bf.scale != nullptr) {   // inference: fused BatchNorm + ReLU epilogue
    L3_REQUIRE(stats == nullptr, "fused activation epilogue: no statistics");
    if (BN == 256) L3_GO(256, 1, EPI_ACT);
    if (BN == 128) L3_GO(128, 2, EPI_ACT);
    L3_GO(64, 4, EPI_ACT);
  }
  if (bf.z != nullptr) {   // dgrad with fused BN-backward statistics (stand-alone op; the step does not use it)
    L3_REQUIRE(stats != nullptr && bias == nullptr && !relu_stats, "fused dgrad statistics: stats, no bias, no relu_first");
    if (BN == 256) L3_GO(256, 1, EPI_BWD);
    if (BN == 128) L3_GO(128, 2, EPI_BWD);
    L3_GO(64, 4, EPI_BWD);
  }
  if (stats == nullptr) {
    if (BN == 256) L3_GO(256, 1, EPI_NONE);
    if (BN == 128) L3_GO(128, 2, EPI_NONE);
    L3_GO(64, 4, EPI_NONE);
  }
  // statistics flavour per tile configuration, chosen by measurement (B = 64, us per launch): BN 256: butterfly 127 vs
  // store-phase 136; BN 128: store-phase 138 vs shared-memory transposition 143; BN 64: carried 216 vs store-phase 225
  if (BN == 256) L3_GO(256, 1, EPI_BUTTERFLY);
  if (BN == 128) L3_GO(128, 2, EPI_FWD2);
  L3_GO(64, 4, EPI_CARRY);
#undef L3_GO
}

// (Measured on B200: the UMMA descriptor "base offset" field must stay 0 for row-shifted operand starts -- the swizzle is
// applied to absolute shared-memory address bits; setting the field to (addr >> 7) & 7 produces wrong results.)
int conv_tc_fuses_stats() { return 1; }

int launch_conv3x3_tc(const bf16* in, const bf16* packed_w, const float* bias, bf16* out, int B, int H, int W, int Cin,
                      int Cout, double* stats, int relu_stats, cudaStream_t s) {
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv_tc: channels must be multiples of 64 (Cin=%d Cout=%d)", Cin, Cout);
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Mp + 4LL * (W + 2) + 1024 < 0x7fffffffLL, "conv_tc: too many pixels for 32-bit TMA coordinates");
  const int BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  return launch_conv3_any(BN, in, packed_w, bias, out, H, W, Cin, Cout, Mp, stats, relu_stats, s);
}

// Inference forward with BatchNorm (scale / shift from the moving statistics) and ReLU in the epilogue (relu_first: the
// Conv -> ReLU -> BN order).  out_padded = 1: `out` is the next layer's zero-haloed padded input (B,H+2,W+2,Cout) whose
// halo must already be zero; 0: un-padded (B,H,W,Cout), e.g. ahead of a max-pool pass.
int launch_conv3x3_tc_act(const bf16* in, const bf16* packed_w, const float* bias, bf16* out, int B, int H, int W, int Cin,
                          int Cout, const float* scale, const float* shift, int relu_first, int out_padded, cudaStream_t s) {
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv_tc: channels must be multiples of 64 (Cin=%d Cout=%d)", Cin, Cout);
  L3_REQUIRE(scale != nullptr && shift != nullptr, "conv_tc_act: scale / shift");
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Mp + 4LL * (W + 2) + 1024 < 0x7fffffffLL, "conv_tc: too many pixels for 32-bit TMA coordinates");
  const int BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  BwdFuse bf{nullptr, scale, shift};
  bf.out_padded = out_padded;
  return launch_conv3_any(BN, in, packed_w, bias, out, H, W, Cin, Cout, Mp, nullptr, relu_first, s, bf);
}

// Pass 1 of the BN/ReLU backward in the dgrad epilogue (EPI_BWD): the z loads sit on the epilogue's critical path, so it
// pays only where the main loop is long enough to hide them -- 128 channels and up (api.cu tower_backward_layer has the
// measured numbers); at 64 channels (356 us fused vs 210 + 135 us separate) the step keeps the separate pass.
// `channels_below` = channels of the layer whose statistics would be fused (= the data gradient's output channels).
int conv_tc_fuses_bwd_stats(int channels_below) { return channels_below >= 128 ? 1 : 0; }

// Parity mode on tensor cores: `in` holds 16-bit SPLIT operands concatenated along the channel axis,
//   [hi | lo | hi] (3*Cin channels, zero-haloed padded) x packed weights [hi ; hi ; lo]  ->  a_hi w_hi + a_lo w_hi + a_hi w_lo
// accumulated in fp32 in TMEM -- a convolution with Cin' = 3*Cin as far as the kernel is concerned.  Output: fp32
// un-padded (B,H,W,Cout) = out_scale * acc + bias.  fp16 = 1: the parts are fp16 (2 x 11 significant bits: forward pass,
// weights pre-scaled by 1/out_scale to stay in fp16's normal range); 0: bf16 (2 x 8 bits, fp32's range: backward pass).
int launch_conv3x3_tc_split(const void* in_split, const void* packed_w_split, const float* bias, float* out, int B, int H, int W,
                            int Cin, int Cout, int fp16, float out_scale, cudaStream_t s) {
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv_tc_split: channels must be multiples of 64 (Cin=%d Cout=%d)", Cin, Cout);
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Mp + 4LL * (W + 2) + 1024 < 0x7fffffffLL, "conv_tc: too many pixels for 32-bit TMA coordinates");
  const int BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  BwdFuse bf{nullptr, nullptr, nullptr};
  bf.f32_out = 1;
  bf.ab_fp16 = fp16;
  bf.out_scale = out_scale;
  return launch_conv3_any(BN, (const bf16*)in_split, (const bf16*)packed_w_split, bias, (bf16*)out, H, W, 3 * Cin, Cout, Mp,
                          nullptr, 0, s, bf);
}

int launch_dgrad3x3_tc_bwdstats(const bf16* dz, const bf16* packed_wt, bf16* da, int B, int H, int W, int Cout, int Cin,
                                const bf16* z_below, const float* scale_below, const float* shift_below, double* sums,
                                cudaStream_t s) {
  // a conv with Cin' = Cout (of the layer), Cout' = Cin (= channels of the layer below)
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "dgrad_tc: channels must be multiples of 64 (Cin=%d Cout=%d)", Cin, Cout);
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Mp + 4LL * (W + 2) + 1024 < 0x7fffffffLL, "conv_tc: too many pixels for 32-bit TMA coordinates");
  const int BN = (Cin % 256 == 0) ? 256 : (Cin % 128 == 0 ? 128 : 64);
  return launch_conv3_any(BN, dz, packed_wt, nullptr, da, H, W, Cout, Cin, Mp, sums, 0, s,
                          BwdFuse{z_below, scale_below, shift_below});
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------------
// One CTA = one "unit" (a set of G accumulator blocks of 128 x BN) over one slice of the pixel axis.
//   Cin >= 128 (PAIR = false): block j = tap (ky*3 + j) for the unit's ky, rows = 128 input channels, BN = 128, G = 3
//   Cin == 64  (PAIR = true) : block j = taps (2j, 2j+1) stacked in M (2 x 64 channels), BN = 64, G = 5
// Per 32-pixel chunk the producer loads, for every block, two [32 px][64 ch] boxes of A (at the taps' row shifts) and
// BN/64 boxes of dZ; both are MN-major UMMA operands.
static const int kWgThreads = 192;
// ---- weight gradient, version 2: shared-halo regions -------------------------------------------------------
// Per 64-pixel chunk a CTA loads, per filter row ky, ONE region of 72 pixel rows x 64 channels; the three kx taps
// are UMMA descriptors starting kx rows into it (MN-major: one pixel = one 128-byte row, so a pixel shift is a
// 128-byte start offset).  Cin == 64 packs two taps into the M = 128 rows of one MMA by pointing the descriptor's
// leading-dimension byte offset at the second tap's rows (128 B away inside a region, or in the next region).
//   Cin >= 128: unit = (ky, 128 input channels, 128 output channels): 2 regions (channel halves) + 2 dz boxes / stage,
//               3 accumulator blocks (kx);   Cin == 64: unit = 64 output channels: 3 regions (ky) + 1 dz box / stage,
//               5 accumulator blocks (tap pairs).  ~2x less L2->SMEM traffic per FLOP than version 1.
static const int kWg2Chunk = 64;
static const int kWg2RegionRows = kWg2Chunk + 8;
static const int kWg2RB = kWg2RegionRows * 128;   // 9216 B
static const int kWg2ZB = kWg2Chunk * 128;        // 8192 B

template <bool PAIR>
struct Wg2Cfg {
  static const int G = PAIR ? 5 : 3;
  static const int BN = PAIR ? 64 : 128;
  static const int NR = PAIR ? 3 : 2;
  static const int kABytes = NR * kWg2RB;
  static const int kBBytes = (BN / 64) * kWg2ZB;
  static const int kStageBytes = kABytes + kBBytes;
  static const int kStages = 5;
  static const int kSmem = kStages * kStageBytes + 1024;
  static const int kTmemCols = 512;
};

template <bool PAIR>
__global__ void __launch_bounds__(kWgThreads, 1)
k_wgrad3x3_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmZ, float* __restrict__ dw,
               int W, int Cin, int Cout, int total_chunks, int chunks_per_slice) {
  using Cfg = Wg2Cfg<PAIR>;
  constexpr int STAGES = Cfg::kStages, G = Cfg::G, BN = Cfg::BN, NR = Cfg::NR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)),
                 "n"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int unit = blockIdx.x;
  int cob, ci_base, ky;
  if (PAIR) { cob = unit; ci_base = 0; ky = 0; }
  else {
    const int n_cob = Cout / BN;
    ky = unit % 3;
    cob = (unit / 3) % n_cob;
    ci_base = (unit / 3 / n_cob) * 128;
  }
  const int co0 = cob * BN;
  const int Wp = W + 2;
  const int c_begin = blockIdx.y * chunks_per_slice;
  const int c_end = min(total_chunks, c_begin + chunks_per_slice);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int m = c * kWg2Chunk;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          // PAIR: region r = filter row r (all 64 channels); else region r = channel half r of filter row ky
          const int rky = PAIR ? r : ky;
          const int rci = PAIR ? 0 : ci_base + r * 64;
          tma_load_2d(&tmA, &full_bar[stage], sa + r * kWg2RB, rci, m + (rky - 1) * Wp - 1);
        }
#pragma unroll
        for (int nb = 0; nb < BN / 64; ++nb)
          tma_load_2d(&tmZ, &full_bar[stage], sa + Cfg::kABytes + nb * kWg2ZB, co0 + nb * 64, m);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp runs the loop (uniform registers), one elected lane issues
    constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
    constexpr uint32_t hi = desc_hi(1024);
    const bool leader = elect_one();
    const uint32_t s_base = smem_addr(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int c = c_begin; c < c_end; ++c) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = s_base + stage * Cfg::kStageBytes;
      if (leader) {
        const uint32_t b_lo0 = desc_lo(sa + Cfg::kABytes, kWg2ZB);
        const uint32_t accum = (c > c_begin) ? 1u : 0u;
#pragma unroll
        for (int j = 0; j < G; ++j) {
          // compile-time byte offset of the block's first tap inside the stage and its leading-dimension offset
          const int t0 = 2 * j, t1 = (2 * j + 1 > 8) ? 8 : 2 * j + 1;
          const int off0 = PAIR ? (t0 / 3) * kWg2RB + (t0 % 3) * 128 : j * 128;
          const int lbo = PAIR ? ((t1 / 3) * kWg2RB + (t1 % 3) * 128) - off0 : kWg2RB;
          const uint32_t a_lo0 = desc_lo(sa + off0, (uint32_t)lbo);
#pragma unroll
          for (int ks = 0; ks < kWg2Chunk / 16; ++ks)   // 16 pixels = 2048 B = 128 sixteen-byte units
            umma_bf16_lh(tmem_base + j * BN, a_lo0 + ks * 128, hi, b_lo0 + ks * 128, hi, idesc, (ks > 0) ? 1u : accum);
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (leader) umma_commit(&tfull_bar);
    __syncwarp();
  } else if (c_end > c_begin) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int h = row >> 6;
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int j = 0; j < G; ++j) {
      const int tap = PAIR ? min(2 * j + h, 8) : ky * 3 + j;
      const bool dup = PAIR && (2 * j + h > 8);
      const int ci = (PAIR ? 0 : ci_base + h * 64) + (row & 63);
      float* dst = dw + ((long long)tap * Cin + ci) * Cout + co0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * BN + c0), v);
        tmem_ld_wait();
        if (!dup) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            red_add_v4(dst + c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                       __uint_as_float(v[i + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
  }
}

// a_pitch / z_pitch: channels per pixel row in memory when the operand is a channel PREFIX of a wider buffer (the split
// [hi | lo | hi] buffers of the parity mode, of which the weight gradient uses [hi | lo]); 0 = dense
template <bool PAIR>
static int launch_wgrad2_cfg(const bf16* a, const bf16* dz, float* dw, int W, int Cin, int Cout, long long Mp,
                             cudaStream_t s, int a_pitch = 0, int z_pitch = 0) {
  using Cfg = Wg2Cfg<PAIR>;
  static PerDeviceOnce once;
  if (once.needed()) {
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad3x3_tc2<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    once.mark();
  }
  CUtensorMap tmA, tmZ;
  if (make_tmap(&tmA, a, a_pitch ? a_pitch : Cin, Mp, kWg2RegionRows)) return -1;
  if (make_tmap(&tmZ, dz, z_pitch ? z_pitch : Cout, Mp, kWg2Chunk)) return -1;
  const int units = PAIR ? Cout / Cfg::BN : (Cin / 128) * (Cout / Cfg::BN) * 3;
  const int total_chunks = (int)((Mp + kWg2Chunk - 1) / kWg2Chunk);
  int slices = (2 * 148) / units;   // units*slices <= 296: exactly two waves of one CTA per SM, never a third
  int max_slices = (total_chunks + 7) / 8;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  int cps = (total_chunks + slices - 1) / slices;
  slices = (total_chunks + cps - 1) / cps;
  dim3 grid(units, slices);
  k_wgrad3x3_tc2<PAIR><<<grid, kWgThreads, Cfg::kSmem, s>>>(tmA, tmZ, dw, W, Cin, Cout, total_chunks, cps);
  L3_CHECK_LAUNCH();
  return 0;
}

// ---- raw input staging for the first-layer kernels ----------------------------------------------------------
// Cin = 1 / 3 rows are 2 / 6 bytes per pixel: no TMA box fits, and gathering the taps with per-thread global loads
// put one DRAM latency on the critical path of every tile (measured: the builder warps sat on their first use of the
// loaded values).  Instead one thread stages, per tile of NPX consecutive flattened pixels, the three contiguous runs
// of (NPX + 2) * C0 input values (filter rows ky = 0..2) into a shared-memory ring with 1-D bulk copies
// (cp.async.bulk, 16-byte aligned windows around each run), several tiles ahead; the builders then read their taps
// from shared memory.  Runs that stick out of the buffer are clipped -- only halo rows ever see the unloaded bytes.
template <int C0, int NPX>
struct RawCfg {
  static const int kLen = 2 * C0 * (NPX + 2);           // bytes of one filter row's run
  static const int kSeg = (kLen + 16 + 15) / 16 * 16;   // segment stride in the stage (room for the alignment slack)
  static const int kStage = 3 * kSeg;
};
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// one thread: the three bulk copies of the tile starting at flattened pixel m0; total16 = input bytes rounded up to 16
template <int C0, int NPX>
__device__ __forceinline__ void raw_issue(const uint8_t* __restrict__ xin, long long m0, int Wp, long long total16,
                                          uint8_t* stage, uint64_t* bar) {
  using R = RawCfg<C0, NPX>;
  long long src[3];
  int nb[3], shift[3];
  uint32_t tx = 0;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const long long start = 2LL * C0 * (m0 + (ky - 1) * Wp - 1);
    const long long a = start & ~15LL;                       // floor to 16 (also for negative starts)
    const long long e = (start + R::kLen + 15) & ~15LL;
    const long long ac = a < 0 ? 0 : a, ec = e > total16 ? total16 : e;
    src[ky] = ac;
    nb[ky] = ec > ac ? (int)(ec - ac) : 0;
    shift[ky] = (int)(ac - a);
    tx += (uint32_t)nb[ky];
  }
  mbar_expect_tx(bar, tx);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
    if (nb[ky] > 0) bulk_g2s(stage + ky * R::kSeg + shift[ky], xin + src[ky], (uint32_t)nb[ky], bar);
}
// byte offset, inside a stage, of tap column 0 of filter row ky for the tile's first pixel
template <int C0, int NPX>
__device__ __forceinline__ int raw_row_off(long long m0, int Wp, int ky) {
  const long long start = 2LL * C0 * (m0 + (ky - 1) * Wp - 1);
  return ky * RawCfg<C0, NPX>::kSeg + (int)(start & 15);
}

// ---- first-layer weight gradient on tensor cores ------------------------------------------------------------
// Cin = 1 / 3 is far too narrow for a TMA box, so the im2col operand is BUILT in shared memory by 8 warps: per
// 64-pixel chunk a [64 px][64] bf16 MN-major tile whose columns are the 9*C0 window taps, the 9 "inside" indicators
// (weight gradient of an all-ones plane, see k_bn0_from_dw) and a constant 1 (bias gradient); the rest stays zero.
// The tile is written in the 128-byte-swizzle pattern UMMA expects (16-byte chunk index XOR pixel row & 7), made
// visible to the async proxy with fence.proxy.async, and multiplied with the TMA-loaded dz chunk:
//   D[col][co] += sum_px tile[px][col] * dz[px][co]        (M = 128 with the upper half aliasing the lower, N = 64)
// so dW (27 or 9 rows), d1 (9 rows) and db (1 row) come out of ONE accumulator.  Replaces a 1.9 ms SIMT reduction.
static const int kFwStages = 4;
static const int kFwRawStages = 8;   // staged input runs (RawCfg<C0, 64>), a few hundred bytes each
static const int kFwThreads = 64 + 256;
static const int kFwTile = 64 * 128;   // bytes: A tile and dz tile

// one 16-byte column chunk Q (columns 8Q .. 8Q+7) of a pixel row as raw bf16 bits; Q is a template parameter so tap /
// channel / address offsets of every column fold to constants.  r0/r1/r2 point at the leftmost tap of the three filter
// rows in the staged input runs (runs are clipped at the buffer ends: halo / out-of-range rows may read stale but
// finite bytes -- their dz row is zero).
template <int C0, int Q>
__device__ __forceinline__ uint4 fw_chunk(const unsigned short* __restrict__ r0, const unsigned short* __restrict__ r1,
                                          const unsigned short* __restrict__ r2, int yp, int xp, int H, int W) {
  constexpr int K = 9 * C0;
  uint32_t pk[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint32_t b2[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = Q * 8 + 2 * e + u;   // compile-time after unrolling
      uint32_t bits = 0u;
      if (i < K) {
        const int tap = i / C0, c = i - tap * C0, ky = tap / 3, kx = tap % 3;
        const unsigned short* rp = ky == 0 ? r0 : (ky == 1 ? r1 : r2);
        bits = (uint32_t)rp[kx * C0 + c];
      } else if (i < K + 9) {
        const int tap = i - K;
        const int yy = yp + tap / 3 - 1, xx = xp + tap % 3 - 1;
        bits = (yy >= 1 && yy <= H && xx >= 1 && xx <= W) ? 0x3F80u : 0u;   // bf16 1.0
      } else if (i == K + 9) {
        bits = 0x3F80u;
      }
      b2[u] = bits;
    }
    pk[e] = b2[0] | (b2[1] << 16);
  }
  return make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

template <int C0>
__global__ void __launch_bounds__(kFwThreads, 1)
k_first_wgrad_tc(const __grid_constant__ CUtensorMap tmZ, const bf16* __restrict__ xin, float* __restrict__ dw,
                 float* __restrict__ db, float* __restrict__ d1, int H, int W, long long Mp, int total_chunks,
                 int chunks_per_cta) {
  constexpr int K = 9 * C0, NQ = (K + 10 + 7) / 8;   // 16-byte column chunks in use: 5 (C0=3) / 3 (C0=1)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  using Raw = RawCfg<C0, 64>;
  uint8_t* smem_in = smem + kFwStages * 2 * kFwTile;   // ring of staged input runs
  __shared__ __align__(8) uint64_t full_bar[kFwStages], empty_bar[kFwStages], tfull_bar, r_full[kFwRawStages],
      r_empty[kFwRawStages];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero every A tile once: the unused columns must read as zeros for the whole kernel; and the input ring, so that
  // bytes a clipped run leaves unwritten are finite
  for (int i = threadIdx.x; i < kFwStages * kFwTile / 16; i += blockDim.x) {
    const int st = i / (kFwTile / 16), o = i % (kFwTile / 16);
    reinterpret_cast<uint4*>(smem + st * 2 * kFwTile)[o] = make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < kFwRawStages * Raw::kStage / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_in)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kFwStages; ++i) { mbar_init(&full_bar[i], 9); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < kFwRawStages; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 8); }
    mbar_init(&tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "n"(64)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int c_begin = blockIdx.x * chunks_per_cta;
  const int c_end = min(total_chunks, c_begin + chunks_per_cta);

  if (warp == 0) {
    if (lane == 0) {
      // producer: dz tiles (2-D TMA) and, kFwRawStages chunks ahead, the input runs the builders gather from
      const uint8_t* xb = reinterpret_cast<const uint8_t*>(xin);
      const long long total16 = (Mp * C0 * 2 + 15) & ~15LL;
      const int Wp = W + 2;
      int stage = 0, rs = 0;
      uint32_t phase = 0, rph = 0;
      for (int c = c_begin; c < c_end && c < c_begin + kFwRawStages; ++c)   // fresh ring: no wait
        raw_issue<C0, 64>(xb, (long long)c * 64, Wp, total16, smem_in + (c - c_begin) * Raw::kStage, &r_full[c - c_begin]);
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], kFwTile);
        tma_load_2d(&tmZ, &full_bar[stage], smem + stage * 2 * kFwTile + kFwTile, 0, c * 64);
        if (++stage == kFwStages) { stage = 0; phase ^= 1; }
        if (c + kFwRawStages < c_end) {   // the slot of chunk c is re-used by chunk c + kFwRawStages
          mbar_wait(&r_empty[rs], rph);
          raw_issue<C0, 64>(xb, (long long)(c + kFwRawStages) * 64, Wp, total16, smem_in + rs * Raw::kStage, &r_full[rs]);
        }
        if (++rs == kFwRawStages) { rs = 0; rph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);
    constexpr uint32_t hi = desc_hi(1024);
    const bool leader = elect_one();
    const uint32_t s_base = smem_addr(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int c = c_begin; c < c_end; ++c) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (leader) {
        const uint32_t a_lo0 = desc_lo(s_base + stage * 2 * kFwTile, 0);             // LBO 0: rows 64..127 alias 0..63
        const uint32_t b_lo0 = desc_lo(s_base + stage * 2 * kFwTile + kFwTile, kFwTile);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16_lh(tmem_base, a_lo0 + ks * 128, hi, b_lo0 + ks * 128, hi, idesc, (c > c_begin || ks > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == kFwStages) { stage = 0; phase ^= 1; }
    }
    if (leader) umma_commit(&tfull_bar);
    __syncwarp();
  } else {
    // ===== builders: 256 threads = 64 pixel rows x 4 parts; part p writes column chunk p (and p + 4) of the im2col
    // tile, gathering its taps from the staged input runs (shared memory: no global latency on this path) =====
    const int bt = threadIdx.x - 64;
    const int row = bt & 63, part = bt >> 6;
    const int Wp = W + 2;
    const unsigned HWp = (unsigned)(H + 2) * (unsigned)Wp;
    int stage = 0, rs = 0;
    uint32_t phase = 0, rph = 0;
    for (int c = c_begin; c < c_end; ++c) {
      const long long m0 = (long long)c * 64;
      const long long m = m0 + row;
      int yp = 0, xp = 0;
      if (m < Mp) {
        const unsigned b = (unsigned)m / HWp;
        const unsigned r = (unsigned)m - b * HWp;
        yp = (int)(r / (unsigned)Wp);
        xp = (int)(r - (unsigned)yp * (unsigned)Wp);
      }
      mbar_wait(&r_full[rs], rph);
      const uint8_t* in_stage = smem_in + rs * Raw::kStage;
      const unsigned short* rp[3];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
        rp[ky] = reinterpret_cast<const unsigned short*>(in_stage + raw_row_off<C0, 64>(m0, Wp, ky)) + row * C0;
      uint4 o0 = make_uint4(0, 0, 0, 0), o1 = make_uint4(0, 0, 0, 0);
      // `part` is warp-uniform (two builder warps per part): the switch does not diverge
      switch (part) {
        case 0:
          o0 = fw_chunk<C0, 0>(rp[0], rp[1], rp[2], yp, xp, H, W);
          if (NQ > 4) o1 = fw_chunk<C0, 4>(rp[0], rp[1], rp[2], yp, xp, H, W);
          break;
        case 1:
          o0 = fw_chunk<C0, 1>(rp[0], rp[1], rp[2], yp, xp, H, W);
          break;
        case 2:
          o0 = fw_chunk<C0, 2>(rp[0], rp[1], rp[2], yp, xp, H, W);
          break;
        default:
          if (NQ > 3) o0 = fw_chunk<C0, 3>(rp[0], rp[1], rp[2], yp, xp, H, W);
          break;
      }
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* tile = smem + stage * 2 * kFwTile;
      if (part < NQ) *reinterpret_cast<uint4*>(tile + row * 128 + ((part ^ (row & 7)) << 4)) = o0;
      if (NQ > 4 && part == 0) *reinterpret_cast<uint4*>(tile + row * 128 + ((4 ^ (row & 7)) << 4)) = o1;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to UMMA
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full_bar[stage]);
        mbar_arrive(&r_empty[rs]);   // the taps were consumed by the stores above
      }
      if (++stage == kFwStages) { stage = 0; phase ^= 1; }
      if (++rs == kFwRawStages) { rs = 0; rph ^= 1; }
    }
    // ===== epilogue (accumulator rows 0..K+9 live in TMEM lanes 0..63 -> quarters 0 and 1) =====
    const int q = warp & 3;
    if (warp < 6 && q < 2 && c_end > c_begin) {
      const int r = q * 32 + lane;
      mbar_wait(&tfull_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        float* dst = nullptr;
        if (r < K) dst = dw + r * 64;
        else if (r < K + 9) dst = d1 ? d1 + (r - K) * 64 : nullptr;
        else if (r == K + 9) dst = db;
        if (dst && r == K + 9) {   // the bias-gradient vector sits at an arbitrary (4-byte aligned) arena offset
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dst + c0 + i, __uint_as_float(v[i]));
        } else if (dst) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            red_add_v4(dst + c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                       __uint_as_float(v[i + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64) : "memory");
  }
}

int launch_first_wgrad_tc(const bf16* xin, const bf16* dz, float* dw, float* db, float* d1, int B, int H, int W, int C0,
                          int Cout, cudaStream_t s) {
  L3_REQUIRE(Cout == 64 && (C0 == 1 || C0 == 3), "first_wgrad_tc: C0=%d Cout=%d", C0, Cout);
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Mp + 1024 < 0x7fffffffLL && Mp >= 3, "first_wgrad_tc: pixel count out of range");
  if (d1) L3_CHECK_CUDA(cudaMemsetAsync(d1, 0, sizeof(float) * 9 * 64, s));
  L3_REQUIRE(((uintptr_t)xin & 15) == 0, "first_wgrad_tc: input must be 16-byte aligned (bulk copies)");
  const int smem = kFwStages * 2 * kFwTile + kFwRawStages * RawCfg<3, 64>::kStage + 1024;
  static PerDeviceOnce once;
  if (once.needed()) {
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_first_wgrad_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_first_wgrad_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    once.mark();
  }
  CUtensorMap tmZ;
  if (make_tmap(&tmZ, dz, 64, Mp, 64)) return -1;
  const int total_chunks = (int)((Mp + 63) / 64);
  int ctas = 148 * 2;
  if (ctas > total_chunks) ctas = total_chunks;
  const int cpc = (total_chunks + ctas - 1) / ctas;
  ctas = (total_chunks + cpc - 1) / cpc;
  if (C0 == 1) k_first_wgrad_tc<1><<<ctas, kFwThreads, smem, s>>>(tmZ, xin, dw, db, d1, H, W, Mp, total_chunks, cpc);
  else k_first_wgrad_tc<3><<<ctas, kFwThreads, smem, s>>>(tmZ, xin, dw, db, d1, H, W, Mp, total_chunks, cpc);
  L3_CHECK_LAUNCH();
  return 0;
}

// ---- first-layer forward on tensor cores -------------------------------------------------------------------
// Cin = 1 / 3: K = 9*C0 = 9 / 27 is too narrow for a TMA box, and the SIMT kernel it replaces ran at a third of the
// fp32 pipe (0.45 ms for a layer whose output write takes 0.06 ms of HBM time).  Four builder warps gather the
// im2col rows of a 128-pixel tile (flattened padded index, like every other conv here) straight from the padded
// bf16 input -- per pixel three contiguous runs of 3*C0 values, prefetched into registers one tile ahead -- and
// write them as a K-major, 128-byte-swizzled UMMA operand ([128 px][KPAD], KPAD = 16 / 32); the weight tile
// ([64 co][KPAD], built once per CTA from the fp32 HWIO kernel) is the B operand.  One elected thread issues KPAD/16
// MMAs (128x64x16) per tile into one of four TMEM accumulators; eight epilogue warps add the fp32 bias, store bf16
// and carry the per-lane BN statistics in registers (one transposing butterfly per kernel).  HBM-bound by design:
// algorithmic bytes per pixel = 2*C0 (input) + 128 (output).
template <int C0>
struct FcCfg {
  static const int K = 9 * C0;
  static const int KPAD = (K + 15) / 16 * 16;   // 16 (C0 = 1) / 32 (C0 = 3)
  static const int NQ = KPAD / 8;               // 16-byte chunks per operand row
  static const int kStages = 4;
  static const int kATile = kBM * 128;          // 128 rows x 128 B (only the first KPAD*2 bytes of a row are used)
  static const int kWTile = 64 * 128;
  static const int kEpiWarps = 16;              // two groups of 8: group g takes the CTA's tiles with local index = g mod 2
  static const int kStoreScratch = kEpiWarps * 32 * 64;   // per-epilogue-warp store transposition scratch (see k_conv3x3_tc3)
  static const int kRawStages = 8;
  static const int kRawStage = RawCfg<C0, kBM>::kStage;
  static const int kSmem = kStages * kATile + kWTile + kStoreScratch + kRawStages * kRawStage + 1024;
  static const int kAccs = 4;                   // TMEM accumulators of 64 columns
};
static const int kFcThreads = 32 + 128 + 512 + 32;   // MMA warp, 4 builder warps, 16 epilogue warps, raw-copy warp

template <int C0>
__global__ void __launch_bounds__(kFcThreads, 1)
k_first_conv_tc(const bf16* __restrict__ xin, const float* __restrict__ w, const float* __restrict__ bias,
                bf16* __restrict__ out, int H, int W, long long Mp, int num_tiles, double* __restrict__ stats,
                FastDiv dHWp, FastDiv dWp, const float* __restrict__ act_scale, const float* __restrict__ act_shift,
                int out_padded) {
  using Cfg = FcCfg<C0>;
  constexpr int K = Cfg::K, KPAD = Cfg::KPAD, NQ = Cfg::NQ, ST = Cfg::kStages, ACCS = Cfg::kAccs, RS = Cfg::kRawStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem + ST * Cfg::kATile;
  uint8_t* smem_scr = smem_w + Cfg::kWTile;
  uint8_t* smem_in = smem_scr + Cfg::kStoreScratch;
  __shared__ __align__(8) uint64_t a_full[ST], a_empty[ST], t_full[ACCS], t_empty[ACCS], r_full[RS], r_empty[RS];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[64];
  __shared__ __align__(16) float s_act[128];   // inference: BN scale | shift applied (with ReLU) in the epilogue

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // weight tile: element (co, k) at co*128 + ((k/8) ^ (co & 7))*16 + (k % 8)*2 ; zero beyond K
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int co = i & 63, k = i >> 6;
    const float v = k < K ? __ldg(w + k * 64 + co) : 0.f;
    *reinterpret_cast<bf16*>(smem_w + co * 128 + (((k >> 3) ^ (co & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
  for (int i = threadIdx.x; i < RS * Cfg::kRawStage / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_in)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 64) s_bias[threadIdx.x] = bias ? __ldg(bias + threadIdx.x) : 0.f;
  if (act_scale != nullptr && threadIdx.x < 64) {
    s_act[threadIdx.x] = __ldg(act_scale + threadIdx.x);
    s_act[64 + threadIdx.x] = __ldg(act_shift + threadIdx.x);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < ACCS; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
    for (int i = 0; i < RS; ++i) { mbar_init(&r_full[i], 1); mbar_init(&r_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)),
                 "n"(ACCS * 64)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weight tile / zeroed ring (generic stores) -> async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int Wp = W + 2;

  if (warp == 5 + Cfg::kEpiWarps) {
    // ===== raw-input producer: bulk copies run up to RS tiles ahead of the builders =====
    if (lane == 0) {
      const long long total16 = (Mp * C0 * 2 + 15) & ~15LL;
      int rs = 0;
      uint32_t rph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&r_empty[rs], rph ^ 1);
        raw_issue<C0, kBM>(reinterpret_cast<const uint8_t*>(xin), (long long)tile * kBM, Wp, total16,
                           smem_in + rs * Cfg::kRawStage, &r_full[rs]);
        if (++rs == RS) { rs = 0; rph ^= 1; }
      }
    }
  } else if (warp == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(kBM, 64, 0, 0);
    constexpr uint32_t hi = desc_hi(1024);
    const bool leader = elect_one();
    const uint32_t sa_base = smem_addr(smem);
    const uint32_t b_lo0 = desc_lo(smem_addr(smem_w), 16);
    int stage = 0, acc = 0;
    uint32_t sph = 0, aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&t_empty[acc], aph ^ 1);
      mbar_wait(&a_full[stage], sph);
      tc_fence_after();
      if (leader) {
        const uint32_t a_lo0 = desc_lo(sa_base + stage * Cfg::kATile, 16);
#pragma unroll
        for (int k = 0; k < KPAD / 16; ++k)
          umma_bf16_lh(tmem_base + acc * 64, a_lo0 + (uint32_t)(k * 2), hi, b_lo0 + (uint32_t)(k * 2), hi, idesc, k ? 1u : 0u);
        umma_commit(&a_empty[stage]);
        umma_commit(&t_full[acc]);
      }
      __syncwarp();
      if (++stage == ST) { stage = 0; sph ^= 1; }
      if (++acc == ACCS) { acc = 0; aph ^= 1; }
    }
  } else if (warp < 5) {
    // ===== builders: thread = pixel row of the tile; taps come from the staged input runs =====
    const int r = threadIdx.x - 32;
    int stage = 0, rs = 0;
    uint32_t sph = 0, rph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long m0 = (long long)tile * kBM;
      mbar_wait(&r_full[rs], rph);
      const uint8_t* in_stage = smem_in + rs * Cfg::kRawStage;
      unsigned short v[K];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const unsigned short* src =
            reinterpret_cast<const unsigned short*>(in_stage + raw_row_off<C0, kBM>(m0, Wp, ky)) + r * C0;
#pragma unroll
        for (int j = 0; j < 3 * C0; ++j) v[ky * 3 * C0 + j] = src[j];
      }
      mbar_wait(&a_empty[stage], sph ^ 1);
      uint8_t* row = smem + stage * Cfg::kATile + r * 128;
#pragma unroll
      for (int qd = 0; qd < NQ; ++qd) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k0 = qd * 8 + 2 * e, k1 = k0 + 1;   // compile-time
          const uint32_t lo = k0 < K ? (uint32_t)v[k0 < K ? k0 : 0] : 0u;
          const uint32_t hi16 = k1 < K ? (uint32_t)v[k1 < K ? k1 : 0] : 0u;
          pk[e] = lo | (hi16 << 16);
        }
        *reinterpret_cast<uint4*>(row + ((qd ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to UMMA
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_full[stage]);
        mbar_arrive(&r_empty[rs]);   // the taps were consumed by the stores above
      }
      if (++stage == ST) { stage = 0; sph ^= 1; }
      if (++rs == RS) { rs = 0; rph ^= 1; }
    }
  } else {
    // ===== epilogue: warp = (tile group, TMEM lane quarter, 32-column half).  One warp's work on a tile is a long
    // dependent chain (wait, TMEM load, pack, transpose, store, statistics: ~1900 cycles measured), so the tiles of
    // a CTA alternate between two independent groups of 8 warps; each accumulator (tile index mod 4) belongs to one
    // group.  Statistics are column sums read back from the store scratch -- no per-lane accumulators, which keeps
    // the 704-thread CTA inside the register file =====
    const int ew = warp - 5;
    const int grp = ew >> 3;
    const int q = warp & 3;
    const int half = (ew >> 2) & 1;
    const int c0 = half * 32;
    uint4* const st_scr = reinterpret_cast<uint4*>(smem_scr) + ew * 128;
    float s1 = 0.f, s2 = 0.f;   // this lane's column (c0 + lane): sum and sum of squares over the warp's tiles
    int it = grp;               // local tile index
    for (int tile = blockIdx.x + grp * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int acc = it & (ACCS - 1);
      const uint32_t aph = (uint32_t)(it / ACCS) & 1u;
      const int opix = out_pixel((uint32_t)tile * kBM + q * 32 + lane, Mp, H, W, dHWp, dWp);
      // inference with out_padded: the row lands at its own padded index = the next layer's zero-haloed input
      const int dpix = (out_padded && opix >= 0) ? (int)((uint32_t)tile * kBM + q * 32 + lane) : opix;
      int spix[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) spix[i] = __shfl_sync(0xffffffffu, dpix, (lane >> 2) + 8 * i);
      mbar_wait(&t_full[acc], aph);
      tc_fence_after();
      uint32_t v[32];
      __syncwarp();
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64 + c0), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);   // accumulator is in registers: hand it back before the math
      uint32_t pk[16];
      if (act_scale != nullptr) {
        // inference: relu(scale * (acc + bias) + shift) on the un-rounded accumulator (the first layer is Conv -> BN -> ReLU)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c0 + 4 * j]);
          const float4 sc = *reinterpret_cast<const float4*>(&s_act[c0 + 4 * j]);
          const float4 sf = *reinterpret_cast<const float4*>(&s_act[64 + c0 + 4 * j]);
          pk[2 * j] = pack_bf16x2(fmaxf(fmaf(__uint_as_float(v[4 * j]) + bb.x, sc.x, sf.x), 0.f),
                                  fmaxf(fmaf(__uint_as_float(v[4 * j + 1]) + bb.y, sc.y, sf.y), 0.f));
          pk[2 * j + 1] = pack_bf16x2(fmaxf(fmaf(__uint_as_float(v[4 * j + 2]) + bb.z, sc.z, sf.z), 0.f),
                                      fmaxf(fmaf(__uint_as_float(v[4 * j + 3]) + bb.w, sc.w, sf.w), 0.f));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c0 + 4 * j]);
          pk[2 * j] = pack_bf16x2(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
          pk[2 * j + 1] = pack_bf16x2(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
        }
      }
      if (opix < 0) {   // halo / out-of-range rows: not stored, and zero in the statistics
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = 0u;
      }
      // coalesced store through the warp's swizzled scratch: one instruction writes 8 rows x 64 contiguous bytes
      uint4* const srow = st_scr + lane * 4;
      const int sw4 = (lane >> 1) & 3;
#pragma unroll
      for (int c = 0; c < 4; ++c) srow[c ^ sw4] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int R = (lane >> 2) + 8 * i;
        const uint4 val = st_scr[R * 4 + ((lane & 3) ^ ((R >> 1) & 3))];
        if (spix[i] >= 0) *reinterpret_cast<uint4*>(out + (long long)spix[i] * 64 + c0 + (lane & 3) * 8) = val;
      }
      if (stats != nullptr) {
        // column `lane` of the 32 x 32 chunk: element (R, lane) sits in 16-byte piece (lane >> 3) ^ ((R >> 1) & 3)
        const unsigned short* sc16 = reinterpret_cast<const unsigned short*>(st_scr);
#pragma unroll
        for (int R = 0; R < 32; ++R) {
          const float x = __uint_as_float((uint32_t)sc16[R * 32 + ((((lane >> 3) ^ ((R >> 1) & 3)) << 3) | (lane & 7))] << 16);
          s1 += x;
          s2 = fmaf(x, x, s2);
        }
      }
      __syncwarp();   // the scratch is rewritten by the next tile
    }
    if (stats != nullptr) {
      atomicAdd(&stats[c0 + lane], (double)s1);
      atomicAdd(&stats[64 + c0 + lane], (double)s2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ACCS * 64) : "memory");
  }
}

int launch_first_conv_tc(const bf16* xin, const float* w, const float* bias, bf16* out, int B, int H, int W, int C0,
                         int Cout, double* stats, cudaStream_t s, const float* act_scale, const float* act_shift,
                         int out_padded) {
  L3_REQUIRE(act_scale == nullptr || (act_shift != nullptr && stats == nullptr), "first_conv_tc: fused activation excludes statistics");
  L3_REQUIRE(!out_padded || act_scale != nullptr, "first_conv_tc: the padded store belongs to the fused-activation mode");
  L3_REQUIRE(Cout == 64 && (C0 == 1 || C0 == 3), "first_conv_tc: C0=%d Cout=%d", C0, Cout);
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Mp + 1024 < 0x7fffffffLL && Mp >= 3, "first_conv_tc: pixel count out of range");
  L3_REQUIRE(((uintptr_t)xin & 15) == 0, "first_conv_tc: input must be 16-byte aligned (bulk copies)");
  static PerDeviceOnce once;
  if (once.needed()) {
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_first_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FcCfg<1>::kSmem));
    L3_CHECK_CUDA(cudaFuncSetAttribute(k_first_conv_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FcCfg<3>::kSmem));
    once.mark();
  }
  if (stats) L3_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * Cout, s));
  const int num_tiles = (int)((Mp + kBM - 1) / kBM);
  const int grid = num_tiles < 148 ? num_tiles : 148;
  const FastDiv dHWp = make_fastdiv((uint32_t)(H + 2) * (uint32_t)(W + 2)), dWp = make_fastdiv((uint32_t)(W + 2));
  if (C0 == 1)
    k_first_conv_tc<1><<<grid, kFcThreads, FcCfg<1>::kSmem, s>>>(xin, w, bias, out, H, W, Mp, num_tiles, stats, dHWp, dWp, act_scale,
                                                              act_shift, out_padded);
  else
    k_first_conv_tc<3><<<grid, kFcThreads, FcCfg<3>::kSmem, s>>>(xin, w, bias, out, H, W, Mp, num_tiles, stats, dHWp, dWp, act_scale,
                                                              act_shift, out_padded);
  L3_CHECK_LAUNCH();
  return 0;
}

// db[co] += sum over all padded rows of dz (halo rows are zero)
__global__ void k_bias_grad(const bf16* __restrict__ dz, long long rows, int C, float* __restrict__ db) {
  extern __shared__ float sh[];
  const int groups = C >> 3;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = blockDim.x / groups;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  float s1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = 0.f;
  for (long long r = (long long)blockIdx.x * lanes + lane; r < rows; r += (long long)gridDim.x * lanes) {
    float v[8];
    load8(dz + r * C + g * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(&sh[g * 8 + i], s1[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&db[i], sh[i]);
}

// dw[tap][ci][co] = sum over the four (operand part) x (operand part) blocks of the split weight gradient
// dw4 (3,3,2*Cin,2*Cout): (a_hi + a_lo)(dz_hi + dz_lo) -- overwrites dw
__global__ void k_fold_split_dw(const float* __restrict__ dw4, float* __restrict__ dw, int Cin, int Cout) {
  const long long n = 9LL * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long r = i / Cout;
    const int ci = (int)(r % Cin);
    const int tap = (int)(r / Cin);
    const float* b = dw4 + ((long long)tap * 2 * Cin + ci) * (2 * Cout) + co;
    const long long up = (long long)Cin * 2 * Cout;
    dw[i] = (b[0] + b[Cout]) + (b[up] + b[up + Cout]);
  }
}

// Weight gradient of the parity mode on tensor cores: a_split (B,H+2,W+2,3*Cin) and dz_split (B,H+2,W+2,3*Cout) hold bf16
// parts [hi | lo | hi]; the kernel runs on the [hi | lo] prefixes as a layer with 2*Cin x 2*Cout channels into the fp32
// scratch dw4 (9 * 2*Cin * 2*Cout floats, zeroed here), whose four blocks are then folded into dw (overwritten).
int launch_wgrad3x3_tc_split(const void* a_split, const void* dz_split, float* dw, float* dw4, int B, int H, int W, int Cin,
                             int Cout, cudaStream_t s) {
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "wgrad_tc_split: channels Cin=%d Cout=%d", Cin, Cout);
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Mp + 4LL * (W + 2) < 0x7fffffffLL, "wgrad_tc: too many pixels for 32-bit TMA coordinates");
  L3_CHECK_CUDA(cudaMemsetAsync(dw4, 0, sizeof(float) * 9 * 4 * (size_t)Cin * Cout, s));
  if (launch_wgrad2_cfg<false>((const bf16*)a_split, (const bf16*)dz_split, dw4, W, 2 * Cin, 2 * Cout, Mp, s, 3 * Cin, 3 * Cout))
    return -1;
  const long long n = 9LL * Cin * Cout;
  int blocks = (int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
  k_fold_split_dw<<<blocks, 256, 0, s>>>(dw4, dw, Cin, Cout);
  L3_CHECK_LAUNCH();
  return 0;
}

int launch_wgrad3x3_tc(const bf16* a, const bf16* dz, float* dw, float* db, int B, int H, int W, int Cin, int Cout,
                       cudaStream_t s) {
  L3_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0 && (Cin == 64 || Cin % 128 == 0) && (Cin == 64 || Cout % 128 == 0),
             "wgrad_tc: unsupported channels Cin=%d Cout=%d", Cin, Cout);
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Mp + 4LL * (W + 2) < 0x7fffffffLL, "wgrad_tc: too many pixels for 32-bit TMA coordinates");
  const int rc = (Cin == 64) ? launch_wgrad2_cfg<true>(a, dz, dw, W, Cin, Cout, Mp, s)
                             : launch_wgrad2_cfg<false>(a, dz, dw, W, Cin, Cout, Mp, s);
  if (rc) return rc;
  if (db) return launch_bias_grad_tc(dz, db, B, H, W, Cout, s);
  return 0;
}

// db[co] += sum of dz over all pixels (dz zero-haloed padded bf16); HBM-bound
int launch_bias_grad_tc(const bf16* dz, float* db, int B, int H, int W, int Cout, cudaStream_t s) {
  const long long Mp = (long long)B * (H + 2) * (W + 2);
  L3_REQUIRE(Cout % 8 == 0 && 256 % (Cout / 8) == 0, "bias_grad: Cout=%d", Cout);
  int lanes = 256 / (Cout / 8);
  long long want = (Mp + (long long)lanes * 16 - 1) / ((long long)lanes * 16);
  int blocks = (int)(want > 148 * 4 ? 148 * 4 : (want < 1 ? 1 : want));
  k_bias_grad<<<blocks, 256, Cout * sizeof(float), s>>>(dz, Mp, Cout, db);
  L3_CHECK_LAUNCH();
  return 0;
}

}  // namespace l3
