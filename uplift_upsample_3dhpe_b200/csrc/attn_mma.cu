// Temporal / strided multi-head attention of the TRAINING step on tensor cores, fp32 data straight from the tape.
//
// reference: common/net/vision_transformer.py:99-130 (MultiHeadAttention.call: scaled dot product, key mask as an
// additive -1e9 term, softmax, weighted sum) and its gradient (TensorFlow autodiff of the same lines).
//
// One CTA per (window, head); q | k | v (and dO) rows of the head sit in shared memory as fp32 with a row stride of
// DH + 4 floats (conflict-free for every fragment pattern below).  All products run on `mma.sync.m16n8k8` with TF32
// operands.  NSPLIT = 3 is the error-compensated form: x = hi + lo with hi = the top 10 mantissa bits, and
// a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi (the dropped lo lo term is 2^-22 relative), i.e. fp32-grade results from the
// tensor core.  NSPLIT = 1 rounds both operands to TF32 (cvt.rna) once.
// The default of the training step is the third form (namespace amb, "nsplit 2"): operands are split into bf16 hi + bf16 lo
// planes ONCE, while the rows are staged into shared memory (the two planes together take the 4 bytes per element the fp32
// rows took), and a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi runs on `mma.sync.m16n8k16` bf16: 16 mantissa bits per operand
// (2^-16 relative per product, 30x finer than TF32) with half the MMA instructions of the compensated TF32 form for the same
// products (ncu: 8.6 cycles per m16n8k16 bf16 MMA and scheduler, tensor pipe 46 % busy in the backward kernel) and without any
// split arithmetic in the inner loops.  Fragment loads are conflict-free 32-bit
// reads (row stride DH / 2 + 4 words) and `ldmatrix.trans` for the operands that are contracted over their ROW index.
//
// Forward: warp w owns queries 16w..16w+15.  S = Q K^T lands in the m16n8 accumulator layout (row g / g+8, columns
// 2t, 2t+1 of every 8-key tile); the softmax runs on those registers (quad shuffles for the row max / sum) and the
// probabilities are re-used AS the A operand of P V without any data movement: the MMA contracts over k, so feeding
// accumulator column 2t as logical k = t and column 2t+1 as k = t+4 only asks for the B fragment rows in the same
// order (V rows 8n+2t and 8n+2t+1).
// Backward, two passes over the same shared-memory tiles, no S x S tile ever written anywhere:
//   pass A (warp = 16 queries): S, P, dP = dO V^T, dot = rowsum(dP o P), dS = P o (dP - dot), dQ = scale dS K;
//           the row statistics (max, 1 / sum, dot) go to shared memory;
//   pass B (warp = 16 keys):    S^T = K Q^T recomputed, P^T from the saved statistics, dV = P^T dO, dP^T = V dO^T,
//           dS^T, dK = scale dS^T Q.
// Recomputing S^T costs 2 of the 9 GEMM passes and removes the two S x S fp32 tiles (66 KB of shared memory per CTA, the
// difference between one and three resident CTAs) and every cross-warp reduction.  Deterministic: no atomics.
#include <algorithm>

#include "common.cuh"
#include "train.cuh"

namespace uu {
namespace am {

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int NS>
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  if constexpr (NS == 1) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    lo = 0u;
  } else {
    hi = __float_as_uint(x) & 0xffffe000u;                 // exactly representable in TF32
    lo = __float_as_uint(x - __uint_as_float(hi));         // exact; the tensor core keeps its top 10 mantissa bits
  }
}
template <int NS>
__device__ __forceinline__ void mma_split(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], float b0f,
                                          float b1f) {
  uint32_t bh0, bl0, bh1, bl1;
  split<NS>(b0f, bh0, bl0);
  split<NS>(b1f, bh1, bl1);
  if constexpr (NS == 3) {
    mma_tf32(c, al, bh0, bh1);
    mma_tf32(c, ah, bl0, bl1);
  }
  mma_tf32(c, ah, bh0, bh1);
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// c[n] (+)= A B^T : A = 16 rows of `arows` (this warp's tile), B = rows 8n .. 8n+7 of `bm`, contraction over DH channels
template <int DH, int NT, int NS>
__device__ __forceinline__ void gemm_abt(float (&c)[NT][4], const float* __restrict__ arows, const float* __restrict__ bm,
                                         int g, int t) {
  constexpr int LD = DH + 4;
#pragma unroll
  for (int kk = 0; kk < DH / 8; ++kk) {
    uint32_t ah[4], al[4];
    split<NS>(arows[g * LD + 8 * kk + t], ah[0], al[0]);
    split<NS>(arows[(g + 8) * LD + 8 * kk + t], ah[1], al[1]);
    split<NS>(arows[g * LD + 8 * kk + t + 4], ah[2], al[2]);
    split<NS>(arows[(g + 8) * LD + 8 * kk + t + 4], ah[3], al[3]);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float* br = bm + (8 * n + g) * LD + 8 * kk + t;
      mma_split<NS>(c[n], ah, al, br[0], br[4]);
    }
  }
}
// c[nc] (+)= P B : P = this warp's 16 x (8 NT) tile in accumulator layout, B = rows of `bm` (one per contraction index)
template <int DH, int NT, int NS>
__device__ __forceinline__ void gemm_pb(float (&c)[DH / 8][4], const float (&p)[NT][4], const float* __restrict__ bm, int g,
                                        int t) {
  constexpr int LD = DH + 4;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    uint32_t ah[4], al[4];
    split<NS>(p[n][0], ah[0], al[0]);      // (row g,   k = t   <-> column 2t)
    split<NS>(p[n][2], ah[1], al[1]);      // (row g+8, k = t)
    split<NS>(p[n][1], ah[2], al[2]);      // (row g,   k = t+4 <-> column 2t+1)
    split<NS>(p[n][3], ah[3], al[3]);
    const float* br = bm + (8 * n + 2 * t) * LD + g;
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) mma_split<NS>(c[nc], ah, al, br[8 * nc], br[LD + 8 * nc]);
  }
}

// rows [0, S) of one head's slice -> shared memory, rows [S, SP) zero.  The trip count is a compile-time constant and ALL
// loads of the array are issued before the first store: with a rolled loop every iteration waited for its own global
// load (24 serialised DRAM latencies per CTA, which made the kernel latency-bound at three CTAs per SM).
template <int DH, int SP, int NTHR>
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, long long ld, int S, int tid) {
  constexpr int LD = DH + 4, TOT = SP * (DH / 4), IT = (TOT + NTHR - 1) / NTHR;
  float4 v[IT];
#pragma unroll
  for (int u = 0; u < IT; ++u) {
    const int i = tid + u * NTHR, j = i / (DH / 4), c = (i - j * (DH / 4)) * 4;
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < TOT && j < S) v[u] = *reinterpret_cast<const float4*>(src + (long long)j * ld + c);
  }
#pragma unroll
  for (int u = 0; u < IT; ++u) {
    const int i = tid + u * NTHR, j = i / (DH / 4), c = (i - j * (DH / 4)) * 4;
    if (i < TOT) *reinterpret_cast<float4*>(dst + j * LD + c) = v[u];
  }
}

__host__ __device__ constexpr int warps_for(int NT) { return (NT + 1) / 2; }

// scores of this warp's 16 queries -> probabilities (in place); returns the row statistics
template <int DH, int NT, int NS>
__device__ __forceinline__ void scores_softmax(float (&s)[NT][4], const float* __restrict__ Qw, const float* __restrict__ Ks,
                                               const float* __restrict__ Km, float scale, int g, int t, float& mx0, float& mx1,
                                               float& inv0, float& inv1) {
#pragma unroll
  for (int n = 0; n < NT; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
  gemm_abt<DH, NT, NS>(s, Qw, Ks, g, t);
  mx0 = -INFINITY; mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const float2 km = *reinterpret_cast<const float2*>(Km + 8 * n + 2 * t);
    s[n][0] = fmaf(s[n][0], scale, km.x); s[n][1] = fmaf(s[n][1], scale, km.y);
    s[n][2] = fmaf(s[n][2], scale, km.x); s[n][3] = fmaf(s[n][3], scale, km.y);
    mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
    mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
  }
  mx0 = quad_max(mx0); mx1 = quad_max(mx1);
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    s[n][0] = __expf(s[n][0] - mx0); s[n][1] = __expf(s[n][1] - mx0);
    s[n][2] = __expf(s[n][2] - mx1); s[n][3] = __expf(s[n][3] - mx1);
    l0 += s[n][0] + s[n][1];
    l1 += s[n][2] + s[n][3];
  }
  inv0 = 1.f / quad_sum(l0); inv1 = 1.f / quad_sum(l1);
#pragma unroll
  for (int n = 0; n < NT; ++n) { s[n][0] *= inv0; s[n][1] *= inv0; s[n][2] *= inv1; s[n][3] *= inv1; }
}

template <int DH, int NT, int NS>
__global__ void __launch_bounds__(32 * warps_for(NT)) k_attn_mma_fwd(const float* __restrict__ qkv, int S, int heads,
                                                                     const uint8_t* __restrict__ mask, int mask_stride,
                                                                     float* __restrict__ out) {
  constexpr int LD = DH + 4, NW = warps_for(NT), SP = 16 * NW;
  extern __shared__ __align__(16) float am_sm[];
  float* Qs = am_sm;
  float* Ks = Qs + SP * LD;
  float* Vs = Ks + SP * LD;
  float* Km = Vs + SP * LD;           // [SP] additive key term: 0 keep, -1e9 masked (vit:118), -inf beyond S
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int d = heads * DH;
  const long long row0 = (long long)b * S;
  const float* base = qkv + row0 * 3 * d + h * DH;
  stage_rows<DH, SP, 32 * NW>(Qs, base, 3 * d, S, tid);
  stage_rows<DH, SP, 32 * NW>(Ks, base + d, 3 * d, S, tid);
  stage_rows<DH, SP, 32 * NW>(Vs, base + 2 * d, 3 * d, S, tid);
  for (int j = tid; j < SP; j += 32 * NW)
    Km[j] = j >= S ? -INFINITY : ((mask && !mask[(long long)b * mask_stride + j]) ? -1e9f : 0.f);
  __syncthreads();
  const float scale = rsqrtf((float)DH);
  float s[NT][4];
  float mx0, mx1, inv0, inv1;
  scores_softmax<DH, NT, NS>(s, Qs + 16 * warp * LD, Ks, Km, scale, g, t, mx0, mx1, inv0, inv1);
  float o[DH / 8][4];
#pragma unroll
  for (int nc = 0; nc < DH / 8; ++nc) o[nc][0] = o[nc][1] = o[nc][2] = o[nc][3] = 0.f;
  gemm_pb<DH, NT, NS>(o, s, Vs, g, t);
  const int r0 = 16 * warp + g, r1 = r0 + 8;
#pragma unroll
  for (int nc = 0; nc < DH / 8; ++nc) {
    if (r0 < S) *reinterpret_cast<float2*>(out + (row0 + r0) * d + h * DH + 8 * nc + 2 * t) = make_float2(o[nc][0], o[nc][1]);
    if (r1 < S) *reinterpret_cast<float2*>(out + (row0 + r1) * d + h * DH + 8 * nc + 2 * t) = make_float2(o[nc][2], o[nc][3]);
  }
}

template <int DH, int NT, int NS>
__global__ void __launch_bounds__(32 * warps_for(NT)) k_attn_mma_bwd(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                                     int S, int heads, const uint8_t* __restrict__ mask,
                                                                     int mask_stride, float* __restrict__ dqkv) {
  constexpr int LD = DH + 4, NW = warps_for(NT), SP = 16 * NW, NTQ = 2 * NW;   // NTQ: 8-query tiles of pass B
  extern __shared__ __align__(16) float am_sm[];
  float* Qs = am_sm;
  float* Ks = Qs + SP * LD;
  float* Vs = Ks + SP * LD;
  float* Gs = Vs + SP * LD;           // dO
  float* Km = Gs + SP * LD;           // [SP]
  float* sM = Km + SP;                // [SP] row max
  float* sI = sM + SP;                // [SP] 1 / row sum
  float* sD = sI + SP;                // [SP] rowsum(dP o P)
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int d = heads * DH;
  const long long row0 = (long long)b * S;
  const float* base = qkv + row0 * 3 * d + h * DH;
  stage_rows<DH, SP, 32 * NW>(Qs, base, 3 * d, S, tid);
  stage_rows<DH, SP, 32 * NW>(Ks, base + d, 3 * d, S, tid);
  stage_rows<DH, SP, 32 * NW>(Vs, base + 2 * d, 3 * d, S, tid);
  stage_rows<DH, SP, 32 * NW>(Gs, dO + row0 * d + h * DH, d, S, tid);
  for (int j = tid; j < SP; j += 32 * NW)
    Km[j] = j >= S ? -INFINITY : ((mask && !mask[(long long)b * mask_stride + j]) ? -1e9f : 0.f);
  __syncthreads();
  const float scale = rsqrtf((float)DH);
  const int r0 = 16 * warp + g, r1 = r0 + 8;
  float* obase = dqkv + row0 * 3 * d + h * DH;
  // ---------------- pass A: this warp's 16 queries
  {
    float p[NT][4];
    float mx0, mx1, inv0, inv1;
    scores_softmax<DH, NT, NS>(p, Qs + 16 * warp * LD, Ks, Km, scale, g, t, mx0, mx1, inv0, inv1);
    float dp[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    gemm_abt<DH, NT, NS>(dp, Gs + 16 * warp * LD, Vs, g, t);
    float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      dot0 = fmaf(dp[n][0], p[n][0], fmaf(dp[n][1], p[n][1], dot0));
      dot1 = fmaf(dp[n][2], p[n][2], fmaf(dp[n][3], p[n][3], dot1));
    }
    dot0 = quad_sum(dot0); dot1 = quad_sum(dot1);
    if (t == 0) {
      sM[r0] = mx0; sI[r0] = inv0; sD[r0] = dot0;
      sM[r1] = mx1; sI[r1] = inv1; sD[r1] = dot1;
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      dp[n][0] = p[n][0] * (dp[n][0] - dot0); dp[n][1] = p[n][1] * (dp[n][1] - dot0);
      dp[n][2] = p[n][2] * (dp[n][2] - dot1); dp[n][3] = p[n][3] * (dp[n][3] - dot1);
    }
    float dq[DH / 8][4];
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) dq[nc][0] = dq[nc][1] = dq[nc][2] = dq[nc][3] = 0.f;
    gemm_pb<DH, NT, NS>(dq, dp, Ks, g, t);
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) {
      if (r0 < S) *reinterpret_cast<float2*>(obase + (long long)r0 * 3 * d + 8 * nc + 2 * t) = make_float2(dq[nc][0] * scale, dq[nc][1] * scale);
      if (r1 < S) *reinterpret_cast<float2*>(obase + (long long)r1 * 3 * d + 8 * nc + 2 * t) = make_float2(dq[nc][2] * scale, dq[nc][3] * scale);
    }
  }
  __syncthreads();
  // ---------------- pass B: this warp's 16 keys (rows r0, r1 are key indices now), all SP queries as columns
  {
    float pt[NTQ][4];
#pragma unroll
    for (int n = 0; n < NTQ; ++n) pt[n][0] = pt[n][1] = pt[n][2] = pt[n][3] = 0.f;
    gemm_abt<DH, NTQ, NS>(pt, Ks + 16 * warp * LD, Qs, g, t);           // S^T
    const float km0 = Km[r0], km1 = Km[r1];
#pragma unroll
    for (int n = 0; n < NTQ; ++n) {
      const float2 m = *reinterpret_cast<const float2*>(sM + 8 * n + 2 * t);
      const float2 il = *reinterpret_cast<const float2*>(sI + 8 * n + 2 * t);
      pt[n][0] = __expf(fmaf(pt[n][0], scale, km0) - m.x) * il.x;
      pt[n][1] = __expf(fmaf(pt[n][1], scale, km0) - m.y) * il.y;
      pt[n][2] = __expf(fmaf(pt[n][2], scale, km1) - m.x) * il.x;
      pt[n][3] = __expf(fmaf(pt[n][3], scale, km1) - m.y) * il.y;
    }
    {
      float dv[DH / 8][4];
#pragma unroll
      for (int nc = 0; nc < DH / 8; ++nc) dv[nc][0] = dv[nc][1] = dv[nc][2] = dv[nc][3] = 0.f;
      gemm_pb<DH, NTQ, NS>(dv, pt, Gs, g, t);                            // dV = P^T dO
#pragma unroll
      for (int nc = 0; nc < DH / 8; ++nc) {
        if (r0 < S) *reinterpret_cast<float2*>(obase + (long long)r0 * 3 * d + 2 * d + 8 * nc + 2 * t) = make_float2(dv[nc][0], dv[nc][1]);
        if (r1 < S) *reinterpret_cast<float2*>(obase + (long long)r1 * 3 * d + 2 * d + 8 * nc + 2 * t) = make_float2(dv[nc][2], dv[nc][3]);
      }
    }
    {
      float dpt[NTQ][4];
#pragma unroll
      for (int n = 0; n < NTQ; ++n) dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
      gemm_abt<DH, NTQ, NS>(dpt, Vs + 16 * warp * LD, Gs, g, t);         // dP^T = V dO^T
#pragma unroll
      for (int n = 0; n < NTQ; ++n) {
        const float2 dt = *reinterpret_cast<const float2*>(sD + 8 * n + 2 * t);
        pt[n][0] *= dpt[n][0] - dt.x; pt[n][1] *= dpt[n][1] - dt.y;
        pt[n][2] *= dpt[n][2] - dt.x; pt[n][3] *= dpt[n][3] - dt.y;
      }
    }
    float dk[DH / 8][4];
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) dk[nc][0] = dk[nc][1] = dk[nc][2] = dk[nc][3] = 0.f;
    gemm_pb<DH, NTQ, NS>(dk, pt, Qs, g, t);                              // dK = scale dS^T Q
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) {
      if (r0 < S) *reinterpret_cast<float2*>(obase + (long long)r0 * 3 * d + d + 8 * nc + 2 * t) = make_float2(dk[nc][0] * scale, dk[nc][1] * scale);
      if (r1 < S) *reinterpret_cast<float2*>(obase + (long long)r1 * 3 * d + d + 8 * nc + 2 * t) = make_float2(dk[nc][2] * scale, dk[nc][3] * scale);
    }
  }
}

template <int DH, int NT, int NS>
cudaError_t run_fwd(const float* qkv, long long B, int S, int heads, const uint8_t* mask, int mask_stride, float* out,
                    cudaStream_t st) {
  constexpr int NW = warps_for(NT), SP = 16 * NW;
  constexpr size_t smem = sizeof(float) * (3 * SP * (DH + 4) + SP);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_attn_mma_fwd<DH, NT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  k_attn_mma_fwd<DH, NT, NS><<<(unsigned)(B * heads), 32 * NW, smem, st>>>(qkv, S, heads, mask, mask_stride, out);
  return cudaGetLastError();
}
template <int DH, int NT, int NS>
cudaError_t run_bwd(const float* qkv, const float* dO, long long B, int S, int heads, const uint8_t* mask, int mask_stride,
                    float* dqkv, cudaStream_t st) {
  constexpr int NW = warps_for(NT), SP = 16 * NW;
  constexpr size_t smem = sizeof(float) * (4 * SP * (DH + 4) + 4 * SP);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_attn_mma_bwd<DH, NT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  k_attn_mma_bwd<DH, NT, NS><<<(unsigned)(B * heads), 32 * NW, smem, st>>>(qkv, dO, S, heads, mask, mask_stride, dqkv);
  return cudaGetLastError();
}

}  // namespace am


// =================================================================================================
// bf16 hi / lo planes, mma.sync.m16n8k16 (see the header comment)
// =================================================================================================
namespace amb {

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
  mma_bf16(c, al, bh0, bh1);
  mma_bf16(c, ah, bl0, bl1);
  mma_bf16(c, ah, bh0, bh1);
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(row)));
}
// (x0, x1) -> bf16x2 hi (x0 in the low half) and bf16x2 lo = the rounded remainders
__device__ __forceinline__ void pack_hi_lo(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float f0 = __uint_as_float(hi << 16), f1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - f0, x1 - f1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
template <int DH> struct Lay { static constexpr int LDW = DH / 2 + 4; };     // row stride of a plane in 32-bit words

// rows [0, S) of one head's slice -> hi / lo planes, rows [S, SP) zero (all loads in flight before the first store, see
// stage_rows above)
template <int DH, int SP, int NTHR>
__device__ __forceinline__ void stage_planes(uint32_t* __restrict__ hi, uint32_t* __restrict__ lo, const float* __restrict__ src,
                                             long long ld, int S, int tid) {
  constexpr int LDW = Lay<DH>::LDW, TOT = SP * (DH / 4), IT = (TOT + NTHR - 1) / NTHR;
  float4 v[IT];
#pragma unroll
  for (int u = 0; u < IT; ++u) {
    const int i = tid + u * NTHR, j = i / (DH / 4), c = (i - j * (DH / 4)) * 4;
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < TOT && j < S) v[u] = *reinterpret_cast<const float4*>(src + (long long)j * ld + c);
  }
#pragma unroll
  for (int u = 0; u < IT; ++u) {
    const int i = tid + u * NTHR, j = i / (DH / 4), c = (i - j * (DH / 4)) * 4;
    if (i < TOT) {
      uint2 h, l;
      pack_hi_lo(v[u].x, v[u].y, h.x, l.x);
      pack_hi_lo(v[u].z, v[u].w, h.y, l.y);
      *reinterpret_cast<uint2*>(hi + j * LDW + c / 2) = h;
      *reinterpret_cast<uint2*>(lo + j * LDW + c / 2) = l;
    }
  }
}
// c[n] (+)= A B^T : A = 16 rows starting at `ah` / `al` (this warp's tile of a plane pair), B = rows 8n .. 8n+7 of (bh, bl)
template <int DH, int NT>
__device__ __forceinline__ void gemm_abt(float (&c)[NT][4], const uint32_t* __restrict__ ah, const uint32_t* __restrict__ al,
                                         const uint32_t* __restrict__ bh, const uint32_t* __restrict__ bl, int g, int t) {
  constexpr int LDW = Lay<DH>::LDW;
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
    uint32_t fh[4], fl[4];
    const int o0 = g * LDW + 8 * kk + t, o1 = (g + 8) * LDW + 8 * kk + t;
    fh[0] = ah[o0]; fh[1] = ah[o1]; fh[2] = ah[o0 + 4]; fh[3] = ah[o1 + 4];
    fl[0] = al[o0]; fl[1] = al[o1]; fl[2] = al[o0 + 4]; fl[3] = al[o1 + 4];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int ob = (8 * n + g) * LDW + 8 * kk + t;
      mma3(c[n], fh, fl, bh[ob], bh[ob + 4], bl[ob], bl[ob + 4]);
    }
  }
}
// c[nc] (+)= P B : P = this warp's 16 x (8 NT) tile in accumulator layout (== the A-fragment layout of m16n8k16),
// B = rows of the plane pair (one per contraction index), read transposed by ldmatrix
template <int DH, int NT>
__device__ __forceinline__ void gemm_pb(float (&c)[DH / 8][4], const float (&p)[NT][4], const uint32_t* __restrict__ bh,
                                        const uint32_t* __restrict__ bl, int lane) {
  constexpr int LDW = Lay<DH>::LDW;
  static_assert(NT % 2 == 0 && (DH / 8) % 2 == 0, "tiles are consumed in pairs");
  // ldmatrix.x4: lane l supplies the row address of matrix l >> 3: (k half = (l >> 3) & 1, column tile = l >> 4), row l & 7
  const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), lcolw = 4 * (lane >> 4);
#pragma unroll
  for (int j = 0; j < NT / 2; ++j) {
    uint32_t fh[4], fl[4];
    pack_hi_lo(p[2 * j][0], p[2 * j][1], fh[0], fl[0]);
    pack_hi_lo(p[2 * j][2], p[2 * j][3], fh[1], fl[1]);
    pack_hi_lo(p[2 * j + 1][0], p[2 * j + 1][1], fh[2], fl[2]);
    pack_hi_lo(p[2 * j + 1][2], p[2 * j + 1][3], fh[3], fl[3]);
#pragma unroll
    for (int nc = 0; nc < DH / 8; nc += 2) {
      uint32_t rh[4], rl[4];
      const int off = (16 * j + lrow) * LDW + 4 * nc + lcolw;
      ldsm_x4_trans(rh, bh + off);
      ldsm_x4_trans(rl, bl + off);
      mma3(c[nc], fh, fl, rh[0], rh[1], rl[0], rl[1]);
      mma3(c[nc + 1], fh, fl, rh[2], rh[3], rl[2], rl[3]);
    }
  }
}

__host__ __device__ constexpr int warps_for(int NT) { return (NT + 1) / 2; }

template <int DH, int NT>
__device__ __forceinline__ void scores_softmax(float (&s)[NT][4], const uint32_t* qh, const uint32_t* ql, const uint32_t* kh,
                                               const uint32_t* kl, const float* __restrict__ Km, float scale, int g, int t,
                                               float& mx0, float& mx1, float& inv0, float& inv1) {
#pragma unroll
  for (int n = 0; n < NT; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
  gemm_abt<DH, NT>(s, qh, ql, kh, kl, g, t);
  mx0 = -INFINITY; mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const float2 km = *reinterpret_cast<const float2*>(Km + 8 * n + 2 * t);
    s[n][0] = fmaf(s[n][0], scale, km.x); s[n][1] = fmaf(s[n][1], scale, km.y);
    s[n][2] = fmaf(s[n][2], scale, km.x); s[n][3] = fmaf(s[n][3], scale, km.y);
    mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
    mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
  }
  mx0 = quad_max(mx0); mx1 = quad_max(mx1);
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    s[n][0] = __expf(s[n][0] - mx0); s[n][1] = __expf(s[n][1] - mx0);
    s[n][2] = __expf(s[n][2] - mx1); s[n][3] = __expf(s[n][3] - mx1);
    l0 += s[n][0] + s[n][1];
    l1 += s[n][2] + s[n][3];
  }
  inv0 = 1.f / quad_sum(l0); inv1 = 1.f / quad_sum(l1);
#pragma unroll
  for (int n = 0; n < NT; ++n) { s[n][0] *= inv0; s[n][1] *= inv0; s[n][2] *= inv1; s[n][3] *= inv1; }
}

template <int DH, int NT>
__global__ void __launch_bounds__(32 * warps_for(NT)) k_attn_bf2_fwd(const float* __restrict__ qkv, int S, int heads,
                                                                     const uint8_t* __restrict__ mask, int mask_stride,
                                                                     float* __restrict__ out) {
  constexpr int LDW = Lay<DH>::LDW, NW = warps_for(NT), SP = 16 * NW, PL = SP * LDW;
  extern __shared__ __align__(16) uint32_t amb_sm[];
  uint32_t *Qh = amb_sm, *Ql = Qh + PL, *Kh = Ql + PL, *Kl = Kh + PL, *Vh = Kl + PL, *Vl = Vh + PL;
  float* Km = reinterpret_cast<float*>(Vl + PL);
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int d = heads * DH;
  const long long row0 = (long long)b * S;
  const float* base = qkv + row0 * 3 * d + h * DH;
  stage_planes<DH, SP, 32 * NW>(Qh, Ql, base, 3 * d, S, tid);
  stage_planes<DH, SP, 32 * NW>(Kh, Kl, base + d, 3 * d, S, tid);
  stage_planes<DH, SP, 32 * NW>(Vh, Vl, base + 2 * d, 3 * d, S, tid);
  for (int j = tid; j < SP; j += 32 * NW)
    Km[j] = j >= S ? -INFINITY : ((mask && !mask[(long long)b * mask_stride + j]) ? -1e9f : 0.f);
  __syncthreads();
  const float scale = rsqrtf((float)DH);
  float s[NT][4];
  float mx0, mx1, inv0, inv1;
  scores_softmax<DH, NT>(s, Qh + 16 * warp * LDW, Ql + 16 * warp * LDW, Kh, Kl, Km, scale, g, t, mx0, mx1, inv0, inv1);
  float o[DH / 8][4];
#pragma unroll
  for (int nc = 0; nc < DH / 8; ++nc) o[nc][0] = o[nc][1] = o[nc][2] = o[nc][3] = 0.f;
  gemm_pb<DH, NT>(o, s, Vh, Vl, lane);
  const int r0 = 16 * warp + g, r1 = r0 + 8;
#pragma unroll
  for (int nc = 0; nc < DH / 8; ++nc) {
    if (r0 < S) *reinterpret_cast<float2*>(out + (row0 + r0) * d + h * DH + 8 * nc + 2 * t) = make_float2(o[nc][0], o[nc][1]);
    if (r1 < S) *reinterpret_cast<float2*>(out + (row0 + r1) * d + h * DH + 8 * nc + 2 * t) = make_float2(o[nc][2], o[nc][3]);
  }
}

template <int DH, int NT>
__global__ void __launch_bounds__(32 * warps_for(NT)) k_attn_bf2_bwd(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                                     int S, int heads, const uint8_t* __restrict__ mask,
                                                                     int mask_stride, float* __restrict__ dqkv) {
  constexpr int LDW = Lay<DH>::LDW, NW = warps_for(NT), SP = 16 * NW, NTQ = 2 * NW, PL = SP * LDW;
  extern __shared__ __align__(16) uint32_t amb_sm[];
  uint32_t *Qh = amb_sm, *Ql = Qh + PL, *Kh = Ql + PL, *Kl = Kh + PL, *Vh = Kl + PL, *Vl = Vh + PL, *Gh = Vl + PL, *Gl = Gh + PL;
  float* Km = reinterpret_cast<float*>(Gl + PL);
  float* sM = Km + SP;
  float* sI = sM + SP;
  float* sD = sI + SP;
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int d = heads * DH;
  const long long row0 = (long long)b * S;
  const float* base = qkv + row0 * 3 * d + h * DH;
  stage_planes<DH, SP, 32 * NW>(Qh, Ql, base, 3 * d, S, tid);
  stage_planes<DH, SP, 32 * NW>(Kh, Kl, base + d, 3 * d, S, tid);
  stage_planes<DH, SP, 32 * NW>(Vh, Vl, base + 2 * d, 3 * d, S, tid);
  stage_planes<DH, SP, 32 * NW>(Gh, Gl, dO + row0 * d + h * DH, d, S, tid);
  for (int j = tid; j < SP; j += 32 * NW)
    Km[j] = j >= S ? -INFINITY : ((mask && !mask[(long long)b * mask_stride + j]) ? -1e9f : 0.f);
  __syncthreads();
  const float scale = rsqrtf((float)DH);
  const int r0 = 16 * warp + g, r1 = r0 + 8, wo = 16 * warp * LDW;
  float* obase = dqkv + row0 * 3 * d + h * DH;
  // ---------------- pass A: this warp's 16 queries
  {
    float p[NT][4];
    float mx0, mx1, inv0, inv1;
    scores_softmax<DH, NT>(p, Qh + wo, Ql + wo, Kh, Kl, Km, scale, g, t, mx0, mx1, inv0, inv1);
    float dp[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    gemm_abt<DH, NT>(dp, Gh + wo, Gl + wo, Vh, Vl, g, t);
    float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      dot0 = fmaf(dp[n][0], p[n][0], fmaf(dp[n][1], p[n][1], dot0));
      dot1 = fmaf(dp[n][2], p[n][2], fmaf(dp[n][3], p[n][3], dot1));
    }
    dot0 = quad_sum(dot0); dot1 = quad_sum(dot1);
    if (t == 0) {
      sM[r0] = mx0; sI[r0] = inv0; sD[r0] = dot0;
      sM[r1] = mx1; sI[r1] = inv1; sD[r1] = dot1;
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      dp[n][0] = p[n][0] * (dp[n][0] - dot0); dp[n][1] = p[n][1] * (dp[n][1] - dot0);
      dp[n][2] = p[n][2] * (dp[n][2] - dot1); dp[n][3] = p[n][3] * (dp[n][3] - dot1);
    }
    float dq[DH / 8][4];
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) dq[nc][0] = dq[nc][1] = dq[nc][2] = dq[nc][3] = 0.f;
    gemm_pb<DH, NT>(dq, dp, Kh, Kl, lane);
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) {
      if (r0 < S) *reinterpret_cast<float2*>(obase + (long long)r0 * 3 * d + 8 * nc + 2 * t) = make_float2(dq[nc][0] * scale, dq[nc][1] * scale);
      if (r1 < S) *reinterpret_cast<float2*>(obase + (long long)r1 * 3 * d + 8 * nc + 2 * t) = make_float2(dq[nc][2] * scale, dq[nc][3] * scale);
    }
  }
  __syncthreads();
  // ---------------- pass B: this warp's 16 keys (rows r0, r1 are key indices now), all SP queries as columns
  {
    float pt[NTQ][4];
#pragma unroll
    for (int n = 0; n < NTQ; ++n) pt[n][0] = pt[n][1] = pt[n][2] = pt[n][3] = 0.f;
    gemm_abt<DH, NTQ>(pt, Kh + wo, Kl + wo, Qh, Ql, g, t);             // S^T
    const float km0 = Km[r0], km1 = Km[r1];
#pragma unroll
    for (int n = 0; n < NTQ; ++n) {
      const float2 m = *reinterpret_cast<const float2*>(sM + 8 * n + 2 * t);
      const float2 il = *reinterpret_cast<const float2*>(sI + 8 * n + 2 * t);
      pt[n][0] = __expf(fmaf(pt[n][0], scale, km0) - m.x) * il.x;
      pt[n][1] = __expf(fmaf(pt[n][1], scale, km0) - m.y) * il.y;
      pt[n][2] = __expf(fmaf(pt[n][2], scale, km1) - m.x) * il.x;
      pt[n][3] = __expf(fmaf(pt[n][3], scale, km1) - m.y) * il.y;
    }
    {
      float dv[DH / 8][4];
#pragma unroll
      for (int nc = 0; nc < DH / 8; ++nc) dv[nc][0] = dv[nc][1] = dv[nc][2] = dv[nc][3] = 0.f;
      gemm_pb<DH, NTQ>(dv, pt, Gh, Gl, lane);                          // dV = P^T dO
#pragma unroll
      for (int nc = 0; nc < DH / 8; ++nc) {
        if (r0 < S) *reinterpret_cast<float2*>(obase + (long long)r0 * 3 * d + 2 * d + 8 * nc + 2 * t) = make_float2(dv[nc][0], dv[nc][1]);
        if (r1 < S) *reinterpret_cast<float2*>(obase + (long long)r1 * 3 * d + 2 * d + 8 * nc + 2 * t) = make_float2(dv[nc][2], dv[nc][3]);
      }
    }
    {
      float dpt[NTQ][4];
#pragma unroll
      for (int n = 0; n < NTQ; ++n) dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
      gemm_abt<DH, NTQ>(dpt, Vh + wo, Vl + wo, Gh, Gl, g, t);          // dP^T = V dO^T
#pragma unroll
      for (int n = 0; n < NTQ; ++n) {
        const float2 dt = *reinterpret_cast<const float2*>(sD + 8 * n + 2 * t);
        pt[n][0] *= dpt[n][0] - dt.x; pt[n][1] *= dpt[n][1] - dt.y;
        pt[n][2] *= dpt[n][2] - dt.x; pt[n][3] *= dpt[n][3] - dt.y;
      }
    }
    float dk[DH / 8][4];
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) dk[nc][0] = dk[nc][1] = dk[nc][2] = dk[nc][3] = 0.f;
    gemm_pb<DH, NTQ>(dk, pt, Qh, Ql, lane);                            // dK = scale dS^T Q
#pragma unroll
    for (int nc = 0; nc < DH / 8; ++nc) {
      if (r0 < S) *reinterpret_cast<float2*>(obase + (long long)r0 * 3 * d + d + 8 * nc + 2 * t) = make_float2(dk[nc][0] * scale, dk[nc][1] * scale);
      if (r1 < S) *reinterpret_cast<float2*>(obase + (long long)r1 * 3 * d + d + 8 * nc + 2 * t) = make_float2(dk[nc][2] * scale, dk[nc][3] * scale);
    }
  }
}

template <int DH, int NT>
cudaError_t run_fwd(const float* qkv, long long B, int S, int heads, const uint8_t* mask, int mask_stride, float* out,
                    cudaStream_t st) {
  constexpr int NW = warps_for(NT), SP = 16 * NW;
  constexpr size_t smem = sizeof(uint32_t) * 6 * SP * Lay<DH>::LDW + sizeof(float) * SP;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_attn_bf2_fwd<DH, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  k_attn_bf2_fwd<DH, NT><<<(unsigned)(B * heads), 32 * NW, smem, st>>>(qkv, S, heads, mask, mask_stride, out);
  return cudaGetLastError();
}
template <int DH, int NT>
cudaError_t run_bwd(const float* qkv, const float* dO, long long B, int S, int heads, const uint8_t* mask, int mask_stride,
                    float* dqkv, cudaStream_t st) {
  constexpr int NW = warps_for(NT), SP = 16 * NW;
  constexpr size_t smem = sizeof(uint32_t) * 8 * SP * Lay<DH>::LDW + sizeof(float) * 4 * SP;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_attn_bf2_bwd<DH, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  k_attn_bf2_bwd<DH, NT><<<(unsigned)(B * heads), 32 * NW, smem, st>>>(qkv, dO, S, heads, mask, mask_stride, dqkv);
  return cudaGetLastError();
}

}  // namespace amb

// S <= 80 keys, head dimension 32 / 48 / 64; B * heads CTAs must fit a 1-D grid
bool attention_mma_ok(long long B, int S, int heads, int dh) {
  return S >= 1 && S <= 80 && (dh == 32 || dh == 48 || dh == 64) && B * heads < (1ll << 31);
}

#define UU_AM_NT(DHV, NSV, CALL)                                   \
  if (S <= 24) return am::CALL<DHV, 3, NSV>;                       \
  if (S <= 48) return am::CALL<DHV, 6, NSV>;                       \
  return am::CALL<DHV, 10, NSV>;
#define UU_AM_DISPATCH(CALL, ARGS)                                                                  \
  auto pick = [&]() -> decltype(&am::CALL<48, 10, 3>) {                                             \
    if (nsplit == 1) {                                                                              \
      if (dh == 32) { UU_AM_NT(32, 1, CALL) }                                                       \
      if (dh == 48) { UU_AM_NT(48, 1, CALL) }                                                       \
      UU_AM_NT(64, 1, CALL)                                                                         \
    }                                                                                               \
    if (dh == 32) { UU_AM_NT(32, 3, CALL) }                                                         \
    if (dh == 48) { UU_AM_NT(48, 3, CALL) }                                                         \
    UU_AM_NT(64, 3, CALL)                                                                           \
  };                                                                                                \
  return pick() ARGS;

#define UU_AMB_DISPATCH(CALL, ARGS)                                                                 \
  auto pickb = [&]() -> decltype(&amb::CALL<48, 10>) {                                              \
    if (dh == 32) return S <= 32 ? amb::CALL<32, 4> : S <= 48 ? amb::CALL<32, 6> : amb::CALL<32, 10>; \
    if (dh == 48) return S <= 32 ? amb::CALL<48, 4> : S <= 48 ? amb::CALL<48, 6> : amb::CALL<48, 10>; \
    return S <= 32 ? amb::CALL<64, 4> : S <= 48 ? amb::CALL<64, 6> : amb::CALL<64, 10>;              \
  };                                                                                                \
  return pickb() ARGS;

cudaError_t launch_attention_mma_fwd(const float* qkv, long long B, int S, int heads, int dh, const uint8_t* mask,
                                     int mask_stride, float* out, int nsplit, cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  if (!attention_mma_ok(B, S, heads, dh) || nsplit < 1 || nsplit > 3) return cudaErrorInvalidValue;
  if (nsplit == 2) { UU_AMB_DISPATCH(run_fwd, (qkv, B, S, heads, mask, mask_stride, out, st)) }
  UU_AM_DISPATCH(run_fwd, (qkv, B, S, heads, mask, mask_stride, out, st))
}
cudaError_t launch_attention_mma_bwd(const float* qkv, const float* dO, long long B, int S, int heads, int dh,
                                     const uint8_t* mask, int mask_stride, float* dqkv, int nsplit, cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  if (!attention_mma_ok(B, S, heads, dh) || nsplit < 1 || nsplit > 3) return cudaErrorInvalidValue;
  if (nsplit == 2) { UU_AMB_DISPATCH(run_bwd, (qkv, dO, B, S, heads, mask, mask_stride, dqkv, st)) }
  UU_AM_DISPATCH(run_bwd, (qkv, dO, B, S, heads, mask, mask_stride, dqkv, st))
}
#undef UU_AM_DISPATCH
#undef UU_AMB_DISPATCH
#undef UU_AM_NT

}  // namespace uu
