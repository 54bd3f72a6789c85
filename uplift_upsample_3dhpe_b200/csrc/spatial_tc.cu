// K2 (bf16 path): fused spatial transformer, everything on tensor cores.
//
// reference: net:313-333 (spatial_transformation), vit:176-195 (TransformerBlock), vit:99-156 (MHA).
// Shapes are tiny (17 joint tokens x 32 channels per frame, 8 heads of dim 4), so the kernel is organised
// around keeping a frame inside ONE warp:
//   * a persistent CTA of 16 warps (4 per scheduler, 128 registers each) owns a group of 15 frames.  Warp w < 15
//     owns joints 0..15 of frame w as one m16 row tile; warp 15 owns joint 16 of all 15 frames.  The residual
//     stream lives in registers (m16n8 accumulator layout) from the key-point embedding to the final LayerNorm;
//   * linears: mma.sync.m16n8k16 (bf16 x bf16 -> fp32) against weights held in shared memory, pre-swizzled at
//     weight-commit time into per-lane B-fragment order; LayerNorm / GELU are applied on accumulator fragments;
//   * attention (17 x 17, head_dim 4) never leaves the warp: Q K^T per head is two m16n8k8 MMAs whose A operand
//     is the q accumulator masked to the head's 4 channels and whose B operand is the k accumulator as-is
//     (C layout == col-major B layout); P V is one m16n8k16 MMA against V^T obtained with movmatrix.trans.  The V
//     projection emits, per head, the columns [v0,1,v1,1,v2,1,v3,1], so the same MMA also yields the softmax
//     row sums.  The 17th key is an extra MMA column / k-step fed from warp 16 through shared memory, the 17th
//     query is evaluated in transposed form (keys as the M dimension) so that all 32 lanes hold live scores;
//   * the softmax scale and log2(e) are folded into W_q, so scores come out of the tensor core in the exp2 domain;
//     LayerNorm gamma/beta of norm1/norm2 are folded into the q|k|v and fc1 weights and biases at pack time;
//   * P, V and the GELU output are fp16 (11-bit mantissa, values are O(1)); GELU runs in packed half2 arithmetic;
//     P V and fc2 are f16 MMAs; q, k and every other operand stay bf16 (range).  (ex2/tanh.approx.f16x2 still cost
//     one MUFU per element on sm_100a, so the softmax exponentials stay fp32.)
// Cross-warp traffic is two point-to-point mbarrier hand-offs per layer (joint-16 q/k/v out, joint-16 attention
// rows back); frame warps never wait for each other.
// HBM traffic: 136 B of key-points in, 1088 B (17x32 bf16) out per frame.
#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"

namespace uu {

namespace st {
constexpr int J = 17, D = 32, HID = 64, HEADS = 8, DEPTH_MAX = 4;
constexpr int FRAMES = 15;                  // frames per group
constexpr int WARPS = 16, THREADS = WARPS * 32;
// B-fragment image of one block: [frag][lane] uint2
constexpr int F_Q = 0, F_K = 8, F_V = 16, F_PROJ = 32, F_FC1 = 40, F_FC2 = 56, F_TOTAL = 72;
// fp32 parameter image of one block
constexpr int P_LN1G = 0, P_LN1B = 32, P_BQKV = 64, P_BP = 192, P_LN2G = 224, P_LN2B = 256, P_B1 = 288, P_B2 = 352,
              P_TOTAL = 384;
// global fp32 parameters
constexpr int G_EK = 0, G_EB = 64, G_PE = 96, G_NG = G_PE + J * D, G_NB = G_NG + 32, G_TOTAL = G_NB + 32;
constexpr float QSCALE = 0.72134752044448170368f;     // (1 / sqrt(4)) * log2(e)
constexpr int STG = 40;                     // bf16 row stride of the per-warp output staging tile (80 B)
// joint-16 exchange area (bytes)
constexpr int X_Q = 0, X_K = 1024, X_V = 2048, X_S = 4096, X_O = 4608, X_BAR = 5632, X_TOTAL = 5648;

// physical channel read by logical k index `kap` (0..15) of k-step kk in the output projection: the attention
// output of head h, dim dd sits in lane t == dd, and heads 4kk..4kk+3 fill the A-fragment slots 2t, 2t+1, 2t+8, 2t+9
__host__ __device__ inline int proj_channel(int kk, int kap) {
  const int hi = kap >= 8, k2 = kap & 7;
  return 4 * (4 * kk + 2 * hi + (k2 & 1)) + (k2 >> 1);
}
// position of (head h, dim dd) in a row laid out in that A-fragment order
__host__ __device__ inline int proj_position(int h, int dd) {
  const int kk = h >> 2, i = h & 3;
  return 16 * kk + ((i < 2) ? 2 * dd + i : 8 + 2 * dd + (i - 2));
}
}  // namespace st

// ---- weight image (built once per uu_set_weight round) -------------------------------------------
// frags: [depth][72][32] uint2 ; params: [depth][384] + [704] floats
__global__ void k_spatial_pack(const float* const* __restrict__ blocks, int depth, const float* __restrict__ embed_k,
                               const float* __restrict__ embed_b, const float* __restrict__ pe,
                               const float* __restrict__ norm_g, const float* __restrict__ norm_b,
                               uint2* __restrict__ frags, float* __restrict__ params) {
  using namespace st;
  const int total_frag = depth * F_TOTAL * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_frag; i += gridDim.x * blockDim.x) {
    const int lane = i & 31, fr = (i >> 5) % F_TOTAL, l = (i >> 5) / F_TOTAL;
    const int g = lane >> 2, t = lane & 3;
    const float* const* tb = blocks + l * 16;
    float w[4];                              // B[k0][n], B[k0+1][n], B[k0+8][n], B[k0+9][n]
    if (fr < F_PROJ) {                       // q | k | v : K = 32 (two k-steps)
      const int f = fr < F_K ? fr - F_Q : (fr < F_V ? fr - F_K : fr - F_V);
      const int j = f >> 1, kk = f & 1;
      for (int e = 0; e < 4; ++e) {           // norm1 gamma folded in: (gamma * xhat) W == xhat (diag(gamma) W)
        const int k = 16 * kk + 2 * t + (e & 1) + 8 * (e >> 1);
        if (fr < F_K) w[e] = tb[0][k] * tb[2][k * 32 + 8 * j + g] * QSCALE;
        else if (fr < F_V) w[e] = tb[0][k] * tb[4][k * 32 + 8 * j + g];
        else w[e] = (g & 1) ? 0.f : tb[0][k] * tb[6][k * 32 + 4 * j + (g >> 1)];       // head j: [v0,1,v1,1,v2,1,v3,1]
      }
    } else if (fr < F_FC1) {                 // projection, permuted k order
      const int f = fr - F_PROJ, j = f >> 1, kk = f & 1;
      for (int e = 0; e < 4; ++e) w[e] = tb[8][proj_channel(kk, 2 * t + (e & 1) + 8 * (e >> 1)) * 32 + 8 * j + g];
    } else if (fr < F_FC2) {
      const int f = fr - F_FC1, j = f >> 1, kk = f & 1;
      for (int e = 0; e < 4; ++e) {
        const int k = 16 * kk + 2 * t + (e & 1) + 8 * (e >> 1);
        w[e] = tb[10][k] * tb[12][k * 64 + 8 * j + g];                      // norm2 gamma folded in
      }
    } else {
      const int f = fr - F_FC2, j = f >> 2, kk = f & 3;
      for (int e = 0; e < 4; ++e) w[e] = tb[14][(16 * kk + 2 * t + (e & 1) + 8 * (e >> 1)) * 32 + 8 * j + g];
    }
    uint2 o;
    if (fr >= F_FC2) {                       // fc2 runs as an f16 MMA (its A operand is the f16 GELU output)
      __half2 b0 = __floats2half2_rn(w[0], w[1]), b1 = __floats2half2_rn(w[2], w[3]);
      o.x = *reinterpret_cast<uint32_t*>(&b0);
      o.y = *reinterpret_cast<uint32_t*>(&b1);
    } else {
      __nv_bfloat162 b0 = __floats2bfloat162_rn(w[0], w[1]);
      __nv_bfloat162 b1 = __floats2bfloat162_rn(w[2], w[3]);
      o.x = *reinterpret_cast<uint32_t*>(&b0);
      o.y = *reinterpret_cast<uint32_t*>(&b1);
    }
    frags[i] = o;
  }
  const int total_p = depth * P_TOTAL + G_TOTAL;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_p; i += gridDim.x * blockDim.x) {
    float v;
    if (i < depth * P_TOTAL) {
      const int l = i / P_TOTAL, o = i % P_TOTAL;
      const float* const* tb = blocks + l * 16;
      if (o < P_LN1B) v = tb[0][o];
      else if (o < P_BQKV) v = tb[1][o - P_LN1B];
      else if (o < P_BP) {                   // bias image of the 16 q|k|v n-tiles
        // (norm1 beta folded in: b' = b + beta W)
        const int c = o - P_BQKV, tile = c >> 3, col = c & 7;
        if (tile < 8) {
          const int which = tile >> 2, n = 8 * (tile & 3) + col;
          v = tb[3 + 2 * which][n];
          for (int k = 0; k < 32; ++k) v += tb[1][k] * tb[2 + 2 * which][k * 32 + n];
          if (which == 0) v *= QSCALE;
        } else if (col & 1) {
          v = 1.f;
        } else {
          const int n = 4 * (tile - 8) + (col >> 1);
          v = tb[7][n];
          for (int k = 0; k < 32; ++k) v += tb[1][k] * tb[6][k * 32 + n];
        }
      }
      else if (o < P_LN2G) v = tb[9][o - P_BP];
      else if (o < P_LN2B) v = tb[10][o - P_LN2G];
      else if (o < P_B1) v = tb[11][o - P_LN2B];
      else if (o < P_B2) {                   // norm2 beta folded in
        const int n = o - P_B1;
        v = tb[13][n];
        for (int k = 0; k < 32; ++k) v += tb[11][k] * tb[12][k * 64 + n];
      }
      else v = tb[15][o - P_B2];
    } else {
      const int o = i - depth * P_TOTAL;
      if (o < G_EB) v = embed_k[o];
      else if (o < G_PE) v = embed_b[o - G_EB];
      else if (o < G_NG) v = pe[o - G_PE];
      else if (o < G_NB) v = norm_g[o - G_NG];
      else v = norm_b[o - G_NB];
    }
    params[i] = v;
  }
}

size_t spatial_tc_frag_bytes(int depth) { return sizeof(uint2) * depth * st::F_TOTAL * 32; }
size_t spatial_tc_param_bytes(int depth) { return sizeof(float) * (depth * st::P_TOTAL + st::G_TOTAL); }

cudaError_t launch_spatial_pack(const float* const* blocks, int depth, const float* embed_k, const float* embed_b,
                                const float* pe, const float* norm_g, const float* norm_b, void* frags, float* params,
                                cudaStream_t s) {
  k_spatial_pack<<<64, 256, 0, s>>>(blocks, depth, embed_k, embed_b, pe, norm_g, norm_b, (uint2*)frags, params);
  return cudaGetLastError();
}

// ---- device helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mma16816h(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688h(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ uint32_t pack2h(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// D = A B with a zero accumulator: separate output registers and literal-zero C inputs let ptxas feed RZ instead of
// materialising four zeros per MMA chain (about a hundred MOVs per layer in the head loop).
__device__ __forceinline__ void mma1688_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
}
__device__ __forceinline__ void mma16816h_z(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
// 8x8 b16 transpose inside the warp: in (row g; cols 2t,2t+1) -> out (row g; cols 2t,2t+1) of the transpose
__device__ __forceinline__ uint32_t movm_t(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
// point-to-point hand-off between the joint-16 warp and the frame warps
__device__ __forceinline__ void sp_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void sp_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {     // a protocol bug must trap, not hang the GPU
      printf("uu3d: spatial mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
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
// reductions over the 8 row groups g (lanes with equal t)
__device__ __forceinline__ float col_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
  return v;
}
__device__ __forceinline__ float col_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf(x / sqrt 2) ~ tanh(x (a + x^2 (b + c x^2))): max |error| 3e-5 over
// all x (fitted against the exact erf form; the clamp keeps the odd polynomial monotone for |x| > 9, where
// tanh is already +-1).  That is ~60x below the bf16 rounding of the value as the fc2 operand, and costs one
// MUFU + 7 FP ops instead of ~20 for an erf evaluation (4352 GELUs per frame).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float x2 = fminf(x * x, 81.0f);
  const float u = x * fmaf(x2, fmaf(x2, -3.58732362e-4f, 3.70503451e-2f), 7.97458471e-1f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// The same GELU on two values in packed fp16 arithmetic (one MUFU, 7 half2 ops per pair): the result feeds the f16
// fc2 MMA directly.  |x| >= 9 saturates through the clamp exactly as above; fp16 keeps 11 mantissa bits.
__device__ __forceinline__ uint32_t gelu_h2(float lo, float hi) {
  const __half2 x = __floats2half2_rn(lo, hi);
  const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(81.0f));
  const __half2 pl = __hfma2(x2, __hfma2(x2, __float2half2_rn(-3.58732362e-4f), __float2half2_rn(3.70503451e-2f)),
                             __float2half2_rn(7.97458471e-1f));
  const __half2 u = __hmul2(x, pl);
  uint32_t tu;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  const __half2 r = __hfma2(hx, *reinterpret_cast<const __half2*>(&tu), hx);
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// LayerNorm of the two rows a thread holds pieces of (x[j][0..1] row g, x[j][2..3] row g+8), Keras form;
// the result is packed straight into the A fragments of the next K=32 MMA.
__device__ __forceinline__ void ln_to_afrag(const float (&x)[4][4], const float* __restrict__ gam,
                                            const float* __restrict__ bet, float eps, int t, uint32_t (&a)[2][4]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3]; }
  const float m0 = quad_sum(s0) * (1.f / 32), m1 = quad_sum(s1) * (1.f / 32);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float d;
    d = x[j][0] - m0; q0 = fmaf(d, d, q0); d = x[j][1] - m0; q0 = fmaf(d, d, q0);
    d = x[j][2] - m1; q1 = fmaf(d, d, q1); d = x[j][3] - m1; q1 = fmaf(d, d, q1);
  }
  const float r0 = rsqrtf(quad_sum(q0) * (1.f / 32) + eps), r1 = rsqrtf(quad_sum(q1) * (1.f / 32) + eps);
  float y[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 g = *reinterpret_cast<const float2*>(gam + 8 * j + 2 * t);
    const float2 b = *reinterpret_cast<const float2*>(bet + 8 * j + 2 * t);
    y[j][0] = x[j][0] * (g.x * r0) + (b.x - m0 * (g.x * r0));
    y[j][1] = x[j][1] * (g.y * r0) + (b.y - m0 * (g.y * r0));
    y[j][2] = x[j][2] * (g.x * r1) + (b.x - m1 * (g.x * r1));
    y[j][3] = x[j][3] * (g.y * r1) + (b.y - m1 * (g.y * r1));
  }
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    a[kk][0] = pack2(y[2 * kk][0], y[2 * kk][1]);
    a[kk][1] = pack2(y[2 * kk][2], y[2 * kk][3]);
    a[kk][2] = pack2(y[2 * kk + 1][0], y[2 * kk + 1][1]);
    a[kk][3] = pack2(y[2 * kk + 1][2], y[2 * kk + 1][3]);
  }
}

// The same without gamma / beta (folded into the following linear layer at pack time): xhat = (x - mean) * rstd.
__device__ __forceinline__ void lnhat_to_afrag(const float (&x)[4][4], float eps, uint32_t (&a)[2][4]) {
  // single pass: sum and sum of squares together (32 values of O(1) in fp32: the cancellation in E[x^2] - mean^2 costs
  // ~1e-7 * mean^2 / var relative, far below the bf16 rounding of the result), then xhat = x * rstd - mean * rstd
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3];
    q0 = fmaf(x[j][0], x[j][0], fmaf(x[j][1], x[j][1], q0));
    q1 = fmaf(x[j][2], x[j][2], fmaf(x[j][3], x[j][3], q1));
  }
  const float m0 = quad_sum(s0) * (1.f / 32), m1 = quad_sum(s1) * (1.f / 32);
  const float v0 = fmaxf(fmaf(quad_sum(q0), 1.f / 32, -m0 * m0), 0.f), v1 = fmaxf(fmaf(quad_sum(q1), 1.f / 32, -m1 * m1), 0.f);
  const float r0 = rsqrtf(v0 + eps), r1 = rsqrtf(v1 + eps);
  const float n0 = -m0 * r0, n1 = -m1 * r1;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    a[kk][0] = pack2(fmaf(x[2 * kk][0], r0, n0), fmaf(x[2 * kk][1], r0, n0));
    a[kk][1] = pack2(fmaf(x[2 * kk][2], r1, n1), fmaf(x[2 * kk][3], r1, n1));
    a[kk][2] = pack2(fmaf(x[2 * kk + 1][0], r0, n0), fmaf(x[2 * kk + 1][1], r0, n0));
    a[kk][3] = pack2(fmaf(x[2 * kk + 1][2], r1, n1), fmaf(x[2 * kk + 1][3], r1, n1));
  }
}

struct SpatialTcParams {
  const float* x2d;      // (B*n_tok, 17, 2)
  const int* list;       // gather list or null
  const int* count;      // device count or null
  const int* src;        // optional token id -> row of x2d (video frame, -1 = zeros): fused sliding-window gather
  const int* flip;       // optional flip augmentation: joint j reads source joint flip[j] with x negated (eval.py:154-159)
  const int* range_lo;   // optional device pointers: process list positions [*range_lo, *range_hi) only
  const int* range_hi;   //   (chunked launches that overlap the host-to-device copy of the next chunk)
  int lo, hi;            // the same as host values when the pointers are null (hi < 0: up to the valid count)
  int max_frames;
  int depth;
  const uint2* frags;    // weight image
  const float* params;
  bf16* out;             // (n_valid, 544) compact
};

__global__ void __launch_bounds__(st::THREADS, 1) k_spatial_tc(SpatialTcParams p) {
  using namespace st;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint2* s_frag = reinterpret_cast<uint2*>(smem_raw);                              // depth*72*32 uint2
  float* s_par = reinterpret_cast<float*>(smem_raw + sizeof(uint2) * p.depth * F_TOTAL * 32);
  uint8_t* s_x = reinterpret_cast<uint8_t*>(s_par + ((p.depth * P_TOTAL + G_TOTAL + 3) & ~3));
  bf16* xq16 = reinterpret_cast<bf16*>(s_x + X_Q);      // [16 rows: 15 frames + 1 unused][32] joint-16 q (exp2 domain)
  bf16* xk16 = reinterpret_cast<bf16*>(s_x + X_K);      // [16][32]
  bf16* xv16 = reinterpret_cast<bf16*>(s_x + X_V);      // [16][8 heads][v0,1,v1,1,v2,1,v3,1]
  float* xs16 = reinterpret_cast<float*>(s_x + X_S);    // [16][8] score of (query 16, key 16)
  bf16* xo16 = reinterpret_cast<bf16*>(s_x + X_O);      // [16][32] joint-16 attention rows, projection A order
  uint64_t* bar_pub = reinterpret_cast<uint64_t*>(s_x + X_BAR);   // joint-16 warp -> frame warps (1 arrival)
  uint64_t* bar_ret = bar_pub + 1;                                // frame warps -> joint-16 warp (FRAMES arrivals)
  bf16* s_stage = reinterpret_cast<bf16*>(s_x + X_TOTAL);   // [16 warps][16 rows][STG]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const bool w16 = warp == WARPS - 1;

  // weights -> shared memory, once per CTA
  {
    const int nf = p.depth * F_TOTAL * 32;
    for (int i = tid; i < nf; i += THREADS) s_frag[i] = p.frags[i];
    const int np = p.depth * P_TOTAL + G_TOTAL;
    for (int i = tid; i < np; i += THREADS) s_par[i] = p.params[i];
    if (tid == 0) {
      sp_mbar_init(bar_pub, 1);
      sp_mbar_init(bar_ret, FRAMES);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  uint32_t phase = 0;                               // one hand-off pair per (group, layer)
  const float* gp = s_par + p.depth * P_TOTAL;      // global params

  const int lo = p.range_lo ? *p.range_lo : p.lo;
  int n_valid;                                      // end of this launch's range of list positions
  if (p.range_hi) n_valid = *p.range_hi;
  else if (p.hi >= 0) n_valid = p.count ? min(p.hi, *p.count) : p.hi;
  else n_valid = p.count ? *p.count : p.max_frames;
  const int n_groups = (n_valid - lo + FRAMES - 1) / FRAMES;
  // rows this thread holds pieces of: (frame, joint) of accumulator rows g and g+8
  // (the joint-16 tile has 15 live rows; row 15 computes on zeros and is never stored)
  const int f0 = w16 ? g : warp, f1 = w16 ? g + 8 : warp;
  const int j0 = w16 ? 16 : g, j1 = w16 ? 16 : g + 8;
  const uint32_t hmask0 = (t < 2) ? 0xffffffffu : 0u, hmask1 = ~hmask0;   // lanes holding the even / odd head of a k8 slice
  bf16* stg = s_stage + warp * 16 * STG;

  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int fbase = lo + grp * FRAMES;
    // ---- S1: key-point embedding + spatial PE (net:321-323), straight into the accumulator layout
    float x[4][4];
    {
      float2 p0 = make_float2(0.f, 0.f), p1 = make_float2(0.f, 0.f);
      if (f0 < FRAMES && fbase + f0 < n_valid) {
        int fr = p.list ? p.list[fbase + f0] : fbase + f0;
        if (p.src) fr = p.src[fr];
        if (fr >= 0) p0 = *reinterpret_cast<const float2*>(p.x2d + ((long long)fr * J + (p.flip ? p.flip[j0] : j0)) * 2);
      }
      if (f1 < FRAMES && fbase + f1 < n_valid) {
        int fr = p.list ? p.list[fbase + f1] : fbase + f1;
        if (p.src) fr = p.src[fr];
        if (fr >= 0) p1 = *reinterpret_cast<const float2*>(p.x2d + ((long long)fr * J + (p.flip ? p.flip[j1] : j1)) * 2);
      }
      if (p.flip) { p0.x = -p0.x; p1.x = -p1.x; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 8 * j + 2 * t;
        const float2 w0 = *reinterpret_cast<const float2*>(gp + G_EK + c);
        const float2 w1 = *reinterpret_cast<const float2*>(gp + G_EK + 32 + c);
        const float2 be = *reinterpret_cast<const float2*>(gp + G_EB + c);
        const float2 e0 = *reinterpret_cast<const float2*>(gp + G_PE + j0 * D + c);
        const float2 e1 = *reinterpret_cast<const float2*>(gp + G_PE + j1 * D + c);
        x[j][0] = (fmaf(p0.y, w1.x, p0.x * w0.x) + be.x) + e0.x;
        x[j][1] = (fmaf(p0.y, w1.y, p0.x * w0.y) + be.y) + e0.y;
        x[j][2] = (fmaf(p1.y, w1.x, p1.x * w0.x) + be.x) + e1.x;
        x[j][3] = (fmaf(p1.y, w1.y, p1.x * w0.y) + be.y) + e1.y;
      }
    }

    for (int l = 0; l < p.depth; ++l, phase ^= 1) {
      const uint2* fr = s_frag + l * F_TOTAL * 32 + lane;
      const float* bp = s_par + l * P_TOTAL;
      uint32_t a[2][4];
      uint32_t ao[2][4];                      // attention output as the A operand of the projection
      // ---- y = LN1(x); q|k|v = y @ Wqkv + b (bias preloaded into the accumulators)
      lnhat_to_afrag(x, 1e-5f, a);
      uint32_t qa[4][2], kb[4][2], vt[8][2];
#pragma unroll
      for (int tile = 0; tile < 16; ++tile) {
        const float2 b = *reinterpret_cast<const float2*>(bp + P_BQKV + 8 * tile + 2 * t);
        float c[4] = {b.x, b.y, b.x, b.y};
        const uint2 w0 = fr[(2 * tile) * 32], w1 = fr[(2 * tile + 1) * 32];
        mma16816(c, a[0], w0.x, w0.y);
        mma16816(c, a[1], w1.x, w1.y);
        // rows g / g+8, cols 2t,2t+1 of the tile: q, k as bf16, v as fp16
        if (tile < 4) { qa[tile][0] = pack2(c[0], c[1]); qa[tile][1] = pack2(c[2], c[3]); }
        else if (tile < 8) { kb[tile - 4][0] = pack2(c[0], c[1]); kb[tile - 4][1] = pack2(c[2], c[3]); }
        else { vt[tile - 8][0] = pack2h(c[0], c[1]); vt[tile - 8][1] = pack2h(c[2], c[3]); }
      }
      if (w16) {
        // joint 16 of the 15 frames (row 15 unused): publish q, k, v and the (query 16, key 16) scores
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint32_t*>(xq16 + g * 32 + 8 * j + 2 * t) = qa[j][0];
          *reinterpret_cast<uint32_t*>(xq16 + (g + 8) * 32 + 8 * j + 2 * t) = qa[j][1];
          *reinterpret_cast<uint32_t*>(xk16 + g * 32 + 8 * j + 2 * t) = kb[j][0];
          *reinterpret_cast<uint32_t*>(xk16 + (g + 8) * 32 + 8 * j + 2 * t) = kb[j][1];
          // q.k over the two channels this lane holds, completed with the neighbour lane (t ^ 1): head 2j + (t >> 1)
          const __nv_bfloat162 q0 = *reinterpret_cast<const __nv_bfloat162*>(&qa[j][0]);
          const __nv_bfloat162 q1 = *reinterpret_cast<const __nv_bfloat162*>(&qa[j][1]);
          const __nv_bfloat162 k0 = *reinterpret_cast<const __nv_bfloat162*>(&kb[j][0]);
          const __nv_bfloat162 k1 = *reinterpret_cast<const __nv_bfloat162*>(&kb[j][1]);
          float d0 = __low2float(q0) * __low2float(k0) + __high2float(q0) * __high2float(k0);
          float d1 = __low2float(q1) * __low2float(k1) + __high2float(q1) * __high2float(k1);
          d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
          d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
          if ((t & 1) == 0) {
            xs16[g * 8 + 2 * j + (t >> 1)] = d0;
            xs16[(g + 8) * 8 + 2 * j + (t >> 1)] = d1;
          }
        }
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          *reinterpret_cast<uint32_t*>(xv16 + g * 64 + 8 * h + 2 * t) = vt[h][0];
          *reinterpret_cast<uint32_t*>(xv16 + (g + 8) * 64 + 8 * h + 2 * t) = vt[h][1];
        }
        __syncwarp();
        if (lane == 0) sp_mbar_arrive(bar_pub);
      } else {
        // ---- attention of frame `warp`: queries/keys 0..15 in registers, key/query 16 from shared memory
#pragma unroll
        for (int h = 0; h < 8; ++h) { vt[h][0] = movm_t(vt[h][0]); vt[h][1] = movm_t(vt[h][1]); }   // -> V^T B fragments
        sp_mbar_wait(bar_pub, phase);
        const bf16* q16 = xq16 + warp * 32;
        const bf16* k16 = xk16 + warp * 32;
        const __half* v16 = reinterpret_cast<const __half*>(xv16) + warp * 64;
        // scores against key 16: s16[0/1] = (row g; heads 2t, 2t+1), s16[2/3] = (row g+8; ...).
        // scores of query 16 (transposed): t16[0/1] = (key g; heads 2t, 2t+1), t16[2/3] = (key g+8; ...).
        float s16[4] = {0.f, 0.f, 0.f, 0.f}, t16[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool mine = g == 2 * j + (t >> 1);           // B[k = 2t,2t+1][n = g] is non-zero iff channel's head == g
          const uint32_t bk = mine ? *reinterpret_cast<const uint32_t*>(k16 + 8 * j + 2 * t) : 0u;
          const uint32_t bq = mine ? *reinterpret_cast<const uint32_t*>(q16 + 8 * j + 2 * t) : 0u;
          mma1688(s16, qa[j][0], qa[j][1], bk);
          mma1688(t16, kb[j][0], kb[j][1], bq);
        }
        // B fragment of the key-16 step of P V: (k = heads 2t,2t+1; n = g) = v16ext[head][g]
        uint32_t bv16;
        {
          const uint16_t lo = *reinterpret_cast<const uint16_t*>(v16 + 8 * (2 * t) + g);
          const uint16_t hi = *reinterpret_cast<const uint16_t*>(v16 + 8 * (2 * t + 1) + g);
          bv16 = (uint32_t)lo | ((uint32_t)hi << 16);
        }
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const int j = h >> 1, e = h & 1;
          const uint32_t hm = e ? hmask1 : hmask0;
          const uint32_t a0 = qa[j][0] & hm, a1 = qa[j][1] & hm;
          float s0[4], s1[4];
          mma1688_z(s0, a0, a1, kb[j][0]);      // keys 0..7
          mma1688_z(s1, a0, a1, kb[j][1]);      // keys 8..15
          const bool owner = t == (h >> 1);     // this lane holds the key-16 score of head h
          float mg = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
          float mh = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
          if (owner) { mg = fmaxf(mg, s16[e]); mh = fmaxf(mh, s16[2 + e]); }
          mg = quad_max(mg);
          mh = quad_max(mh);
          uint32_t pa[4];
          pa[0] = pack2h(ex2_approx(s0[0] - mg), ex2_approx(s0[1] - mg));
          pa[1] = pack2h(ex2_approx(s0[2] - mh), ex2_approx(s0[3] - mh));
          pa[2] = pack2h(ex2_approx(s1[0] - mg), ex2_approx(s1[1] - mg));
          pa[3] = pack2h(ex2_approx(s1[2] - mh), ex2_approx(s1[3] - mh));
          const float pg = owner ? ex2_approx(s16[e] - mg) : 0.f;
          const float ph = owner ? ex2_approx(s16[2 + e] - mh) : 0.f;
          float o[4];
          mma16816h_z(o, pa, vt[h][0], vt[h][1]);  // o[0] = sum_k p v[ch t], o[1] = sum_k p   (row g); o[2], o[3]: row g+8
          mma1688h(o, e ? pack2h(0.f, pg) : pack2h(pg, 0.f), e ? pack2h(0.f, ph) : pack2h(ph, 0.f), bv16);   // key 16
          const float og = o[0] * rcp_approx(o[1]), oh = o[2] * rcp_approx(o[3]);
          // heads 4kk..4kk+3 fill slots (2t,2t+1 | 2t+8,2t+9) of k-step kk: pairs (h, h+1) pack into one register
          if (e == 0) {
            ao[h >> 2][(h & 2) ? 2 : 0] = __float_as_uint(og);      // parked as fp32 until the odd head arrives
            ao[h >> 2][(h & 2) ? 3 : 1] = __float_as_uint(oh);
          } else {
            ao[h >> 2][(h & 2) ? 2 : 0] = pack2(__uint_as_float(ao[h >> 2][(h & 2) ? 2 : 0]), og);
            ao[h >> 2][(h & 2) ? 3 : 1] = pack2(__uint_as_float(ao[h >> 2][(h & 2) ? 3 : 1]), oh);
          }
        }
        // ---- query 16 (joint 16 of this frame): softmax over the 17 keys with keys as the M dimension
        {
          const float sA = xs16[warp * 8 + 2 * t], sB = xs16[warp * 8 + 2 * t + 1];
          const float mA = fmaxf(col_max(fmaxf(t16[0], t16[2])), sA);
          const float mB = fmaxf(col_max(fmaxf(t16[1], t16[3])), sB);
          float pA0 = ex2_approx(t16[0] - mA), pA1 = ex2_approx(t16[2] - mA), pA16 = ex2_approx(sA - mA);
          float pB0 = ex2_approx(t16[1] - mB), pB1 = ex2_approx(t16[3] - mB), pB16 = ex2_approx(sB - mB);
          const float iA = rcp_approx(col_sum(pA0 + pA1) + pA16), iB = rcp_approx(col_sum(pB0 + pB1) + pB16);
          pA0 *= iA; pA1 *= iA; pA16 *= iA;
          pB0 *= iB; pB1 *= iB; pB16 *= iB;
          // (key g; heads 2t,2t+1) -> transpose -> (head g; keys 2t,2t+1): A fragments with heads as rows
          const uint32_t pt0 = movm_t(pack2h(pA0, pB0)), pt1 = movm_t(pack2h(pA1, pB1));
          float e16[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int h = 0; h < 8; ++h) {        // row h of the product with head h's V^T; other rows masked to zero
            const uint32_t aa[4] = {g == h ? pt0 : 0u, 0u, g == h ? pt1 : 0u, 0u};
            mma16816h(e16, aa, vt[h][0], vt[h][1]);
          }
          // e16[0] = (head g, dim t) over keys 0..15; add key 16 with the normalised weight of head g
          const float v0 = __shfl_sync(0xffffffffu, pA16, g >> 1), v1 = __shfl_sync(0xffffffffu, pB16, g >> 1);
          const float pn = (g & 1) ? v1 : v0;
          const float out = fmaf(pn, __half2float(v16[8 * g + 2 * t]), e16[0]);
          xo16[warp * 32 + proj_position(g, t)] = __float2bfloat16_rn(out);
        }
        __syncwarp();
        if (lane == 0) sp_mbar_arrive(bar_ret);
      }
      if (w16) {
        sp_mbar_wait(bar_ret, phase);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          ao[kk][0] = *reinterpret_cast<const uint32_t*>(xo16 + g * 32 + 16 * kk + 2 * t);
          ao[kk][1] = *reinterpret_cast<const uint32_t*>(xo16 + (g + 8) * 32 + 16 * kk + 2 * t);
          ao[kk][2] = *reinterpret_cast<const uint32_t*>(xo16 + g * 32 + 16 * kk + 8 + 2 * t);
          ao[kk][3] = *reinterpret_cast<const uint32_t*>(xo16 + (g + 8) * 32 + 16 * kk + 8 + 2 * t);
        }
      }
      // ---- x += attn @ Wp + bp   (projection rows permuted to the A order above)
#pragma unroll
      for (int j = 0; j < 4; ++j) {             // the residual stream is the accumulator
        const float2 b = *reinterpret_cast<const float2*>(bp + P_BP + 8 * j + 2 * t);
        x[j][0] += b.x; x[j][1] += b.y; x[j][2] += b.x; x[j][3] += b.y;
        const uint2 w0 = fr[(F_PROJ + 2 * j) * 32], w1 = fr[(F_PROJ + 2 * j + 1) * 32];
        mma16816(x[j], ao[0], w0.x, w0.y);
        mma16816(x[j], ao[1], w1.x, w1.y);
      }
      // ---- x += fc2(gelu(fc1(LN2(x))))
      lnhat_to_afrag(x, 1e-5f, a);
      uint32_t ah[4][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 b = *reinterpret_cast<const float2*>(bp + P_B1 + 8 * j + 2 * t);
        float c[4] = {b.x, b.y, b.x, b.y};
        const uint2 w0 = fr[(F_FC1 + 2 * j) * 32], w1 = fr[(F_FC1 + 2 * j + 1) * 32];
        mma16816(c, a[0], w0.x, w0.y);
        mma16816(c, a[1], w1.x, w1.y);
        // accumulator n-tile j -> A fragment of k-step j/2 (cols 16*(j/2) + 8*(j&1) + 2t)
        ah[j >> 1][(j & 1) * 2] = gelu_h2(c[0], c[1]);
        ah[j >> 1][(j & 1) * 2 + 1] = gelu_h2(c[2], c[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 b = *reinterpret_cast<const float2*>(bp + P_B2 + 8 * j + 2 * t);
        x[j][0] += b.x; x[j][1] += b.y; x[j][2] += b.x; x[j][3] += b.y;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint2 w = fr[(F_FC2 + 4 * j + kk) * 32];
          mma16816h(x[j], ah[kk], w.x, w.y);
        }
      }
    }

    // ---- spatial_norm (eps 1e-6, net:238), joint-major flatten (net:330): frame row = 17 x 32 bf16
    {
      uint32_t a[2][4];
      ln_to_afrag(x, gp + G_NG, gp + G_NB, 1e-6f, t, a);
      __syncwarp();
      uint32_t* sw = reinterpret_cast<uint32_t*>(stg);             // per-warp tile [16 rows][STG/2 words]
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        sw[g * (STG / 2) + 8 * kk + t] = a[kk][0];                  // cols 16kk + 2t
        sw[(g + 8) * (STG / 2) + 8 * kk + t] = a[kk][1];
        sw[g * (STG / 2) + 8 * kk + 4 + t] = a[kk][2];              // cols 16kk + 8 + 2t
        sw[(g + 8) * (STG / 2) + 8 * kk + 4 + t] = a[kk][3];
      }
      __syncwarp();
      // 16 rows x 64 B = 64 chunks of 16 B, two per lane.  Frame warp w: one contiguous 1 KB run of frame w;
      // joint-16 warp: the last 64 B of each of the 15 frames.
#pragma unroll
      for (int i = lane; i < 64; i += 32) {
        const int r = i >> 2, c = i & 3;
        const int frame = w16 ? fbase + r : fbase + warp;
        const int joint = w16 ? 16 : r;
        if (frame < n_valid && r < (w16 ? FRAMES : 16))
          *reinterpret_cast<uint4*>(p.out + ((long long)frame * J + joint) * D + c * 8) =
              *reinterpret_cast<const uint4*>(stg + r * STG + c * 8);
      }
    }
  }
}

size_t spatial_tc_smem_bytes(int depth) {
  using namespace st;
  return sizeof(uint2) * depth * F_TOTAL * 32 + sizeof(float) * ((depth * P_TOTAL + G_TOTAL + 3) & ~3) + X_TOTAL +
         sizeof(bf16) * WARPS * 16 * STG;
}

cudaError_t launch_spatial_tc(const float* x2d, const int* list, const int* count, int max_frames, int depth,
                              const void* frags, const float* params, bf16* out, int num_sms, cudaStream_t s,
                              const int* src, const int* range_lo, const int* range_hi, int lo, int hi, const int* flip) {
  if (max_frames == 0) return cudaSuccess;
  if (depth > st::DEPTH_MAX) return cudaErrorInvalidValue;
  const size_t smem = spatial_tc_smem_bytes(depth);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_spatial_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  SpatialTcParams p;
  p.x2d = x2d; p.list = list; p.count = count; p.src = src; p.max_frames = max_frames; p.depth = depth;
  p.range_lo = range_lo; p.range_hi = range_hi; p.lo = lo; p.hi = hi; p.flip = flip;
  p.frags = (const uint2*)frags; p.params = params; p.out = out;
  const int groups = (max_frames + st::FRAMES - 1) / st::FRAMES;
  k_spatial_tc<<<std::min(groups, num_sms), st::THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace uu
