// K2 (bf16 path): fused spatial transformer on tensor cores.
//
// reference: net:313-333 (spatial_transformation), vit:176-195 (TransformerBlock), vit:99-156 (MHA).
// Shapes are tiny (17 joint tokens x 32 channels per frame, head_dim 4), so the kernel is built around
// keeping everything on chip rather than around one big MMA:
//   * a persistent CTA of 17 warps owns a group of 16 frames = 272 token rows = 17 row tiles of 16;
//     warp w owns row tile w for the whole network, its residual stream lives in registers
//     (m16n8 accumulator layout) from the key-point embedding to the final LayerNorm;
//   * the 32/64/96-wide linears run on mma.sync.m16n8k16 (bf16 x bf16 -> fp32); LayerNorm and GELU
//     are applied on the accumulator fragments, which convert to the next A operand without shuffles;
//   * all four blocks' weights sit in shared memory, pre-swizzled at weight-commit time into
//     per-lane B-fragment order (one conflict-free 8-byte load per MMA);
//   * only q/k/v (fp32) and the attention output (bf16) go through shared memory, because the
//     17x17 attention of a frame needs rows owned by two different warps.
// HBM traffic: 136 B of key-points in, 1088 B (17x32 bf16) out per frame.
#include <algorithm>

#include "common.cuh"

namespace uu {

namespace st {
constexpr int J = 17, D = 32, HID = 64, HEADS = 8, DEPTH_MAX = 4;
constexpr int FRAMES = 16;                  // frames per group
constexpr int ROWS = FRAMES * J;            // 272 = 17 tiles of 16
constexpr int WARPS = 17, THREADS = WARPS * 32;
constexpr int QS = 100;                     // fp32 q|k|v row stride (96 + pad)
constexpr int OS = 40;                      // bf16 attention-output row stride (32 + pad): conflict-free A loads
// B-fragment image of one block: [frag][lane] uint2
constexpr int F_QKV = 0, F_PROJ = 24, F_FC1 = 32, F_FC2 = 48, F_TOTAL = 64;
// fp32 parameter image of one block
constexpr int P_LN1G = 0, P_LN1B = 32, P_BQKV = 64, P_BP = 160, P_LN2G = 192, P_LN2B = 224, P_B1 = 256, P_B2 = 320,
              P_TOTAL = 352;
// global fp32 parameters
constexpr int G_EK = 0, G_EB = 64, G_PE = 96, G_NG = G_PE + J * D, G_NB = G_NG + 32, G_TOTAL = G_NB + 32;
}  // namespace st

// ---- weight image (built once per uu_set_weight round) -------------------------------------------
// frags: [depth][64][32] uint2 ; params: [depth][352] + [704] floats
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
    // which matrix, n-tile j, k-step kk
    const float* Wm; int N, j, kk, ncol0 = 0;
    if (fr < F_PROJ) { j = fr >> 1; kk = fr & 1; N = 32; const int which = j >> 2;   // q | k | v, 4 n-tiles each
      Wm = tb[2 + 2 * which]; ncol0 = (j & 3) * 8; }
    else if (fr < F_FC1) { const int f = fr - F_PROJ; j = f >> 1; kk = f & 1; N = 32; Wm = tb[8]; ncol0 = j * 8; }
    else if (fr < F_FC2) { const int f = fr - F_FC1; j = f >> 1; kk = f & 1; N = 64; Wm = tb[12]; ncol0 = j * 8; }
    else { const int f = fr - F_FC2; j = f >> 2; kk = f & 3; N = 32; Wm = tb[14]; ncol0 = j * 8; }
    const int n = ncol0 + g, k0 = 16 * kk + 2 * t;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(Wm[(k0)*N + n], Wm[(k0 + 1) * N + n]);
    __nv_bfloat162 b1 = __floats2bfloat162_rn(Wm[(k0 + 8) * N + n], Wm[(k0 + 9) * N + n]);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&b0);
    o.y = *reinterpret_cast<uint32_t*>(&b1);
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
      else if (o < P_BP) { const int c = o - P_BQKV; v = tb[3 + 2 * (c >> 5)][c & 31]; }
      else if (o < P_LN2G) v = tb[9][o - P_BP];
      else if (o < P_LN2B) v = tb[10][o - P_LN2G];
      else if (o < P_B1) v = tb[11][o - P_LN2B];
      else if (o < P_B2) v = tb[13][o - P_B1];
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
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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

// One (frame, head, query) item of the 17x17 attention: softmax(q k^T / 2) v in fp32, result as 4 bf16.
__device__ __forceinline__ void attention_item(const float* __restrict__ s_qkv, bf16* __restrict__ s_o, int it) {
  using namespace st;
  const int pr = it / J, i = it - pr * J;
  const int f = pr >> 3, h = pr & 7;
  const float* base = s_qkv + f * J * QS + h * 4;
  const float4 q = *reinterpret_cast<const float4*>(base + i * QS);
  float s[J];
  float mx = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < J; ++jj) {
    const float4 k4 = *reinterpret_cast<const float4*>(base + jj * QS + 32);
    s[jj] = fmaf(q.w, k4.w, fmaf(q.z, k4.z, fmaf(q.y, k4.y, q.x * k4.x)));
    mx = fmaxf(mx, s[jj]);
  }
  // softmax(s / sqrt(4)): exp((s - max)/2) = exp2(s * c - max * c), c = 0.5*log2(e)
  const float sc = 0.72134752044448170368f;
  const float nm = -mx * sc;
  float sum = 0.f;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int jj = 0; jj < J; ++jj) {
    const float pj = ex2_approx(fmaf(s[jj], sc, nm));
    sum += pj;
    const float4 v4 = *reinterpret_cast<const float4*>(base + jj * QS + 64);
    o.x = fmaf(pj, v4.x, o.x); o.y = fmaf(pj, v4.y, o.y);
    o.z = fmaf(pj, v4.z, o.z); o.w = fmaf(pj, v4.w, o.w);
  }
  const float inv = __frcp_rn(sum);
  uint2 pk;
  pk.x = pack2(o.x * inv, o.y * inv);
  pk.y = pack2(o.z * inv, o.w * inv);
  *reinterpret_cast<uint2*>(s_o + (f * J + i) * OS + h * 4) = pk;       // heads merged: channel = 4h + d
}

struct SpatialTcParams {
  const float* x2d;      // (B*n_tok, 17, 2)
  const int* list;       // gather list or null
  const int* count;      // device count or null
  int max_frames;
  int depth;
  const uint2* frags;    // weight image
  const float* params;
  bf16* out;             // (n_valid, 544) compact
};

__global__ void __launch_bounds__(st::THREADS, 1) k_spatial_tc(SpatialTcParams p) {
  using namespace st;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint2* s_frag = reinterpret_cast<uint2*>(smem_raw);                              // depth*64*32 uint2
  float* s_par = reinterpret_cast<float*>(smem_raw + sizeof(uint2) * p.depth * F_TOTAL * 32);
  float* s_qkv = s_par + ((p.depth * P_TOTAL + G_TOTAL + 3) & ~3);                 // ROWS * QS floats
  bf16* s_o = reinterpret_cast<bf16*>(s_qkv + ROWS * QS);                          // ROWS * OS bf16
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  // weights -> shared memory, once per CTA
  {
    const int nf = p.depth * F_TOTAL * 32;
    for (int i = tid; i < nf; i += THREADS) s_frag[i] = p.frags[i];
    const int np = p.depth * P_TOTAL + G_TOTAL;
    for (int i = tid; i < np; i += THREADS) s_par[i] = p.params[i];
  }
  __syncthreads();
  const float* gp = s_par + p.depth * P_TOTAL;      // global params

  const int n_valid = p.count ? *p.count : p.max_frames;
  const int n_groups = (n_valid + FRAMES - 1) / FRAMES;
  const int r0 = warp * 16 + g, r1 = r0 + 8;        // the two token rows this thread holds pieces of
  const int f0 = r0 / J, j0 = r0 - f0 * J, f1 = r1 / J, j1 = r1 - f1 * J;

  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int fbase = grp * FRAMES;
    // ---- S1: key-point embedding + spatial PE (net:321-323), straight into the accumulator layout
    float x[4][4];
    {
      float2 p0 = make_float2(0.f, 0.f), p1 = make_float2(0.f, 0.f);
      if (fbase + f0 < n_valid) {
        const int fr = p.list ? p.list[fbase + f0] : fbase + f0;
        p0 = *reinterpret_cast<const float2*>(p.x2d + ((long long)fr * J + j0) * 2);
      }
      if (fbase + f1 < n_valid) {
        const int fr = p.list ? p.list[fbase + f1] : fbase + f1;
        p1 = *reinterpret_cast<const float2*>(p.x2d + ((long long)fr * J + j1) * 2);
      }
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

    for (int l = 0; l < p.depth; ++l) {
      const uint2* fr = s_frag + l * F_TOTAL * 32 + lane;
      const float* bp = s_par + l * P_TOTAL;
      uint32_t a[2][4];
      // ---- y = LN1(x); q|k|v = y @ Wqkv + b -> shared memory (fp32)
      ln_to_afrag(x, bp + P_LN1G, bp + P_LN1B, 1e-5f, t, a);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(c, a[0], fr[(F_QKV + 2 * j) * 32]);
        mma16816(c, a[1], fr[(F_QKV + 2 * j + 1) * 32]);
        const float2 b = *reinterpret_cast<const float2*>(bp + P_BQKV + 8 * j + 2 * t);
        *reinterpret_cast<float2*>(s_qkv + r0 * QS + 8 * j + 2 * t) = make_float2(c[0] + b.x, c[1] + b.y);
        *reinterpret_cast<float2*>(s_qkv + r1 * QS + 8 * j + 2 * t) = make_float2(c[2] + b.x, c[3] + b.y);
      }
      __syncthreads();
      // ---- attention: 16 frames x 8 heads x 17 queries = 2176 items, 4 per thread (vit:117-129);
      //      two independent items per iteration keep more shared-memory loads in flight
#pragma unroll 1
      for (int it = tid; it < FRAMES * HEADS * J; it += 2 * THREADS) {
        attention_item(s_qkv, s_o, it);
        attention_item(s_qkv, s_o, it + THREADS);
      }
      __syncthreads();
      // ---- x += attn @ Wp + bp
      {
        uint32_t ao[2][4];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          ao[kk][0] = *reinterpret_cast<const uint32_t*>(s_o + r0 * OS + 16 * kk + 2 * t);
          ao[kk][1] = *reinterpret_cast<const uint32_t*>(s_o + r1 * OS + 16 * kk + 2 * t);
          ao[kk][2] = *reinterpret_cast<const uint32_t*>(s_o + r0 * OS + 16 * kk + 8 + 2 * t);
          ao[kk][3] = *reinterpret_cast<const uint32_t*>(s_o + r1 * OS + 16 * kk + 8 + 2 * t);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float c[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(c, ao[0], fr[(F_PROJ + 2 * j) * 32]);
          mma16816(c, ao[1], fr[(F_PROJ + 2 * j + 1) * 32]);
          const float2 b = *reinterpret_cast<const float2*>(bp + P_BP + 8 * j + 2 * t);
          x[j][0] += c[0] + b.x; x[j][1] += c[1] + b.y; x[j][2] += c[2] + b.x; x[j][3] += c[3] + b.y;
        }
      }
      // ---- x += fc2(gelu(fc1(LN2(x))))
      ln_to_afrag(x, bp + P_LN2G, bp + P_LN2B, 1e-5f, t, a);
      uint32_t ah[4][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(c, a[0], fr[(F_FC1 + 2 * j) * 32]);
        mma16816(c, a[1], fr[(F_FC1 + 2 * j + 1) * 32]);
        const float2 b = *reinterpret_cast<const float2*>(bp + P_B1 + 8 * j + 2 * t);
        const float h0 = gelu_erf_fast(c[0] + b.x), h1 = gelu_erf_fast(c[1] + b.y);
        const float h2 = gelu_erf_fast(c[2] + b.x), h3 = gelu_erf_fast(c[3] + b.y);
        // accumulator n-tile j -> A fragment of k-step j/2 (cols 16*(j/2) + 8*(j&1) + 2t)
        ah[j >> 1][(j & 1) * 2] = pack2(h0, h1);
        ah[j >> 1][(j & 1) * 2 + 1] = pack2(h2, h3);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma16816(c, ah[kk], fr[(F_FC2 + 4 * j + kk) * 32]);
        const float2 b = *reinterpret_cast<const float2*>(bp + P_B2 + 8 * j + 2 * t);
        x[j][0] += c[0] + b.x; x[j][1] += c[1] + b.y; x[j][2] += c[2] + b.x; x[j][3] += c[3] + b.y;
      }
    }

    // ---- spatial_norm (eps 1e-6, net:238), joint-major flatten (net:330): 16 frames x 544 bf16 contiguous
    {
      uint32_t a[2][4];
      ln_to_afrag(x, gp + G_NG, gp + G_NB, 1e-6f, t, a);
      // stage in shared memory (the q|k|v buffer is free: every warp is past the last attention)
      uint32_t* stage = reinterpret_cast<uint32_t*>(s_qkv);     // [ROWS][16] words (32 bf16 per row)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        stage[r0 * 16 + 8 * kk + t] = a[kk][0];          // cols 16kk + 2t
        stage[r1 * 16 + 8 * kk + t] = a[kk][1];
        stage[r0 * 16 + 8 * kk + 4 + t] = a[kk][2];      // cols 16kk + 8 + 2t
        stage[r1 * 16 + 8 * kk + 4 + t] = a[kk][3];
      }
      __syncthreads();
      const int frames_here = min(FRAMES, n_valid - fbase);
      const int n16 = frames_here * J * D * 2 / 16;                 // 16-byte chunks
      const uint4* src = reinterpret_cast<const uint4*>(s_qkv);
      uint4* dst = reinterpret_cast<uint4*>(p.out + (long long)fbase * J * D);
      for (int i = tid; i < n16; i += THREADS) dst[i] = src[i];
      __syncthreads();                                              // staging is reused as q|k|v by the next group
    }
  }
}

size_t spatial_tc_smem_bytes(int depth) {
  using namespace st;
  return sizeof(uint2) * depth * F_TOTAL * 32 + sizeof(float) * ((depth * P_TOTAL + G_TOTAL + 3) & ~3) +
         sizeof(float) * ROWS * QS + sizeof(bf16) * ROWS * OS;
}

cudaError_t launch_spatial_tc(const float* x2d, const int* list, const int* count, int max_frames, int depth,
                              const void* frags, const float* params, bf16* out, int num_sms, cudaStream_t s) {
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
  p.x2d = x2d; p.list = list; p.count = count; p.max_frames = max_frames; p.depth = depth;
  p.frags = (const uint2*)frags; p.params = params; p.out = out;
  const int groups = (max_frames + st::FRAMES - 1) / st::FRAMES;
  k_spatial_tc<<<std::min(groups, num_sms), st::THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace uu
