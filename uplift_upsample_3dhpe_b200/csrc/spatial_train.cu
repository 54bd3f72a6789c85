// Training step, spatial stage: ONE spatial transformer block forward per launch, with the tape the backward pass reads.
//
// reference: net:313-333 (spatial_transformation), vit:176-195 (TransformerBlock.call incl. the two drop_path_layer calls),
// vit:99-156 (MultiHeadAttention).
//
// The unfused forward of a block is 11 launches over (frames x 17) rows of 32 / 64 / 96 floats (LayerNorm, the q | k | v
// linear, attention, projection, residual, LayerNorm, fc1, GELU, fc2, residual) and every one of them is memory-bound: each
// re-reads what the previous one wrote.  Here a warp owns a frame (17 joint tokens x 32 channels, lane == channel): the block's
// weights sit in shared memory once per CTA, the frame's activations in the warp's shared-memory slot, and global memory sees
// one read of x0 and one write of each tape tensor (y1, qkv, o, x1, y2, hpre, hact, x2 = 384 floats per token).  Persistent
// CTAs of 16 warps walk the frame list.  fp32 CUDA-core arithmetic (the 32 / 64-wide layers are far below any tensor-core
// tile; the stage is bound by the 26 KB of tape it writes per frame), exact-erf GELU and expf as the unfused kernels use.
#include <algorithm>

#include "common.cuh"
#include "train.cuh"

namespace uu {
namespace spt {

constexpr int J = 17, D = 32, HID = 64, HEADS = 8;
constexpr int WARPS = 16;
constexpr int YS = 36;        // row stride of the LayerNorm / attention-output buffer (float4 aligned, conflict-free)
constexpr int QS = 100;       // row stride of the q | k | v / hidden buffer
// weight image of one block in shared memory (floats)
constexpr int W_LN1G = 0, W_LN1B = 32, W_QKV = 64, W_BQKV = W_QKV + D * 96, W_P = W_BQKV + 96, W_BP = W_P + D * D,
              W_LN2G = W_BP + 32, W_LN2B = W_LN2G + 32, W_FC1 = W_LN2B + 32, W_B1 = W_FC1 + D * HID, W_FC2 = W_B1 + HID,
              W_B2 = W_FC2 + HID * D, W_TOTAL = W_B2 + 32;
constexpr int SLOT = J * D + J * YS + J * QS;      // per-warp activations

struct Params {
  const float* x0;            // [frames * 17, 32]
  const float* w[16];         // Keras order of a block: ln1 g, b | wq, bq, wk, bk, wv, bv | wp, bp | ln2 g, b | w1, b1 | w2, b2
  const float* scale;         // per-frame stochastic-depth factor of the attention branch (null = 1)
  const float* scale2;        // ... of the MLP branch
  long long frames;
  float *y1, *qkv, *o, *x1, *y2, *hpre, *hact, *x2;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[r][c] = bias[lane + 32 c] + sum_k in[r][k] W[k][lane + 32 c]: lane owns NC output columns of all 17 rows
template <int K, int NC>
__device__ __forceinline__ void warp_linear(const float* __restrict__ in, int in_stride, const float* __restrict__ W,
                                            const float* __restrict__ bias, float (&acc)[J][NC], int lane) {
  constexpr int NOUT = NC * 32;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float b = bias[lane + 32 * c];
#pragma unroll
    for (int r = 0; r < J; ++r) acc[r][c] = b;
  }
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float w[4][NC];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int c = 0; c < NC; ++c) w[kk][c] = W[(k + kk) * NOUT + lane + 32 * c];
#pragma unroll
    for (int r = 0; r < J; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(in + r * in_stride + k);      // broadcast read
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        acc[r][c] = fmaf(a.x, w[0][c], acc[r][c]);
        acc[r][c] = fmaf(a.y, w[1][c], acc[r][c]);
        acc[r][c] = fmaf(a.z, w[2][c], acc[r][c]);
        acc[r][c] = fmaf(a.w, w[3][c], acc[r][c]);
      }
    }
  }
}

// 32-column layers (projection, fc2): with one column per lane every lane re-reads all 17 rows (17 broadcast 16-byte reads per
// four k for 68 FMAs: these two layers were 60 % of the kernel's shared-memory traffic for a third of its FMAs).  Here a half-warp
// takes half of the rows (0..8 | 9..16) and a lane two adjacent columns: 9 row reads per four k for 72 FMAs.
// acc[i][0..1] = bias + sum_k in[r0 + i][k] W[k][c0 .. c0 + 1],  r0 = 9 * (lane >> 4), c0 = 2 * (lane & 15); row r0 + 8 of the
// upper half does not exist (its accumulator is computed on row 16 again and never used).
template <int K>
__device__ __forceinline__ void halfwarp_linear32(const float* __restrict__ in, int in_stride, const float* __restrict__ W,
                                                  const float* __restrict__ bias, float (&acc)[9][2], int lane) {
  const int r0 = 9 * (lane >> 4), c0 = 2 * (lane & 15);
  const float2 b = *reinterpret_cast<const float2*>(bias + c0);
#pragma unroll
  for (int i = 0; i < 9; ++i) { acc[i][0] = b.x; acc[i][1] = b.y; }
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float2 w[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float2*>(W + (k + kk) * 32 + c0);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int r = min(r0 + i, J - 1);
      const float4 a = *reinterpret_cast<const float4*>(in + r * in_stride + k);
      acc[i][0] = fmaf(a.x, w[0].x, acc[i][0]); acc[i][1] = fmaf(a.x, w[0].y, acc[i][1]);
      acc[i][0] = fmaf(a.y, w[1].x, acc[i][0]); acc[i][1] = fmaf(a.y, w[1].y, acc[i][1]);
      acc[i][0] = fmaf(a.z, w[2].x, acc[i][0]); acc[i][1] = fmaf(a.z, w[2].y, acc[i][1]);
      acc[i][0] = fmaf(a.w, w[3].x, acc[i][0]); acc[i][1] = fmaf(a.w, w[3].y, acc[i][1]);
    }
  }
}

// Keras LayerNormalization over the 32 channels of each row (lane == channel); result to shared memory and to the tape
__device__ __forceinline__ void warp_ln_rows(const float* __restrict__ xs, float* __restrict__ ys, float g, float b,
                                             float* __restrict__ tape, int lane) {
#pragma unroll
  for (int r = 0; r < J; ++r) {
    const float v = xs[r * D + lane];
    const float mean = warp_sum(v) * (1.f / D);
    const float d = v - mean;
    const float var = warp_sum(d * d) * (1.f / D);
    const float y = d * rsqrtf(var + 1e-5f) * g + b;
    ys[r * YS + lane] = y;
    tape[r * D + lane] = y;
  }
}

__global__ void __launch_bounds__(WARPS * 32, 1) k_spatial_block_fwd_tape(Params p) {
  extern __shared__ __align__(16) float smem[];
  float* wbuf = smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xs = smem + W_TOTAL + warp * SLOT;      // [17][32] residual stream of the frame
  float* ys = xs + J * D;                        // [17][YS]
  float* qs = ys + J * YS;                       // [17][QS]
  // weights -> shared memory, once per CTA
  for (int i = threadIdx.x; i < 32; i += blockDim.x) {
    wbuf[W_LN1G + i] = p.w[0][i]; wbuf[W_LN1B + i] = p.w[1][i];
    wbuf[W_BQKV + i] = p.w[3][i]; wbuf[W_BQKV + 32 + i] = p.w[5][i]; wbuf[W_BQKV + 64 + i] = p.w[7][i];
    wbuf[W_BP + i] = p.w[9][i];
    wbuf[W_LN2G + i] = p.w[10][i]; wbuf[W_LN2B + i] = p.w[11][i];
    wbuf[W_B2 + i] = p.w[15][i];
  }
  for (int i = threadIdx.x; i < HID; i += blockDim.x) wbuf[W_B1 + i] = p.w[13][i];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    const int k = i >> 5, c = i & 31;
    wbuf[W_QKV + k * 96 + c] = p.w[2][i];
    wbuf[W_QKV + k * 96 + 32 + c] = p.w[4][i];
    wbuf[W_QKV + k * 96 + 64 + c] = p.w[6][i];
    wbuf[W_P + i] = p.w[8][i];
  }
  for (int i = threadIdx.x; i < D * HID; i += blockDim.x) {
    wbuf[W_FC1 + i] = p.w[12][i];
    wbuf[W_FC2 + i] = p.w[14][i];
  }
  __syncthreads();

  for (long long f = (long long)blockIdx.x * WARPS + warp; f < p.frames; f += (long long)gridDim.x * WARPS) {
    const long long row0 = f * J;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < J; ++r) xs[r * D + lane] = p.x0[(row0 + r) * D + lane];
    __syncwarp();
    // y1 = LN1(x0)
    warp_ln_rows(xs, ys, wbuf[W_LN1G + lane], wbuf[W_LN1B + lane], p.y1 + row0 * D, lane);
    __syncwarp();
    {  // q | k | v = y1 @ [Wq Wk Wv] + b
      float acc[J][3];
      warp_linear<D, 3>(ys, YS, wbuf + W_QKV, wbuf + W_BQKV, acc, lane);
      float* tq = p.qkv + row0 * 96;
#pragma unroll
      for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          qs[r * QS + lane + 32 * c] = acc[r][c];
          tq[r * 96 + lane + 32 * c] = acc[r][c];
        }
    }
    __syncwarp();
    // attention (vit:117-129, scale 1/2): lane = (head h, key group g) keeps the k / v rows of its keys j = g, g + 4, ... in
    // registers, the warp walks the 17 queries and combines the softmax statistics and the output over the four key groups
    // with shuffles (see k_attn_small_fwd in train_kernels.cu: shared memory is read once per query instead of 34 times)
    {
      constexpr int NK = (J + 3) / 4;
      const int h = lane & 7, grp = lane >> 3;
      float4 kk[NK], vv[NK];
#pragma unroll
      for (int u = 0; u < NK; ++u) {
        const int j = grp + 4 * u;
        kk[u] = j < J ? *reinterpret_cast<const float4*>(qs + j * QS + 32 + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
        vv[u] = j < J ? *reinterpret_cast<const float4*>(qs + j * QS + 64 + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll 2
      for (int i = 0; i < J; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(qs + i * QS + 4 * h);
        float pr[NK];
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < NK; ++u) {
          pr[u] = (grp + 4 * u < J) ? fmaf(q.w, kk[u].w, fmaf(q.z, kk[u].z, fmaf(q.y, kk[u].y, q.x * kk[u].x))) * 0.5f : -INFINITY;
          mx = fmaxf(mx, pr[u]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        float sum = 0.f;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < NK; ++u) {
          pr[u] = expf(pr[u] - mx);
          sum += pr[u];
          o.x = fmaf(pr[u], vv[u].x, o.x); o.y = fmaf(pr[u], vv[u].y, o.y); o.z = fmaf(pr[u], vv[u].z, o.z); o.w = fmaf(pr[u], vv[u].w, o.w);
        }
#pragma unroll
        for (int sh = 8; sh <= 16; sh <<= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, sh);
          o.x += __shfl_xor_sync(0xffffffffu, o.x, sh); o.y += __shfl_xor_sync(0xffffffffu, o.y, sh);
          o.z += __shfl_xor_sync(0xffffffffu, o.z, sh); o.w += __shfl_xor_sync(0xffffffffu, o.w, sh);
        }
        if (grp == 0) {
          const float inv = 1.f / sum;
          o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
          *reinterpret_cast<float4*>(ys + i * YS + 4 * h) = o;             // heads merged: channel = 4 h + dim
          *reinterpret_cast<float4*>(p.o + (row0 + i) * D + 4 * h) = o;
        }
      }
    }
    __syncwarp();
    {  // x1 = x0 + scale * (o @ Wp + bp)
      float acc[9][2];
      halfwarp_linear32<D>(ys, YS, wbuf + W_P, wbuf + W_BP, acc, lane);
      const float sc = p.scale ? p.scale[f] : 1.f;
      const int r0 = 9 * (lane >> 4), c0 = 2 * (lane & 15);
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int r = r0 + i;
        if (r < J) {
          float2 v = *reinterpret_cast<const float2*>(xs + r * D + c0);
          v.x = fmaf(sc, acc[i][0], v.x); v.y = fmaf(sc, acc[i][1], v.y);
          *reinterpret_cast<float2*>(xs + r * D + c0) = v;
          *reinterpret_cast<float2*>(p.x1 + (row0 + r) * D + c0) = v;
        }
      }
    }
    __syncwarp();
    warp_ln_rows(xs, ys, wbuf[W_LN2G + lane], wbuf[W_LN2B + lane], p.y2 + row0 * D, lane);
    __syncwarp();
    {  // hpre = y2 @ W1 + b1 ; hact = gelu_erf(hpre)
      float acc[J][2];
      warp_linear<D, 2>(ys, YS, wbuf + W_FC1, wbuf + W_B1, acc, lane);
      float* tp = p.hpre + row0 * HID;
      float* ta = p.hact + row0 * HID;
#pragma unroll
      for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float v = acc[r][c];
          const float a = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
          tp[r * HID + lane + 32 * c] = v;
          ta[r * HID + lane + 32 * c] = a;
          qs[r * QS + lane + 32 * c] = a;
        }
    }
    __syncwarp();
    {  // x2 = x1 + scale2 * (hact @ W2 + b2)
      float acc[9][2];
      halfwarp_linear32<HID>(qs, QS, wbuf + W_FC2, wbuf + W_B2, acc, lane);
      const float sc = p.scale2 ? p.scale2[f] : 1.f;
      const int r0 = 9 * (lane >> 4), c0 = 2 * (lane & 15);
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int r = r0 + i;
        if (r < J) {
          const float2 x = *reinterpret_cast<const float2*>(xs + r * D + c0);
          *reinterpret_cast<float2*>(p.x2 + (row0 + r) * D + c0) = make_float2(fmaf(sc, acc[i][0], x.x), fmaf(sc, acc[i][1], x.y));
        }
      }
    }
  }
}

}  // namespace spt

bool spatial_block_fused_ok(int J, int d, int h, int heads, int act) {
  return J == spt::J && d == spt::D && h == spt::HID && heads == spt::HEADS && act == 1;
}

// weights: the 16 tensors of the block in Keras order (device pointers); scale / scale2 may be null
cudaError_t launch_spatial_block_fwd_tape(const float* x0, const float* const* weights, const float* scale, const float* scale2,
                                          long long frames, float* y1, float* qkv, float* o, float* x1, float* y2, float* hpre,
                                          float* hact, float* x2, int num_sms, cudaStream_t st) {
  if (frames == 0) return cudaSuccess;
  spt::Params p;
  p.x0 = x0;
  for (int i = 0; i < 16; ++i) p.w[i] = weights[i];
  p.scale = scale; p.scale2 = scale2; p.frames = frames;
  p.y1 = y1; p.qkv = qkv; p.o = o; p.x1 = x1; p.y2 = y2; p.hpre = hpre; p.hact = hact; p.x2 = x2;
  const size_t smem = sizeof(float) * (spt::W_TOTAL + spt::WARPS * spt::SLOT);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(spt::k_spatial_block_fwd_tape, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const unsigned grid = (unsigned)std::min<long long>((frames + spt::WARPS - 1) / spt::WARPS, num_sms);
  spt::k_spatial_block_fwd_tape<<<grid, spt::WARPS * 32, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace uu
