// K5 (bf16 path): fused softmax attention per (window, head) on tensor cores.
//   S = Q K^T / sqrt(dh) (+ key mask * -1e9, literal fp32 arithmetic of vit:117-123), softmax, O = P V,
//   heads merged on store.  A whole window (<= 128 tokens) lives in one CTA: K and V^T of the head sit in
//   shared memory (padded strides, conflict-free B-fragment loads), every warp owns 16 query rows, the
//   score tile stays in registers (mma.sync m16n8k16 accumulators) and is re-used as the A operand of P V
//   without leaving the register file.  The (B,8,S,S) score tensor of the reference never exists.
#include "common.cuh"

namespace uu {

__device__ __forceinline__ void mma16816_att(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2_att(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ST = number of 16-row tiles covering the sequence (S <= 16*ST); blockDim = 32*ST.
template <int DH, int ST>
__global__ void __launch_bounds__(32 * ST) k_attention_tc(const bf16* __restrict__ qkv, int S, int heads,
                                                          const uint8_t* __restrict__ mask, int mask_stride,
                                                          bf16* __restrict__ out) {
  constexpr int SP = 16 * ST;            // padded sequence
  constexpr int NT = SP / 8;             // key n-tiles of the score matrix
  constexpr int KS = DH + 8;             // K row stride (bf16): (KS/2) % 8 == 4 -> conflict-free fragment loads
  constexpr int VS = SP + 8;             // V^T row stride (bf16)
  __shared__ __align__(16) bf16 Ks[SP * KS];
  __shared__ __align__(16) bf16 Vt[DH * VS];
  __shared__ float Km[SP];               // additive key term: 0, -1e9 (masked key) or -inf (padding)
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int d = heads * DH;
  const long long row0 = (long long)b * S;
  constexpr int CH = DH / 8;             // 16-byte chunks per head row
  for (int i = tid; i < SP * CH; i += 32 * ST) {
    const int key = i / CH, c = i - key * CH;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (key < S) {
      const bf16* r = qkv + (row0 + key) * 3 * d + h * DH + c * 8;
      kv = *reinterpret_cast<const uint4*>(r + d);
      vv = *reinterpret_cast<const uint4*>(r + 2 * d);
    }
    *reinterpret_cast<uint4*>(Ks + key * KS + c * 8) = kv;
    const bf16* ve = reinterpret_cast<const bf16*>(&vv);
#pragma unroll
    for (int e = 0; e < 8; ++e) Vt[(c * 8 + e) * VS + key] = ve[e];
  }
  for (int j = tid; j < SP; j += 32 * ST)
    Km[j] = j >= S ? -INFINITY : ((mask && !mask[(long long)b * mask_stride + j]) ? -1e9f : 0.f);
  __syncthreads();

  const int q0 = warp * 16 + g, q1 = q0 + 8;         // query rows of this thread
  // Q fragments straight from global memory (each row is read exactly once)
  uint32_t aq[DH / 16][4];
  {
    const bf16* qr0 = qkv + (row0 + min(q0, S - 1)) * 3 * d + h * DH;
    const bf16* qr1 = qkv + (row0 + min(q1, S - 1)) * 3 * d + h * DH;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      aq[kk][0] = *reinterpret_cast<const uint32_t*>(qr0 + 16 * kk + 2 * t);
      aq[kk][1] = *reinterpret_cast<const uint32_t*>(qr1 + 16 * kk + 2 * t);
      aq[kk][2] = *reinterpret_cast<const uint32_t*>(qr0 + 16 * kk + 8 + 2 * t);
      aq[kk][3] = *reinterpret_cast<const uint32_t*>(qr1 + 16 * kk + 8 + 2 * t);
    }
  }
  float sc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
    const bf16* kr = Ks + (8 * j + g) * KS + 2 * t;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk)
      mma16816_att(sc[j], aq[kk], *reinterpret_cast<const uint32_t*>(kr + 16 * kk),
                   *reinterpret_cast<const uint32_t*>(kr + 16 * kk + 8));
  }
  // logits = s / sqrt(dh) + key term (fp32, literal), row max, exp, row sum
  const float scale = 1.0f / sqrtf((float)DH);
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const float2 km = *reinterpret_cast<const float2*>(Km + 8 * j + 2 * t);
    sc[j][0] = sc[j][0] * scale + km.x; sc[j][1] = sc[j][1] * scale + km.y;
    sc[j][2] = sc[j][2] * scale + km.x; sc[j][3] = sc[j][3] * scale + km.y;
    m0 = fmaxf(m0, fmaxf(sc[j][0], sc[j][1]));
    m1 = fmaxf(m1, fmaxf(sc[j][2], sc[j][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
  const float LOG2E = 1.4426950408889634f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    sc[j][0] = exp2f((sc[j][0] - m0) * LOG2E); sc[j][1] = exp2f((sc[j][1] - m0) * LOG2E);
    sc[j][2] = exp2f((sc[j][2] - m1) * LOG2E); sc[j][3] = exp2f((sc[j][3] - m1) * LOG2E);
    l0 += sc[j][0] + sc[j][1];
    l1 += sc[j][2] + sc[j][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // O = P V : the score accumulators are the A fragments of the second MMA
  float o[DH / 8][4];
#pragma unroll
  for (int jn = 0; jn < DH / 8; ++jn) o[jn][0] = o[jn][1] = o[jn][2] = o[jn][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < ST; ++kk) {
    uint32_t ap[4];
    ap[0] = pack2_att(sc[2 * kk][0], sc[2 * kk][1]);
    ap[1] = pack2_att(sc[2 * kk][2], sc[2 * kk][3]);
    ap[2] = pack2_att(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
    ap[3] = pack2_att(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
    for (int jn = 0; jn < DH / 8; ++jn) {
      const bf16* vr = Vt + (8 * jn + g) * VS + 16 * kk + 2 * t;
      mma16816_att(o[jn], ap, *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
    }
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  if (q0 < S) {
    bf16* orow = out + (row0 + q0) * d + h * DH + 2 * t;
#pragma unroll
    for (int jn = 0; jn < DH / 8; ++jn) *reinterpret_cast<uint32_t*>(orow + 8 * jn) = pack2_att(o[jn][0] * i0, o[jn][1] * i0);
  }
  if (q1 < S) {
    bf16* orow = out + (row0 + q1) * d + h * DH + 2 * t;
#pragma unroll
    for (int jn = 0; jn < DH / 8; ++jn) *reinterpret_cast<uint32_t*>(orow + 8 * jn) = pack2_att(o[jn][2] * i1, o[jn][3] * i1);
  }
}

template <int DH>
static cudaError_t att_tc_dh(const bf16* qkv, int B, int S, int heads, const uint8_t* mask, int mask_stride, bf16* out,
                             cudaStream_t st) {
  const int tiles = (S + 15) / 16;
  dim3 grid(B, heads);
#define UU_ATT_CASE(T)                                                                               \
  case T:                                                                                            \
    k_attention_tc<DH, T><<<grid, 32 * T, 0, st>>>(qkv, S, heads, mask, mask_stride, out);           \
    break;
  switch (tiles) {
    UU_ATT_CASE(1) UU_ATT_CASE(2) UU_ATT_CASE(3) UU_ATT_CASE(4)
    UU_ATT_CASE(5) UU_ATT_CASE(6) UU_ATT_CASE(7) UU_ATT_CASE(8)
    default: return cudaErrorInvalidValue;
  }
#undef UU_ATT_CASE
  return cudaGetLastError();
}

cudaError_t launch_attention_tc(const bf16* qkv, int B, int S, int heads, int dh, const uint8_t* mask, int mask_stride,
                                bf16* out, cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  if (S < 1 || S > 128) return cudaErrorInvalidValue;
  switch (dh) {
    case 16: return att_tc_dh<16>(qkv, B, S, heads, mask, mask_stride, out, st);
    case 32: return att_tc_dh<32>(qkv, B, S, heads, mask, mask_stride, out, st);
    case 48: return att_tc_dh<48>(qkv, B, S, heads, mask, mask_stride, out, st);
    case 64: return att_tc_dh<64>(qkv, B, S, heads, mask, mask_stride, out, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace uu
