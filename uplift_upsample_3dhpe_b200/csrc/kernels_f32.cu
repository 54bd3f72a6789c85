// CUDA-core kernels of the uu3d hot path (sm_100a): stride-mask gather list, fused spatial
// transformer, upsampling-token fill, LayerNorm, softmax attention and the fp32 reference-precision
// GEMM.  Together they are the complete fp32 ("exact") forward path; the bf16 path swaps the GEMMs
// for the tcgen05 kernel in gemm_tc.cu and keeps the memory-bound kernels.
#include <algorithm>

#include "common.cuh"
#include "epilogue.cuh"

namespace uu {

// =================================================================================================
// K1a: valid-frame gather list from the stride mask (bit-exact index work).
// reference semantics: frames with mask==0 never influence the output (their spatial result is
// multiplied by 0, net:350), so the spatial stage only runs on list[0..count).
// =================================================================================================
__global__ void k_mask_count(const uint8_t* __restrict__ mask, int B, int n_tok, int* __restrict__ counts) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B) return;
  int c = 0;
  for (int n0 = 0; n0 < n_tok; n0 += 32) {
    int n = n0 + lane;
    bool v = n < n_tok && mask[(long long)w * n_tok + n] != 0;
    c += __popc(__ballot_sync(0xffffffffu, v));
  }
  if (lane == 0) counts[w] = c;
}

// single-CTA exclusive scan of counts[0..B) in place; counts[B] = total; count_out[0] = total
__global__ void k_mask_scan(int* __restrict__ counts, int B, int* __restrict__ count_out) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < B; base += blockDim.x) {
    int i = base + tid;
    int v = i < B ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;   // inclusive
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + (wid ? warp_sums[wid - 1] : 0) + x - v;
    if (i < B) counts[i] = excl;
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = carry + warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
  if (tid == 0) {
    counts[B] = carry_s;
    count_out[0] = carry_s;
  }
}

__global__ void k_mask_fill(const uint8_t* __restrict__ mask, int B, int n_tok, const int* __restrict__ offs,
                            int* __restrict__ list) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B) return;
  int base = offs[w];
  for (int n0 = 0; n0 < n_tok; n0 += 32) {
    int n = n0 + lane;
    bool v = n < n_tok && mask[(long long)w * n_tok + n] != 0;
    unsigned bal = __ballot_sync(0xffffffffu, v);
    if (v) list[base + __popc(bal & ((1u << lane) - 1u))] = w * n_tok + n;
    base += __popc(bal);
  }
}

cudaError_t launch_build_gather(const uint8_t* mask, int B, int n_tok, int* scratch, int* list, int* count_out,
                                cudaStream_t st) {
  int blocks = (B * 32 + 127) / 128;
  k_mask_count<<<blocks, 128, 0, st>>>(mask, B, n_tok, scratch);
  k_mask_scan<<<1, 1024, 0, st>>>(scratch, B, count_out);
  k_mask_fill<<<blocks, 128, 0, st>>>(mask, B, n_tok, scratch, list);
  return cudaGetLastError();
}

// =================================================================================================
// K2 (fp32): fused spatial transformer.  One warp owns one frame (17 joint tokens x 32 channels,
// lane == channel); a CTA of 8 warps shares one block's weights in shared memory.  Nothing but the
// final (J*32) row per frame ever reaches HBM.
// reference: net:313-333 (spatial_transformation), vit:176-195 (block), vit:99-156 (MHA).
// =================================================================================================
constexpr int SP_J = 17;        // joints
constexpr int SP_D = 32;        // SPATIAL_EMBED_DIM
constexpr int SP_HID = 64;      // hidden = d * MLP_RATIO
constexpr int SP_HEADS = 8;     // head_dim 4
constexpr int SP_F = 8;         // frames (warps) per CTA
constexpr int SP_YS = 36;       // row stride of the LN/attention-output buffer (float4 aligned)
constexpr int SP_QS = 100;      // row stride of the qkv / hidden buffer
// per-block weight image in shared memory (floats)
constexpr int W_LN1G = 0, W_LN1B = 32, W_QKV = 64, W_BQKV = W_QKV + 32 * 96, W_P = W_BQKV + 96,
              W_BP = W_P + 32 * 32, W_LN2G = W_BP + 32, W_LN2B = W_LN2G + 32, W_FC1 = W_LN2B + 32,
              W_B1 = W_FC1 + 32 * 64, W_FC2 = W_B1 + 64, W_B2 = W_FC2 + 64 * 32, W_TOTAL = W_B2 + 32;

template <int K, int NC>
__device__ __forceinline__ void warp_linear(const float* __restrict__ in, int in_stride, const float* __restrict__ W,
                                            const float* __restrict__ bias, float (&acc)[SP_J][NC], int lane) {
  constexpr int NOUT = NC * 32;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float b = bias[lane + 32 * c];
#pragma unroll
    for (int r = 0; r < SP_J; ++r) acc[r][c] = b;
  }
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float w[4][NC];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int c = 0; c < NC; ++c) w[kk][c] = W[(k + kk) * NOUT + lane + 32 * c];
#pragma unroll
    for (int r = 0; r < SP_J; ++r) {
      float4 a = *reinterpret_cast<const float4*>(in + r * in_stride + k);
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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm over the 32 channels of each of the 17 rows (lane == channel), Keras non-fused form.
__device__ __forceinline__ void warp_ln_rows(const float* __restrict__ xs, float* __restrict__ ys, float g, float b,
                                             float eps, int lane) {
#pragma unroll
  for (int r = 0; r < SP_J; ++r) {
    float v = xs[r * SP_D + lane];
    float mean = warp_sum(v) * (1.f / SP_D);
    float d = v - mean;
    float var = warp_sum(d * d) * (1.f / SP_D);
    float inv = g * rsqrtf(var + eps);
    ys[r * SP_YS + lane] = v * inv + (b - mean * inv);
  }
}

template <typename TOut>
__global__ void __launch_bounds__(SP_F * 32, 1) k_spatial_f32(SpatialParams p) {
  extern __shared__ __align__(16) float smem[];
  float* wbuf = smem;                                     // W_TOTAL
  float* xs_all = wbuf + W_TOTAL;                         // SP_F * 17*32
  float* ys_all = xs_all + SP_F * SP_J * SP_D;            // SP_F * 17*36
  float* qs_all = ys_all + SP_F * SP_J * SP_YS;           // SP_F * 17*100
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_valid = p.count ? *p.count : p.max_frames;
  const int g0 = blockIdx.x * SP_F;
  if (g0 >= n_valid) return;                              // whole CTA past the gather list
  const int g = g0 + warp;
  const bool active = g < n_valid;
  float* xs = xs_all + warp * SP_J * SP_D;
  float* ys = ys_all + warp * SP_J * SP_YS;
  float* qs = qs_all + warp * SP_J * SP_QS;

  // S1: key-point embedding + spatial positional encoding (net:321-323)
  if (active) {
    const int tok = p.list ? p.list[g] : g;
    const int fr = p.src ? p.src[tok] : tok;              // window token -> video frame (fused window gather)
    const float* x = p.x2d + (long long)(fr < 0 ? 0 : fr) * SP_J * 2;
    const float w0 = p.embed_k[lane], w1 = p.embed_k[SP_D + lane], be = p.embed_b[lane];
#pragma unroll
    for (int j = 0; j < SP_J; ++j) {
      // flip augmentation (eval.py:154-159): joint j reads source joint flip[j] with the x coordinate negated
      float2 xy = *reinterpret_cast<const float2*>(x + 2 * (p.flip ? p.flip[j] : j));
      if (p.flip) xy.x = -xy.x;
      if (fr < 0) xy = make_float2(0.f, 0.f);             // zero padding outside the video
      // Dense = x @ W + b (sum over k in order), then + PE
      xs[j * SP_D + lane] = (fmaf(xy.y, w1, xy.x * w0) + be) + p.pe[j * SP_D + lane];
    }
  }

  for (int l = 0; l < p.depth; ++l) {
    __syncthreads();                                      // everyone done with the previous block's weights
    const float* const* t = p.blocks + l * 16;
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
      wbuf[W_LN1G + i] = t[0][i]; wbuf[W_LN1B + i] = t[1][i];
      wbuf[W_BQKV + i] = t[3][i]; wbuf[W_BQKV + 32 + i] = t[5][i]; wbuf[W_BQKV + 64 + i] = t[7][i];
      wbuf[W_BP + i] = t[9][i];
      wbuf[W_LN2G + i] = t[10][i]; wbuf[W_LN2B + i] = t[11][i];
      wbuf[W_B2 + i] = t[15][i];
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) wbuf[W_B1 + i] = t[13][i];
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
      int k = i >> 5, c = i & 31;
      wbuf[W_QKV + k * 96 + c] = t[2][i];
      wbuf[W_QKV + k * 96 + 32 + c] = t[4][i];
      wbuf[W_QKV + k * 96 + 64 + c] = t[6][i];
      wbuf[W_P + i] = t[8][i];
    }
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
      wbuf[W_FC1 + i] = t[12][i];
      wbuf[W_FC2 + i] = t[14][i];
    }
    __syncthreads();
    if (!active) continue;

    // y = LN1(x)
    warp_ln_rows(xs, ys, wbuf[W_LN1G + lane], wbuf[W_LN1B + lane], 1e-5f, lane);
    __syncwarp();
    {  // q | k | v = y @ [Wq Wk Wv] + b
      float acc[SP_J][3];
      warp_linear<SP_D, 3>(ys, SP_YS, wbuf + W_QKV, wbuf + W_BQKV, acc, lane);
#pragma unroll
      for (int r = 0; r < SP_J; ++r) {
        qs[r * SP_QS + lane] = acc[r][0];
        qs[r * SP_QS + 32 + lane] = acc[r][1];
        qs[r * SP_QS + 64 + lane] = acc[r][2];
      }
    }
    __syncwarp();
    // attention: 8 heads x 17 queries = 136 items over 32 lanes (vit:117-129), head_dim 4, scale 1/2
    for (int it = lane; it < SP_HEADS * SP_J; it += 32) {
      const int h = it / SP_J, i = it - h * SP_J;
      const float4 q = *reinterpret_cast<const float4*>(qs + i * SP_QS + h * 4);
      float s[SP_J];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < SP_J; ++j) {
        const float4 kk = *reinterpret_cast<const float4*>(qs + j * SP_QS + 32 + h * 4);
        float d = fmaf(q.w, kk.w, fmaf(q.z, kk.z, fmaf(q.y, kk.y, q.x * kk.x)));
        s[j] = d * 0.5f;
        mx = fmaxf(mx, s[j]);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < SP_J; ++j) {
        s[j] = expf(s[j] - mx);
        sum += s[j];
      }
      const float inv = 1.f / sum;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < SP_J; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(qs + j * SP_QS + 64 + h * 4);
        const float pj = s[j] * inv;
        o.x = fmaf(pj, v.x, o.x); o.y = fmaf(pj, v.y, o.y); o.z = fmaf(pj, v.z, o.z); o.w = fmaf(pj, v.w, o.w);
      }
      *reinterpret_cast<float4*>(ys + i * SP_YS + h * 4) = o;     // heads merged: channel = h*4 + d
    }
    __syncwarp();
    {  // x += attn @ Wp + bp
      float acc[SP_J][1];
      warp_linear<SP_D, 1>(ys, SP_YS, wbuf + W_P, wbuf + W_BP, acc, lane);
#pragma unroll
      for (int r = 0; r < SP_J; ++r) xs[r * SP_D + lane] += acc[r][0];
    }
    __syncwarp();
    warp_ln_rows(xs, ys, wbuf[W_LN2G + lane], wbuf[W_LN2B + lane], 1e-5f, lane);
    __syncwarp();
    {  // h = gelu_erf(z @ W1 + b1)
      float acc[SP_J][2];
      warp_linear<SP_D, 2>(ys, SP_YS, wbuf + W_FC1, wbuf + W_B1, acc, lane);
#pragma unroll
      for (int r = 0; r < SP_J; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v = acc[r][c];
          qs[r * SP_QS + lane + 32 * c] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        }
    }
    __syncwarp();
    {  // x += h @ W2 + b2
      float acc[SP_J][1];
      warp_linear<SP_HID, 1>(qs, SP_QS, wbuf + W_FC2, wbuf + W_B2, acc, lane);
#pragma unroll
      for (int r = 0; r < SP_J; ++r) xs[r * SP_D + lane] += acc[r][0];
    }
    __syncwarp();
  }
  if (!active) return;
  // spatial_norm (eps 1e-6, net:238) and joint-major flatten (net:330): out[g][j*32 + c]
  {
    const float gn = p.norm_g[lane], bn = p.norm_b[lane];
    TOut* out = reinterpret_cast<TOut*>(p.out) + (long long)g * SP_J * SP_D;
#pragma unroll
    for (int r = 0; r < SP_J; ++r) {
      float v = xs[r * SP_D + lane];
      float mean = warp_sum(v) * (1.f / SP_D);
      float d = v - mean;
      float var = warp_sum(d * d) * (1.f / SP_D);
      float inv = gn * rsqrtf(var + 1e-6f);
      store_out(out + r * SP_D + lane, v * inv + (bn - mean * inv));
    }
  }
}

cudaError_t launch_spatial_f32(const SpatialParams& p, cudaStream_t st) {
  if (p.J != SP_J) return cudaErrorInvalidValue;
  const size_t smem = sizeof(float) * (W_TOTAL + SP_F * SP_J * (SP_D + SP_YS + SP_QS));
  const int grid = (p.max_frames + SP_F - 1) / SP_F;
  if (grid == 0) return cudaSuccess;
  cudaError_t e;
  if (p.out_bf16) {
    e = cudaFuncSetAttribute(k_spatial_f32<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_spatial_f32<bf16><<<grid, SP_F * 32, smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(k_spatial_f32<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_spatial_f32<float><<<grid, SP_F * 32, smem, st>>>(p);
  }
  return cudaGetLastError();
}

// =================================================================================================
// K0: sliding-window index + globally aligned stride mask for windows cut from ONE video.
// reference: common/dataset/uplifiting_dataset.py:341-394.  Window b is centred on video frame c = centers[b];
// token k looks at frame f = (k - n_tok/2) * s_out + c.  Frames outside [0, T) are padded: "copy" (np.pad mode
// "edge" on the strided sequence) repeats the first / last in-range strided sample, "zeros" yields zeros (src -1).
// mask[b, k] = (f mod s_in == 0) with floor-mod on the UNclamped index.  Integer arithmetic only: bit-exact.
// =================================================================================================
__global__ void k_window_index(const int* __restrict__ centers, int B, int n_tok, int s_out, int s_in, int T,
                               int pad_copy, int* __restrict__ src, uint8_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_tok) return;
  const int b = i / n_tok, k = i - b * n_tok;
  const long long c = centers[b];
  const long long f = (long long)(k - n_tok / 2) * s_out + c;
  long long r = f % s_in;
  if (r < 0) r += s_in;
  mask[i] = r == 0;
  long long sf = f;
  if (f < 0 || f >= T) {
    if (!pad_copy) {
      sf = -1;
    } else {
      // first / last token whose frame is inside the video (the centre token always is)
      const long long f0 = c - (long long)(n_tok / 2) * s_out;           // frame of token 0
      const long long k_min = f0 >= 0 ? 0 : (-f0 + s_out - 1) / s_out;
      const long long k_max = (T - 1 - f0) / s_out;                       // floor, T - 1 - f0 >= 0
      const long long kc = f < 0 ? k_min : (k_max < n_tok - 1 ? k_max : n_tok - 1);
      sf = f0 + kc * s_out;
    }
  }
  src[i] = (int)sf;
}

cudaError_t launch_window_index(const int* centers, int B, int n_tok, int s_out, int s_in, int T, int pad_copy, int* src,
                                uint8_t* mask, cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  const int n = B * n_tok;
  k_window_index<<<(n + 255) / 256, 256, 0, st>>>(centers, B, n_tok, s_out, s_in, T, pad_copy, src, mask);
  return cudaGetLastError();
}

// materialised windows (tests / callers that want the reference's tensors): x[b,k] = video[src[b,k]] (or 0)
__global__ void k_window_copy(const float2* __restrict__ video, const int* __restrict__ src, int n_tokens, int J,
                              float2* __restrict__ x) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n_tokens * J) return;
  const int tok = (int)(i / J), j = (int)(i - (long long)tok * J);
  const int f = src[tok];
  x[i] = f < 0 ? make_float2(0.f, 0.f) : video[(long long)f * J + j];
}

cudaError_t launch_window_copy(const float* video, const int* src, int n_tokens, int J, float* x, cudaStream_t st) {
  if (n_tokens == 0) return cudaSuccess;
  const long long n = (long long)n_tokens * J;
  k_window_copy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(video), src, n_tokens, J,
                                                             reinterpret_cast<float2*>(x));
  return cudaGetLastError();
}

// =================================================================================================
// Evaluation glue on the device.
// k_flip_average: test-time flip augmentation, eval.py:161-180 — pred = (pred + unflip(pred_flipped)) / 2 where
//   unflip negates the x coordinate and gathers joints by AUGM_FLIP_KEYPOINT_ORDER.
// k_keyframe_interp: common/dataset/action_wise_eval.py:76-100 — frames whose index is a multiple of the key-frame
//   stride keep their prediction; frames between two key frames of the same video are interpolated linearly in LIST
//   position; frames after the last key frame of a video copy it.  A video ends where the frame index does not
//   increase.  (A video that starts on a non-key frame makes the reference raise; here such frames keep their value.)
// =================================================================================================
__global__ void k_flip_average(float* __restrict__ a, const float* __restrict__ b, const int* __restrict__ perm,
                               long long n_poses, int J) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_poses * J * 3) return;
  const int c = (int)(i % 3);
  const int j = (int)((i / 3) % J);
  const long long pose = i / (3LL * J);
  float v = b[(pose * J + perm[j]) * 3 + c];
  if (c == 0) v = -v;
  a[i] = (a[i] + v) / 2.f;
}
cudaError_t launch_flip_average(float* a, const float* b, const int* perm, long long n_poses, int J, cudaStream_t st) {
  if (n_poses == 0) return cudaSuccess;
  const long long n = n_poses * J * 3;
  k_flip_average<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, perm, n_poses, J);
  return cudaGetLastError();
}

__global__ void k_keyframe_interp(const float* __restrict__ pred, const int* __restrict__ fidx, int n, int stride, int V,
                                  float* __restrict__ out) {
  const int i = blockIdx.x;                    // one CTA per frame, threads over the V values of a pose
  if (i >= n) return;
  const int f = fidx[i];
  int L = -1, N = -1;
  if (f % stride != 0) {
    // last key frame at or before i inside this video
    for (int k = i; k >= 0; --k) {
      if (fidx[k] % stride == 0) { L = k; break; }
      if (k == 0 || fidx[k] <= fidx[k - 1]) break;         // k starts a video
    }
    // next key frame after i inside this video
    for (int k = i + 1; k < n; ++k) {
      if (fidx[k] <= fidx[k - 1]) break;                   // k starts the next video
      if (fidx[k] % stride == 0) { N = k; break; }
    }
  }
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float r = pred[(long long)i * V + v];
    if (L >= 0 && N >= 0) {
      const float d_left = (float)(i - L), d_right = (float)(N - i), d_sum = d_left + d_right;
      r = pred[(long long)L * V + v] * (d_right / d_sum) + pred[(long long)N * V + v] * (d_left / d_sum);
    } else if (L >= 0) {
      r = pred[(long long)L * V + v];
    }
    out[(long long)i * V + v] = r;
  }
}
cudaError_t launch_keyframe_interp(const float* pred, const int* fidx, int n, int stride, int V, float* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_keyframe_interp<<<n, 64, 0, st>>>(pred, fidx, n, stride, V, out);
  return cudaGetLastError();
}

// =================================================================================================
// K1b: upsampling-token fill for frames without 2-D input: x[row] = token + PE[row % n_tok].
// (net:350-352; rows with mask==1 are written by the 544->384 GEMM epilogue.)  float4, coalesced.
// =================================================================================================
__global__ void k_token_fill(const uint8_t* __restrict__ mask, int rows, int n_tok, int d4,
                             const float4* __restrict__ token, const float4* __restrict__ pe, float4* __restrict__ x) {
  const int per_block = blockDim.x / 32;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * per_block) {
    if (mask[row]) continue;
    const int n = row % n_tok;
    for (int c = lane; c < d4; c += 32) {
      float4 t = token[c], q = pe[(long long)n * d4 + c];
      // m*x + (1-m)*tok with m == 0 is exactly tok (x finite), then + PE
      x[(long long)row * d4 + c] = make_float4(t.x + q.x, t.y + q.y, t.z + q.z, t.w + q.w);
    }
  }
}

cudaError_t launch_token_fill(const uint8_t* mask, int rows, int n_tok, int d, const float* token, const float* pe,
                              float* x, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (d % 4) return cudaErrorInvalidValue;
  int grid = (rows + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  k_token_fill<<<grid, 256, 0, st>>>(mask, rows, n_tok, d / 4, reinterpret_cast<const float4*>(token),
                                     reinterpret_cast<const float4*>(pe), reinterpret_cast<float4*>(x));
  return cudaGetLastError();
}

// =================================================================================================
// K4: LayerNorm over rows of d = 128*V channels, one warp per row, fp32 statistics (two-pass in
// registers), Keras non-fused form.  Optional fused "x += table[row % period]" written back to x
// (the strided blocks' positional encoding, net:126-128).
// =================================================================================================
template <int V, typename TOut>
__global__ void k_layernorm(float* __restrict__ x, int rows, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float eps, const float* __restrict__ table, int period,
                            TOut* __restrict__ y) {
  constexpr int D = V * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 v[V];
  float4* xr = reinterpret_cast<float4*>(x + (long long)row * D);
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = xr[lane + 32 * i];
  if (table) {
    const float4* tr = reinterpret_cast<const float4*>(table + (long long)(row % period) * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 t = tr[lane + 32 * i];
      v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
      xr[lane + 32 * i] = v[i];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  TOut* yr = y + (long long)row * D;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 g = g4[lane + 32 * i], b = b4[lane + 32 * i];
    float o0 = v[i].x * (g.x * rstd) + (b.x - mean * (g.x * rstd));
    float o1 = v[i].y * (g.y * rstd) + (b.y - mean * (g.y * rstd));
    float o2 = v[i].z * (g.z * rstd) + (b.z - mean * (g.z * rstd));
    float o3 = v[i].w * (g.w * rstd) + (b.w - mean * (g.w * rstd));
    if constexpr (sizeof(TOut) == 4) {
      reinterpret_cast<float4*>(yr)[lane + 32 * i] = make_float4(o0, o1, o2, o3);
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1), hi = __floats2bfloat162_rn(o2, o3);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(yr)[lane + 32 * i] = pk;
    }
  }
}

template <int V>
static cudaError_t ln_dispatch(float* x, int rows, const float* g, const float* b, float eps, const float* table,
                               int period, void* y, int y_bf16, cudaStream_t st) {
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  if (y_bf16)
    k_layernorm<V, bf16><<<grid, wpb * 32, 0, st>>>(x, rows, g, b, eps, table, period, (bf16*)y);
  else
    k_layernorm<V, float><<<grid, wpb * 32, 0, st>>>(x, rows, g, b, eps, table, period, (float*)y);
  return cudaGetLastError();
}

cudaError_t launch_layernorm(float* x, int rows, int d, const float* gamma, const float* beta, float eps,
                             const float* table, int period, void* y, int y_bf16, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  switch (d) {
    case 128: return ln_dispatch<1>(x, rows, gamma, beta, eps, table, period, y, y_bf16, st);
    case 256: return ln_dispatch<2>(x, rows, gamma, beta, eps, table, period, y, y_bf16, st);
    case 384: return ln_dispatch<3>(x, rows, gamma, beta, eps, table, period, y, y_bf16, st);
    case 512: return ln_dispatch<4>(x, rows, gamma, beta, eps, table, period, y, y_bf16, st);
    case 768: return ln_dispatch<6>(x, rows, gamma, beta, eps, table, period, y, y_bf16, st);
    case 1024: return ln_dispatch<8>(x, rows, gamma, beta, eps, table, period, y, y_bf16, st);
    default: return cudaErrorInvalidValue;
  }
}

// =================================================================================================
// bf16-resident residual stream (bf16 schedule): streaming kernels with X stored as bf16 between kernels
// (statistics, adds and the LayerNorm arithmetic stay fp32), one warp per row:
//   v = xsrc[map(r)] (+ bf16 update[r])      (update = strided-conv / projection output)
//   xcast[r] = bf16(v)                       (optional: operand of the regression heads)
//   v += table[r % period]                   (optional: positional encoding)
//   xdst[r] = v ; y[r] = bf16(LN(v))         (both optional)
// xsrc may be a different, longer sequence than xdst (strided identity path x[:, c0::s], net:146-152).
// =================================================================================================
__device__ __forceinline__ float4 ld_bf16x4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}
__device__ __forceinline__ void st_bf16x4(bf16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

__global__ void k_token_fill_bx(const uint8_t* __restrict__ mask, int rows, int n_tok, int d4,
                                const float4* __restrict__ token, const float4* __restrict__ pe, bf16* __restrict__ x,
                                float* __restrict__ stats, int slots) {
  const int per_block = blockDim.x / 32;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * per_block) {
    if (mask[row]) continue;
    const int n = row % n_tok;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 t = token[c];
      const float4 q = pe ? pe[(long long)n * d4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);   // pe == null: added downstream
      const float a = t.x + q.x, b = t.y + q.y, cc = t.z + q.z, dd = t.w + q.w;
      st_bf16x4(x + ((long long)row * d4 + c) * 4, a, b, cc, dd);
      s1 += (a + b) + (cc + dd);
      s2 += (a * a + b * b) + (cc * cc + dd * dd);
    }
    if (stats) {      // row partials for a GEMM that folds the next LayerNorm: everything in slot 0
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane < slots)
        reinterpret_cast<float2*>(stats)[(long long)row * slots + lane] = lane == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
    }
  }
}
cudaError_t launch_token_fill_bx(const uint8_t* mask, int rows, int n_tok, int d, const float* token, const float* pe,
                                 bf16* x, cudaStream_t st, float* stats, int slots) {
  if (rows == 0) return cudaSuccess;
  if (d % 4 || (stats && (slots < 1 || slots > 32))) return cudaErrorInvalidValue;
  int grid = (rows + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  k_token_fill_bx<<<grid, 256, 0, st>>>(mask, rows, n_tok, d / 4, reinterpret_cast<const float4*>(token),
                                        reinterpret_cast<const float4*>(pe), x, stats, slots);
  return cudaGetLastError();
}

// v = xsrc[map(r)] (+ upd[r]) ; xcast[r] = bf16(v) (opt) ; v += table[r % period] (opt) ; xdst[r] = bf16(v) (opt) ;
// y[r] = bf16(LN(v)) (opt).  upd == null: plain LayerNorm of the stream (first temporal block).
template <int V>
__global__ void k_residual_ln_bx(const bf16* __restrict__ xsrc, RowMap smap, const bf16* __restrict__ upd,
                                 bf16* __restrict__ xdst, int rows, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, const float* __restrict__ table, int period,
                                 bf16* __restrict__ y, bf16* __restrict__ xcast, float* __restrict__ stats,
                                 int slots) {
  constexpr int D = V * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long srow = map_row(smap, row);
  const bf16* xr = xsrc + srow * D;
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = ld_bf16x4(xr + (lane + 32 * i) * 4);
  if (upd) {
    const bf16* ur = upd + (long long)row * D;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 u = ld_bf16x4(ur + (lane + 32 * i) * 4);
      v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
    }
  }
  if (xcast) {
#pragma unroll
    for (int i = 0; i < V; ++i) st_bf16x4(xcast + (long long)row * D + (lane + 32 * i) * 4, v[i].x, v[i].y, v[i].z, v[i].w);
  }
  if (table) {
    const float4* tr = reinterpret_cast<const float4*>(table + (long long)(row % period) * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 t = tr[lane + 32 * i];
      v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
    }
  }
  if (xdst) {
#pragma unroll
    for (int i = 0; i < V; ++i) st_bf16x4(xdst + (long long)row * D + (lane + 32 * i) * 4, v[i].x, v[i].y, v[i].z, v[i].w);
  }
  if (stats) {
    // Row statistics for the GEMMs that fold the next LayerNorm (EPI_LNFOLD): (sum, sum of squares) of the row AS
    // STORED (bf16-rounded), in slot 0 of the row's partials; the other slots are cleared.
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float a = __bfloat162float(__float2bfloat16_rn(v[i].x)), b = __bfloat162float(__float2bfloat16_rn(v[i].y));
      const float c = __bfloat162float(__float2bfloat16_rn(v[i].z)), d = __bfloat162float(__float2bfloat16_rn(v[i].w));
      s1 += (a + b) + (c + d);
      s2 += (a * a + b * b) + (c * c + d * d);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane < slots)
      reinterpret_cast<float2*>(stats)[(long long)row * slots + lane] = lane == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
  }
  if (!y) return;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 g = g4[lane + 32 * i], b = b4[lane + 32 * i];
    st_bf16x4(y + (long long)row * D + (lane + 32 * i) * 4,
              v[i].x * (g.x * rstd) + (b.x - mean * (g.x * rstd)), v[i].y * (g.y * rstd) + (b.y - mean * (g.y * rstd)),
              v[i].z * (g.z * rstd) + (b.z - mean * (g.z * rstd)), v[i].w * (g.w * rstd) + (b.w - mean * (g.w * rstd)));
  }
}

cudaError_t launch_residual_ln_bx(const bf16* xsrc, const RowMap& smap, const bf16* upd, bf16* xdst, int rows, int d,
                                  const float* gamma, const float* beta, float eps, const float* table, int period,
                                  bf16* y, bf16* xcast, cudaStream_t st, float* stats, int slots) {
  if (rows == 0) return cudaSuccess;
  if (stats && (slots < 1 || slots > 32)) return cudaErrorInvalidValue;
  const int wpb = 8, grid = (rows + wpb - 1) / wpb;
#define UU_RLN_CASE(VV)                                                                                           \
  case VV * 128:                                                                                                  \
    k_residual_ln_bx<VV><<<grid, wpb * 32, 0, st>>>(xsrc, smap, upd, xdst, rows, gamma, beta, eps, table, period, \
                                                    y, xcast, stats, slots);                                      \
    break;
  switch (d) {
    UU_RLN_CASE(1) UU_RLN_CASE(2) UU_RLN_CASE(3) UU_RLN_CASE(4) UU_RLN_CASE(6) UU_RLN_CASE(8)
    default: return cudaErrorInvalidValue;
  }
#undef UU_RLN_CASE
  return cudaGetLastError();
}

// =================================================================================================
// Virtual-camera projection for on-device training-data synthesis (uplifiting_dataset.py:669-761):
// world -> camera with the inverse of the unit quaternion cam[0:4] after subtracting the translation cam[4:7]
// (tf_world_to_cam / tf_qrot), then the Human3.6M projection with radial and tangential distortion
// (tf_project_to_2d: x/z clamped to [-1, 1]).  One thread per point; cams is (B, 18), points_per_sample points share
// a camera.
// =================================================================================================
__global__ void k_world_to_cam_2d(const float* __restrict__ x3d, const float* __restrict__ cams, long long n_points,
                                  int points_per_sample, float* __restrict__ cam3d, float* __restrict__ p2d) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const float* cam = cams + (i / points_per_sample) * 18;
  const float w = cam[0], qx = -cam[1], qy = -cam[2], qz = -cam[3];          // conjugate = inverse rotation
  const float vx = x3d[i * 3] - cam[4], vy = x3d[i * 3 + 1] - cam[5], vz = x3d[i * 3 + 2] - cam[6];
  const float ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;            // q x v
  const float wx = qy * uz - qz * uy, wy = qz * ux - qx * uz, wz = qx * uy - qy * ux;            // q x (q x v)
  const float cx = vx + 2.f * (w * ux + wx), cy = vy + 2.f * (w * uy + wy), cz = vz + 2.f * (w * uz + wz);
  if (cam3d) { cam3d[i * 3] = cx; cam3d[i * 3 + 1] = cy; cam3d[i * 3 + 2] = cz; }
  if (p2d) {
    const float* in = cam + 7;     // res (2), focal (2), centre (2), radial (3), tangential (2)
    const float xx = fminf(fmaxf(cx / cz, -1.f), 1.f), yy = fminf(fmaxf(cy / cz, -1.f), 1.f);
    const float r2 = xx * xx + yy * yy;
    const float radial = 1.f + in[6] * r2 + in[7] * r2 * r2 + in[8] * r2 * r2 * r2;
    const float tan = in[9] * xx + in[10] * yy;
    p2d[i * 2] = in[2] * (xx * (radial + tan) + in[9] * r2) + in[4];
    p2d[i * 2 + 1] = in[3] * (yy * (radial + tan) + in[10] * r2) + in[5];
  }
}
cudaError_t launch_world_to_cam_2d(const float* x3d, const float* cams, long long n_points, int points_per_sample,
                                   float* cam3d, float* p2d, cudaStream_t st) {
  if (n_points == 0) return cudaSuccess;
  k_world_to_cam_2d<<<(unsigned)((n_points + 255) / 256), 256, 0, st>>>(x3d, cams, n_points, points_per_sample, cam3d, p2d);
  return cudaGetLastError();
}

// =================================================================================================
// Evaluation metrics on the device: root-aligned MPJPE and N-MPJPE (root alignment + per-pose
// optimal scale s = <p, g> / <p, p> over the valid joints), common/dataset/metrics.py:13-81, :120-133.
// One warp per pose, lane == joint (J <= 32).  per-joint outputs are -1 where the ground truth is invalid
// (normalize=False form); sums[pose] = (sum of valid MPJPE distances, sum of valid N-MPJPE distances, valid count),
// reduced in double by one block in pose order (deterministic).
// =================================================================================================
__global__ void k_pose_metrics(const float* __restrict__ pred, const float* __restrict__ gt, int n, int J, int root,
                               float* __restrict__ jpe, float* __restrict__ njpe, float* __restrict__ sums) {
  const int lane = threadIdx.x & 31;
  const int pose = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pose >= n) return;
  const float* p = pred + (long long)pose * J * 3;
  const float* g = gt + (long long)pose * J * 4;
  const bool in = lane < J;
  const float pr0 = p[root * 3], pr1 = p[root * 3 + 1], pr2 = p[root * 3 + 2];
  const float gr0 = g[root * 4], gr1 = g[root * 4 + 1], gr2 = g[root * 4 + 2];
  float px = 0.f, py = 0.f, pz = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
  bool valid = false;
  if (in) {
    px = p[lane * 3] - pr0; py = p[lane * 3 + 1] - pr1; pz = p[lane * 3 + 2] - pr2;
    gx = g[lane * 4] - gr0; gy = g[lane * 4 + 1] - gr1; gz = g[lane * 4 + 2] - gr2;
    valid = g[lane * 4 + 3] > 0.f;
  }
  const float d0 = sqrtf((px - gx) * (px - gx) + (py - gy) * (py - gy) + (pz - gz) * (pz - gz));
  const float nom = warp_sum(valid ? px * gx + py * gy + pz * gz : 0.f);
  const float den = warp_sum(valid ? px * px + py * py + pz * pz : 0.f);
  const float sc = nom / den;
  const float d1 = sqrtf((sc * px - gx) * (sc * px - gx) + (sc * py - gy) * (sc * py - gy) + (sc * pz - gz) * (sc * pz - gz));
  if (in) {
    if (jpe) jpe[(long long)pose * J + lane] = valid ? d0 : -1.f;
    if (njpe) njpe[(long long)pose * J + lane] = valid ? d1 : -1.f;
  }
  const float s0 = warp_sum(valid ? d0 : 0.f), s1 = warp_sum(valid ? d1 : 0.f), cnt = warp_sum(valid ? 1.f : 0.f);
  if (lane == 0) { sums[pose * 3] = s0; sums[pose * 3 + 1] = s1; sums[pose * 3 + 2] = cnt; }
}
__global__ void k_pose_metrics_reduce(const float* __restrict__ sums, int n, double* __restrict__ out) {
  __shared__ double sh[3][256];
  double a = 0, b = 0, c = 0;
  // contiguous slices per thread, then a fixed-order tree: deterministic
  const int per = (n + 255) / 256, lo = threadIdx.x * per, hi = min(n, lo + per);
  for (int i = lo; i < hi; ++i) { a += sums[i * 3]; b += sums[i * 3 + 1]; c += sums[i * 3 + 2]; }
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b; sh[2][threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh[0][0] / sh[2][0]; out[1] = sh[1][0] / sh[2][0]; out[2] = sh[2][0]; }
}
cudaError_t launch_pose_metrics(const float* pred, const float* gt, int n, int J, int root, float* jpe, float* njpe,
                                float* sums, double* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_pose_metrics<<<(n + 7) / 8, 256, 0, st>>>(pred, gt, n, J, root, jpe, njpe, sums);
  k_pose_metrics_reduce<<<1, 256, 0, st>>>(sums, n, out);
  return cudaGetLastError();
}

// =================================================================================================
// K5 (CUDA-core version): softmax attention per (window, head), S <= 128 keys, thread == query.
// Literal reference arithmetic: logits = q.k / sqrt(dh) + keymask * -1e9 in fp32 (vit:117-123), so
// an all-masked window reproduces the reference's uniform attention.
// =================================================================================================
__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const bf16* p) { return __bfloat162float(*p); }

template <int DH, typename T>
__global__ void __launch_bounds__(128) k_attention(const T* __restrict__ qkv, int S, int heads,
                                                   const uint8_t* __restrict__ mask, int mask_stride,
                                                   T* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                    // [S][DH]
  float* Vs = Ks + S * DH;           // [S][DH]
  float* Km = Vs + S * DH;           // [S] additive key-mask term
  float* Sc = Km + ((S + 3) & ~3);   // [S][128] scores, thread-major columns
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int d = heads * DH;
  const long long row0 = (long long)b * S;
  for (int i = tid; i < S * DH; i += 128) {
    int j = i / DH, c = i - j * DH;
    const T* r = qkv + (row0 + j) * 3 * d + h * DH + c;
    Ks[i] = ld_as_float(r + d);
    Vs[i] = ld_as_float(r + 2 * d);
  }
  for (int j = tid; j < S; j += 128)
    Km[j] = mask ? (1.0f - (mask[(long long)b * mask_stride + j] ? 1.0f : 0.0f)) * -1e9f : 0.0f;
  __syncthreads();
  if (tid >= S) return;
  float q[DH];
  const T* qr = qkv + (row0 + tid) * 3 * d + h * DH;
#pragma unroll
  for (int c = 0; c < DH; ++c) q[c] = ld_as_float(qr + c);
  const float scale = 1.0f / sqrtf((float)DH);
  float mx = -INFINITY;
  for (int j = 0; j < S; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(Ks + j * DH);
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      float4 k4 = kr[c];
      a = fmaf(q[4 * c], k4.x, a); a = fmaf(q[4 * c + 1], k4.y, a);
      a = fmaf(q[4 * c + 2], k4.z, a); a = fmaf(q[4 * c + 3], k4.w, a);
    }
    a = a * scale + Km[j];
    Sc[j * 128 + tid] = a;
    mx = fmaxf(mx, a);
  }
  float o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = 0.f;
  float sum = 0.f;
  for (int j = 0; j < S; ++j) {
    const float pj = expf(Sc[j * 128 + tid] - mx);
    sum += pj;
    const float4* vr = reinterpret_cast<const float4*>(Vs + j * DH);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      float4 v4 = vr[c];
      o[4 * c] = fmaf(pj, v4.x, o[4 * c]); o[4 * c + 1] = fmaf(pj, v4.y, o[4 * c + 1]);
      o[4 * c + 2] = fmaf(pj, v4.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pj, v4.w, o[4 * c + 3]);
    }
  }
  const float inv = 1.f / sum;
  T* orow = out + (row0 + tid) * d + h * DH;
#pragma unroll
  for (int c = 0; c < DH; ++c) store_out(orow + c, o[c] * inv);
}

template <int DH, typename T>
static cudaError_t attn_dispatch(const void* qkv, int B, int S, int heads, const uint8_t* mask, int mask_stride,
                                 void* out, cudaStream_t st) {
  size_t smem = sizeof(float) * (2 * S * DH + ((S + 3) & ~3) + S * 128);
  cudaError_t e = cudaFuncSetAttribute(k_attention<DH, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_attention<DH, T><<<dim3(B, heads), 128, smem, st>>>((const T*)qkv, S, heads, mask, mask_stride, (T*)out);
  return cudaGetLastError();
}

cudaError_t launch_attention(const void* qkv, int is_bf16, int B, int S, int heads, int dh, const uint8_t* mask,
                             int mask_stride, void* out, cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  if (S > 128 || S < 1) return cudaErrorInvalidValue;
  if (is_bf16)   // the bf16 path has its own tensor-core kernel (attention_tc.cu)
    return launch_attention_tc((const bf16*)qkv, B, S, heads, dh, mask, mask_stride, (bf16*)out, st);
#define UU_ATTN_CASE(DHV) \
  case DHV:               \
    return attn_dispatch<DHV, float>(qkv, B, S, heads, mask, mask_stride, out, st);
  switch (dh) {
    UU_ATTN_CASE(4)
    UU_ATTN_CASE(16)
    UU_ATTN_CASE(32)
    UU_ATTN_CASE(48)
    UU_ATTN_CASE(64)
    default: return cudaErrorInvalidValue;
  }
#undef UU_ATTN_CASE
}

// =================================================================================================
// K3 (fp32): CUDA-core GEMM with the shared fused epilogue, 64x64x16 tiles, 4x4 micro-tiles.
// This is the exact-precision path (<= 1e-4 against the fp32 oracle); throughput work goes through
// the tcgen05 kernel.
// =================================================================================================
constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

template <typename TA, typename TC>
__global__ void __launch_bounds__(256) k_gemm_simt(const TA* __restrict__ A, long long lda,
                                                   const float* __restrict__ W, int M, int N, int K, Epilogue epi,
                                                   TC* __restrict__ C, long long ldc) {
  __shared__ __align__(16) float As[GB_K][GB_M + 4];
  __shared__ __align__(16) float Bs[GB_K][GB_N];
  const int m_eff = epi.m_dev ? min(M, *epi.m_dev) : M;
  const int row0 = blockIdx.x * GB_M, col0 = blockIdx.y * GB_N;
  if (row0 >= m_eff) return;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += GB_K) {
    {  // A tile (K % 4 == 0 and lda % 4 == 0 are checked by the launcher)
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int r = row0 + a_row, k = k0 + a_k;
      if (r < m_eff && k < K) {
        const TA* src = A + (long long)r * lda + k;
        if constexpr (sizeof(TA) == 4) {
          float4 t = *reinterpret_cast<const float4*>(src);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
          uint2 t = *reinterpret_cast<const uint2*>(src);
          __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&t.x), hi = *reinterpret_cast<__nv_bfloat162*>(&t.y);
          v[0] = __low2float(lo); v[1] = __high2float(lo); v[2] = __low2float(hi); v[3] = __high2float(hi);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[a_k + i][a_row] = v[i];
    }
    {  // W tile
      const int k = k0 + b_k;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = col0 + b_n + i;
        Bs[b_k][b_n + i] = (k < K && n < N) ? W[(long long)k * N + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GB_K; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= m_eff) continue;
    const EpiRow er = epi_row(epi, r);
    if (er.crow < 0) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c < N) store_out(C + er.crow * ldc + c, epi_value(epi, er, acc[i][j], c, N));
    }
  }
}

cudaError_t launch_gemm_simt(const void* A, int a_bf16, long long lda, const float* W, int M, int N, int K,
                             const Epilogue& epi, void* C, int c_bf16, long long ldc, cudaStream_t st) {
  if (M == 0) return cudaSuccess;
  if ((K % 4) || (lda % 4)) return cudaErrorInvalidValue;
  dim3 grid((M + GB_M - 1) / GB_M, (N + GB_N - 1) / GB_N);
  if (a_bf16) {
    if (c_bf16) k_gemm_simt<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)A, lda, W, M, N, K, epi, (bf16*)C, ldc);
    else k_gemm_simt<bf16, float><<<grid, 256, 0, st>>>((const bf16*)A, lda, W, M, N, K, epi, (float*)C, ldc);
  } else {
    if (c_bf16) k_gemm_simt<float, bf16><<<grid, 256, 0, st>>>((const float*)A, lda, W, M, N, K, epi, (bf16*)C, ldc);
    else k_gemm_simt<float, float><<<grid, 256, 0, st>>>((const float*)A, lda, W, M, N, K, epi, (float*)C, ldc);
  }
  return cudaGetLastError();
}

}  // namespace uu
