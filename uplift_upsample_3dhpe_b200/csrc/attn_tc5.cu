// T3 (bf16 schedule): fused softmax attention on the 5th-generation tensor cores, fed by TMA.
//   reference: vit:99-130 (q k^T / sqrt(dh) + (1 - mask) * -1e9, softmax, . v, heads merged).
// One persistent CTA per SM walks the windows.  A window's q | k | v rows arrive as two "head-quad" stages (4 heads =
// 192 columns = three 64-column TMA boxes for q and for k, one box per head for v, 128B swizzle, S rounded up to 16 rows),
// double-buffered so the next 100 KB are in flight while a stage is being consumed: the kernel is HBM-bound (2304 B in, 768 B out per token).
// Per head, everything stays on chip:
//   S = Q_h K_h^T      tcgen05.mma M = 128 (query rows, S valid), N = SP (keys), K = 48: both operands are 16-column
//                      slices of the swizzled boxes (descriptor start address inside the 128-byte row), accumulator in TMEM;
//   softmax            eight warps (two per TMEM lane quarter, each alternating between two of the four heads in flight so
//                      that the latency of one head's P V hides behind the next head's softmax): tcgen05.ld the row in
//                      16-key chunks (pass 1: row maximum, pass 2: exp2-domain softmax with the reference's literal -1e9
//                      key term), probabilities rounded to bf16 and written back over the scores with tcgen05.st (two per
//                      32-bit column);
//   O_h = P V_h        tcgen05.mma with A = P from TENSOR MEMORY and B = V_h straight from the row-major box through an
//                      MN-major descriptor (one N = 48 MMA per 16 keys; V arrives as one box per head) — no transposed copy of V, P never
//                      touches shared memory;
//   O_h / rowsum       tcgen05.ld, normalised in fp32, bf16, 96 contiguous bytes per token to global memory.
// Four score and four output accumulators (head index mod 4): the four heads of a stage are in flight at once, each on
// its own warp of every lane quarter, and the tensor pipe issues the next stage's scores for a buffer as soon as its
// P V has been issued.
#include "tc_ptx.cuh"

namespace uu {

constexpr int A5_DH = 48, A5_HEADS = 8, A5_D = A5_DH * A5_HEADS, A5_QUAD = 4;
constexpr int A5_THREADS = 320;          // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 softmax / output
constexpr int A5_MAX_SP = 80;
constexpr int A5_TM_S = 0, A5_TM_O = 320;        // TMEM columns: four score buffers of 80, four output buffers of 48

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem descriptor]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// bf16 operand in a [rows][64 columns] 128B-swizzled box, read MN-major (the row index is the contraction index):
// 8 rows 128 B apart form an atom, atoms 1024 B apart (SBO); LBO (next 64-column group) is never used by a 16-column slice
__device__ __forceinline__ uint64_t make_sw128_mn_bf16_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(int M, int N) {      // A K-major (TMEM), B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float ex2_a5(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Attn5Args {
  int B, S, SP;              // windows, tokens per window, tokens rounded up to 16
  const uint8_t* mask;       // [B, mask_stride] key mask or null
  int mask_stride;
  bf16* out;                 // [B * S, 384]
};

__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

__global__ void __launch_bounds__(A5_THREADS, 1) k_attention_tc5(const __grid_constant__ CUtensorMap map_qkv, Attn5Args a) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int SP = a.SP, S = a.S;
  const int box_bytes = SP * 128;                 // one 64-column box
  const int stage_bytes = 10 * box_bytes;         // q | k of four heads (3 + 3 boxes), v as one 64-column box PER HEAD
  float* s_km = reinterpret_cast<float*>(smem + 2 * stage_bytes + 8192);        // (+8 KB: M = 128 reads past the last box)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_km + 2 * A5_MAX_SP);
  uint64_t* stage_full = bars;          // [2]
  uint64_t* stage_empty = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;          // [4] scores of head (index mod 4) ready
  uint64_t* p_full = bars + 8;          // [4] probabilities written (one warp per lane quarter)
  uint64_t* o_full = bars + 12;         // [4] P V done
  uint64_t* o_empty = bars + 16;        // [4] output accumulator read (4 warps)
  uint64_t* km_full = bars + 20;        // [2] key term of a window written (per window parity)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_win = a.B > (int)blockIdx.x ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
    for (int i = 0; i < 2; ++i) { mbar_init(stage_full + i, 1); mbar_init(stage_empty + i, 1); mbar_init(km_full + i, 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(s_full + i, 1); mbar_init(p_full + i, 4);      // (4: one warp per lane quarter)
      mbar_init(o_full + i, 1); mbar_init(o_empty + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---------------- TMA producer: per window two head-quad stages of nine boxes; the window's key term ----------------
    // (km[parity][j]: 0 for a kept key, -1e9 log2 e for a masked one, -inf for the padding keys j >= S.  Buffer `parity`
    // is rewritten for window i + 2 only after every head of window i has been issued its P V, which the stage_empty wait
    // of that window's first stage implies.)
    const float LOG2E = 1.4426950408889634f;
    uint32_t n = 0;
    for (int i = 0; i < n_win; ++i) {
      const int w = (int)blockIdx.x + i * (int)gridDim.x;
      for (int quad = 0; quad < 2; ++quad, ++n) {
        const int st = n & 1;
        mbar_wait(stage_empty + st, ((n >> 1) & 1) ^ 1);
        if (quad == 0) {
          float* km = s_km + (i & 1) * A5_MAX_SP;
          for (int j = lane; j < SP; j += 32)
            km[j] = j >= S ? -INFINITY : ((a.mask && !a.mask[(long long)w * a.mask_stride + j]) ? -1e9f * LOG2E : 0.f);
          __syncwarp();
          if (lane == 0) mbar_arrive(km_full + (i & 1));
        }
        if (lane == 0) {
          mbar_expect_tx(stage_full + st, stage_bytes);
          uint8_t* dst = smem + st * stage_bytes;
          for (int op = 0; op < 2; ++op)
            for (int bx = 0; bx < 3; ++bx)
              tma_load_2d(dst + (op * 3 + bx) * box_bytes, &map_qkv, stage_full + st, op * A5_D + quad * 192 + bx * 64, w * S);
          // v: box h starts at the head's first column (columns 48 .. 63 of the box belong to the next head and are not
          // read), so that P V is ONE N = 48 MMA per 16 keys instead of three N = 16 slices
          for (int hq = 0; hq < A5_QUAD; ++hq)
            tma_load_2d(dst + (6 + hq) * box_bytes, &map_qkv, stage_full + st, 2 * A5_D + quad * 192 + hq * A5_DH, w * S);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc_s = make_idesc_bf16(128, SP);
    constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, A5_DH);
    const int total = n_win * A5_HEADS;                 // heads this CTA processes, in order
    auto issue_qk = [&](int hc) {                       // S[hc % 4] = Q_h K_h^T   (hc-th head overall)
      const int sidx = hc >> 2, st = sidx & 1, hq = hc & 3;
      if (hq == 0) {
        mbar_wait(stage_full + st, (sidx >> 1) & 1);
        tcgen05_fence_after();
      }
      if (elect_one()) {
        const uint32_t qbase = smem_u32(smem + st * stage_bytes), kbase = qbase + 3 * box_bytes;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int c = hq * A5_DH + 16 * i;
          const uint32_t off = (c >> 6) * box_bytes + (c & 63) * 2;
          umma_bf16(tmem_base + A5_TM_S + hq * 80, make_sw128_desc(qbase + off), make_sw128_desc(kbase + off), idesc_s, i != 0);
        }
        umma_commit(s_full + hq);
      }
      __syncwarp();
    };
    for (int hc = 0; hc < total && hc < 4; ++hc) issue_qk(hc);
    for (int hc = 0; hc < total; ++hc) {
      const int sidx = hc >> 2, st = sidx & 1, hq = hc & 3;
      const uint32_t use = (uint32_t)sidx & 1;           // parity of this buffer's use
      mbar_wait(p_full + hq, use);
      mbar_wait(o_empty + hq, use ^ 1);
      tcgen05_fence_after();
      if (elect_one()) {                                // O[hq] = P[hq] . V_h: one N = 48 MMA per 16 keys
        const uint64_t vd = make_sw128_mn_bf16_desc(smem_u32(smem + st * stage_bytes + (6 + hq) * box_bytes));
        for (int t = 0; t < SP / 16; ++t)               // 16 rows of the box = 2048 B, 8 TMEM columns of P
          umma_bf16_ts(tmem_base + A5_TM_O + hq * 48, tmem_base + A5_TM_S + hq * 80 + 8 * t, vd + (uint64_t)(t * (2048 >> 4)),
                       idesc_o, t != 0);
        umma_commit(o_full + hq);
        if (hq == 3) umma_commit(stage_empty + st);     // every MMA that reads this stage has been issued
      }
      __syncwarp();
      if (hc + 4 < total) issue_qk(hc + 4);             // this buffer's next head (the next stage)
    }
  } else {
    // ---------------- softmax / output warps: lane quarter q, head slots par and par + 2 of every head-quad ----------------
    // A warp alternates between two accumulator slots: while the tensor pipe runs P V of slot par (issue, five MMAs,
    // commit: ~1000 cycles of latency) the warp computes the softmax of slot par + 2, then drains both outputs.
    const int e = warp - 2, q = warp & 3, par = e >> 2;
    const int row = q * 32 + lane;
    const bool live = q * 32 < S;                        // a lane quarter made of padding rows only does no arithmetic
    const float LOG2E = 1.4426950408889634f;
    const float scale = LOG2E / sqrtf((float)A5_DH);
    const uint32_t t_q = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t km_s = smem_u32(s_km);
    const int nch = SP / 16;
    for (int i = 0; i < n_win; ++i) {
      const int w = (int)blockIdx.x + i * (int)gridDim.x;
      mbar_wait(km_full + (i & 1), (i >> 1) & 1);
      const uint32_t km_a = km_s + (i & 1) * (A5_MAX_SP * 4);
      for (int quad = 0; quad < 2; ++quad) {
        const uint32_t use = (uint32_t)(2 * i + quad) & 1;
        float lsum[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int hs = par + 2 * k;
          const uint32_t t_s = t_q + A5_TM_S + hs * 80;
          mbar_wait(s_full + hs, use);
          tcgen05_fence_after();
          float l = 0.f;
          if (live) {
            // pass 1: row maximum of the logits (exp2 domain).  (Keeping the 80 logits of a row in registers for a single
            // pass over tensor memory measured slower: 330 vs 265 us per launch, register pressure.)
            float m = -INFINITY;
            for (int c = 0; c < nch; ++c) {
              uint32_t v[16];
              tmem_ld_x16(t_s + 16 * c, v);
              tmem_wait_ld();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 k4 = lds_f4(km_a + (16 * c + 4 * g) * 4);
                m = fmaxf(m, fmaxf(fmaxf(fmaf(__uint_as_float(v[4 * g]), scale, k4.x), fmaf(__uint_as_float(v[4 * g + 1]), scale, k4.y)),
                                   fmaxf(fmaf(__uint_as_float(v[4 * g + 2]), scale, k4.z), fmaf(__uint_as_float(v[4 * g + 3]), scale, k4.w))));
              }
            }
            // pass 2: probabilities, row sum, bf16 pairs written over the scores (8 columns per 16 keys)
            for (int c = 0; c < nch; ++c) {
              uint32_t v[16], pk[8];
              tmem_ld_x16(t_s + 16 * c, v);
              tmem_wait_ld();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 k4 = lds_f4(km_a + (16 * c + 4 * g) * 4);
                const float p0 = ex2_a5(fmaf(__uint_as_float(v[4 * g]), scale, k4.x) - m);
                const float p1 = ex2_a5(fmaf(__uint_as_float(v[4 * g + 1]), scale, k4.y) - m);
                const float p2 = ex2_a5(fmaf(__uint_as_float(v[4 * g + 2]), scale, k4.z) - m);
                const float p3 = ex2_a5(fmaf(__uint_as_float(v[4 * g + 3]), scale, k4.w) - m);
                l += (p0 + p1) + (p2 + p3);
                __nv_bfloat162 a01 = __floats2bfloat162_rn(p0, p1), a23 = __floats2bfloat162_rn(p2, p3);
                pk[2 * g] = *reinterpret_cast<uint32_t*>(&a01);
                pk[2 * g + 1] = *reinterpret_cast<uint32_t*>(&a23);
              }
              tmem_st_x8(t_s + 8 * c, pk);       // (columns 8c .. 8c+7 hold scores of keys this thread has already consumed)
            }
            tmem_wait_st();
          }
          lsum[k] = l;
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full + hs);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int hs = par + 2 * k;
          const uint32_t t_o = t_q + A5_TM_O + hs * 48;
          mbar_wait(o_full + hs, use);
          tcgen05_fence_after();
          uint32_t o0[32], o1[16];
          if (live) {
            tmem_ld_32x32b_x32(t_o, o0);
            tmem_ld_x16(t_o + 32, o1);
            tmem_wait_ld();
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty + hs);
          if (live && row < S) {
            const float inv = 1.f / lsum[k];
            uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)w * S + row) * A5_D + (quad * A5_QUAD + hs) * A5_DH);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              uint32_t wv[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int kk = 8 * c + 2 * j;
                const float x0 = __uint_as_float(kk < 32 ? o0[kk] : o1[kk - 32]) * inv;
                const float x1 = __uint_as_float(kk + 1 < 32 ? o0[kk + 1] : o1[kk + 1 - 32]) * inv;
                __nv_bfloat162 pb = __floats2bfloat162_rn(x0, x1);
                wv[j] = *reinterpret_cast<uint32_t*>(&pb);
              }
              dst[c] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

bool attention_tc5_ok(int S, int heads, int dh) { return heads == A5_HEADS && dh == A5_DH && S >= 1 && S <= A5_MAX_SP; }

cudaError_t launch_attention_tc5(const bf16* qkv, int B, int S, const uint8_t* mask, int mask_stride, bf16* out, int num_sms,
                                 cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  const int SP = (S + 15) / 16 * 16;
  CUtensorMap map;
  if (encode_2d(&map, qkv, 3 * A5_D, (uint64_t)B * S, 3 * A5_D, 64, (uint32_t)SP)) return cudaErrorInvalidValue;
  const int smem = 2 * 10 * SP * 128 + 8192 + 2 * A5_MAX_SP * 4 + 256 + 1024;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(k_attention_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 10 * A5_MAX_SP * 128 + 8192 + 2 * A5_MAX_SP * 4 + 256 + 1024);
    if (e != cudaSuccess) return e;
    attr_smem = 2 * 10 * A5_MAX_SP * 128 + 8192 + 2 * A5_MAX_SP * 4 + 256 + 1024;
  }
  Attn5Args a;
  a.B = B; a.S = S; a.SP = SP; a.mask = mask; a.mask_stride = mask_stride; a.out = out;
  return launch_pdl(k_attention_tc5, dim3(B < num_sms ? B : num_sms), dim3(A5_THREADS), (size_t)smem, st, map, a);
}

}  // namespace uu
