// T3 (bf16 schedule): fused softmax attention on the 5th-generation tensor cores, fed by TMA.
//   reference: vit:99-130 (q k^T / sqrt(dh) + (1 - mask) * -1e9, softmax, . v, heads merged).
// One persistent CTA per SM walks the windows.  A window's q | k | v rows arrive as two "head-quad" stages (4 heads =
// 192 columns = three 64-column TMA boxes per operand, 128B swizzle, S rounded up to 16 rows), double-buffered so the next
// 90 KB are in flight while a stage is being consumed: the kernel is HBM-bound (2304 B in, 768 B out per token).
// Per head, everything stays on chip:
//   S = Q_h K_h^T      tcgen05.mma M = 128 (query rows, S valid), N = SP (keys), K = 48: both operands are 16-column
//                      slices of the swizzled boxes (descriptor start address inside the 128-byte row), accumulator in TMEM;
//   softmax            eight warps (two per TMEM lane quarter, alternating heads): tcgen05.ld the row, exp2-domain softmax
//                      with the reference's literal -1e9 key term, probabilities rounded to bf16 and written back over
//                      the scores with tcgen05.st (two per 32-bit column);
//   O_h = P V_h        tcgen05.mma with A = P from TENSOR MEMORY and B = V_h straight from the row-major box through an
//                      MN-major descriptor (three N = 16 slices per 16-key step) — no transposed copy of V, P never
//                      touches shared memory;
//   O_h / rowsum       tcgen05.ld, normalised in fp32, bf16, 96 contiguous bytes per token to global memory.
// Two score and two output accumulators (even / odd heads) let the tensor pipe run head h + 1 while the softmax warps
// work on head h.
#include "tc_ptx.cuh"

namespace uu {

constexpr int A5_DH = 48, A5_HEADS = 8, A5_D = A5_DH * A5_HEADS, A5_QUAD = 4;
constexpr int A5_THREADS = 320;          // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 softmax / output
constexpr int A5_MAX_SP = 80;
constexpr int A5_TM_S = 0, A5_TM_O = 256;        // TMEM columns: S buffers at 0 / 128, O buffers at 256 / 320

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

__global__ void __launch_bounds__(A5_THREADS, 1) k_attention_tc5(const __grid_constant__ CUtensorMap map_qkv, Attn5Args a) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int SP = a.SP, S = a.S;
  const int box_bytes = SP * 128;                 // one 64-column box
  const int stage_bytes = 9 * box_bytes;          // q | k | v of four heads
  float* s_km = reinterpret_cast<float*>(smem + 2 * stage_bytes + 8192);        // (+8 KB: M = 128 reads past the last box)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_km + 8 * A5_MAX_SP);
  uint64_t* stage_full = bars;          // [2]
  uint64_t* stage_empty = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;          // [2] scores of an even / odd head ready
  uint64_t* p_full = bars + 6;          // [2] probabilities written (4 warps)
  uint64_t* o_full = bars + 8;          // [2] P V done
  uint64_t* o_empty = bars + 10;        // [2] output accumulator read (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(stage_full + i, 1); mbar_init(stage_empty + i, 1);
      mbar_init(s_full + i, 1); mbar_init(p_full + i, 4);
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
    // ---------------- TMA producer: per window two head-quad stages of nine boxes ----------------
    if (lane == 0) {
      uint32_t n = 0;
      for (int w = blockIdx.x; w < a.B; w += gridDim.x)
        for (int quad = 0; quad < 2; ++quad, ++n) {
          const int st = n & 1;
          mbar_wait(stage_empty + st, ((n >> 1) & 1) ^ 1);
          mbar_expect_tx(stage_full + st, stage_bytes);
          uint8_t* dst = smem + st * stage_bytes;
          for (int op = 0; op < 3; ++op)
            for (int bx = 0; bx < 3; ++bx)
              tma_load_2d(dst + (op * 3 + bx) * box_bytes, &map_qkv, stage_full + st, op * A5_D + quad * 192 + bx * 64, w * S);
        }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc_s = make_idesc_bf16(128, SP);
    constexpr uint32_t idesc_o = make_idesc_bf16_bmn(128, 16);
    uint32_t n = 0, hcount = 0;       // stages consumed, heads issued
    auto issue_pv = [&](int st, int hq, uint32_t hc) {      // O[b] = P[b] . V_h (head hq of the stage, hc-th head overall)
      const int b = hc & 1;
      mbar_wait(p_full + b, (hc >> 1) & 1);
      mbar_wait(o_empty + b, ((hc >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t vbase = smem_u32(smem + st * stage_bytes + 6 * box_bytes);
        for (int i = 0; i < 3; ++i) {                       // three 16-column slices of the head's 48 value columns
          const int c = hq * A5_DH + 16 * i;
          const uint64_t vd = make_sw128_mn_bf16_desc(vbase + (c >> 6) * box_bytes + (c & 63) * 2);
          for (int t = 0; t < SP / 16; ++t)                 // 16 keys per step: 16 rows of the box = 2048 B, 8 TMEM columns of P
            umma_bf16_ts(tmem_base + A5_TM_O + b * 64 + 16 * i, tmem_base + A5_TM_S + b * 128 + 8 * t,
                         vd + (uint64_t)((t * 2048) >> 4), idesc_o, t != 0);
        }
        umma_commit(o_full + b);
      }
      __syncwarp();
    };
    uint32_t pend_st = 0, pend_hq = 0, pend_hc = 0;
    bool pending = false;
    for (int w = blockIdx.x; w < a.B; w += gridDim.x)
      for (int quad = 0; quad < 2; ++quad, ++n) {
        const int st = n & 1;
        mbar_wait(stage_full + st, (n >> 1) & 1);
        tcgen05_fence_after();
        for (int hq = 0; hq < A5_QUAD; ++hq, ++hcount) {
          const int b = hcount & 1;
          // S[b] = Q_h K_h^T.  (S[b] was last read by the softmax of head hcount - 2, whose P V was issued before this.)
          if (elect_one()) {
            const uint32_t qbase = smem_u32(smem + st * stage_bytes), kbase = qbase + 3 * box_bytes;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int c = hq * A5_DH + 16 * i;
              const uint32_t off = (c >> 6) * box_bytes + (c & 63) * 2;
              umma_bf16(tmem_base + A5_TM_S + b * 128, make_sw128_desc(qbase + off), make_sw128_desc(kbase + off), idesc_s, i != 0);
            }
            umma_commit(s_full + b);
          }
          __syncwarp();
          if (pending) {                 // P V of the previous head, behind this head's scores
            issue_pv(pend_st, pend_hq, pend_hc);
            if (pend_hq == A5_QUAD - 1) {          // last reads of that stage
              if (elect_one()) umma_commit(stage_empty + pend_st);
              __syncwarp();
            }
          }
          pend_st = st; pend_hq = hq; pend_hc = hcount; pending = true;
        }
      }
    if (pending) {
      issue_pv(pend_st, pend_hq, pend_hc);
      if (elect_one()) umma_commit(stage_empty + pend_st);
      __syncwarp();
    }
  } else {
    // ---------------- softmax / output warps: quarter q rows, heads of parity par ----------------
    const int e = warp - 2, q = warp & 3, par = e >> 2;
    const int row = q * 32 + lane;
    float* km = s_km + e * A5_MAX_SP;
    const float LOG2E = 1.4426950408889634f;
    const float scale = LOG2E / sqrtf((float)A5_DH);
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t hc = par;                   // running head counter of this parity: hc, hc + 2, ...
    for (int w = blockIdx.x; w < a.B; w += gridDim.x) {
      __syncwarp();
      for (int j = lane; j < SP; j += 32)
        km[j] = j >= S ? -INFINITY : ((a.mask && !a.mask[(long long)w * a.mask_stride + j]) ? -1e9f * LOG2E : 0.f);
      __syncwarp();
      for (int h = par; h < A5_HEADS; h += 2, hc += 2) {
        const uint32_t ph = (hc >> 1) & 1;
        const uint32_t t_s = tmem_base + lane_sel + A5_TM_S + par * 128;
        mbar_wait(s_full + par, ph);
        tcgen05_fence_after();
        float sc[A5_MAX_SP];
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_s, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) sc[i] = __uint_as_float(v[i]);
          if (SP > 32) {
            tmem_ld_32x32b_x32(t_s + 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) sc[32 + i] = __uint_as_float(v[i]);
          }
          if (SP > 64) {
            uint32_t u[16];
            tmem_ld_x16(t_s + 64, u);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) sc[64 + i] = __uint_as_float(u[i]);
          }
        }
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < A5_MAX_SP; ++j)
          if (j < SP) {
            sc[j] = fmaf(sc[j], scale, km[j]);
            m = fmaxf(m, sc[j]);
          }
        float l = 0.f;
        uint32_t pk[A5_MAX_SP / 2];
#pragma unroll
        for (int j = 0; j < A5_MAX_SP; j += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (j < SP) { p0 = ex2_a5(sc[j] - m); p1 = ex2_a5(sc[j + 1] - m); }
          l += p0 + p1;
          __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&pb);
        }
        // P over the scores: two bf16 per 32-bit column, 8 columns per 16-key step
#pragma unroll
        for (int c = 0; c < A5_MAX_SP / 2; c += 8)
          if (2 * c < SP) tmem_st_x8(t_s + c, pk + c);
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full + par);
        // O_h
        const uint32_t t_o = tmem_base + lane_sel + A5_TM_O + par * 64;
        mbar_wait(o_full + par, ph);
        tcgen05_fence_after();
        uint32_t o0[32], o1[16];
        tmem_ld_32x32b_x32(t_o, o0);
        tmem_ld_x16(t_o + 32, o1);
        tmem_wait_ld();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty + par);
        if (row < S) {
          const float inv = 1.f / l;
          uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)w * S + row) * A5_D + h * A5_DH);
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            uint32_t wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = 8 * c + 2 * i;
              const float x0 = __uint_as_float(k < 32 ? o0[k] : o1[k - 32]) * inv;
              const float x1 = __uint_as_float(k + 1 < 32 ? o0[k + 1] : o1[k + 1 - 32]) * inv;
              __nv_bfloat162 pb = __floats2bfloat162_rn(x0, x1);
              wv[i] = *reinterpret_cast<uint32_t*>(&pb);
            }
            dst[c] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
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
  const int smem = 2 * 9 * SP * 128 + 8192 + 8 * A5_MAX_SP * 4 + 256 + 1024;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(k_attention_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 9 * A5_MAX_SP * 128 + 8192 + 8 * A5_MAX_SP * 4 + 256 + 1024);
    if (e != cudaSuccess) return e;
    attr_smem = 2 * 9 * A5_MAX_SP * 128 + 8192 + 8 * A5_MAX_SP * 4 + 256 + 1024;
  }
  Attn5Args a;
  a.B = B; a.S = S; a.SP = SP; a.mask = mask; a.mask_stride = mask_stride; a.out = out;
  return launch_pdl(k_attention_tc5, dim3(B < num_sms ? B : num_sms), dim3(A5_THREADS), (size_t)smem, st, map, a);
}

}  // namespace uu
