// Weight gradients of the training step on the 5th-generation tensor cores:  dW[Kd, Nd] (+)= X^T dY  with
// X [R, Kd] and dY [R, Nd] fp32 row-major exactly as the tape / the backward pass hold them (train.py:493-499 computes
// these products inside tape.gradient).  The contraction runs over the ROW index, i.e. both operands are "MN-major" for
// the tensor core: TMA (3-D maps {32 floats, rows, column groups}, 128B swizzle) drops [group][row][32 floats] tiles
// into shared memory and tcgen05.mma kind::tf32 reads them through MN-major descriptors (a_major = b_major = 1), so
// no transposed copy of an activation is ever made.  Split-K over row ranges: CTA (tile, split) writes its fp32 partial
// tile, a second kernel sums the partials in split order — deterministic, no atomics.
#include <algorithm>

#include "tc_ptx.cuh"

namespace uu {

constexpr int WG_BKR = 32;         // contraction rows per pipeline stage (4 MMAs of K = 8)
constexpr int WG_THREADS = 192;    // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 epilogue

template <int BN>
struct WgCfg {
  static constexpr int A_BYTES = 4 * WG_BKR * 128;            // 128 columns of X = 4 groups of 32 floats
  static constexpr int B_BYTES = (BN / 32) * WG_BKR * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (176 * 1024) / STAGE_BYTES < 8 ? (176 * 1024) / STAGE_BYTES : 8;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// MN-major TF32 operand: the only layout the tensor core accepts is the 128B swizzle with 32-byte atomicity
// (descriptor layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32 fp32 (128 B) contiguous along M/N, 4
// contraction rows 128 B apart form one atom (32-byte units of a row XOR-ed with row mod 4);
// LBO = distance between 32-column groups, SBO = distance between 4-row atoms.
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;        // SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {     // D f32, A = B = TF32, both MN-major
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_wg(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

#ifdef WG_DEBUG
__device__ uint32_t g_wg_lbo = WG_BKR * 128, g_wg_sbo = 512, g_wg_idesc_xor = 0;
#endif
// grid = (m_tiles * n_tiles, splits).  partial: [splits][m_tiles * 128][n_tiles * BN] fp32.
template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) k_wgrad_tc(const __grid_constant__ CUtensorMap map_x,
                                                            const __grid_constant__ CUtensorMap map_dy, int R,
                                                            int rows_per_split, int n_tiles, float* __restrict__ partial) {
  using Cfg = WgCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* acc_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tm = blockIdx.x / n_tiles, tn = blockIdx.x % n_tiles, split = blockIdx.y;
  const int r0 = split * rows_per_split, r1 = min(R, r0 + rows_per_split);
  const int num_kb = r1 > r0 ? (r1 - r0 + WG_BKR - 1) / WG_BKR : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar + s, ph ^ 1);
        uint8_t* dst = smem + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(full_bar + s, Cfg::STAGE_BYTES);
        tma_load_3d(dst, &map_x, full_bar + s, 0, r0 + kb * WG_BKR, tm * 4);
        tma_load_3d(dst + Cfg::A_BYTES, &map_dy, full_bar + s, 0, r0 + kb * WG_BKR, tn * (BN / 32));
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
#ifdef WG_DEBUG
    const uint32_t idesc = make_idesc_tf32_mn(128, BN) ^ g_wg_idesc_xor;
#else
    constexpr uint32_t idesc = make_idesc_tf32_mn(128, BN);
#endif
    int s = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(full_bar + s, ph);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
#ifdef WG_DEBUG
        const uint64_t a_desc = make_sw128_mn_desc(a_addr, g_wg_lbo, g_wg_sbo);
        const uint64_t b_desc = make_sw128_mn_desc(a_addr + Cfg::A_BYTES, g_wg_lbo, g_wg_sbo);
#else
        const uint64_t a_desc = make_sw128_mn_desc(a_addr, WG_BKR * 128, 512);
        const uint64_t b_desc = make_sw128_mn_desc(a_addr + Cfg::A_BYTES, WG_BKR * 128, 512);
#endif
#pragma unroll
        for (int k = 0; k < WG_BKR / 8; ++k)       // two 4-row atoms (1024 B) per MMA
          umma_tf32_wg(tmem_base, a_desc + (uint64_t)(k * (1024 >> 4)), b_desc + (uint64_t)(k * (1024 >> 4)), idesc, (kb | k) != 0);
        umma_commit(empty_bar + s);
        if (kb == num_kb - 1) umma_commit(acc_bar);
      }
      __syncwarp();
      if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    // epilogue: each warp drains its TMEM lane quarter (rows of dW), 32 columns at a time
    const int q = warp & 3;
    const int m = tm * 128 + q * 32 + lane;
    const long long ldp = (long long)n_tiles * BN;
    float* dst = partial + ((long long)split * (gridDim.x / n_tiles) * 128 + m) * ldp + (long long)tn * BN;
    if (num_kb > 0) {
      mbar_wait(acc_bar, 0);
      tcgen05_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(dst + c + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                  __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// dW[m][n] (+)= sum_s partial[s][m][n], splits summed in index order
__global__ void k_wgrad_reduce(const float* __restrict__ partial, int splits, int m_pad, long long ldp, int Kd, int Nd,
                               float* __restrict__ dW, int accumulate) {
  const int n4 = Nd >> 2;
  const long long total = (long long)Kd * n4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / n4), c = (int)(i - (long long)m * n4) * 4;
    float4 acc = accumulate ? *reinterpret_cast<const float4*>(dW + (long long)m * Nd + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 p = *reinterpret_cast<const float4*>(partial + ((long long)s * m_pad + m) * ldp + c);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    *reinterpret_cast<float4*>(dW + (long long)m * Nd + c) = acc;
  }
}

static int encode_3d_f32(CUtensorMap* map, const float* base, int cols, long long rows, long long ld, int box_rows, int box_groups) {
  PFN_encodeTiled enc = get_encoder();
  UU_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  UU_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 4) % 16 == 0, "wgrad operand must be 16-byte aligned");
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)((cols + 31) / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_groups};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UU_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed (code " + std::to_string((int)r) + ")");
  return 0;
}

size_t wgrad_tc_scratch_bytes() { return (size_t)48 << 20; }

bool wgrad_tc_ok(const float* X, long long ldx, const float* dY, long long ldy, long long R, int Kd, int Nd) {
  return R >= 256 && Kd >= 128 && Kd % 32 == 0 && Nd >= 64 && Nd % 64 == 0 && ldx % 4 == 0 && ldy % 4 == 0 &&
         ((uintptr_t)X & 15) == 0 && ((uintptr_t)dY & 15) == 0;
}

template <int BN>
static int wgrad_launch(const CUtensorMap& mx, const CUtensorMap& my, int R, int rows_per_split, int m_tiles, int n_tiles,
                        int splits, float* scratch, cudaStream_t st) {
  using Cfg = WgCfg<BN>;
  static bool attr = false;
  if (!attr) {
    UU_CUDA(cudaFuncSetAttribute(k_wgrad_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr = true;
  }
  k_wgrad_tc<BN><<<dim3(m_tiles * n_tiles, splits), WG_THREADS, Cfg::SMEM_BYTES, st>>>(mx, my, R, rows_per_split, n_tiles, scratch);
  UU_CUDA(cudaGetLastError());
  return 0;
}

// dW [Kd, Nd] (row-major, contiguous) (+)= X^T dY.  scratch: wgrad_tc_scratch_bytes() bytes of device memory.
int wgrad_tc(const float* X, long long ldx, const float* dY, long long ldy, long long R, int Kd, int Nd, float* dW,
             int accumulate, float* scratch, int num_sms, cudaStream_t st) {
  UU_CHECK(wgrad_tc_ok(X, ldx, dY, ldy, R, Kd, Nd), "shape not supported by the tcgen05 wgrad kernel");
  const int bn = (Nd % 192 == 0) ? 192 : (Nd % 128 == 0) ? 128 : 64;
  const int m_tiles = (Kd + 127) / 128, n_tiles = Nd / bn, tiles = m_tiles * n_tiles;
  const long long kblocks = (R + WG_BKR - 1) / WG_BKR;
  // ONE wave: the kernel keeps ~176 KB of operand stages, so a single CTA is resident per SM; "about two CTAs per SM" (the first
  // choice) ran 306 CTAs of the q|k|v wgrad as two full waves plus a third with 10 CTAs
  int splits = std::max(1, num_sms / tiles);
  splits = (int)std::min<long long>(splits, std::max<long long>(1, kblocks / 8));     // at least 8 k-blocks per CTA
  const size_t per_split = (size_t)m_tiles * 128 * Nd * sizeof(float);
  splits = (int)std::min<size_t>(splits, std::max<size_t>(1, wgrad_tc_scratch_bytes() / per_split));
  const int rows_per_split = (int)(((kblocks + splits - 1) / splits) * WG_BKR);
  splits = (int)((R + rows_per_split - 1) / rows_per_split);
  CUtensorMap mx, my;
  if (encode_3d_f32(&mx, X, Kd, R, ldx, WG_BKR, 4) || encode_3d_f32(&my, dY, Nd, R, ldy, WG_BKR, bn / 32)) return 1;
  int rc;
  switch (bn) {
    case 192: rc = wgrad_launch<192>(mx, my, (int)R, rows_per_split, m_tiles, n_tiles, splits, scratch, st); break;
    case 128: rc = wgrad_launch<128>(mx, my, (int)R, rows_per_split, m_tiles, n_tiles, splits, scratch, st); break;
    default: rc = wgrad_launch<64>(mx, my, (int)R, rows_per_split, m_tiles, n_tiles, splits, scratch, st); break;
  }
  if (rc) return 1;
  const long long total4 = (long long)Kd * (Nd / 4);
  const int blocks = (int)std::min<long long>((total4 + 255) / 256, 4 * num_sms);
  k_wgrad_reduce<<<blocks, 256, 0, st>>>(scratch, splits, m_tiles * 128, (long long)Nd, Kd, Nd, dW, accumulate);
  UU_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace uu
