// K3: bf16 GEMM on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared-memory ring -> tcgen05.mma (accumulator in
//   TMEM) -> tcgen05.ld -> fused epilogue (bias / ReLU / residual / positional table / row scatter).
// Warp roles in a 320-thread persistent CTA (one per SM): warp 0 = TMA producer, warp 1 = TMEM allocator +
// MMA issuer, warps 2..9 = epilogue.  The accumulator is double-buffered in TMEM so the epilogue of tile i
// overlaps the MMA main loop of tile i+1.
//
// A  : bf16 row-major [M, K] with leading dimension lda (activations; the strided Conv1D reads its
//      zero-padded input as a [B*L_out, 3*C] matrix with lda = stride*C — implicit GEMM, no im2col).
// Wt : bf16 [N_pad, K] = W^T, i.e. both operands are K-major.
#include <cstdlib>

#include <cuda.h>   // CUtensorMap types only; the encoder is fetched through the runtime (no -lcuda)

#include "common.cuh"
#include "epilogue.cuh"
#include "tc_ptx.cuh"

namespace uu {

// Timing-experiment switches (DESIGN.md "GEMM analysis") exist only in builds with -DUU_EXPERIMENT: most of them
// produce wrong results on purpose, so the shipped library never reads them from the environment.
#ifdef UU_EXPERIMENT
#define UU_EXP_FLAG(epi, bit) (((epi).flags & (bit)) != 0)
#else
#define UU_EXP_FLAG(epi, bit) false
#endif

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;          // 64 bf16 = 128 B = one swizzle atom row
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 320;        // TMA warp, MMA warp, 8 epilogue warps
constexpr int TC_TSTRIDE = 36;          // fp32 row stride of the epilogue transpose tiles (conflict-free float4)
#ifndef UU_TC_EPI_NBUF
#define UU_TC_EPI_NBUF 2        // TMA-store staging tiles per epilogue warp
#endif
#ifndef UU_TC_RING_KB
#define UU_TC_RING_KB 160       // operand ring budget
#endif
constexpr int TC_EPI_SMEM = 8 * UU_TC_EPI_NBUF * 32 * 128 + 1024;   // max(fp32 transpose tiles 36 KB, per-warp TMA staging + align)

// Epilogue of one warp for one tile (bf16 output): sub-tiles first, first+2, ... of its 32 accumulator rows go
// TMEM -> registers -> bias / ReLU / bf16 -> private swizzled staging tile -> cp.async.bulk.tensor store of a
// [32 x 64] box.  `release` is called right after the warp's last TMEM read of the tile.
// Row scatter (epi.c_rowidx: the compact valid-frame rows of the 544 -> 384 GEMM go to their token rows): the list is
// ascending, so the 32 rows of a warp almost always land on 32 consecutive destination rows and still leave as one
// TMA box; otherwise (a masked frame inside the run, or the ragged end of the list) every lane writes its own row.
// Global-memory inputs of one warp's epilogue for one tile, fetched BEFORE the warp waits for the accumulator so their
// latency hides behind the MMA main loop: the row statistics of a folded LayerNorm and the residual line of the first
// sub-tile (the following sub-tiles' lines are fetched while the previous one is being stored).
struct EpiPre {
  float mu = 0.f, rstd = 1.f;
  uint4 rres[8];
};
__device__ __forceinline__ void epi_load_residual(EpiPre& pre, const Epilogue& epi, int row, int col, int m_eff, long long ldc) {
  const uint4* rp = reinterpret_cast<const uint4*>(epi.res_bf16 + (long long)row * ldc + col);
#ifdef UU_EXPERIMENT
  const bool ok = row < m_eff && !(epi.flags & 1024);     // (flag 1024: timing experiment UU_GEMM_NORES)
#else
  const bool ok = row < m_eff;
#endif
#pragma unroll
  for (int g = 0; g < 8; ++g) pre.rres[g] = ok ? rp[g] : make_uint4(0u, 0u, 0u, 0u);
}
template <int EMODE>
__device__ __forceinline__ void epi_prefetch(EpiPre& pre, const Epilogue& epi, int row_q0, int col0, int first, int lane,
                                             int m_eff, long long ldc) {
  const int my_row = row_q0 + lane;
  if constexpr (EMODE == 1) {
    float s1 = 0.f, s2 = 0.f;
    if (my_row < m_eff) {
      const float2* sp = reinterpret_cast<const float2*>(epi.ln_stats) + (long long)my_row * epi.ln_slots;
      for (int i = 0; i < epi.ln_slots; ++i) {
        const float2 t = sp[i];
        s1 += t.x; s2 += t.y;
      }
    }
    pre.mu = s1 * epi.ln_inv_k;
    pre.rstd = rsqrtf(fmaxf(s2 * epi.ln_inv_k - pre.mu * pre.mu, 0.f) + epi.ln_eps);
  }
  if constexpr (EMODE == 2) epi_load_residual(pre, epi, my_row, col0 + first * 64, m_eff, ldc);
}

// Packed fp32 pairs (sm_100 FFMA2 / FADD2): the epilogue is issue-bound on 8 warps, these halve its arithmetic.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

// EMODE selects the epilogue at compile time (runtime flag tests inside the 8-column loop doubled its instruction count):
//   EMODE_PLAIN  out = acc (+ bias) (ReLU)
//   EMODE_LNFOLD out = rstd * acc + ((-rstd * mean) * csum + bias') (ReLU)      two FFMA2 per column pair
//   EMODE_RESID  out = acc + bias + residual, row partial sums of out and out^2 for the next folded LayerNorm
//   EMODE_TABLE  out = acc + bias + table[dst_row % period] (fp32 positional table), same row partial sums; rows may
//                be scattered (c_rowidx): table row and statistics slot follow the DESTINATION row
constexpr int EMODE_PLAIN = 0, EMODE_LNFOLD = 1, EMODE_RESID = 2, EMODE_TABLE = 3;

template <int BLOCK_N, int NBUF, int EMODE, typename Release>
__device__ __forceinline__ void epi_warp_store_tile(uint32_t tmem_acc, int first, uint8_t* my_stage, uint32_t& my_count,
                                                    const Epilogue& epi, const CUtensorMap* map_c, int row_q0, int col0,
                                                    int lane, Release release, EpiPre& pre, int m_eff, bf16* c_ptr,
                                                    long long ldc) {
  constexpr int NSUB = BLOCK_N / 64;
  int dst_row = row_q0;                      // first destination row of the TMA box
  bool boxed = true;
  int my_dst = -1;
  if (epi.c_rowidx || epi.cmap.rpb != 0x7fffffff) {      // scatter list, or a (batch, position) row map that drops rows
    const int r = row_q0 + lane;
    my_dst = r < m_eff ? (epi.c_rowidx ? epi.c_rowidx[r] : (int)map_row(epi.cmap, r)) : -1;
    dst_row = __shfl_sync(0xffffffffu, my_dst, 0);
    boxed = __all_sync(0xffffffffu, my_dst >= 0 && my_dst == dst_row + lane);
  }
  const int my_row = row_q0 + lane;
  // destination row of this lane (EMODE_TABLE: positional-table row and statistics slot), -1 = dropped
  const int out_row = (epi.c_rowidx || epi.cmap.rpb != 0x7fffffff) ? my_dst : (my_row < m_eff ? my_row : -1);
  const float* trow = nullptr;
  if constexpr (EMODE == EMODE_TABLE)
    trow = epi.table + (long long)(out_row < 0 ? 0 : out_row % epi.table_period) * (long long)ldc;
  const float2 a1 = make_float2(pre.rstd, pre.rstd), a2 = make_float2(-pre.rstd * pre.mu, -pre.rstd * pre.mu);
  const bool relu = (epi.flags & EPI_RELU) != 0;
  const bool has_bias = epi.bias != nullptr;
#pragma unroll 1
  for (int sub = first; sub < NSUB; sub += 2, ++my_count) {
    uint8_t* sbuf = my_stage + (my_count % NBUF) * (32 * 128);
    float2 st_s = make_float2(0.f, 0.f), st_q = make_float2(0.f, 0.f);
    uint32_t v0[32], v1[32];
    tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(sub * 64), v0);
    tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(sub * 64 + 32), v1);
    if (sub + 2 >= NSUB) release();         // last TMEM read of this warp for this tile: hand the accumulator back
    // the store issued NBUF sub-tiles ago read this staging tile: it must have finished reading
    if (lane == 0) {
      if constexpr (NBUF == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
      else if constexpr (NBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
#ifdef UU_EXPERIMENT
    if (!(epi.flags & 512))                 // (flag 512: timing experiment UU_GEMM_NOEPI — no conversion)
#endif
#pragma unroll
    for (int g = 0; g < 8; ++g) {           // 8 columns -> one 16-byte chunk
      const int cb = col0 + sub * 64 + 8 * g;
      float2 o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = g < 4 ? make_float2(__uint_as_float(v0[8 * g + 2 * i]), __uint_as_float(v0[8 * g + 2 * i + 1]))
                     : make_float2(__uint_as_float(v1[8 * (g - 4) + 2 * i]), __uint_as_float(v1[8 * (g - 4) + 2 * i + 1]));
      if constexpr (EMODE == EMODE_LNFOLD) {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(epi.ln_csum + cb));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(epi.ln_csum + cb + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + cb));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + cb + 4));
        o[0] = ffma2(a1, o[0], ffma2(a2, make_float2(c0.x, c0.y), make_float2(b0.x, b0.y)));
        o[1] = ffma2(a1, o[1], ffma2(a2, make_float2(c0.z, c0.w), make_float2(b0.z, b0.w)));
        o[2] = ffma2(a1, o[2], ffma2(a2, make_float2(c1.x, c1.y), make_float2(b1.x, b1.y)));
        o[3] = ffma2(a1, o[3], ffma2(a2, make_float2(c1.z, c1.w), make_float2(b1.z, b1.w)));
      } else if (EMODE == EMODE_RESID || EMODE == EMODE_TABLE || has_bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + cb));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + cb + 4));
        o[0] = fadd2(o[0], make_float2(b0.x, b0.y)); o[1] = fadd2(o[1], make_float2(b0.z, b0.w));
        o[2] = fadd2(o[2], make_float2(b1.x, b1.y)); o[3] = fadd2(o[3], make_float2(b1.z, b1.w));
      }
      if constexpr (EMODE == EMODE_PLAIN || EMODE == EMODE_LNFOLD) {
        if (relu) {
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = make_float2(fmaxf(o[i].x, 0.f), fmaxf(o[i].y, 0.f));
        }
      } else {
        if constexpr (EMODE == EMODE_RESID) {
          const uint4 r = pre.rres[g];
          o[0] = fadd2(o[0], bf16x2_to_f2(r.x)); o[1] = fadd2(o[1], bf16x2_to_f2(r.y));
          o[2] = fadd2(o[2], bf16x2_to_f2(r.z)); o[3] = fadd2(o[3], bf16x2_to_f2(r.w));
        } else {
          const float4 t0 = __ldg(reinterpret_cast<const float4*>(trow + cb));
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(trow + cb + 4));
          o[0] = fadd2(o[0], make_float2(t0.x, t0.y)); o[1] = fadd2(o[1], make_float2(t0.z, t0.w));
          o[2] = fadd2(o[2], make_float2(t1.x, t1.y)); o[3] = fadd2(o[3], make_float2(t1.z, t1.w));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          st_s = fadd2(st_s, o[i]);
          st_q = ffma2(o[i], o[i], st_q);
        }
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0].x, o[0].y), p1 = __floats2bfloat162_rn(o[1].x, o[1].y);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(o[2].x, o[2].y), p3 = __floats2bfloat162_rn(o[3].x, o[3].y);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
      pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(sbuf + lane * 128 + ((g ^ (lane & 7)) << 4)) = pk;
    }

    if constexpr (EMODE == EMODE_RESID || EMODE == EMODE_TABLE) {
      // (statistics of the fp32 sums before the bf16 rounding: the difference to the stored values is far below the
      // bf16 resolution of the normalised output)
      if (epi.stats_out && out_row >= 0)
        reinterpret_cast<float2*>(epi.stats_out)[(long long)out_row * epi.ln_slots + ((col0 >> 6) + sub)] =
            make_float2(st_s.x + st_s.y, st_q.x + st_q.y);
    }
    if constexpr (EMODE == EMODE_RESID) {
      if (sub + 2 < NSUB) epi_load_residual(pre, epi, my_row, col0 + (sub + 2) * 64, m_eff, ldc);   // next sub-tile's line
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy
    __syncwarp();
    if (boxed) {
#ifdef UU_EXPERIMENT
      if (lane == 0 && !(epi.flags & 64)) {   // (flag 64: timing experiment, UU_GEMM_NOSTORE)
#else
      if (lane == 0) {
#endif
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map_c),
                     "r"(smem_u32(sbuf)), "r"(col0 + sub * 64), "r"(dst_row)
                     : "memory");
      }
    } else if (my_dst >= 0) {                // scattered rows: 128 bytes per lane straight from the staging tile
      bf16* dst = c_ptr + (long long)my_dst * ldc + col0 + sub * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(dst + 8 * g) = *reinterpret_cast<const uint4*>(sbuf + lane * 128 + ((g ^ (lane & 7)) << 4));
    }
    if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// Instruction descriptor, kind::tf32: D = f32, A = B = TF32 (format 2), both K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BLOCK_N>
struct TcCfg {
  static constexpr int A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;   // 16 KB
  static constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue scratch: per-warp TMA-store staging, 8 warps x NBUF x 4 KB (+ alignment); the generic fp32 transpose
  // tiles (36 KB) fit in the same region.
  static constexpr int EPI_NBUF = UU_TC_EPI_NBUF;
  static constexpr int EPI_BYTES = TC_EPI_SMEM;
  static constexpr int STAGES = (UU_TC_RING_KB * 1024) / STAGE_BYTES < 8 ? (UU_TC_RING_KB * 1024) / STAGE_BYTES : 8;
  static constexpr int ACC_STAGES = 2;                           // double-buffered accumulator in TMEM
  static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N <= 32 ? 32 : ACC_STAGES * BLOCK_N <= 64 ? 64
                                   : ACC_STAGES * BLOCK_N <= 128 ? 128 : ACC_STAGES * BLOCK_N <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  static_assert(ACC_STAGES * BLOCK_N <= 512, "accumulator stages exceed TMEM");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment for the 128B swizzle");
};

// Persistent, warp-specialised: every CTA walks the tile list t = blockIdx.x, += gridDim.x with
// (m_blk, n_blk) = (t / n_tiles, t % n_tiles), so CTAs running concurrently share the same A rows in L2.
// Three pipelines: smem ring (TMA -> MMA), TMEM accumulator ring (MMA -> epilogue), tile list.
template <int BLOCK_N, typename TC, bool TMA_OUT, int EMODE = 0, bool TF32 = false>
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc(const __grid_constant__ CUtensorMap map_a,
                                                           const __grid_constant__ CUtensorMap map_b,
                                                           const __grid_constant__ CUtensorMap map_c, int M, int N,
                                                           int n_tiles, int K, Epilogue epi, TC* __restrict__ C,
                                                           long long ldc) {
  using Cfg = TcCfg<BLOCK_N>;
  constexpr int BK = TF32 ? TC_BLOCK_K / 2 : TC_BLOCK_K;      // elements per 128-byte k-block row: 32 fp32 or 64 bf16
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  const int m_eff = epi.m_dev ? min(M, *epi.m_dev) : M;
  const int m_tiles = (m_eff + TC_BLOCK_M - 1) / TC_BLOCK_M;
  const int total_tiles = m_tiles * n_tiles;
  const int tile_first = blockIdx.x;
  const int tile_step = gridDim.x;

  uint8_t* smem_al = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem = smem_al;                                // operand ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + Cfg::ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + Cfg::ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    if constexpr (TMA_OUT) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < Cfg::ACC_STAGES; ++s) {
      mbar_init(tmem_full_bar + s, 1);
      mbar_init(tmem_empty_bar + s, 8);      // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // everything above overlapped the previous kernel's tail; its results are needed from here on

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t it = 0;                        // running k-block counter across tiles -> stage / phase
      int s = 0;
      uint32_t ph = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const int row0 = (tile / n_tiles) * TC_BLOCK_M, col0 = (tile % n_tiles) * BLOCK_N;
        // A rows are streamed from HBM exactly once but needed by all n_tiles column blocks at about the same time,
        // so every stage used to pay the full DRAM latency (~1.9 us under load, more than the ring covers).  The CTA
        // that will own column block 0 of a row block pulls that row block into L2 one tile ahead.
        {
          const int nt = tile + tile_step;
          if (nt < total_tiles && (nt % n_tiles) == 0 && !UU_EXP_FLAG(epi, 256)) {
            const int prow = (nt / n_tiles) * TC_BLOCK_M;
            for (int kb = 0; kb < num_kb; ++kb) tma_prefetch_2d(&map_a, kb * BK, prow);
          }
          if constexpr (TMA_OUT && EMODE == EMODE_RESID) {
            // residual rows of this CTA's next tile (the output map describes the same matrix): the epilogue's
            // row-per-lane loads then hit L2 instead of paying the DRAM latency in front of a sub-tile
            if (nt < total_tiles && !UU_EXP_FLAG(epi, 256)) {
              const int prow = (nt / n_tiles) * TC_BLOCK_M, pcol = (nt % n_tiles) * BLOCK_N;
              for (int r = 0; r < TC_BLOCK_M; r += 32)
                for (int c = 0; c < BLOCK_N; c += 64) tma_prefetch_2d(&map_c, pcol + c, prow + r);
            }
          }
        }
#ifdef UU_EXPERIMENT
        if ((epi.flags & 128) && (epi.flags & 4096)) continue;
#endif
        for (int kb = 0; kb < num_kb; ++kb, ++it, s = (s + 1 == Cfg::STAGES ? 0 : s + 1), ph ^= (s == 0)) {
          mbar_wait(empty_bar + s, ph ^ 1);
          uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
#ifdef UU_EXPERIMENT
          if (epi.flags & 128) {                 // (flag 128: timing experiment UU_GEMM_NOLOAD — MMA on stale smem)
            mbar_arrive(full_bar + s);
            continue;
          }
#endif
          mbar_expect_tx(full_bar + s, Cfg::STAGE_BYTES);
          tma_load_2d(a_dst, &map_a, full_bar + s, kb * BK, row0);
          tma_load_2d(a_dst + Cfg::A_BYTES, &map_b, full_bar + s, kb * BK, col0);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // The whole warp walks the loop (warp-uniform control flow, every lane polls the barriers) and one elected
    // lane issues the tcgen05 instructions.  The loop body is kept to a handful of instructions per MMA: a
    // 128 x BLOCK_N x 16 MMA retires in BLOCK_N / 2 cycles, so with 64-wide k-blocks the issue loop, not the tensor
    // pipe, was the pacer of these short-K GEMMs (descriptors are now one add from a per-kernel base, stage and
    // phase are carried incrementally).
    {
      constexpr uint32_t idesc = TF32 ? make_idesc_tf32(TC_BLOCK_M, BLOCK_N) : make_idesc_bf16(TC_BLOCK_M, BLOCK_N);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(smem));
      const uint64_t b_desc0 = make_sw128_desc(smem_u32(smem + Cfg::A_BYTES));
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tcount) {
        const int as = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(tmem_empty_bar + as, aph ^ 1);          // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
#ifdef UU_EXPERIMENT
        if ((epi.flags & 128) && (epi.flags & 4096)) {     // (timing experiment UU_GEMM_NOLOAD + UU_GEMM_ONECOMMIT: the
                                                            // tile's MMAs back to back on stale smem, one commit, no ring)
          if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb)
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
                umma_bf16(tmem_d, a_desc0 + 2 * k, b_desc0 + 2 * k, idesc, (kb | k) != 0);
            umma_commit(tmem_full_bar + as);
          }
          __syncwarp();
          continue;
        }
#endif
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + s, ph);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = a_desc0 + (uint64_t)((s * Cfg::STAGE_BYTES) >> 4);
            const uint64_t b_desc = b_desc0 + (uint64_t)((s * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (>>4) address field
              if constexpr (TF32) umma_tf32(tmem_d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              else umma_bf16(tmem_d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit(empty_bar + s);           // frees the smem stage when these MMAs retire
            if (kb == num_kb - 1) umma_commit(tmem_full_bar + as);        // accumulator complete
          }
          __syncwarp();
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> smem transpose -> coalesced global ----------------
    // 8 warps; warps e and e+4 share a TMEM lane quarter and alternate over the 32-column chunks.  The
    // tcgen05.ld layout is row-per-thread; a 32x32 fp32 transpose through a private, padded smem tile turns
    // every global access (residual read, positional table read, output write) into full 128-byte rows.
    const int e = warp - 2;                   // 0..7
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int hsel = e >> 2;                  // chunk parity handled by this warp
    if constexpr (TMA_OUT) {
      // ---- bf16 output through shared memory + TMA store --------------------------------------------------
      // Every epilogue warp is self-contained: it owns 32 accumulator rows (its TMEM lane quarter), converts whole
      // 64-column sub-tiles of them (bias, ReLU, bf16 pack) into a private 32 x 128 B staging tile in the
      // 128B-swizzle layout (conflict-free 16-byte stores) and issues its own cp.async.bulk.tensor store of that
      // [32 x 64] box; two private tiles alternate so a store overlaps the next conversion.  The two warps that
      // share a lane quarter take alternate sub-tiles.  No cross-warp barrier, no single issuing thread.
      // Rows past M are clipped by TMA.
      uint8_t* stage_base = reinterpret_cast<uint8_t*>(
          (reinterpret_cast<uintptr_t>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256) + 1023) & ~uintptr_t(1023));
      uint8_t* my_stage = stage_base + e * (Cfg::EPI_NBUF * 32 * 128);          // NBUF x 4 KB per warp
      constexpr int NSUB = BLOCK_N / 64;
      uint32_t tcount = 0, my_count = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tcount) {
        const int row0 = (tile / n_tiles) * TC_BLOCK_M, col0 = (tile % n_tiles) * BLOCK_N;
        const int as = tcount % Cfg::ACC_STAGES;
        const uint32_t aph = (tcount / Cfg::ACC_STAGES) & 1;
        // sub-tiles of this warp: those with (sub + tcount) % 2 == hsel (alternating start balances odd NSUB)
        const int first = (hsel + (int)tcount) & 1;
        EpiPre pre;
        if (first < NSUB) epi_prefetch<EMODE>(pre, epi, row0 + q * 32, col0, first, lane, m_eff, ldc);
        mbar_wait(tmem_full_bar + as, aph);
        tcgen05_fence_after();
        const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N);
        if (first >= NSUB) {                    // nothing to do for this tile (only possible for NSUB == 1)
          tcgen05_fence_before();
          if (lane == 0) mbar_arrive(tmem_empty_bar + as);
          continue;
        }
        epi_warp_store_tile<BLOCK_N, Cfg::EPI_NBUF, EMODE>(tmem_acc, first, my_stage, my_count, epi, &map_c, row0 + q * 32, col0, lane, [&] {
          tcgen05_fence_before();
          if (lane == 0) mbar_arrive(tmem_empty_bar + as);
        }, pre, m_eff, reinterpret_cast<bf16*>(C), ldc);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
    float* tbuf = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256) + e * (32 * TC_TSTRIDE);
    const int rsub = lane >> 3, g4 = (lane & 7) * 4;
    const bool vec_ok = (N % 4 == 0) && (ldc % 4 == 0) && (!(epi.flags & EPI_RESIDUAL) || (epi.ldr % 4 == 0));
    uint32_t tcount = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tcount) {
      const int row0 = (tile / n_tiles) * TC_BLOCK_M, col0 = (tile % n_tiles) * BLOCK_N;
      const int as = tcount % Cfg::ACC_STAGES;
      const uint32_t aph = (tcount / Cfg::ACC_STAGES) & 1;
      const int r = row0 + q * 32 + lane;
      int my_crow = -1, my_rrow = 0, my_trow = 0;
      if (r < m_eff) {
        const EpiRow er = epi_row(epi, r);
        my_crow = (int)er.crow; my_rrow = (int)er.rrow; my_trow = er.trow;
      }
      int crow[8], rrow[8];                   // rows this lane stores in the transposed pass: 4*it + rsub
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        crow[it] = __shfl_sync(0xffffffffu, my_crow, it * 4 + rsub);
        rrow[it] = __shfl_sync(0xffffffffu, my_rrow, it * 4 + rsub);
      }
      mbar_wait(tmem_full_bar + as, aph);
      tcgen05_fence_after();
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N);
#pragma unroll 1
      for (int ch = hsel; ch < BLOCK_N / 32; ch += 2) {
        const int cbase = col0 + ch * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(ch * 32), v);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(tbuf + lane * TC_TSTRIDE + 4 * g) =
              make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                          __uint_as_float(v[4 * g + 3]));
        __syncwarp();
        if (vec_ok) {
          const int c = cbase + g4;
          const bool col_ok = c < N;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (epi.bias && col_ok) b4 = *reinterpret_cast<const float4*>(epi.bias + c);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (!col_ok || crow[it] < 0) continue;
            const float4 a = *reinterpret_cast<const float4*>(tbuf + (it * 4 + rsub) * TC_TSTRIDE + g4);
            float o0 = a.x + b4.x, o1 = a.y + b4.y, o2 = a.z + b4.z, o3 = a.w + b4.w;
            if (epi.flags & EPI_RELU) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); o2 = fmaxf(o2, 0.f); o3 = fmaxf(o3, 0.f); }
            if (epi.flags & EPI_RESIDUAL) {
              const float4 t = *reinterpret_cast<const float4*>(epi.res + (long long)rrow[it] * epi.ldr + c);
              o0 += t.x; o1 += t.y; o2 += t.z; o3 += t.w;
            }
            if (epi.flags & EPI_ROWTABLE) {
              const float4 t = *reinterpret_cast<const float4*>(epi.table + (long long)(crow[it] % epi.table_period) * N + c);
              o0 += t.x; o1 += t.y; o2 += t.z; o3 += t.w;
            }
            TC* dst = C + (long long)crow[it] * ldc + c;
            if constexpr (sizeof(TC) == 4) {
              *reinterpret_cast<float4*>(dst) = make_float4(o0, o1, o2, o3);
            } else {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
              *reinterpret_cast<uint2*>(dst) = pk;
            }
          }
        } else {
          // generic path (e.g. the 51-wide heads): lane == column, one row per step
          const int c = cbase + lane;
#pragma unroll 4
          for (int it = 0; it < 32; ++it) {
            const int cr = __shfl_sync(0xffffffffu, my_crow, it), rr = __shfl_sync(0xffffffffu, my_rrow, it);
            const int tr = __shfl_sync(0xffffffffu, my_trow, it);      // (one modulo per row, not per element)
            if (cr < 0 || c >= N) continue;
            EpiRow er;
            er.crow = cr; er.rrow = rr; er.trow = tr;
            store_out(C + (long long)cr * ldc + c, epi_value(epi, er, tbuf[it * TC_TSTRIDE + lane], c, N));
          }
        }
        __syncwarp();   // the transpose tile is reused by the next chunk; also reconverges before tcgen05.ld
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld): hand the accumulator back to the MMA warp
      tcgen05_fence_before();
      if (lane == 0) mbar_arrive(tmem_empty_bar + as);
    }
    }   // !TMA_OUT
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BLOCK_N tile.
//   * each CTA loads its own 128 rows of A and HALF of the B tile (BLOCK_N/2 rows of W^T) per k-block, so the
//     L2 -> SM operand traffic per output drops by a third against the 128-row tile and a stage is 28-32 KB
//     (5 stages of 64-wide k-blocks instead of 3-4);
//   * the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256: the hardware reads A and B
//     from both CTAs' shared memory and writes each CTA's 128 accumulator rows into its own TMEM;
//   * barriers: TMA loads of both CTAs complete on the leader's full barrier; tcgen05.commit multicasts the
//     "stage free" and "accumulator ready" arrivals to both CTAs; the epilogue warps of both CTAs hand the
//     accumulator back by arriving on the leader's barrier.
// Only the hot-path epilogue (bias, ReLU, bf16, TMA store) exists in this variant.
// ================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(const void* p, uint32_t rank) {
  uint32_t d;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(smem_u32(p)), "r"(rank));
  return d;
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {      // arrives on `bar` of BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int BLOCK_N>
struct Tc2Cfg {
  static constexpr int A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;          // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BLOCK_N / 2) * TC_BLOCK_K * 2;       // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (UU_TC_RING_KB * 1024) / STAGE_BYTES < 8 ? (UU_TC_RING_KB * 1024) / STAGE_BYTES : 8;
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N <= 128 ? 128 : ACC_STAGES * BLOCK_N <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + TC_EPI_SMEM;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  static_assert(ACC_STAGES * BLOCK_N <= 512, "accumulator stages exceed TMEM");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment for the 128B swizzle");
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "UMMA M=256 needs N % 16 == 0, N <= 256");
};

template <int BLOCK_N, int EMODE = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    k_gemm_tc2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, int M, int n_tiles, int K, Epilogue epi,
               bf16* __restrict__ C, long long ldc) {
  using Cfg = Tc2Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  const int m_pairs = (M + 2 * TC_BLOCK_M - 1) / (2 * TC_BLOCK_M);
  const int total_tiles = m_pairs * n_tiles;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + Cfg::ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + Cfg::ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + TC_BLOCK_K - 1) / TC_BLOCK_K;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar + s, 1);            // leader's producer arms it; bytes of both CTAs complete it
      mbar_init(empty_bar + s, 1);           // one multicast commit per use
    }
    for (int s = 0; s < Cfg::ACC_STAGES; ++s) {
      mbar_init(tmem_full_bar + s, 1);       // one multicast commit per tile
      mbar_init(tmem_empty_bar + s, 16);     // 8 epilogue warps of each CTA (only the leader's copy is waited on)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                        // peer barriers are initialised before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const int row0 = (tile / n_tiles) * (2 * TC_BLOCK_M) + (int)rank * TC_BLOCK_M;
        const int col0 = (tile % n_tiles) * BLOCK_N + (int)rank * (BLOCK_N / 2);
        {   // pull this CTA's rows of the cluster's next tile into L2 (see the single-CTA kernel)
          const int nt = tile + n_clusters;
          if (nt < total_tiles && (nt % n_tiles) == 0 && !UU_EXP_FLAG(epi, 256)) {
            const int prow = (nt / n_tiles) * (2 * TC_BLOCK_M) + (int)rank * TC_BLOCK_M;
            for (int kb = 0; kb < num_kb; ++kb) tma_prefetch_2d(&map_a, kb * TC_BLOCK_K, prow);
          }
          if constexpr (EMODE == EMODE_RESID) {      // residual rows of this CTA's half of the next tile
            if (nt < total_tiles && !UU_EXP_FLAG(epi, 256)) {
              const int prow = (nt / n_tiles) * (2 * TC_BLOCK_M) + (int)rank * TC_BLOCK_M, pcol = (nt % n_tiles) * BLOCK_N;
              for (int r = 0; r < TC_BLOCK_M; r += 32)
                for (int c = 0; c < BLOCK_N; c += 64) tma_prefetch_2d(&map_c, pcol + c, prow + r);
            }
          }
        }
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::STAGES;
          const uint32_t ph = (it / Cfg::STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);                 // own slot free (multicast commit of the leader)
          uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + Cfg::A_BYTES;
          const uint32_t lbar = mapa_rank(full_bar + s, 0);
          if (leader) mbar_expect_tx(full_bar + s, 2 * Cfg::STAGE_BYTES);
          tma_load_2d_2sm(a_dst, &map_a, lbar, kb * TC_BLOCK_K, row0);
          tma_load_2d_2sm(b_dst, &map_b, lbar, kb * TC_BLOCK_K, col0);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues) ----------------
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * TC_BLOCK_M, BLOCK_N);
      const uint64_t a_desc0 = make_sw128_desc(smem_u32(smem));
      const uint64_t b_desc0 = make_sw128_desc(smem_u32(smem + Cfg::A_BYTES));
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++tcount) {
        const int as = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(tmem_empty_bar + as, aph ^ 1);            // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + s, ph);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t off = (uint64_t)((s * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
              umma_bf16_2sm(tmem_d, a_desc0 + off + 2 * k, b_desc0 + off + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2sm(empty_bar + s);       // frees the stage in both CTAs
            if (kb == num_kb - 1) umma_commit_2sm(tmem_full_bar + as);    // accumulator complete in both CTAs
          }
          __syncwarp();
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): TMEM -> bias/ReLU -> bf16 -> swizzled smem -> TMA store ----------------
    const int e = warp - 2;
    const int q = warp & 3;
    const int hsel = e >> 2;
    uint8_t* stage_base = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256) + 1023) & ~uintptr_t(1023));
    uint8_t* my_stage = stage_base + e * (UU_TC_EPI_NBUF * 32 * 128);
    constexpr int NSUB = BLOCK_N / 64;
    uint32_t tcount = 0, my_count = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++tcount) {
      const int row0 = (tile / n_tiles) * (2 * TC_BLOCK_M) + (int)rank * TC_BLOCK_M;
      const int col0 = (tile % n_tiles) * BLOCK_N;
      const int as = tcount % Cfg::ACC_STAGES;
      const uint32_t aph = (tcount / Cfg::ACC_STAGES) & 1;
      const int first = (hsel + (int)tcount) & 1;
      EpiPre pre;
      if (first < NSUB) epi_prefetch<EMODE>(pre, epi, row0 + q * 32, col0, first, lane, M, ldc);
      mbar_wait(tmem_full_bar + as, aph);
      tcgen05_fence_after();
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N);
      auto release = [&] {                     // hand the accumulator back to the leader's MMA thread
        tcgen05_fence_before();
        if (lane == 0) mbar_arrive_cluster(mapa_rank(tmem_empty_bar + as, 0));
      };
      if (first >= NSUB) { release(); continue; }
      epi_warp_store_tile<BLOCK_N, UU_TC_EPI_NBUF, EMODE>(tmem_acc, first, my_stage, my_count, epi, &map_c, row0 + q * 32, col0, lane, release,
                                      pre, M, C, ldc);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                        // nobody exits (or frees TMEM) while the peer may still touch it
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps and launch
// ------------------------------------------------------------------------------------------------
struct TcGemmPlan {
  CUtensorMap map_a, map_b, map_c;
  CUtensorMap map_b2;               // 2-CTA variant: box of block_n / 2 rows of W^T
  int M, N, N_pad, K, block_n;
  bool tf32 = false;                // fp32 operands, kind::tf32 (training)
  const void* c_ptr = nullptr;      // output the store map was encoded for
  long long c_ld = 0;
};

int tc_gemm_plan_create(TcGemmPlan** out, const bf16* A, long long lda, int M, int K, const bf16* Wt, int N_pad, int N) {
  UU_CHECK(M > 0 && N > 0 && K > 0 && N <= N_pad, "bad GEMM shape");
  TcGemmPlan* p = new TcGemmPlan();
  p->M = M; p->N = N; p->N_pad = N_pad; p->K = K;
  // widest tile that divides the padded N: 256 (fc1), 192 (q|k|v, 384-wide outputs), 128, 64 (heads)
  p->block_n = (N_pad % 256 == 0) ? 256 : (N_pad % 192 == 0) ? 192 : (N_pad % 128 == 0) ? 128 : 64;
#ifdef UU_EXPERIMENT
  if (const char* e = getenv("UU_GEMM_BN")) {           // tile-width experiment
    const int bn = atoi(e);
    if ((bn == 64 || bn == 128 || bn == 192 || bn == 256) && N_pad % bn == 0) p->block_n = bn;
  }
#endif
  if (N_pad % p->block_n != 0) {
    delete p;
    set_error("tcgen05 GEMM needs the packed weight rows padded to a multiple of 64");
    return 1;
  }
  if (encode_2d(&p->map_a, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TC_BLOCK_K, TC_BLOCK_M) ||
      encode_2d(&p->map_b, Wt, (uint64_t)K, (uint64_t)N_pad, (uint64_t)K, TC_BLOCK_K, (uint32_t)p->block_n)) {
    delete p;
    return 1;
  }
  if (encode_2d(&p->map_b2, Wt, (uint64_t)K, (uint64_t)N_pad, (uint64_t)K, TC_BLOCK_K, (uint32_t)p->block_n / 2)) {
    delete p;
    return 1;
  }
  *out = p;
  return 0;
}

// fp32 operands for kind::tf32: A (M, K) row-major with pitch lda, Bt (N, K) row-major with pitch ldb (= W^T, or for a
// dgrad the weight matrix itself); N % 64 == 0.  fp32 output through the generic epilogue only.
int tc_gemm_plan_create_tf32(TcGemmPlan** out, const float* A, long long lda, int M, int K, const float* Bt, long long ldb,
                             int N) {
  UU_CHECK(M > 0 && N > 0 && K > 0 && N % 64 == 0 && K % 4 == 0, "bad tf32 GEMM shape");
  TcGemmPlan* p = new TcGemmPlan();
  p->M = M; p->N = N; p->N_pad = N; p->K = K; p->tf32 = true;
  p->block_n = (N % 256 == 0) ? 256 : (N % 192 == 0) ? 192 : (N % 128 == 0) ? 128 : 64;
  if (encode_2d(&p->map_a, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TC_BLOCK_K / 2, TC_BLOCK_M, 4) ||
      encode_2d(&p->map_b, Bt, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, TC_BLOCK_K / 2, (uint32_t)p->block_n, 4)) {
    delete p;
    return 1;
  }
  p->map_c = p->map_a; p->map_b2 = p->map_b;     // unused by the generic epilogue / single-CTA kernel
  *out = p;
  return 0;
}

void tc_gemm_plan_destroy(TcGemmPlan* p) { delete p; }

static int g_num_sms = 0;

// Measured (h36m_351, one box, CUDA-graph replay): B = 512 windows per call 1.432 -> 1.365 ms (+4.9 %), B = 4096
// 9.45 -> 9.6-9.9 ms (slower), so the schedules ask for it only for small batches (pdl_set_auto); UU_PDL=0 / 1 forces it.
static bool g_pdl_auto = false;
void pdl_set_auto(bool on) { g_pdl_auto = on; }
bool pdl_enabled() {
  static int mode = -1;     // 0 off, 1 on, 2 auto
  if (mode < 0) { const char* e = getenv("UU_PDL"); mode = !e ? 2 : (e[0] == '0' ? 0 : 1); }
  return mode == 1 || (mode == 2 && g_pdl_auto);
}

template <int BLOCK_N, typename TC, bool TMA_OUT, int EMODE = 0, bool TF32 = false>
static cudaError_t tc_launch_t(const TcGemmPlan* p, const Epilogue& epi, void* C, long long ldc, cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BLOCK_N, TC, TMA_OUT, EMODE, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_tiles = p->N_pad / BLOCK_N;
  const int total = ((p->M + TC_BLOCK_M - 1) / TC_BLOCK_M) * n_tiles;
  const int grid = total < g_num_sms ? total : g_num_sms;          // persistent: one CTA per SM
  return launch_pdl(k_gemm_tc<BLOCK_N, TC, TMA_OUT, EMODE, TF32>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, p->map_a,
                    p->map_b, p->map_c, p->M, p->N, n_tiles, p->K, epi, reinterpret_cast<TC*>(C), (long long)ldc);
}

template <int BLOCK_N, int EMODE = 0>
static cudaError_t tc2_launch_t(const TcGemmPlan* p, const Epilogue& epi, void* C, long long ldc, cudaStream_t st) {
  using Cfg = Tc2Cfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc2<BLOCK_N, EMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_tiles = p->N_pad / BLOCK_N;
  const int total = ((p->M + 2 * TC_BLOCK_M - 1) / (2 * TC_BLOCK_M)) * n_tiles;
  const int clusters = total < g_num_sms / 2 ? total : g_num_sms / 2;      // persistent: one CTA pair per TPC
  return launch_pdl(k_gemm_tc2<BLOCK_N, EMODE>, dim3(2 * clusters), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, p->map_a, p->map_b2,
                    p->map_c, p->M, n_tiles, p->K, epi, reinterpret_cast<bf16*>(C), (long long)ldc);
}

static int g_use_2cta = -1;     // UU_GEMM_2CTA=0/1 forces the single-CTA / 2-CTA kernel (A/B comparison)

// bf16 output, plain row mapping, no residual / table / scatter, full 64-column sub-tiles: TMA-store epilogue
// (a positional table is allowed when the caller also asks for row statistics: EMODE_TABLE, table pitch == ldc)
static bool tma_out_eligible(const TcGemmPlan* p, const Epilogue& epi, int c_bf16, long long ldc) {
  return c_bf16 && !(epi.flags & EPI_RESIDUAL) && (!(epi.flags & EPI_ROWTABLE) || epi.stats_out) && p->N == p->N_pad &&
         (p->N % 64) == 0 && (ldc % 8) == 0 && p->block_n >= 64;
}

cudaError_t tc_gemm_launch(TcGemmPlan* p, const Epilogue& epi_in, void* C, int c_bf16, long long ldc, cudaStream_t st) {
  if ((epi_in.flags & (EPI_LNFOLD | EPI_RESID_BF16)) && !tma_out_eligible(p, epi_in, c_bf16, ldc))
    return cudaErrorInvalidValue;     // folded-LayerNorm / bf16-residual epilogues exist on the TMA-store path only
  Epilogue epi = epi_in;
#ifdef UU_EXPERIMENT
  {
    auto on = [](const char* name) { const char* e = getenv(name); return e && e[0] == '1'; };
    static const int exp_flags = (on("UU_GEMM_NOSTORE") ? 64 : 0) | (on("UU_GEMM_NOLOAD") ? 128 : 0) |
                                 (on("UU_GEMM_NOPREFETCH") ? 256 : 0) | (on("UU_GEMM_NOEPI") ? 512 : 0) |
                                 (on("UU_GEMM_NORES") ? 1024 : 0) | (getenv("UU_GEMM_ONECOMMIT") ? 4096 : 0);
    epi.flags |= exp_flags;
  }
#endif
  if (p->tf32) {
    if (c_bf16 || (epi.flags & (EPI_LNFOLD | EPI_RESID_BF16))) return cudaErrorInvalidValue;
    switch (p->block_n) {
      case 256: return tc_launch_t<256, float, false, 0, true>(p, epi, C, ldc, st);
      case 192: return tc_launch_t<192, float, false, 0, true>(p, epi, C, ldc, st);
      case 128: return tc_launch_t<128, float, false, 0, true>(p, epi, C, ldc, st);
      default: return tc_launch_t<64, float, false, 0, true>(p, epi, C, ldc, st);
    }
  }
  if (tma_out_eligible(p, epi, c_bf16, ldc)) {
    if (p->c_ptr != C || p->c_ld != ldc) {
      // rows of the output matrix: M, or with a (batch, position) row map the extent it can reach
      uint64_t c_rows = (uint64_t)p->M;
      if (epi.cmap.rpb != 0x7fffffff)
        c_rows = (uint64_t)((p->M + epi.cmap.rpb - 1) / epi.cmap.rpb) * (uint64_t)epi.cmap.batch_rows;
      if (encode_2d(&p->map_c, C, (uint64_t)p->N, c_rows, (uint64_t)ldc, 64, 32)) return cudaErrorInvalidValue;
      p->c_ptr = C; p->c_ld = ldc;
    }
    if (g_use_2cta < 0) {
      g_use_2cta = 2;                               // default (2): only where it measured faster, the K >= 768 GEMMs
#ifdef UU_EXPERIMENT
      if (const char* e = getenv("UU_GEMM_2CTA")) g_use_2cta = e[0] == '1' ? 1 : 0;
#endif
    }
    const bool scatter = epi.c_rowidx || epi.m_dev || epi.cmap.rpb != 0x7fffffff;
    const bool two_cta = !scatter && (g_use_2cta == 1 || (g_use_2cta == 2 && p->K >= 768)) && p->M >= 512;
    // folded-LayerNorm / residual epilogues (temporal blocks): 128-, 192- and 256-wide tiles only
    if (epi.flags & EPI_LNFOLD) {
      if (!epi.bias || !epi.ln_stats || !epi.ln_csum || (epi.flags & EPI_RESID_BF16)) return cudaErrorInvalidValue;
      switch (p->block_n) {
        case 256: return tc_launch_t<256, bf16, true, EMODE_LNFOLD>(p, epi, C, ldc, st);
        case 192: return tc_launch_t<192, bf16, true, EMODE_LNFOLD>(p, epi, C, ldc, st);
        case 128: return tc_launch_t<128, bf16, true, EMODE_LNFOLD>(p, epi, C, ldc, st);
        default: return cudaErrorInvalidValue;
      }
    }
    if (epi.flags & EPI_ROWTABLE) {      // bias + positional table + row statistics (the 544 -> 384 GEMM)
      if (!epi.bias || !epi.table || (epi.flags & (EPI_RELU | EPI_RESID_BF16)) || ldc != p->N) return cudaErrorInvalidValue;
      switch (p->block_n) {
        case 192: return tc_launch_t<192, bf16, true, EMODE_TABLE>(p, epi, C, ldc, st);
        case 128: return tc_launch_t<128, bf16, true, EMODE_TABLE>(p, epi, C, ldc, st);
        default: return cudaErrorInvalidValue;
      }
    }
    if (epi.flags & EPI_RESID_BF16) {
      // (in place only: the producer warp prefetches the residual through the output's tensor map)
      if (!epi.bias || epi.res_bf16 != C || scatter || (epi.flags & EPI_RELU)) return cudaErrorInvalidValue;
      switch (p->block_n) {
        case 256: return two_cta ? tc2_launch_t<256, EMODE_RESID>(p, epi, C, ldc, st) : tc_launch_t<256, bf16, true, EMODE_RESID>(p, epi, C, ldc, st);
        case 192: return two_cta ? tc2_launch_t<192, EMODE_RESID>(p, epi, C, ldc, st) : tc_launch_t<192, bf16, true, EMODE_RESID>(p, epi, C, ldc, st);
        case 128: return two_cta ? tc2_launch_t<128, EMODE_RESID>(p, epi, C, ldc, st) : tc_launch_t<128, bf16, true, EMODE_RESID>(p, epi, C, ldc, st);
        default: return cudaErrorInvalidValue;
      }
    }
    if (two_cta && (p->block_n == 256 || p->block_n == 192 || p->block_n == 128)) {
      switch (p->block_n) {
        case 256: return tc2_launch_t<256>(p, epi, C, ldc, st);
        case 192: return tc2_launch_t<192>(p, epi, C, ldc, st);
        default: return tc2_launch_t<128>(p, epi, C, ldc, st);
      }
    }
    switch (p->block_n) {
      case 256: return tc_launch_t<256, bf16, true>(p, epi, C, ldc, st);
      case 192: return tc_launch_t<192, bf16, true>(p, epi, C, ldc, st);
      case 128: return tc_launch_t<128, bf16, true>(p, epi, C, ldc, st);
      default: return tc_launch_t<64, bf16, true>(p, epi, C, ldc, st);
    }
  }
  switch (p->block_n) {
    case 256: return c_bf16 ? tc_launch_t<256, bf16, false>(p, epi, C, ldc, st) : tc_launch_t<256, float, false>(p, epi, C, ldc, st);
    case 192: return c_bf16 ? tc_launch_t<192, bf16, false>(p, epi, C, ldc, st) : tc_launch_t<192, float, false>(p, epi, C, ldc, st);
    case 128: return c_bf16 ? tc_launch_t<128, bf16, false>(p, epi, C, ldc, st) : tc_launch_t<128, float, false>(p, epi, C, ldc, st);
    default: return c_bf16 ? tc_launch_t<64, bf16, false>(p, epi, C, ldc, st) : tc_launch_t<64, float, false>(p, epi, C, ldc, st);
  }
}

}  // namespace uu

#include "mlp_tc.cuh"
