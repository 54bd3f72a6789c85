// Fused transformer MLP of the temporal blocks (vit:190-195: x += fc2(ReLU(fc1(LN2(x)))) — the reference's MLP of the
// temporal blocks uses ReLU, net:262-268) as ONE tcgen05 kernel: the 768-wide hidden activation never leaves the SM.
//
// A cluster of two CTAs (cta_group::2) owns 256 rows of the bf16 residual stream X:
//   * TMA drops each CTA's 128 x 384 row block into shared memory once (6 k-blocks, 128B swizzle);
//   * the hidden layer is produced in chunks of 64 columns: fc1 chunk j = X . (gamma (.) W1)^T[64 j .. 64 j + 64) runs as
//     tcgen05.mma M = 256, N = 64 into one of two TMEM accumulators (LayerNorm folded into the weights);
//   * four epilogue warps per CTA finish the LayerNorm (rstd * acc - rstd * mean * csum + b'), apply ReLU, round to bf16
//     and write the chunk as a 128 x 64 K-major operand tile (128B-swizzle layout) into one of two shared-memory buffers;
//   * fc2 consumes that tile straight from shared memory: acc2[256 x 384] += H_j . W2^T[:, 64 j ..) as two N = 192 MMAs
//     per 16-column k-step; fc2 of chunk j - 1 is issued behind fc1 of chunk j, so the tensor pipe works while the
//     chunk epilogue runs;
//   * eight more epilogue warps per CTA drain acc2 (two 192-column halves with their own barriers): + bias + residual
//     row, LayerNorm statistics of the result, bf16, TMA store in place (the EMODE_RESID epilogue of gemm_tc.cu).
// Weights stream through two small TMA rings; with the CTA pair every weight tile is fetched from L2 once per 256 rows
// (each CTA loads half of its rows), which is what keeps the rings (60 KB) ahead of the tensor pipe.
// TMEM: acc2 columns [0, 384), acc1 buffers [384, 448) and [448, 512).
#pragma once

namespace uu {

constexpr int ML_D = 384, ML_KB = ML_D / 64, ML_NC = 64;
constexpr int ML_THREADS = 512;          // warp 0 TMA (X, W1), warp 1 fc1 MMA, warps 2..5 chunk epilogue, 6..13 output epilogue,
                                         // warp 14 fc2 MMA, warp 15 TMA (W2)
constexpr int ML_W1_STAGES = 2, ML_W2_STAGES = 3;     // a W1 stage = the chunk's six k-block tiles (one barrier, one wait)
constexpr int ML_XKB_BYTES = 128 * 128;  // one k-block of this CTA's rows
constexpr int ML_W1_KB = 32 * 128;       // this CTA's 32 of the chunk's 64 weight rows, one k-block
constexpr int ML_W1_SLOT = ML_KB * ML_W1_KB;
constexpr int ML_W2_SLOT = 96 * 128;     // this CTA's 96 of a 192-row half of W2^T
constexpr int ML_OFF_W1 = ML_KB * ML_XKB_BYTES;
constexpr int ML_OFF_W2 = ML_OFF_W1 + ML_W1_STAGES * ML_W1_SLOT;
constexpr int ML_OFF_STG = ML_OFF_W2 + ML_W2_STAGES * ML_W2_SLOT;
constexpr int ML_OFF_PAR = ML_OFF_STG + 8 * 32 * 128;     // fp32 csum1[h] | bias1[h] (h <= ML_MAX_H)
constexpr int ML_MAX_H = 768;
constexpr int ML_OFF_BAR = ML_OFF_PAR + 2 * ML_MAX_H * 4;
constexpr int ML_SMEM_BYTES = ML_OFF_BAR + 512 + 1024;
static_assert(ML_SMEM_BYTES <= 227 * 1024, "fused MLP shared memory budget");

__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
// "buffer drained" arrivals only order tensor-memory reads (tcgen05.fence::before_thread_sync), not memory: relaxed
// semantics keep the MEMBAR + ERRBAR of a release out of the epilogue loops
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// "operand tile written" arrival on the leader's barrier: release at CTA scope (the PTX default, what CUTLASS's
// ClusterBarrier::arrive(cta_id) emits); the writes were already pushed to the async proxy by fence.proxy.async.  The
// .release.cluster form compiles to MEMBAR.ALL.CTA + ERRBAR and was a third of the chunk epilogue's time.
__device__ __forceinline__ void mbar_arrive_remote_cta_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ML_THREADS, 1)
    k_mlp_tc2(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
              const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_out, MlpArgs a) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ML_OFF_BAR);
  uint64_t* x_full = bars;                       // [6]
  uint64_t* x_empty = x_full + ML_KB;            // [6]
  uint64_t* w1_full = x_empty + ML_KB;           // [6]
  uint64_t* w1_empty = w1_full + ML_W1_STAGES;   // [6]
  uint64_t* w2_full = w1_empty + ML_W1_STAGES;   // [3]
  uint64_t* w2_empty = w2_full + ML_W2_STAGES;   // [3]
  uint64_t* acc1_full = w2_empty + ML_W2_STAGES; // [2]
  uint64_t* acc1_empty = acc1_full + 2;          // [2]
  uint64_t* h_full = acc1_empty + 2;             // [2]
  uint64_t* h_empty = h_full + 2;                // [2]
  uint64_t* acc2_full = h_empty + 2;             // [2]
  uint64_t* acc2_empty = acc2_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int total_tiles = (a.M + 255) / 256;
  const int n_chunks = a.n_chunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    for (int i = 0; i < ML_KB; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 1); }
    for (int i = 0; i < ML_W1_STAGES; ++i) { mbar_init(w1_full + i, 1); mbar_init(w1_empty + i, 1); }
    for (int i = 0; i < ML_W2_STAGES; ++i) { mbar_init(w2_full + i, 1); mbar_init(w2_empty + i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc1_full + i, 1);      // multicast commit
      mbar_init(acc1_empty + i, 1);     // multicast commit of fc2: the hidden chunk held in this accumulator has been consumed
      mbar_init(h_full + i, 8);         // 4 chunk-epilogue warps of each CTA (leader's copy is the one waited on)
      mbar_init(h_empty + i, 1);        // (unused)
      mbar_init(acc2_full + i, 1);      // multicast commit
      mbar_init(acc2_empty + i, 16);    // 8 output-epilogue warps of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  float* s_csum = reinterpret_cast<float*>(smem + ML_OFF_PAR);
  float* s_bias = s_csum + ML_MAX_H;
  if (warp >= 2 && warp < 6) {        // folded-LayerNorm column constants of fc1, read by the chunk epilogue
    for (int i = threadIdx.x - 64; i < n_chunks * ML_NC; i += 128) { s_csum[i] = a.csum1[i]; s_bias[i] = a.bias1[i]; }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }

  if (warp == 0) {
    // ---------------- TMA producer 1 (both CTAs): X row block and the fc1 weight tiles ----------------
    // (fc2's weight tiles have their own producer warp: one thread walking both rings blocked on a full W2 ring while
    // W1 slots were free, and the 4 KB W1 tiles need many loads in flight to cover the L2 latency)
    if (lane == 0) {
      int s1 = 0;
      uint32_t ph1 = 0, tcnt = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++tcnt) {
        const int row0 = tile * 256 + (int)rank * 128;
        for (int kb = 0; kb < ML_KB; ++kb) {
          mbar_wait(x_empty + kb, (tcnt & 1) ^ 1);
          if (leader) mbar_expect_tx(x_full + kb, 2 * ML_XKB_BYTES);
          tma_load_2d_2sm(smem + kb * ML_XKB_BYTES, &map_x, mapa_rank(x_full + kb, 0), kb * 64, row0);
        }
        if (tile + n_clusters < total_tiles)           // the next row block of this CTA: pull it into L2 meanwhile
          for (int kb = 0; kb < ML_KB; ++kb) tma_prefetch_2d(&map_x, kb * 64, row0 + n_clusters * 256);
        for (int j = 0; j < n_chunks; ++j) {
          mbar_wait(w1_empty + s1, ph1 ^ 1);
          if (leader) mbar_expect_tx(w1_full + s1, 2 * ML_W1_SLOT);
          for (int kb = 0; kb < ML_KB; ++kb)
            tma_load_2d_2sm(smem + ML_OFF_W1 + s1 * ML_W1_SLOT + kb * ML_W1_KB, &map_w1, mapa_rank(w1_full + s1, 0), kb * 64,
                            j * ML_NC + (int)rank * 32);
          if (++s1 == ML_W1_STAGES) { s1 = 0; ph1 ^= 1; }
        }
      }
    }
  } else if (warp == 15) {
    // ---------------- TMA producer 2 (both CTAs): the fc2 weight tiles ----------------
    if (lane == 0) {
      int s2 = 0;
      uint32_t ph2 = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters)
        for (int j = 0; j < n_chunks; ++j)
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(w2_empty + s2, ph2 ^ 1);
            if (leader) mbar_expect_tx(w2_full + s2, 2 * ML_W2_SLOT);
            tma_load_2d_2sm(smem + ML_OFF_W2 + s2 * ML_W2_SLOT, &map_w2, mapa_rank(w2_full + s2, 0), j * ML_NC,
                            hf * 192 + (int)rank * 96);
            if (++s2 == ML_W2_STAGES) { s2 = 0; ph2 ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ---------------- fc1 MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) ----------------
    // fc1 and fc2 are issued by two different warps: with 64-column chunks an MMA retires in 32 cycles, and one warp
    // walking both loops (barrier polls, elect, commits) was the pacer of the kernel.  (Splitting fc1 itself over two
    // issuers — even / odd chunks — measured no further gain: the kernel is then bound by shared-memory bandwidth, the
    // N = 64 MMAs re-read their 4 KB A slice every 32 cycles.)
    if (leader) {
      constexpr uint32_t idesc1 = make_idesc_bf16(256, ML_NC);
      const uint64_t x_desc0 = make_sw128_desc(smem_u32(smem));
      const uint64_t w1_desc0 = make_sw128_desc(smem_u32(smem + ML_OFF_W1));
      int s1 = 0;
      uint32_t ph1 = 0, c1 = 0, tcnt = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++tcnt) {
        for (int kb = 0; kb < ML_KB; ++kb) mbar_wait(x_full + kb, tcnt & 1);
        for (int j = 0; j < n_chunks; ++j, ++c1) {
          const int b = c1 & 1;
          mbar_wait(acc1_empty + b, ((c1 >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          mbar_wait(w1_full + s1, ph1);                  // the chunk's six weight tiles (one barrier)
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t bd0 = w1_desc0 + (uint64_t)((s1 * ML_W1_SLOT) >> 4);
#pragma unroll
            for (int kb = 0; kb < ML_KB; ++kb) {
              const uint64_t ad = x_desc0 + (uint64_t)((kb * ML_XKB_BYTES) >> 4);
              const uint64_t bd = bd0 + (uint64_t)((kb * ML_W1_KB) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(tmem_base + (uint32_t)(ML_D + b * ML_NC), ad + 2 * k, bd + 2 * k, idesc1, (kb | k) != 0);
              if (j == n_chunks - 1) umma_commit_2sm(x_empty + kb);      // the row block may be refilled
            }
            umma_commit_2sm(w1_empty + s1);
            umma_commit_2sm(acc1_full + b);
          }
          __syncwarp();
          if (++s1 == ML_W1_STAGES) { s1 = 0; ph1 ^= 1; }
        }
      }
    }
  } else if (warp == 14) {
    // ---------------- fc2 MMA issuer (leader CTA): acc2 += H_j . W2^T[:, 64 j ..) as two 192-column halves ----------------
    if (leader) {
      constexpr uint32_t idesc2 = make_idesc_bf16(256, 192);
      const uint64_t w2_desc0 = make_sw128_desc(smem_u32(smem + ML_OFF_W2));
      int s2 = 0;
      uint32_t ph2 = 0, c2 = 0, tcnt = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++tcnt) {
        for (int j = 0; j < n_chunks; ++j, ++c2) {
          const int b = c2 & 1;
          mbar_wait(h_full + b, (c2 >> 1) & 1);
          tcgen05_fence_after();
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {
            if (j == 0) {               // the previous tile's output epilogue has drained this half
              mbar_wait(acc2_empty + hf, (tcnt & 1) ^ 1);
              tcgen05_fence_after();
            }
            mbar_wait(w2_full + s2, ph2);
            tcgen05_fence_after();
            if (elect_one()) {
              // A = the hidden chunk in TENSOR MEMORY (bf16 pairs over the fc1 accumulator, 8 columns per 16-wide k-step)
              const uint32_t ta = tmem_base + (uint32_t)(ML_D + b * ML_NC);
              const uint64_t bd = w2_desc0 + (uint64_t)((s2 * ML_W2_SLOT) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ts_2sm(tmem_base + (uint32_t)(hf * 192), ta + 8 * k, bd + 2 * k, idesc2, (j | k) != 0);
              umma_commit_2sm(w2_empty + s2);
              if (hf == 1) umma_commit_2sm(acc1_empty + b);       // accumulator / hidden chunk free for fc1 of chunk j + 2
              if (j == n_chunks - 1) umma_commit_2sm(acc2_full + hf);
            }
            __syncwarp();
            if (++s2 == ML_W2_STAGES) { s2 = 0; ph2 ^= 1; }
          }
        }
      }
    }
  } else if (warp < 6) {
    // ---------------- chunk epilogue (4 warps per CTA): acc1 -> LayerNorm fold, ReLU, bf16 -> H tile ----------------
    const int q = warp & 3;
    uint32_t c = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      const int my_row = tile * 256 + (int)rank * 128 + q * 32 + lane;
      float mu = 0.f, rstd = 1.f;
      {
        float s1 = 0.f, s2 = 0.f;
        if (my_row < a.M) {
          const float2* sp = reinterpret_cast<const float2*>(a.ln_stats) + (long long)my_row * a.ln_slots;
          for (int i = 0; i < a.ln_slots; ++i) { const float2 t = sp[i]; s1 += t.x; s2 += t.y; }
        }
        mu = s1 * a.ln_inv_k;
        rstd = rsqrtf(fmaxf(s2 * a.ln_inv_k - mu * mu, 0.f) + a.ln_eps);
      }
      const float2 a1 = make_float2(rstd, rstd), a2 = make_float2(-rstd * mu, -rstd * mu);
      for (int j = 0; j < n_chunks; ++j, ++c) {
        const int b = c & 1;
        const uint32_t use = (c >> 1) & 1;
        mbar_wait(acc1_full + b, use);
        tcgen05_fence_after();
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ML_D + b * ML_NC);
        tmem_ld_32x32b_x32(taddr, v0);
        tmem_ld_32x32b_x32(taddr + 32, v1);
        const int cb0 = j * ML_NC;
        uint32_t hk[32];                     // the thread's 64 hidden values as bf16 pairs
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int cb = cb0 + 8 * g;
          float2 o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = g < 4 ? make_float2(__uint_as_float(v0[8 * g + 2 * i]), __uint_as_float(v0[8 * g + 2 * i + 1]))
                         : make_float2(__uint_as_float(v1[8 * (g - 4) + 2 * i]), __uint_as_float(v1[8 * (g - 4) + 2 * i + 1]));
          const float4 c0 = *reinterpret_cast<const float4*>(s_csum + cb);
          const float4 c1 = *reinterpret_cast<const float4*>(s_csum + cb + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + cb);
          const float4 b1 = *reinterpret_cast<const float4*>(s_bias + cb + 4);
          o[0] = ffma2(a1, o[0], ffma2(a2, make_float2(c0.x, c0.y), make_float2(b0.x, b0.y)));
          o[1] = ffma2(a1, o[1], ffma2(a2, make_float2(c0.z, c0.w), make_float2(b0.z, b0.w)));
          o[2] = ffma2(a1, o[2], ffma2(a2, make_float2(c1.x, c1.y), make_float2(b1.x, b1.y)));
          o[3] = ffma2(a1, o[3], ffma2(a2, make_float2(c1.z, c1.w), make_float2(b1.z, b1.w)));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 pb = __floats2bfloat162_rn(fmaxf(o[i].x, 0.f), fmaxf(o[i].y, 0.f));
            hk[4 * g + i] = *reinterpret_cast<uint32_t*>(&pb);
          }
        }
        // the hidden chunk goes back into TENSOR MEMORY over the accumulator it came from (columns 0 .. 31 of the 64): fc2
        // reads it as its A operand from there, so the hidden activation touches neither shared memory nor HBM
        tmem_st_32x32b_x32(taddr, hk);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_rank(h_full + b, 0));
      }
    }
  } else if (warp < 14) {
    // ---------------- output epilogue (8 warps per CTA): acc2 + b2 + residual -> statistics, bf16, TMA store ----------------
    const int e = warp - 6;
    const int q = warp & 3;
    const int hsel = e >> 2;
    uint8_t* my_stage = smem + ML_OFF_STG + e * (32 * 128);
    uint32_t tcnt = 0, my_count = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++tcnt) {
      const int row0 = tile * 256 + (int)rank * 128;
      for (int hf = 0; hf < 2; ++hf) {
        const int first = (hsel + hf) & 1;
        EpiPre pre;
        epi_prefetch<EMODE_RESID>(pre, a.epi2, row0 + q * 32, hf * 192, first, lane, a.M, a.ldx);
        mbar_wait(acc2_full + hf, tcnt & 1);
        tcgen05_fence_after();
        const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 192);
        auto release = [&] {
          tcgen05_fence_before();
          if (lane == 0) mbar_arrive_cluster_relaxed(mapa_rank(acc2_empty + hf, 0));
        };
        epi_warp_store_tile<192, 1, EMODE_RESID>(tmem_acc, first, my_stage, my_count, a.epi2, &map_out, row0 + q * 32, hf * 192,
                                                 lane, release, pre, a.M, a.X, a.ldx);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

struct MlpPlan {
  CUtensorMap map_x, map_w1, map_w2, map_out;
  int M, h;
};

// X [M, 384] bf16 (pitch ldx), W1t = bf16((gamma (.) W1)^T) [h, 384], W2t = bf16(W2^T) [384, h]
int mlp_plan_create(MlpPlan** out, const bf16* X, long long ldx, int M, int d, int h, const bf16* W1t, const bf16* W2t) {
  UU_CHECK(d == ML_D && h % (2 * ML_NC) == 0 && h >= 2 * ML_NC && h <= ML_MAX_H && M > 0, "fused MLP kernel: width 384, hidden multiple of 64 (<= 768)");
  MlpPlan* p = new MlpPlan();
  p->M = M; p->h = h;
  if (encode_2d(&p->map_x, X, (uint64_t)d, (uint64_t)M, (uint64_t)ldx, 64, 128) ||
      encode_2d(&p->map_w1, W1t, (uint64_t)d, (uint64_t)h, (uint64_t)d, 64, 32) ||
      encode_2d(&p->map_w2, W2t, (uint64_t)h, (uint64_t)d, (uint64_t)h, 64, 96) ||
      encode_2d(&p->map_out, X, (uint64_t)d, (uint64_t)M, (uint64_t)ldx, 64, 32)) {
    delete p;
    return 1;
  }
  *out = p;
  return 0;
}
void mlp_plan_destroy(MlpPlan* p) { delete p; }

cudaError_t mlp_launch(const MlpPlan* p, const MlpArgs& args, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int total = (p->M + 255) / 256;
  const int clusters = total < g_num_sms / 2 ? total : g_num_sms / 2;
  return launch_pdl(k_mlp_tc2, dim3(2 * clusters), dim3(ML_THREADS), ML_SMEM_BYTES, st, p->map_x, p->map_w1, p->map_w2, p->map_out, args);
}

}  // namespace uu
