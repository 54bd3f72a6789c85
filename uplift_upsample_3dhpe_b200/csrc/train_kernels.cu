// Kernels of the training step (train.py:464-506): generic fp32 GEMM with transposed operands and
// split-K (dgrad / wgrad), reductions for bias / positional-encoding / token gradients, LayerNorm,
// attention and activation backward, the MPJPE loss with its gradient, stochastic depth, and the fused
// multi-tensor AdamW (tfa.optimizers.AdamW semantics) + EMA update.
// The reference relies on TensorFlow autodiff for all of this; every formula below
// is the hand-derived adjoint of the forward op it names.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "train.cuh"

namespace uu {

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- deterministic cross-CTA reductions ---------------------------------------------------------------------------
// Gradients that sum over rows (biases, LayerNorm gamma / beta, positional tables, split-K weight gradients) are reduced
// in two passes: every CTA writes its partial sums into a scratch slab [part][n], k_reduce_partials adds the slabs to the
// gradient in slab order.  No floating-point atomics: two runs of a step give bit-identical gradients.
static float* g_red_scratch = nullptr;
static size_t g_red_floats = 0;
void train_reduce_scratch(float* p, size_t n_floats) { g_red_scratch = p; g_red_floats = n_floats; }
size_t train_reduce_scratch_floats() { return (size_t)6 << 20; }

// out[i] += sum_p part[p * n + i]; elements i >= n0 go to out1[i - n0] (two destination tensors, e.g. gamma | beta).
// A CTA owns 32 elements; RP_ROWS thread rows each sum every RP_ROWS-th slab in slab order, then the sums are added in row order:
// the association is fixed by (nparts), never by timing.
constexpr int RP_ROWS = 32;      // slab rows summed in parallel by one CTA (1024 threads: short dependent chains)
__device__ __forceinline__ void reduce_chunk(const float* __restrict__ part, int nparts, long long n, float* __restrict__ out0,
                                             long long n0, float* __restrict__ out1, long long chunk, float (*sm)[32]) {
  const int e = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const long long i = chunk * 32 + e;
  float s = 0.f;
  if (i < n)
    for (int p = pl; p < nparts; p += RP_ROWS) s += part[(long long)p * n + i];
  sm[pl][e] = s;
  __syncthreads();
  if (pl == 0 && i < n) {
#pragma unroll
    for (int q = 1; q < RP_ROWS; ++q) s += sm[q][e];
    float* dst = i < n0 ? out0 + i : out1 + (i - n0);
    *dst += s;
  }
}
__global__ void __launch_bounds__(32 * RP_ROWS) k_reduce_partials(const float* __restrict__ part, int nparts, long long n,
                                                         float* __restrict__ out0, long long n0, float* __restrict__ out1) {
  __shared__ float sm[RP_ROWS][32];
  reduce_chunk(part, nparts, n, out0, n0, out1, blockIdx.x, sm);
}

// Deferred mode (the backward pass of a training step): every producer takes a FRESH region of a large arena and its
// second pass is only recorded; train_reduce_flush() then runs all recorded reductions in ONE launch (a step had ~90
// second-pass launches of ~5 us each).  Same per-element association as the immediate form.  One process per GPU: the state
// below is per process.
struct RedJob {
  const float* part; float* out0; float* out1;
  long long n, n0;
  int nparts, chunk0;
};
__global__ void __launch_bounds__(32 * RP_ROWS) k_reduce_jobs(const RedJob* __restrict__ jobs, int njobs) {
  __shared__ float sm[RP_ROWS][32];
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {                                   // last job whose first chunk is <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].chunk0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const RedJob j = jobs[lo];
  reduce_chunk(j.part, j.nparts, j.n, j.out0, j.n0, j.out1, (long long)blockIdx.x - j.chunk0, sm);
}
constexpr int RED_MAX_JOBS = 256, RED_SLOTS = 64;
static bool g_defer = false;
static float* g_arena = nullptr;
static size_t g_arena_floats = 0, g_arena_used = 0;
static std::vector<RedJob> g_jobs;
static int g_job_chunks = 0, g_job_slot = 0;
static RedJob *g_jobs_host = nullptr, *g_jobs_dev = nullptr;      // RED_SLOTS x RED_MAX_JOBS, pinned / device
cudaError_t train_reduce_flush(cudaStream_t st) {
  g_arena_used = 0;                                   // regions are re-used in stream order behind the reduction
  if (g_jobs.empty()) return cudaSuccess;
  if (!g_jobs_host) {
    cudaError_t e = cudaMallocHost(&g_jobs_host, sizeof(RedJob) * RED_SLOTS * RED_MAX_JOBS);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&g_jobs_dev, sizeof(RedJob) * RED_SLOTS * RED_MAX_JOBS);
    if (e != cudaSuccess) return e;
  }
  // a slot is re-used RED_SLOTS = 64 flushes later.  A step flushes about six times, and the host cannot run further ahead of
  // the device than the driver's launch queue (~1000 launches, two steps), so the copy that last read the slot has completed
  const int slot = g_job_slot++ % RED_SLOTS, nj = (int)g_jobs.size();
  RedJob* h = g_jobs_host + (size_t)slot * RED_MAX_JOBS;
  RedJob* dv = g_jobs_dev + (size_t)slot * RED_MAX_JOBS;
  std::copy(g_jobs.begin(), g_jobs.end(), h);
  cudaError_t e = cudaMemcpyAsync(dv, h, sizeof(RedJob) * nj, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  k_reduce_jobs<<<(unsigned)g_job_chunks, 32 * RP_ROWS, 0, st>>>(dv, nj);
  g_jobs.clear();
  g_job_chunks = 0;
  return cudaGetLastError();
}
void train_reduce_defer_begin(float* arena, size_t arena_floats) {
  g_defer = true; g_arena = arena; g_arena_floats = arena_floats; g_arena_used = 0;
  g_jobs.clear(); g_job_chunks = 0;
}
cudaError_t train_reduce_defer_end(cudaStream_t st) {
  cudaError_t e = train_reduce_flush(st);
  g_defer = false;
  return e;
}
void train_reduce_defer_abort() { g_defer = false; g_jobs.clear(); g_job_chunks = 0; g_arena_used = 0; }
// scratch region for n floats of partial slabs (immediate mode: the one shared scratch buffer)
static float* red_region(size_t n, cudaStream_t st) {
  if (!g_defer) return n <= g_red_floats + 4096 ? g_red_scratch : nullptr;
  n = (n + 31) & ~(size_t)31;
  if (n > g_arena_floats) return nullptr;
  if (g_arena_used + n > g_arena_floats && train_reduce_flush(st) != cudaSuccess) return nullptr;
  float* p = g_arena + g_arena_used;
  g_arena_used += n;
  return p;
}
static cudaError_t reduce_partials(const float* part, int nparts, long long n, float* out0, long long n0, float* out1,
                                   cudaStream_t st) {
  if (g_defer) {
    RedJob j;
    j.part = part; j.out0 = out0; j.out1 = out1; j.n = n; j.n0 = n0; j.nparts = nparts; j.chunk0 = g_job_chunks;
    g_jobs.push_back(j);
    g_job_chunks += (int)((n + 31) / 32);
    return (int)g_jobs.size() >= RED_MAX_JOBS ? train_reduce_flush(st) : cudaSuccess;
  }
  k_reduce_partials<<<(unsigned)((n + 31) / 32), 32 * RP_ROWS, 0, st>>>(part, nparts, n, out0, n0, out1);
  return cudaGetLastError();
}

// =================================================================================================
// Generic fp32 GEMM: C[crow(r)][n] (+)= sum_k opA(r,k) * opB(k,n)   (+ bias[n])
//   opA(r,k) = TA ? A[k*lda + r] : A[arow(r)*lda + k]   (arow via RowMap, dropped rows read as 0)
//   opB(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
//   split-K over gridDim.z into partial slabs reduced in split order (used for weight gradients).
// =================================================================================================
// BM x BN x 16 tiles, 256 threads, 8 x TN outputs per thread (4-wide strips so that every shared-memory read is a
// conflict-free float4), 16-byte global loads with a scalar fallback for unaligned / ragged edges, next tile
// prefetched into registers while the current one is multiplied.  Tile shapes: 128 x 128 (TN 8) for the d = 384
// layers, 128 x 64 and 256 x 32 (TN 4) for the 32 / 64-wide spatial layers and the 51-wide heads, so narrow outputs
// do not pay for columns that do not exist.
constexpr int GG_K = 16;

// One operand tile of RC rows/cols x 16 k.  "Along K" operands (A row-major, B^T) are transposed on the way into
// shared memory; "along M/N" operands (A^T, B row-major) are stored as they are loaded.
template <bool ALONG_K, int RC>
struct GgLoader {
  static constexpr int NV = (RC * GG_K / 4 + 255) / 256;          // float4 per thread (RC = 32: half the threads idle)
  static constexpr int ACTIVE = RC * GG_K / 4 < 256 ? RC * GG_K / 4 : 256;
  static constexpr int PER_ROW = RC / 4;                           // float4 per k-row (along M/N)
  float4 v[NV];
  // ALONG_K: element (rc, k) at base[map(rc)*ld + k]; else element (k, rc) at base[k*ld + rc]
  __device__ __forceinline__ void load(const float* __restrict__ base, long long ld, const RowMap* map, int rc0, int rc_end,
                                       int k0, int kend, bool vec_ok, int tid) {
#pragma unroll
    for (int h = 0; h < NV; ++h) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      if (tid < ACTIVE) {
        if constexpr (ALONG_K) {
          const int rc = rc0 + (tid >> 2) + 64 * h, k = k0 + (tid & 3) * 4;
          if (rc < rc_end && k < kend) {
            const long long rr = map ? map_row(*map, rc) : rc;
            if (rr >= 0) {
              const float* src = base + rr * ld + k;
              if (k + 3 < kend && vec_ok) {
                const float4 q = *reinterpret_cast<const float4*>(src);
                t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
              } else {
                for (int i = 0; i < 4 && k + i < kend; ++i) t[i] = src[i];
              }
            }
          }
        } else {
          const int k = k0 + tid / PER_ROW + (256 / PER_ROW) * h, rc = rc0 + (tid % PER_ROW) * 4;
          if (k < kend && rc < rc_end) {
            const float* src = base + (long long)k * ld + rc;
            if (rc + 3 < rc_end && vec_ok) {
              const float4 q = *reinterpret_cast<const float4*>(src);
              t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
            } else {
              for (int i = 0; i < 4 && rc + i < rc_end; ++i) t[i] = src[i];
            }
          }
        }
      }
      v[h] = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
  __device__ __forceinline__ void store(float (*S)[RC + 4], int tid) const {
    if (tid >= ACTIVE) return;
#pragma unroll
    for (int h = 0; h < NV; ++h) {
      if constexpr (ALONG_K) {
        const int rc = (tid >> 2) + 64 * h, k = (tid & 3) * 4;
        S[k][rc] = v[h].x; S[k + 1][rc] = v[h].y; S[k + 2][rc] = v[h].z; S[k + 3][rc] = v[h].w;
      } else {
        *reinterpret_cast<float4*>(&S[tid / PER_ROW + (256 / PER_ROW) * h][(tid % PER_ROW) * 4]) = v[h];
      }
    }
  }
};

template <bool TA, bool TB, int BM, int BN>
__global__ void __launch_bounds__(256, 2) k_gemm_gen(GemmGen g) {
  constexpr int TM = 8, TN = BM * BN / (TM * 256), NTX = BN / TN;
  static_assert(TN == 4 || TN == 8, "thread tile is 8 x 4 or 8 x 8");
  __shared__ __align__(16) float As[GG_K][BM + 4];
  __shared__ __align__(16) float Bs[GG_K][BN + 4];
  const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  const int tid = threadIdx.x, tx = tid % NTX, ty = tid / NTX;
  const int kchunk = (((g.K + gridDim.z - 1) / gridDim.z) + GG_K - 1) / GG_K * GG_K;
  const int kbeg = blockIdx.z * kchunk, kend = min(g.K, kbeg + kchunk);
  const bool avec = (g.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;   // 16-byte loads allowed
  const bool bvec = (g.ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0;
  const bool amapped = !TA && g.amap.rpb != 0x7fffffff;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  GgLoader<!TA, BM> la;     // A row-major [M, K] is read along K; A^T [K, M] along M
  GgLoader<TB, BN> lb;      // B^T [N, K] is read along K; B row-major [K, N] along N
  if (kbeg < kend) {
    la.load(g.A, g.lda, amapped ? &g.amap : nullptr, row0, g.M, kbeg, kend, avec, tid);
    lb.load(g.B, g.ldb, nullptr, col0, g.N, kbeg, kend, bvec, tid);
  }
  for (int k0 = kbeg; k0 < kend; k0 += GG_K) {
    la.store(As, tid);
    lb.store(Bs, tid);
    __syncthreads();
    if (k0 + GG_K < kend) {          // prefetch the next tile while this one is multiplied
      la.load(g.A, g.lda, amapped ? &g.amap : nullptr, row0, g.M, k0 + GG_K, kend, avec, tid);
      lb.load(g.B, g.ldb, nullptr, col0, g.N, k0 + GG_K, kend, bvec, tid);
    }
#pragma unroll
    for (int kk = 0; kk < GG_K; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][BM / 2 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
      if constexpr (TN == 8) {
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][BN / 2 + tx * 4]);
        bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = row0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
    if (r >= g.M) continue;
    const long long cr = map_row(g.cmap, r);
    if (cr < 0) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = col0 + (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4));
      if (c >= g.N) continue;
      float v = acc[i][j];
      float* dst = g.C + cr * g.ldc + c;
      if (gridDim.z > 1) {       // split-K: partial slab [split][M][N] (plain output rows, ldc == N), reduced in split order
        g.partial[((long long)blockIdx.z * g.M + r) * g.N + c] = v;
      } else {
        if (g.bias) v += g.bias[c];
        if (g.relu) v = fmaxf(v, 0.f);
        *dst = g.accumulate ? *dst + v : v;
      }
    }
  }
}

// ---- narrow weight gradients (spatial blocks: 32 / 64-wide layers over B * n_tok * 17 rows) ----------------------
// dW[KD][ND] += X^T dY and db[ND] += colsum(dY) in ONE pass over the rows: these products are memory-bound (one read of
// X and dY), so every CTA streams a contiguous row range through shared memory, keeps its KD x ND partial sums in
// registers (thread (k, n-chunk) owns KD/32 x ND/8 of them) and writes one partial slab; slabs are reduced in order.
// A lane's ND/8 columns are 4-wide chunks interleaved over the 8 lanes (chunk c of lane ln = columns 32 c + 4 ln ..): the
// 16-byte reads of a quarter-warp are then 128 contiguous bytes (lane-contiguous column blocks of 8 were a 2-way conflict).
template <int KD, int ND>
__global__ void __launch_bounds__(256) k_wgrad_skinny(const float* __restrict__ X, long long ldx, const float* __restrict__ dY,
                                                      long long ldy, long long rows, float* __restrict__ partial) {
  // Every warp takes every 8th row of a 64-row tile and accumulates the WHOLE KD x ND product of its rows: lane
  // (lk, ln) = (lane / 8, lane % 8) owns the KD/4 x ND/8 block at (lk * KD/4, ln * ND/8), so a row costs two or three
  // float4 shared-memory reads per operand for KD/4 * ND/8 FMAs.  The 8 warps are summed in warp order at the end.
  constexpr int TR = 64, KPL = KD / 4, NPL = ND / 8;
  __shared__ __align__(16) float Xs[TR][KD];
  __shared__ __align__(16) float Ys[TR][ND];
  __shared__ __align__(16) float Rs[KD * ND + ND];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lk = lane >> 3, ln = lane & 7;
  const long long per = ((rows + gridDim.x - 1) / gridDim.x + TR - 1) / TR * TR;
  const long long r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  float acc[KPL][NPL], bs[NPL];
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    bs[j] = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) acc[i][j] = 0.f;
  }
  // register-staged, software-pipelined tile loads: the loads of tile t + 1 are in flight while tile t is multiplied (a
  // rolled load loop made every iteration wait for its own DRAM latency: the kernel ran at 1 - 2 TB/s)
  constexpr int ITX = TR * KD / 4 / 256, ITY = TR * ND / 4 / 256;
  static_assert(TR * KD / 4 % 256 == 0 && TR * ND / 4 % 256 == 0, "tile loads are whole passes of the CTA");
  float4 xr[ITX], yr[ITY];
  auto fetch = [&](long long t0) {
#pragma unroll
    for (int u = 0; u < ITX; ++u) {
      const int i = tid + 256 * u, r = i / (KD / 4), c = (i - r * (KD / 4)) * 4;
      xr[u] = (t0 + r < r1) ? *reinterpret_cast<const float4*>(X + (t0 + r) * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < ITY; ++u) {
      const int i = tid + 256 * u, r = i / (ND / 4), c = (i - r * (ND / 4)) * 4;
      yr[u] = (t0 + r < r1) ? *reinterpret_cast<const float4*>(dY + (t0 + r) * ldy + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (r0 < r1) fetch(r0);
  for (long long t0 = r0; t0 < r1; t0 += TR) {
#pragma unroll
    for (int u = 0; u < ITX; ++u) {
      const int i = tid + 256 * u, r = i / (KD / 4), c = (i - r * (KD / 4)) * 4;
      *reinterpret_cast<float4*>(&Xs[r][c]) = xr[u];
    }
#pragma unroll
    for (int u = 0; u < ITY; ++u) {
      const int i = tid + 256 * u, r = i / (ND / 4), c = (i - r * (ND / 4)) * 4;
      *reinterpret_cast<float4*>(&Ys[r][c]) = yr[u];
    }
    __syncthreads();
    if (t0 + TR < r1) fetch(t0 + TR);
#pragma unroll 2
    for (int r = warp; r < TR; r += 8) {
      float xv[KPL], y[NPL];
#pragma unroll
      for (int i = 0; i < KPL; i += 4) {
        const float4 q = *reinterpret_cast<const float4*>(&Xs[r][lk * KPL + i]);
        xv[i] = q.x; xv[i + 1] = q.y; xv[i + 2] = q.z; xv[i + 3] = q.w;
      }
#pragma unroll
      for (int j = 0; j < NPL; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(&Ys[r][(j / 4) * 32 + ln * 4]);     // j is a multiple of 4
        y[j] = q.x; y[j + 1] = q.y; y[j + 2] = q.z; y[j + 3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < KPL; ++i)
#pragma unroll
        for (int j = 0; j < NPL; ++j) acc[i][j] = fmaf(xv[i], y[j], acc[i][j]);
#pragma unroll
      for (int j = 0; j < NPL; ++j) bs[j] += y[j];
    }
    __syncthreads();
  }
  for (int w = 0; w < 8; ++w) {            // warps add their blocks in warp order
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < KPL; ++i)
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          float* p = &Rs[(lk * KPL + i) * ND + (j / 4) * 32 + ln * 4 + (j & 3)];
          *p = w == 0 ? acc[i][j] : *p + acc[i][j];
        }
      if (lk == 0) {
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          float* p = &Rs[KD * ND + (j / 4) * 32 + ln * 4 + (j & 3)];
          *p = w == 0 ? bs[j] : *p + bs[j];
        }
      }
    }
    __syncthreads();
  }
  float* out = partial + (long long)blockIdx.x * (KD * ND + ND);
  for (int i = tid; i < KD * ND + ND; i += 256) out[i] = Rs[i];
}
// CTAs of one full wave: SM count x resident CTAs per SM of this kernel (queried once).  The persistent skinny kernels split the
// rows evenly over the grid, so a grid that is a whole number of waves leaves no partly filled last wave (592 CTAs of the
// 32 x 32 wgrad were 1.33 waves at three resident CTAs per SM).
template <typename Kern>
static int wave_ctas(Kern kern, int threads, int* cache) {
  if (*cache == 0) {
    int dev = 0, sms = 148, occ = 1;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0) != cudaSuccess || occ < 1) occ = 1;
    *cache = sms * occ;
  }
  return *cache;
}
template <int KD, int ND> static int wgrad_skinny_wave() { static int c = 0; return wave_ctas(k_wgrad_skinny<KD, ND>, 256, &c); }
template <int KI, int NO> static int dgrad_skinny_wave();      // (defined behind the kernel)

bool wgrad_skinny_ok(const float* X, long long ldx, const float* dY, long long ldy, long long rows, int K, int N) {
  return rows >= 4096 && ((K == 32 && (N == 32 || N == 64 || N == 96)) || (K == 64 && N == 32)) && ldx % 4 == 0 && ldy % 4 == 0 &&
         ((uintptr_t)X & 15) == 0 && ((uintptr_t)dY & 15) == 0;
}
// dW [K, N] += X^T dY, db [N] += colsum(dY) (db may be null)
cudaError_t launch_wgrad_skinny(const float* X, long long ldx, const float* dY, long long ldy, long long rows, int K, int N,
                                float* dW, float* db, cudaStream_t st) {
  if (K == 64 && N == 32) {
    // two 32 x 32 passes over the halves of X (dW rows 0..31 | 32..63 are contiguous): the 64 x 32 instantiation keeps 64
    // accumulators per thread and is slower than both passes together (104 vs 2 x 44 us)
    cudaError_t e = launch_wgrad_skinny(X, ldx, dY, ldy, rows, 32, 32, dW, db, st);
    if (e != cudaSuccess) return e;
    return launch_wgrad_skinny(X + 32, ldx, dY, ldy, rows, 32, 32, dW + 32 * 32, nullptr, st);
  }
  const int wave = (K == 32 && N == 32) ? wgrad_skinny_wave<32, 32>() : (K == 32 && N == 64) ? wgrad_skinny_wave<32, 64>()
                   : (K == 32 && N == 96) ? wgrad_skinny_wave<32, 96>() : wgrad_skinny_wave<64, 32>();
  unsigned grid = (unsigned)std::min<long long>(wave, (rows + 255) / 256);
  grid = (unsigned)std::min<size_t>(grid, g_red_floats / (size_t)(K * N + N));
  if (grid == 0) return cudaErrorInvalidValue;
  float* scr = red_region((size_t)grid * (K * N + N) + N, st);      // (+ N: dump area of the column sums when db is null)
  if (!scr) return cudaErrorInvalidValue;
#define UU_WS(KD, ND) if (K == KD && N == ND) k_wgrad_skinny<KD, ND><<<grid, 256, 0, st>>>(X, ldx, dY, ldy, rows, scr); else
  UU_WS(32, 32) UU_WS(32, 64) UU_WS(32, 96) UU_WS(64, 32) return cudaErrorInvalidValue;
#undef UU_WS
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (db) return reduce_partials(scr, (int)grid, (long long)K * N + N, dW, (long long)K * N, db, st);
  // without a bias gradient the slab still carries the N column sums: they are added onto a dump area behind the slabs
  return reduce_partials(scr, (int)grid, (long long)K * N + N, dW, (long long)K * N, scr + (size_t)grid * (K * N + N), st);
}

// ---- narrow input gradients (spatial blocks): dX[R, NO] (+)= dY[R, KI] W^T with W (NO, KI) as the forward layer stores it ----
// Memory-bound like the skinny wgrad above (one read of dY, one write of dX), so the same shape: persistent CTAs stream
// 128-row tiles through shared memory with the next tile's loads in flight, W^T sits in shared memory once per CTA, a thread
// owns 4 rows x CPT columns (rows tr, tr + 32, ...: the 16-byte row reads of a warp fall on distinct banks through the + 4
// padding) and every global access is a coalesced 16-byte one.  The generic GEMM kernel ran these at 1.2 - 1.4 TB/s.
template <int KI, int NO>
__global__ void __launch_bounds__(256) k_dgrad_skinny(const float* __restrict__ dY, long long ldy, const float* __restrict__ Wm,
                                                      long long rows, float* __restrict__ dX, long long ldx, int accumulate) {
  constexpr int TR = KI > 64 ? 64 : 128, RPT = TR / 32, LDS_X = KI + 4, CPT = NO / 8, ITX = TR * KI / 4 / 256;   // (static shared memory <= 48 KB)
  static_assert(TR * KI / 4 % 256 == 0 && NO % 32 == 0, "shape");
  __shared__ __align__(16) float Xs[TR * LDS_X];
  __shared__ __align__(16) float Bs[KI * NO];          // Bs[k][n] = W[n][k]
  const int tid = threadIdx.x, tr = tid >> 3, tc = tid & 7;
  for (int i = tid; i < KI * NO; i += 256) {
    const int n = i / KI, k = i - n * KI;
    Bs[k * NO + n] = Wm[i];
  }
  const long long n_tiles = (rows + TR - 1) / TR;
  float4 xr[ITX];
  auto fetch = [&](long long t0) {
#pragma unroll
    for (int u = 0; u < ITX; ++u) {
      const int i = tid + 256 * u, r = i / (KI / 4), c = (i - r * (KI / 4)) * 4;
      xr[u] = (t0 + r < rows) ? *reinterpret_cast<const float4*>(dY + (t0 + r) * ldy + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  long long tile = blockIdx.x;
  if (tile < n_tiles) fetch(tile * TR);
  for (; tile < n_tiles; tile += gridDim.x) {
    const long long t0 = tile * TR;
    __syncthreads();                                    // previous tile consumed (and Bs written, first pass)
#pragma unroll
    for (int u = 0; u < ITX; ++u) {
      const int i = tid + 256 * u, r = i / (KI / 4), c = (i - r * (KI / 4)) * 4;
      *reinterpret_cast<float4*>(&Xs[r * LDS_X + c]) = xr[u];
    }
    __syncthreads();
    if (tile + gridDim.x < n_tiles) fetch((tile + gridDim.x) * TR);
    float acc[RPT][CPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < KI; k += 4) {
      float4 xv[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) xv[i] = *reinterpret_cast<const float4*>(&Xs[(tr + 32 * i) * LDS_X + k]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float w[CPT];
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          const float4 q = *reinterpret_cast<const float4*>(&Bs[(k + kk) * NO + (j / 4) * 32 + tc * 4]);   // j multiple of 4
          w[j] = q.x; w[j + 1] = q.y; w[j + 2] = q.z; w[j + 3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const float x = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
#pragma unroll
          for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(x, w[j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const long long r = t0 + tr + 32 * i;
      if (r < rows) {
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          float4* dst = reinterpret_cast<float4*>(dX + r * ldx + (j / 4) * 32 + tc * 4);
          float4 v = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
          if (accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
          *dst = v;
        }
      }
    }
  }
}
template <int KI, int NO> static int dgrad_skinny_wave() { static int c = 0; return wave_ctas(k_dgrad_skinny<KI, NO>, 256, &c); }
// dX [rows, NO] (+)= dY [rows, KI] . W^T, W (NO, KI) row-major contiguous
bool dgrad_skinny_ok(const float* dY, long long ldy, long long rows, int KI, int NO, const float* dX, long long ldx) {
  // (the 96 -> 32 shape, the packed q | k | v input gradient, is instantiated but measured slower than the generic kernel:
  // 99 vs 76 us with its 64-row tile, so it is not routed here)
  const bool shape = (KI == 32 && NO == 32) || (KI == 64 && NO == 32) || (KI == 32 && NO == 64);
  return shape && rows >= 4096 && ldy % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)dY & 15) == 0 && ((uintptr_t)dX & 15) == 0;
}
cudaError_t launch_dgrad_skinny(const float* dY, long long ldy, const float* Wm, long long rows, int KI, int NO, float* dX,
                                long long ldx, int accumulate, cudaStream_t st) {
  const int tr = KI > 64 ? 64 : 128;
  const int wave = (KI == 32 && NO == 32) ? dgrad_skinny_wave<32, 32>() : (KI == 64 && NO == 32) ? dgrad_skinny_wave<64, 32>()
                   : (KI == 96 && NO == 32) ? dgrad_skinny_wave<96, 32>() : dgrad_skinny_wave<32, 64>();
  const unsigned grid = (unsigned)std::min<long long>((rows + tr - 1) / tr, wave);
#define UU_DS(KIV, NOV) if (KI == KIV && NO == NOV) k_dgrad_skinny<KIV, NOV><<<grid, 256, 0, st>>>(dY, ldy, Wm, rows, dX, ldx, accumulate); else
  UU_DS(32, 32) UU_DS(64, 32) UU_DS(96, 32) UU_DS(32, 64) return cudaErrorInvalidValue;
#undef UU_DS
  return cudaGetLastError();
}

void gemm_gen_tile(int N, int* bm, int* bn) {
  if (N > 64) { *bm = 128; *bn = 128; }
  else if (N > 32) { *bm = 128; *bn = 64; }
  else { *bm = 256; *bn = 32; }
}

template <int BM, int BN>
static void gg_launch(const GemmGen& g, int splits, cudaStream_t st) {
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN, splits);
  if (g.transA) {
    if (g.transB) k_gemm_gen<true, true, BM, BN><<<grid, 256, 0, st>>>(g);
    else k_gemm_gen<true, false, BM, BN><<<grid, 256, 0, st>>>(g);
  } else {
    if (g.transB) k_gemm_gen<false, true, BM, BN><<<grid, 256, 0, st>>>(g);
    else k_gemm_gen<false, false, BM, BN><<<grid, 256, 0, st>>>(g);
  }
}

cudaError_t launch_gemm_gen(const GemmGen& g, cudaStream_t st) {
  if (g.M == 0 || g.N == 0 || g.K == 0) return cudaSuccess;
  int splits = g.split_k < 1 ? 1 : g.split_k;
  if (splits > 1 && (g.bias || g.relu || !g.accumulate || g.cmap.rpb != 0x7fffffff || g.ldc != g.N)) return cudaErrorInvalidValue;
  GemmGen gg = g;
  if (splits > 1) {
    splits = (int)std::min<size_t>(splits, g_red_floats / ((size_t)g.M * g.N));
    if (splits < 1) return cudaErrorInvalidValue;
    gg.partial = red_region((size_t)splits * g.M * g.N, st);
    if (!gg.partial) return cudaErrorInvalidValue;
  }
  int bm, bn;
  gemm_gen_tile(g.N, &bm, &bn);
  if (bn == 128) gg_launch<128, 128>(gg, splits, st);
  else if (bn == 64) gg_launch<128, 64>(gg, splits, st);
  else gg_launch<256, 32>(gg, splits, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || splits == 1) return e;
  return reduce_partials(gg.partial, splits, (long long)g.M * g.N, g.C, (long long)g.M * g.N, nullptr, st);
}

// out[n] += sum_r Y[r*ld + n]   (bias gradients)
__global__ void k_colsum(const float* __restrict__ Y, int M, int N, long long ld, float* __restrict__ out) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5, nr = blockDim.x >> 5;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + rl; r < r1; r += nr) s += Y[(long long)r * ld + n];
  __shared__ float sm[8][33];
  sm[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && n < N) {
    for (int i = 1; i < nr; ++i) s += sm[i][threadIdx.x & 31];
    out[(long long)blockIdx.y * N + n] = s;          // partial slab of this row range
  }
}
// 16-byte variant: a warp covers 128 columns, the 8 warps of a CTA take every 8th row of the CTA's row range, four rows per
// thread in flight; warps are summed in warp order through shared memory (fixed association).
__global__ void __launch_bounds__(256) k_colsum4(const float* __restrict__ Y, int M, int N, long long ld, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + 4 * lane;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
    int r = r0 + rl;
    for (; r + 24 < r1; r += 32) {
      const float4 a = *reinterpret_cast<const float4*>(Y + (long long)r * ld + n);
      const float4 b = *reinterpret_cast<const float4*>(Y + (long long)(r + 8) * ld + n);
      const float4 c = *reinterpret_cast<const float4*>(Y + (long long)(r + 16) * ld + n);
      const float4 d = *reinterpret_cast<const float4*>(Y + (long long)(r + 24) * ld + n);
      s.x += (a.x + b.x) + (c.x + d.x); s.y += (a.y + b.y) + (c.y + d.y);
      s.z += (a.z + b.z) + (c.z + d.z); s.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < r1; r += 8) {
      const float4 a = *reinterpret_cast<const float4*>(Y + (long long)r * ld + n);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  __shared__ __align__(16) float4 sm4[8][32];
  sm4[rl][lane] = s;
  __syncthreads();
  if (rl == 0 && n < N) {
    for (int i = 1; i < 8; ++i) { const float4 t = sm4[i][lane]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
    *reinterpret_cast<float4*>(out + (long long)blockIdx.y * N + n) = s;          // partial slab of this row range
  }
}
cudaError_t launch_colsum(const float* Y, int M, int N, long long ld, float* out, cudaStream_t st) {
  if (M == 0 || N == 0) return cudaSuccess;
  const int splits = std::max(1, std::min(128, M / 256));
  if ((size_t)splits * N > g_red_floats) return cudaErrorInvalidValue;
  float* scr = red_region((size_t)splits * N, st);
  if (!scr) return cudaErrorInvalidValue;
  if (N % 4 == 0 && ld % 4 == 0 && ((uintptr_t)Y & 15) == 0 && ((uintptr_t)scr & 15) == 0)
    k_colsum4<<<dim3((N + 127) / 128, splits), 256, 0, st>>>(Y, M, N, ld, scr);
  else
    k_colsum<<<dim3((N + 31) / 32, splits), 256, 0, st>>>(Y, M, N, ld, scr);
  return reduce_partials(scr, splits, N, out, N, nullptr, st);
}

// out[(r % period)*d + c] += X[r*d + c]  (positional-encoding gradients: sum over the batch);
// with `rowmask`: only rows where (rowmask[r] != 0) == want are summed (upsampling-token gradient uses period 1).
__global__ void k_period_sum(const float* __restrict__ X, long long rows, int period, int d,
                             const uint8_t* __restrict__ rowmask, int want, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (c >= d) return;
  const long long nb = rows / period;
  const long long b0 = (long long)blockIdx.z * ((nb + gridDim.z - 1) / gridDim.z);
  const long long b1 = min(nb, b0 + (nb + gridDim.z - 1) / gridDim.z);
  float s = 0.f;
  for (long long b = b0; b < b1; ++b) {
    const long long r = b * period + p;
    if (rowmask && ((rowmask[r] != 0) != (want != 0))) continue;
    s += X[r * d + c];
  }
  out[((long long)blockIdx.z * gridDim.y + p) * d + c] = s;      // partial slab of this batch range
}
cudaError_t launch_period_sum(const float* X, long long rows, int period, int d, const uint8_t* rowmask, int want,
                              float* out, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  // without a row mask this is the column sum of X viewed as (rows / period, period * d): the 16-byte kernel (the loop below
  // ran at 1 TB/s on the positional tables)
  if (!rowmask && rows % period == 0 && ((long long)period * d) % 4 == 0 && (long long)period * d < (1ll << 30) &&
      rows / period < (1ll << 31) && ((uintptr_t)X & 15) == 0)
    return launch_colsum(X, (int)(rows / period), period * d, (long long)period * d, out, st);
  const int bx = d >= 128 ? 128 : 32;
  const long long nb = rows / period;
  long long nz = std::max<long long>(1, std::min<long long>(64, nb / 8));      // (3 x 3 x 16 CTAs of 128 threads ran at 1 TB/s)
  nz = std::max<long long>(1, std::min<long long>(nz, (long long)(g_red_floats / ((size_t)period * d))));
  dim3 grid((d + bx - 1) / bx, period, (unsigned)nz);
  if ((size_t)period * d > g_red_floats) return cudaErrorInvalidValue;
  float* scr = red_region((size_t)nz * period * d, st);
  if (!scr) return cudaErrorInvalidValue;
  k_period_sum<<<grid, bx, 0, st>>>(X, rows, period, d, rowmask, want, scr);
  return reduce_partials(scr, (int)nz, (long long)period * d, out, (long long)period * d, nullptr, st);
}

// =================================================================================================
// LayerNorm forward / backward for any width d (one warp per row).
//   y = (x - mean) * rstd * gamma + beta ;  dx = rstd * (dyg - mean(dyg) - xhat * mean(dyg * xhat)), dyg = dy*gamma
//   dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy.  dx is added to `dx_acc` when accumulate != 0.
// =================================================================================================
__global__ void k_ln_fwd_gen(const float* __restrict__ x, long long rows, int d, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float eps, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += xr[c];
  const float mean = warp_sum_t(s) / d;
  float q = 0.f;
  for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; q += t * t; }
  const float rstd = rsqrtf(warp_sum_t(q) / d + eps);
  float* yr = y + row * d;
  for (int c = lane; c < d; c += 32) {
    const float inv = gamma[c] * rstd;
    yr[c] = xr[c] * inv + (beta[c] - mean * inv);
  }
}
// d == 4 * LPR * VPL: a row is LPR lanes x VPL float4 (32 / LPR rows per warp), held in registers between the statistics and
// the normalisation: one 16-byte read and one 16-byte write per element quad (the scalar kernel above re-read the row three
// times through 4-byte accesses and ran at 1.1 TB/s).
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) k_ln_fwd_v4(const float* __restrict__ x, long long rows, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float eps, float* __restrict__ y) {
  constexpr int D = 4 * LPR * VPL, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
  const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + sub;
  const bool ok = row < rows;
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = ok ? *reinterpret_cast<const float4*>(x + row * D + 4 * (l + LPR * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + b * b) + (c * c + e * e);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / D) + eps);
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (l + LPR * i);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = (v[i].x - mean) * (g.x * rstd) + b.x; o.y = (v[i].y - mean) * (g.y * rstd) + b.y;
    o.z = (v[i].z - mean) * (g.z * rstd) + b.z; o.w = (v[i].w - mean) * (g.w * rstd) + b.w;
    *reinterpret_cast<float4*>(y + row * D + c) = o;
  }
}
static inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
cudaError_t launch_ln_fwd_gen(const float* x, long long rows, int d, const float* gamma, const float* beta, float eps,
                              float* y, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (al16(x) && al16(y) && al16(gamma) && al16(beta)) {
#define UU_LNF(LPR, VPL)                                                                                                  \
  if (d == 4 * LPR * VPL) {                                                                                               \
    const long long rpc = 8 * (32 / LPR);                                                                                 \
    k_ln_fwd_v4<LPR, VPL><<<(unsigned)((rows + rpc - 1) / rpc), 256, 0, st>>>(x, rows, gamma, beta, eps, y);              \
    return cudaGetLastError();                                                                                            \
  }
    UU_LNF(8, 1) UU_LNF(16, 1) UU_LNF(32, 1) UU_LNF(32, 2) UU_LNF(32, 3) UU_LNF(32, 4)
#undef UU_LNF
  }
  k_ln_fwd_gen<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, d, gamma, beta, eps, y);
  return cudaGetLastError();
}

// One warp per row; lane l owns columns l, l + 32, ... (at most LN_CPL of them): its share of dgamma / dbeta stays in
// registers over all rows of the warp, warps of a CTA are then summed through shared memory in warp order and the CTA
// writes one partial slab [2 d] (no atomics; launch_ln_bwd_gen reduces the slabs in order).
template <int LN_CPL>          // columns per lane: d <= 32 * LN_CPL
__global__ void __launch_bounds__(256) k_ln_bwd_gen(const float* __restrict__ x, const float* __restrict__ dy, long long rows, int d,
                             const float* __restrict__ gamma, float eps, float* __restrict__ dx, int accumulate,
                             float* __restrict__ partial) {
  extern __shared__ float sm[];            // [warps][2 d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float pg[LN_CPL], pb[LN_CPL], gm[LN_CPL];
#pragma unroll
  for (int i = 0; i < LN_CPL; ++i) {
    pg[i] = 0.f; pb[i] = 0.f;
    gm[i] = (lane + 32 * i) < d ? gamma[lane + 32 * i] : 0.f;
  }
  for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
    const float* xr = x + row * d;
    const float* dyr = dy + row * d;
    float xv[LN_CPL], gv[LN_CPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_CPL; ++i) {
      const int c = lane + 32 * i;
      xv[i] = c < d ? xr[c] : 0.f;
      gv[i] = c < d ? dyr[c] : 0.f;
      s += xv[i];
    }
    const float mean = warp_sum_t(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_CPL; ++i) { const float t = (lane + 32 * i) < d ? xv[i] - mean : 0.f; q += t * t; }
    const float rstd = rsqrtf(warp_sum_t(q) / d + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_CPL; ++i) {
      const float xh = (xv[i] - mean) * rstd, g = gv[i] * gm[i];
      xv[i] = (lane + 32 * i) < d ? xh : 0.f;
      m1 += g; m2 += g * xv[i];
    }
    m1 = warp_sum_t(m1) / d; m2 = warp_sum_t(m2) / d;
    float* dxr = dx + row * d;
#pragma unroll
    for (int i = 0; i < LN_CPL; ++i) {
      const int c = lane + 32 * i;
      if (c < d) {
        const float v = rstd * (gv[i] * gm[i] - m1 - xv[i] * m2);
        dxr[c] = accumulate ? dxr[c] + v : v;
        pg[i] = fmaf(gv[i], xv[i], pg[i]);
        pb[i] += gv[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LN_CPL; ++i) {
    const int c = lane + 32 * i;
    if (c < d) { sm[warp * 2 * d + c] = pg[i]; sm[warp * 2 * d + d + c] = pb[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < wpb; ++w) t += sm[w * 2 * d + c];
    partial[(long long)blockIdx.x * 2 * d + c] = t;
  }
}
// d == 32 (spatial blocks): a row is 8 lanes x float4, a warp handles 4 rows at a time and two such groups per loop
// iteration, so that enough loads are in flight (one warp per 128-byte row was latency-bound at 1.2 TB/s).
__global__ void __launch_bounds__(256) k_ln_bwd_d32(const float* __restrict__ x, const float* __restrict__ dy, long long rows,
                                                    const float* __restrict__ gamma, float eps, float* __restrict__ dx,
                                                    int accumulate, float* __restrict__ partial) {
  __shared__ float sm[32][64];             // [warp-quarter = 8 warps x 4 row slots][gamma 32 | beta 32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane >> 3, c4 = (lane & 7) * 4;
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c4);
  float4 pg = make_float4(0.f, 0.f, 0.f, 0.f), pb = pg;
  const long long stride = (long long)gridDim.x * 64;       // rows per grid pass: 8 warps x 4 rows x 2
  for (long long wb = (long long)blockIdx.x * 64 + warp * 8; wb < rows; wb += stride) {     // (warp-uniform bound)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long row = wb + sub + 4 * u;
      const bool ok = row < rows;
      const float4 xv = ok ? *reinterpret_cast<const float4*>(x + row * 32 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 gv = ok ? *reinterpret_cast<const float4*>(dy + row * 32 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float s = xv.x + xv.y + xv.z + xv.w;
      s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
      const float mean = s * (1.f / 32);
      const float d0 = xv.x - mean, d1 = xv.y - mean, d2 = xv.z - mean, d3 = xv.w - mean;
      float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
      const float rstd = rsqrtf(q * (1.f / 32) + eps);
      const float h0 = d0 * rstd, h1 = d1 * rstd, h2 = d2 * rstd, h3 = d3 * rstd;
      const float g0 = gv.x * gm.x, g1 = gv.y * gm.y, g2 = gv.z * gm.z, g3 = gv.w * gm.w;
      float m1 = g0 + g1 + g2 + g3, m2 = g0 * h0 + g1 * h1 + g2 * h2 + g3 * h3;
      m1 += __shfl_xor_sync(0xffffffffu, m1, 1); m1 += __shfl_xor_sync(0xffffffffu, m1, 2); m1 += __shfl_xor_sync(0xffffffffu, m1, 4);
      m2 += __shfl_xor_sync(0xffffffffu, m2, 1); m2 += __shfl_xor_sync(0xffffffffu, m2, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
      m1 *= (1.f / 32); m2 *= (1.f / 32);
      if (ok) {
        float4 v = make_float4(rstd * (g0 - m1 - h0 * m2), rstd * (g1 - m1 - h1 * m2), rstd * (g2 - m1 - h2 * m2),
                               rstd * (g3 - m1 - h3 * m2));
        float4* dst = reinterpret_cast<float4*>(dx + row * 32 + c4);
        if (accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *dst = v;
        pg.x = fmaf(gv.x, h0, pg.x); pg.y = fmaf(gv.y, h1, pg.y); pg.z = fmaf(gv.z, h2, pg.z); pg.w = fmaf(gv.w, h3, pg.w);
        pb.x += gv.x; pb.y += gv.y; pb.z += gv.z; pb.w += gv.w;
      }
    }
  }
  *reinterpret_cast<float4*>(&sm[warp * 4 + sub][c4]) = pg;
  *reinterpret_cast<float4*>(&sm[warp * 4 + sub][32 + c4]) = pb;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += sm[w][threadIdx.x];
    partial[(long long)blockIdx.x * 64 + threadIdx.x] = t;
  }
}
// d == 128 * VPL (temporal / strided blocks): lane l owns the float4 columns 4 l + 128 i; x, dy (and dx when accumulating)
// move as 16-byte accesses, the gamma / beta partials of the lane stay in registers over all rows of the warp.
template <int VPL>
__global__ void __launch_bounds__(256) k_ln_bwd_v4(const float* __restrict__ x, const float* __restrict__ dy, long long rows,
                                                   const float* __restrict__ gamma, float eps, float* __restrict__ dx,
                                                   int accumulate, float* __restrict__ partial) {
  constexpr int D = 128 * VPL;
  extern __shared__ __align__(16) float sm[];            // [warps][2 D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float4 pg[VPL], pb[VPL], gm[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    pg[i] = make_float4(0.f, 0.f, 0.f, 0.f); pb[i] = pg[i];
    gm[i] = *reinterpret_cast<const float4*>(gamma + 4 * lane + 128 * i);
  }
  for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
    float4 xv[VPL], gv[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i] = *reinterpret_cast<const float4*>(x + row * D + 4 * lane + 128 * i);
      gv[i] = *reinterpret_cast<const float4*>(dy + row * D + 4 * lane + 128 * i);
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    const float mean = warp_sum_t(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
    const float rstd = rsqrtf(warp_sum_t(q) * (1.f / D) + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;          // xhat
      const float g0 = gv[i].x * gm[i].x, g1 = gv[i].y * gm[i].y, g2 = gv[i].z * gm[i].z, g3 = gv[i].w * gm[i].w;
      m1 += (g0 + g1) + (g2 + g3);
      m2 += (g0 * xv[i].x + g1 * xv[i].y) + (g2 * xv[i].z + g3 * xv[i].w);
    }
    m1 = warp_sum_t(m1) * (1.f / D); m2 = warp_sum_t(m2) * (1.f / D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 v;
      v.x = rstd * (gv[i].x * gm[i].x - m1 - xv[i].x * m2); v.y = rstd * (gv[i].y * gm[i].y - m1 - xv[i].y * m2);
      v.z = rstd * (gv[i].z * gm[i].z - m1 - xv[i].z * m2); v.w = rstd * (gv[i].w * gm[i].w - m1 - xv[i].w * m2);
      float4* dst = reinterpret_cast<float4*>(dx + row * D + 4 * lane + 128 * i);
      if (accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *dst = v;
      pg[i].x = fmaf(gv[i].x, xv[i].x, pg[i].x); pg[i].y = fmaf(gv[i].y, xv[i].y, pg[i].y);
      pg[i].z = fmaf(gv[i].z, xv[i].z, pg[i].z); pg[i].w = fmaf(gv[i].w, xv[i].w, pg[i].w);
      pb[i].x += gv[i].x; pb[i].y += gv[i].y; pb[i].z += gv[i].z; pb[i].w += gv[i].w;
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    *reinterpret_cast<float4*>(sm + warp * 2 * D + 4 * lane + 128 * i) = pg[i];
    *reinterpret_cast<float4*>(sm + warp * 2 * D + D + 4 * lane + 128 * i) = pb[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < wpb; ++w) t += sm[w * 2 * D + c];
    partial[(long long)blockIdx.x * 2 * D + c] = t;
  }
}
cudaError_t launch_ln_bwd_gen(const float* x, const float* dy, long long rows, int d, const float* gamma, float eps,
                              float* dx, int accumulate, float* dgamma, float* dbeta, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (d > 512) return cudaErrorInvalidValue;
  unsigned grid = (unsigned)std::min<long long>((rows + 7) / 8, 148 * 4);
  grid = (unsigned)std::min<size_t>(grid, g_red_floats / (2 * (size_t)d));
  if (grid == 0) return cudaErrorInvalidValue;
  const size_t smem = 8 * 2 * d * sizeof(float);
  float* scr = red_region((size_t)grid * 2 * d, st);
  if (!scr) return cudaErrorInvalidValue;
  if (d == 32 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)gamma & 15) == 0)
    k_ln_bwd_d32<<<grid, 256, 0, st>>>(x, dy, rows, gamma, eps, dx, accumulate, scr);
  else if (d % 128 == 0 && al16(x) && al16(dy) && al16(dx) && al16(gamma)) {
    switch (d / 128) {
      case 1: k_ln_bwd_v4<1><<<grid, 256, smem, st>>>(x, dy, rows, gamma, eps, dx, accumulate, scr); break;
      case 2: k_ln_bwd_v4<2><<<grid, 256, smem, st>>>(x, dy, rows, gamma, eps, dx, accumulate, scr); break;
      case 3: k_ln_bwd_v4<3><<<grid, 256, smem, st>>>(x, dy, rows, gamma, eps, dx, accumulate, scr); break;
      default: k_ln_bwd_v4<4><<<grid, 256, smem, st>>>(x, dy, rows, gamma, eps, dx, accumulate, scr); break;
    }
  }
  else if (d <= 32) k_ln_bwd_gen<1><<<grid, 256, smem, st>>>(x, dy, rows, d, gamma, eps, dx, accumulate, scr);
  else if (d <= 64) k_ln_bwd_gen<2><<<grid, 256, smem, st>>>(x, dy, rows, d, gamma, eps, dx, accumulate, scr);
  else if (d <= 384) k_ln_bwd_gen<12><<<grid, 256, smem, st>>>(x, dy, rows, d, gamma, eps, dx, accumulate, scr);
  else k_ln_bwd_gen<16><<<grid, 256, smem, st>>>(x, dy, rows, d, gamma, eps, dx, accumulate, scr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return reduce_partials(scr, (int)grid, 2LL * d, dgamma, d, dbeta, st);
}

// =================================================================================================
// Attention backward per (sample, head), S <= 128, thread == query row, then thread == key row.
//   A = softmax(q k^T * scale + keyterm) (recomputed); dV = A^T dO; dA = dO V^T;
//   dS = A o (dA - rowsum(dA o A)); dQ = dS K * scale; dK = dS^T Q * scale.   (vit:99-130)
// qkv / dqkv rows are [q | k | v] (3*d); dO rows are d wide (merged heads).
// =================================================================================================
// One CTA per (sample, head), thread == query in the first phase and == key in the second.  The thread's own q / dO
// row lives in registers (a [S][DH] shared row read with stride DH is a 16-way bank conflict for DH = 48); K, V, Q, dO
// rows needed by everyone are read as broadcast float4; the S x S weight / gradient tiles use an odd row stride so
// that both the row-wise writes and the column-wise reads are conflict-free.  blockDim = S rounded up to a warp.
template <int DH>
__global__ void __launch_bounds__(128) k_attention_bwd(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                       int S, int heads, const uint8_t* __restrict__ mask,
                                                       int mask_stride, float* __restrict__ dqkv) {
  extern __shared__ __align__(16) float sm[];
  const int ST = S | 1;            // odd row stride of the S x S tiles
  float* Qs = sm;                  // [S][DH]
  float* Ks = Qs + S * DH;
  float* Vs = Ks + S * DH;
  float* Gs = Vs + S * DH;         // dO
  float* Km = Gs + S * DH;         // [S]
  float* As = Km + ((S + 3) & ~3); // [S][ST]  attention weights
  float* Ds = As + S * ST;         // [S][ST]  dS
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const int d = heads * DH;
  const long long row0 = (long long)b * S;
  for (int i = tid; i < S * (DH / 4); i += nt) {
    const int j = i / (DH / 4), c = (i - j * (DH / 4)) * 4;
    const float* r = qkv + (row0 + j) * 3 * d + h * DH + c;
    *reinterpret_cast<float4*>(Qs + j * DH + c) = *reinterpret_cast<const float4*>(r);
    *reinterpret_cast<float4*>(Ks + j * DH + c) = *reinterpret_cast<const float4*>(r + d);
    *reinterpret_cast<float4*>(Vs + j * DH + c) = *reinterpret_cast<const float4*>(r + 2 * d);
    *reinterpret_cast<float4*>(Gs + j * DH + c) = *reinterpret_cast<const float4*>(dO + (row0 + j) * d + h * DH + c);
  }
  for (int j = tid; j < S; j += nt)
    Km[j] = mask ? (1.0f - (mask[(long long)b * mask_stride + j] ? 1.0f : 0.0f)) * -1e9f : 0.0f;
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)DH);
  if (tid < S) {
    const int i = tid;
    float* Ai = As + i * ST;
    float* Di = Ds + i * ST;
    float mx = -INFINITY;
    {   // scores and softmax of query i
      float q[DH];
#pragma unroll
      for (int c = 0; c < DH; ++c) q[c] = Qs[i * DH + c];
      for (int j = 0; j < S; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + j * DH);
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 k4 = kr[c];
          a = fmaf(q[4 * c], k4.x, a); a = fmaf(q[4 * c + 1], k4.y, a);
          a = fmaf(q[4 * c + 2], k4.z, a); a = fmaf(q[4 * c + 3], k4.w, a);
        }
        a = a * scale + Km[j];
        Ai[j] = a;
        mx = fmaxf(mx, a);
      }
    }
    float sum = 0.f;
    for (int j = 0; j < S; ++j) { const float p = expf(Ai[j] - mx); Ai[j] = p; sum += p; }
    const float inv = 1.f / sum;
    float dot = 0.f;
    {   // dA = dO V^T, dot = sum_j dA_j p_j
      float gq[DH];
#pragma unroll
      for (int c = 0; c < DH; ++c) gq[c] = Gs[i * DH + c];
      for (int j = 0; j < S; ++j) {
        const float p = Ai[j] * inv;
        const float4* vr = reinterpret_cast<const float4*>(Vs + j * DH);
        float da = 0.f;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 v4 = vr[c];
          da = fmaf(gq[4 * c], v4.x, da); da = fmaf(gq[4 * c + 1], v4.y, da);
          da = fmaf(gq[4 * c + 2], v4.z, da); da = fmaf(gq[4 * c + 3], v4.w, da);
        }
        Ai[j] = p;
        Di[j] = da;
        dot = fmaf(da, p, dot);
      }
    }
    float dq[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) dq[c] = 0.f;
    for (int j = 0; j < S; ++j) {
      const float ds = Ai[j] * (Di[j] - dot);
      Di[j] = ds;
      const float4* kr = reinterpret_cast<const float4*>(Ks + j * DH);
#pragma unroll
      for (int c = 0; c < DH / 4; ++c) {
        const float4 k4 = kr[c];
        dq[4 * c] = fmaf(ds, k4.x, dq[4 * c]); dq[4 * c + 1] = fmaf(ds, k4.y, dq[4 * c + 1]);
        dq[4 * c + 2] = fmaf(ds, k4.z, dq[4 * c + 2]); dq[4 * c + 3] = fmaf(ds, k4.w, dq[4 * c + 3]);
      }
    }
    float* o = dqkv + (row0 + i) * 3 * d + h * DH;
#pragma unroll
    for (int c = 0; c < DH; c += 4)
      *reinterpret_cast<float4*>(o + c) = make_float4(dq[c] * scale, dq[c + 1] * scale, dq[c + 2] * scale, dq[c + 3] * scale);
  }
  __syncthreads();
  if (tid < S) {
    const int j = tid;
    float dk[DH], dv[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
    for (int i = 0; i < S; ++i) {
      const float ds = Ds[i * ST + j], p = As[i * ST + j];
      const float4* qr = reinterpret_cast<const float4*>(Qs + i * DH);
      const float4* gr = reinterpret_cast<const float4*>(Gs + i * DH);
#pragma unroll
      for (int c = 0; c < DH / 4; ++c) {
        const float4 q4 = qr[c], g4 = gr[c];
        dk[4 * c] = fmaf(ds, q4.x, dk[4 * c]); dk[4 * c + 1] = fmaf(ds, q4.y, dk[4 * c + 1]);
        dk[4 * c + 2] = fmaf(ds, q4.z, dk[4 * c + 2]); dk[4 * c + 3] = fmaf(ds, q4.w, dk[4 * c + 3]);
        dv[4 * c] = fmaf(p, g4.x, dv[4 * c]); dv[4 * c + 1] = fmaf(p, g4.y, dv[4 * c + 1]);
        dv[4 * c + 2] = fmaf(p, g4.z, dv[4 * c + 2]); dv[4 * c + 3] = fmaf(p, g4.w, dv[4 * c + 3]);
      }
    }
    float* o = dqkv + (row0 + j) * 3 * d + h * DH;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      *reinterpret_cast<float4*>(o + d + c) = make_float4(dk[c] * scale, dk[c + 1] * scale, dk[c + 2] * scale, dk[c + 3] * scale);
      *reinterpret_cast<float4*>(o + 2 * d + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
    }
  }
}

// =================================================================================================
// Spatial attention of the training step (17 joint tokens, 8 heads of dimension 4, no key mask; vit:99-130 inside
// net:313-333): the generic kernels spend one CTA per (frame, head) on a 17 x 17 problem.  Here one WARP owns a frame:
// its q | k | v rows (17 x 96 floats, contiguous in the tape) are staged in the warp's shared-memory slot, lane = (head h,
// key group g) with h = lane & 7, g = lane >> 3, and the lane keeps the k and v rows of ITS keys (j = g, g + 4, ..., five at
// most) of head h in registers for the whole frame.  The warp then walks the 17 queries: every lane scores the query against
// its own keys, the softmax statistics (max, sum, and in the backward pass sum_j dA_ij A_ij) are combined over the four key
// groups with two shuffles each, and the lane accumulates what belongs to its keys — the forward output (reduced over the
// groups), or dQ_i (reduced) plus dK_j and dV_j (private, written once per frame).  One pass, no 17 x 17 tile, no
// recomputation, and shared memory is read twice per query (q_i, dO_i) instead of ~50 times per (query, head) pair: the
// earlier (query, head)-per-lane form was bound by shared-memory bandwidth (every 16-byte key / value chunk was fetched again
// by each quarter-warp: 76 % of the LSU peak, 186 us per layer).  S is a compile-time constant; heads * 4 == 32.
// =================================================================================================
constexpr int SA_WARPS = 4;
// N float4 from global to the warp's shared-memory slot, all loads in flight before the first store
template <int N>
__device__ __forceinline__ void sa_stage(float4* __restrict__ dst, const float4* __restrict__ src, int lane) {
  constexpr int IT = (N + 31) / 32;
  float4 v[IT];
#pragma unroll
  for (int u = 0; u < IT; ++u)
    if (lane + 32 * u < N) v[u] = src[lane + 32 * u];
#pragma unroll
  for (int u = 0; u < IT; ++u)
    if (lane + 32 * u < N) dst[lane + 32 * u] = v[u];
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x))); }
__device__ __forceinline__ float grp_sum(float v) {          // over the four key groups (lanes h, h + 8, h + 16, h + 24)
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ float grp_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
  return v;
}
template <int S>
__global__ void __launch_bounds__(SA_WARPS * 32, 6) k_attn_small_fwd(const float* __restrict__ qkv, long long frames,
                                                                      float* __restrict__ out) {
  constexpr int NK = (S + 3) / 4;
  __shared__ __align__(16) float sa_sm[SA_WARPS][S * 96];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane & 7, grp = lane >> 3;
  float* sq = sa_sm[warp];
  const float scale = 0.5f;                            // 1 / sqrt(4)
  for (long long f = (long long)blockIdx.x * SA_WARPS + warp; f < frames; f += (long long)gridDim.x * SA_WARPS) {
    __syncwarp();
    sa_stage<S * 24>(reinterpret_cast<float4*>(sq), reinterpret_cast<const float4*>(qkv + f * S * 96), lane);
    __syncwarp();
    float4 kk[NK], vv[NK];
#pragma unroll
    for (int u = 0; u < NK; ++u) {
      const int j = grp + 4 * u;
      kk[u] = j < S ? *reinterpret_cast<const float4*>(sq + j * 96 + 32 + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
      vv[u] = j < S ? *reinterpret_cast<const float4*>(sq + j * 96 + 64 + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 2
    for (int i = 0; i < S; ++i) {
      const float4 q = *reinterpret_cast<const float4*>(sq + i * 96 + 4 * h);
      float p[NK];
      float m = -INFINITY;
#pragma unroll
      for (int u = 0; u < NK; ++u) {
        p[u] = (grp + 4 * u < S) ? dot4(q, kk[u]) * scale : -INFINITY;
        m = fmaxf(m, p[u]);
      }
      m = grp_max(m);
      float l = 0.f;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < NK; ++u) {
        p[u] = __expf(p[u] - m);
        l += p[u];
        o.x = fmaf(p[u], vv[u].x, o.x); o.y = fmaf(p[u], vv[u].y, o.y); o.z = fmaf(p[u], vv[u].z, o.z); o.w = fmaf(p[u], vv[u].w, o.w);
      }
      const float inv = 1.f / grp_sum(l);
      o.x = grp_sum(o.x); o.y = grp_sum(o.y); o.z = grp_sum(o.z); o.w = grp_sum(o.w);
      if (grp == 0)
        *reinterpret_cast<float4*>(out + f * S * 32 + i * 32 + 4 * h) = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
    }
  }
}

template <int S>
__global__ void __launch_bounds__(SA_WARPS * 32, 4) k_attn_small_bwd(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                                      long long frames, float* __restrict__ dqkv) {
  constexpr int NK = (S + 3) / 4, SLOT = S * 96 + S * 32;      // q | k | v rows, dO rows
  __shared__ __align__(16) float sa_sm[SA_WARPS][SLOT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane & 7, grp = lane >> 3;
  float* sq = sa_sm[warp];
  float* sd = sq + S * 96;
  const float scale = 0.5f;
  for (long long f = (long long)blockIdx.x * SA_WARPS + warp; f < frames; f += (long long)gridDim.x * SA_WARPS) {
    float* dst = dqkv + f * S * 96;
    __syncwarp();
    sa_stage<S * 24>(reinterpret_cast<float4*>(sq), reinterpret_cast<const float4*>(qkv + f * S * 96), lane);
    sa_stage<S * 8>(reinterpret_cast<float4*>(sd), reinterpret_cast<const float4*>(dO + f * S * 32), lane);
    __syncwarp();
    float4 kk[NK], vv[NK], dk[NK], dv[NK];
#pragma unroll
    for (int u = 0; u < NK; ++u) {
      const int j = grp + 4 * u;
      kk[u] = j < S ? *reinterpret_cast<const float4*>(sq + j * 96 + 32 + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
      vv[u] = j < S ? *reinterpret_cast<const float4*>(sq + j * 96 + 64 + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
      dk[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      dv[u] = dk[u];
    }
#pragma unroll 1
    for (int i = 0; i < S; ++i) {
      const float4 q = *reinterpret_cast<const float4*>(sq + i * 96 + 4 * h);
      const float4 g = *reinterpret_cast<const float4*>(sd + i * 32 + 4 * h);
      float a[NK], da[NK];
      float m = -INFINITY;
#pragma unroll
      for (int u = 0; u < NK; ++u) {
        a[u] = (grp + 4 * u < S) ? dot4(q, kk[u]) * scale : -INFINITY;
        m = fmaxf(m, a[u]);
      }
      m = grp_max(m);
      float l = 0.f;
#pragma unroll
      for (int u = 0; u < NK; ++u) { a[u] = __expf(a[u] - m); l += a[u]; }
      const float inv = 1.f / grp_sum(l);
      float dot = 0.f;
#pragma unroll
      for (int u = 0; u < NK; ++u) {
        da[u] = dot4(g, vv[u]);
        a[u] *= inv;
        dot = fmaf(da[u], a[u], dot);
      }
      dot = grp_sum(dot);
      float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < NK; ++u) {
        const float ds = a[u] * (da[u] - dot);
        dq.x = fmaf(ds, kk[u].x, dq.x); dq.y = fmaf(ds, kk[u].y, dq.y); dq.z = fmaf(ds, kk[u].z, dq.z); dq.w = fmaf(ds, kk[u].w, dq.w);
        dk[u].x = fmaf(ds, q.x, dk[u].x); dk[u].y = fmaf(ds, q.y, dk[u].y); dk[u].z = fmaf(ds, q.z, dk[u].z); dk[u].w = fmaf(ds, q.w, dk[u].w);
        dv[u].x = fmaf(a[u], g.x, dv[u].x); dv[u].y = fmaf(a[u], g.y, dv[u].y); dv[u].z = fmaf(a[u], g.z, dv[u].z); dv[u].w = fmaf(a[u], g.w, dv[u].w);
      }
      dq.x = grp_sum(dq.x); dq.y = grp_sum(dq.y); dq.z = grp_sum(dq.z); dq.w = grp_sum(dq.w);
      if (grp == 0)
        *reinterpret_cast<float4*>(dst + i * 96 + 4 * h) = make_float4(dq.x * scale, dq.y * scale, dq.z * scale, dq.w * scale);
    }
#pragma unroll
    for (int u = 0; u < NK; ++u) {
      const int j = grp + 4 * u;
      if (j < S) {
        *reinterpret_cast<float4*>(dst + j * 96 + 32 + 4 * h) = make_float4(dk[u].x * scale, dk[u].y * scale, dk[u].z * scale, dk[u].w * scale);
        *reinterpret_cast<float4*>(dst + j * 96 + 64 + 4 * h) = dv[u];
      }
    }
  }
}
bool attention_small_ok(int S, int heads, int dh, const uint8_t* mask) { return dh == 4 && heads == 8 && S == 17 && !mask; }
cudaError_t launch_attention_small_fwd(const float* qkv, long long frames, int S, float* out, cudaStream_t st) {
  if (S != 17) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)std::min<long long>((frames + SA_WARPS - 1) / SA_WARPS, 148 * 8);
  k_attn_small_fwd<17><<<grid, SA_WARPS * 32, 0, st>>>(qkv, frames, out);
  return cudaGetLastError();
}
cudaError_t launch_attention_small_bwd(const float* qkv, const float* dO, long long frames, int S, float* dqkv, cudaStream_t st) {
  if (S != 17) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)std::min<long long>((frames + SA_WARPS - 1) / SA_WARPS, 148 * 4);
  k_attn_small_bwd<17><<<grid, SA_WARPS * 32, 0, st>>>(qkv, dO, frames, dqkv);
  return cudaGetLastError();
}

cudaError_t launch_attention_bwd(const float* qkv, const float* dO, long long B, int S, int heads, int dh,
                                 const uint8_t* mask, int mask_stride, float* dqkv, cudaStream_t st) {
  if (B == 0) return cudaSuccess;
  if (S < 1 || S > 128) return cudaErrorInvalidValue;
  if (attention_small_ok(S, heads, dh, mask)) return launch_attention_small_bwd(qkv, dO, B, S, dqkv, st);
  const size_t smem = sizeof(float) * (4 * S * dh + ((S + 3) & ~3) + 2 * S * (S | 1));
  dim3 grid((unsigned)B, heads);
  const int threads = (S + 31) / 32 * 32;
#define UU_AB_CASE(DHV)                                                                                        \
  case DHV: {                                                                                                  \
    static size_t attr_smem = 0;                                                                               \
    if (smem > attr_smem) {                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(k_attention_bwd<DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           (int)smem);                                                         \
      if (e != cudaSuccess) return e;                                                                          \
      attr_smem = smem;                                                                                        \
    }                                                                                                          \
    k_attention_bwd<DHV><<<grid, threads, smem, st>>>(qkv, dO, S, heads, mask, mask_stride, dqkv);             \
    break;                                                                                                     \
  }
  switch (dh) {
    UU_AB_CASE(4) UU_AB_CASE(16) UU_AB_CASE(32) UU_AB_CASE(48) UU_AB_CASE(64)
    default: return cudaErrorInvalidValue;
  }
#undef UU_AB_CASE
  return cudaGetLastError();
}

// =================================================================================================
// Element-wise pieces
// =================================================================================================
// act: 0 = ReLU (uses the activation output), 1 = exact-erf GELU (uses the pre-activation)
__global__ void k_act_fwd(const float* __restrict__ pre, long long n, int act, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = pre[i];
    out[i] = act == 0 ? fmaxf(v, 0.f) : 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  }
}
// dpre = dout * act'(pre).  `dout` may be read through a row map (strided block: gradient of the zero-padded
// conv input gathered back to sequence rows; dropped rows have zero gradient).
__global__ void k_act_bwd(const float* __restrict__ pre, const float* __restrict__ dout, RowMap dmap, long long ldd,
                          long long rows, int cols, int act, float* __restrict__ dpre) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const long long dr = map_row(dmap, (int)r);
    const float gout = dr < 0 ? 0.f : dout[dr * ldd + c];
    const float v = pre[i];
    float gp;
    if (act == 0) gp = v > 0.f ? 1.f : 0.f;     // TF: relu'(0) = 0
    else gp = 0.5f * (1.f + erff(v * 0.70710678118654752440f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);
    dpre[i] = gout * gp;
  }
}
// ReLU backward when both the activation and its incoming gradient live in the zero-padded conv-input layout:
// dpre[r][c] = (hp[map r][c] > 0) * dhp[map r][c]; rows the map drops get 0.
__global__ void k_act_bwd_mapped(const float* __restrict__ hp, const float* __restrict__ dhp, RowMap map, long long ld,
                                 long long rows, int cols, float* __restrict__ dpre) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const long long pr = map_row(map, (int)r);
    dpre[i] = (pr >= 0 && hp[pr * ld + c] > 0.f) ? dhp[pr * ld + c] : 0.f;
  }
}
// ---- 16-byte variants of the element-wise kernels: one float4 per thread and step, ONE 32-bit index division per four
// elements (the scalar kernels above spend a 64-bit division per element and ran at 2.0 - 2.6 TB/s); used whenever the
// widths are multiples of 4, the pointers 16-byte aligned and the element count fits 32 bits.
__device__ __forceinline__ float act_fwd1(float v, int act) {
  return act == 0 ? fmaxf(v, 0.f) : 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
}
__device__ __forceinline__ float act_grad1(float v, int act) {
  if (act == 0) return v > 0.f ? 1.f : 0.f;     // TF: relu'(0) = 0
  return 0.5f * (1.f + erff(v * 0.70710678118654752440f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);
}
__global__ void __launch_bounds__(256) k_act_fwd4(const float4* __restrict__ pre, uint32_t n4, int act, float4* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const float4 v = pre[i];
    out[i] = make_float4(act_fwd1(v.x, act), act_fwd1(v.y, act), act_fwd1(v.z, act), act_fwd1(v.w, act));
  }
}
__global__ void __launch_bounds__(256) k_act_bwd4(const float4* __restrict__ pre, const float* __restrict__ dout, RowMap dmap,
                                                  long long ldd, uint32_t n4, uint32_t cols4, int act, float4* __restrict__ dpre) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / cols4, c = (i - r * cols4) * 4;
    const long long dr = map_row(dmap, (int)r);
    const float4 g = dr < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(dout + dr * ldd + c);
    const float4 v = pre[i];
    dpre[i] = make_float4(g.x * act_grad1(v.x, act), g.y * act_grad1(v.y, act), g.z * act_grad1(v.z, act),
                          g.w * act_grad1(v.w, act));
  }
}
__global__ void __launch_bounds__(256) k_residual4(const float* __restrict__ base, RowMap bmap, const float4* __restrict__ y,
                                                   const float* __restrict__ scale, uint32_t rows_per_sample,
                                                   const float* __restrict__ table, uint32_t period, uint32_t n4, uint32_t d4,
                                                   float4* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / d4, c = (i - r * d4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y) {
      v = y[i];
      if (scale) { const float sc = scale[r / rows_per_sample]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
    }
    if (base) {
      const float4 b = *reinterpret_cast<const float4*>(base + map_row(bmap, (int)r) * (long long)(4 * d4) + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (table) {
      const float4 t = *reinterpret_cast<const float4*>(table + (long long)(r % period) * (4 * d4) + c);
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    out[i] = v;
  }
}
__global__ void __launch_bounds__(256) k_scale_rows4(const float4* __restrict__ src, const float* __restrict__ scale, uint32_t rps,
                                                     uint32_t n4, uint32_t d4, float4* __restrict__ dst) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 v = src[i];
    if (scale) { const float sc = scale[(i / d4) / rps]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
    dst[i] = v;
  }
}
static inline unsigned ew_grid(long long n) { return (unsigned)std::min<long long>((n + 255) / 256, 148 * 32); }
static inline bool fits_u32(long long n) { return n > 0 && n < (1ll << 31); }
cudaError_t launch_act_fwd(const float* pre, long long n, int act, float* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (n % 4 == 0 && fits_u32(n) && al16(pre) && al16(out))
    k_act_fwd4<<<ew_grid(n / 4), 256, 0, st>>>((const float4*)pre, (uint32_t)(n / 4), act, (float4*)out);
  else
    k_act_fwd<<<ew_grid(n), 256, 0, st>>>(pre, n, act, out);
  return cudaGetLastError();
}
cudaError_t launch_act_bwd(const float* pre, const float* dout, const RowMap& dmap, long long ldd, long long rows, int cols,
                           int act, float* dpre, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (cols % 4 == 0 && ldd % 4 == 0 && fits_u32(rows * cols) && al16(pre) && al16(dout) && al16(dpre))
    k_act_bwd4<<<ew_grid(rows * cols / 4), 256, 0, st>>>((const float4*)pre, dout, dmap, ldd, (uint32_t)(rows * cols / 4),
                                                         (uint32_t)(cols / 4), act, (float4*)dpre);
  else
    k_act_bwd<<<ew_grid(rows * cols), 256, 0, st>>>(pre, dout, dmap, ldd, rows, cols, act, dpre);
  return cudaGetLastError();
}

// out[r] = (base ? base[map(r)] : 0) + scale[r / rows_per_sample] * y[r] (+ table[r % period])
// residual add with the stochastic-depth factor of vit:16-28 (scale = mask / keep_prob, or null = 1).
__global__ void k_residual(const float* __restrict__ base, RowMap bmap, const float* __restrict__ y,
                           const float* __restrict__ scale, int rows_per_sample, const float* __restrict__ table,
                           int period, long long rows, int d, float* __restrict__ out) {
  const long long n = rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const int c = (int)(i - r * d);
    float v = y ? y[i] * (scale ? scale[r / rows_per_sample] : 1.f) : 0.f;
    if (base) v += base[map_row(bmap, (int)r) * d + c];
    if (table) v += table[(r % period) * d + c];
    out[i] = v;
  }
}
__global__ void __launch_bounds__(256) k_act_bwd_mapped4(const float* __restrict__ hp, const float* __restrict__ dhp, RowMap map,
                                                         long long ld, uint32_t n4, uint32_t cols4, float4* __restrict__ dpre) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / cols4, c = (i - r * cols4) * 4;
    const long long pr = map_row(map, (int)r);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pr >= 0) {
      const float4 h = *reinterpret_cast<const float4*>(hp + pr * ld + c);
      const float4 g = *reinterpret_cast<const float4*>(dhp + pr * ld + c);
      v = make_float4(h.x > 0.f ? g.x : 0.f, h.y > 0.f ? g.y : 0.f, h.z > 0.f ? g.z : 0.f, h.w > 0.f ? g.w : 0.f);
    }
    dpre[i] = v;
  }
}
cudaError_t launch_act_bwd_mapped(const float* hp, const float* dhp, const RowMap& map, long long ld, long long rows,
                                  int cols, float* dpre, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (cols % 4 == 0 && ld % 4 == 0 && fits_u32(rows * cols) && al16(hp) && al16(dhp) && al16(dpre))
    k_act_bwd_mapped4<<<ew_grid(rows * cols / 4), 256, 0, st>>>(hp, dhp, map, ld, (uint32_t)(rows * cols / 4),
                                                                (uint32_t)(cols / 4), (float4*)dpre);
  else
    k_act_bwd_mapped<<<ew_grid(rows * cols), 256, 0, st>>>(hp, dhp, map, ld, rows, cols, dpre);
  return cudaGetLastError();
}
cudaError_t launch_residual(const float* base, const RowMap& bmap, const float* y, const float* scale,
                            int rows_per_sample, const float* table, int period, long long rows, int d, float* out,
                            cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (d % 4 == 0 && fits_u32(rows * d) && al16(base) && al16(y) && al16(table) && al16(out))
    k_residual4<<<ew_grid(rows * d / 4), 256, 0, st>>>(base, bmap, (const float4*)y, scale, (uint32_t)rows_per_sample, table,
                                                       (uint32_t)period, (uint32_t)(rows * d / 4), (uint32_t)(d / 4), (float4*)out);
  else
    k_residual<<<ew_grid(rows * d), 256, 0, st>>>(base, bmap, y, scale, rows_per_sample, table, period, rows, d, out);
  return cudaGetLastError();
}

// dst[r] = scale[r / rps] * src[r]       (gradient through the stochastic-depth factor)
__global__ void k_scale_rows(const float* __restrict__ src, const float* __restrict__ scale, int rps, long long rows,
                             int d, float* __restrict__ dst) {
  const long long n = rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i] * (scale ? scale[(i / d) / rps] : 1.f);
}
cudaError_t launch_scale_rows(const float* src, const float* scale, int rps, long long rows, int d, float* dst,
                              cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (d % 4 == 0 && fits_u32(rows * d) && al16(src) && al16(dst))
    k_scale_rows4<<<ew_grid(rows * d / 4), 256, 0, st>>>((const float4*)src, scale, (uint32_t)rps, (uint32_t)(rows * d / 4),
                                                         (uint32_t)(d / 4), (float4*)dst);
  else
    k_scale_rows<<<ew_grid(rows * d), 256, 0, st>>>(src, scale, rps, rows, d, dst);
  return cudaGetLastError();
}

// dst[map(r)] += src[r]   (adjoint of the strided identity gather; map is injective)
__global__ void k_scatter_add(const float* __restrict__ src, RowMap dmap, long long rows, int d, float* __restrict__ dst) {
  const long long n = rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const long long dr = map_row(dmap, (int)r);
    if (dr >= 0) dst[dr * d + (i - r * d)] += src[i];
  }
}
cudaError_t launch_scatter_add(const float* src, const RowMap& dmap, long long rows, int d, float* dst, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  k_scatter_add<<<ew_grid(rows * d), 256, 0, st>>>(src, dmap, rows, d, dst);
  return cudaGetLastError();
}

// Key-point embedding (K = 2): e[r][c] = x[r][0] W[0][c] + x[r][1] W[1][c] + b[c] + pe[r % J][c]  (net:321-323);
// frames without 2-D input see zeros (the caller-side mask multiply, train.py:474).
// With `list` (the valid-frame gather list of the stride mask) output frame f reads input frame list[f]: the spatial stage of
// the training step only runs on frames whose result survives the token fill (masked frames have zero gradient, net:350).
__global__ void k_embed_fwd(const float* __restrict__ x2d, const uint8_t* __restrict__ mask, const int* __restrict__ list, int J,
                            long long rows, int d, const float* __restrict__ Wk, const float* __restrict__ b,
                            const float* __restrict__ pe, float* __restrict__ out) {
  const long long n = rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const int c = (int)(i - r * d);
    const long long f = r / J, j = r - f * J;
    const long long sr = list ? (long long)list[f] * J + j : r;
    const bool ok = list || !mask || mask[f];
    const float x0 = ok ? x2d[2 * sr] : 0.f, x1 = ok ? x2d[2 * sr + 1] : 0.f;
    out[i] = (fmaf(x1, Wk[d + c], x0 * Wk[c]) + b[c]) + pe[j * d + c];
  }
}
cudaError_t launch_embed_fwd(const float* x2d, const uint8_t* mask, const int* list, int J, long long rows, int d,
                             const float* Wk, const float* b, const float* pe, float* out, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  k_embed_fwd<<<ew_grid(rows * d), 256, 0, st>>>(x2d, mask, list, J, rows, d, Wk, b, pe, out);
  return cudaGetLastError();
}
// dW[k][c] += sum_r x[r][k] * de[r][c]   (k in {0,1})
__global__ void k_embed_wgrad(const float* __restrict__ x2d, const uint8_t* __restrict__ mask, const int* __restrict__ list, int J,
                              const float* __restrict__ de, long long rows, int d, float* __restrict__ dW) {
  const int c = threadIdx.x % d;
  const int sub = threadIdx.x / d, nsub = blockDim.x / d;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  float s0 = 0.f, s1 = 0.f;
  for (long long r = r0 + sub; r < r1; r += nsub) {
    const long long f = r / J;
    if (!list && mask && !mask[f]) continue;
    const long long sr = list ? (long long)list[f] * J + (r - f * J) : r;
    const float g = de[r * d + c];
    s0 = fmaf(x2d[2 * sr], g, s0);
    s1 = fmaf(x2d[2 * sr + 1], g, s1);
  }
  __shared__ float red[2][256];
  red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1;
  __syncthreads();
  if (sub == 0) {
    for (int i = 1; i < nsub; ++i) { s0 += red[0][i * d + c]; s1 += red[1][i * d + c]; }
    dW[(long long)blockIdx.x * 2 * d + c] = s0;                 // partial slab of this row range
    dW[(long long)blockIdx.x * 2 * d + d + c] = s1;
  }
}
cudaError_t launch_embed_wgrad(const float* x2d, const uint8_t* mask, const int* list, int J, const float* de, long long rows,
                               int d, float* dW, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (256 % d) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)std::min<long long>(148 * 4, (rows + 255) / 256);
  if ((size_t)grid * 2 * d > g_red_floats) return cudaErrorInvalidValue;
  float* scr = red_region((size_t)grid * 2 * d, st);
  if (!scr) return cudaErrorInvalidValue;
  k_embed_wgrad<<<grid, 256, 0, st>>>(x2d, mask, list, J, de, rows, d, scr);
  return reduce_partials(scr, (int)grid, 2LL * d, dW, 2LL * d, nullptr, st);
}

// Token fill (net:350-352), dense form used in training:
//   x[r] = m ? keep[r] * s[pos[r]] : token ; x += pe[r % n_tok].  Backward: ds[i] = keep[list[i]] * dx[list[i]].
// `pos` maps a token row to its row in the compact (valid frames only) output of the spatial stage and `list` is its
// inverse; both null: s / ds are dense (models without a stride mask).  `keep` = the random token-masking factors
// (net:336-338), null when TOKEN_MASK_RATE is 0.
__global__ void k_fill_fwd(const float* __restrict__ s, const uint8_t* __restrict__ mask, const int* __restrict__ pos,
                           const float* __restrict__ keep, const float* __restrict__ token, const float* __restrict__ pe,
                           int n_tok, long long rows, int d, float* __restrict__ x) {
  const long long n = rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const int c = (int)(i - r * d);
    float v;
    if (!mask || mask[r]) v = s[(pos ? (long long)pos[r] : r) * d + c] * (keep ? keep[r] : 1.f);
    else v = token[c];
    x[i] = v + pe[(r % n_tok) * d + c];
  }
}
// rows = rows of ds: the compact count with a list, all token rows otherwise
__global__ void k_fill_bwd(const float* __restrict__ dx, const uint8_t* __restrict__ mask, const int* __restrict__ list,
                           const float* __restrict__ keep, long long rows, int d, float* __restrict__ ds) {
  const long long n = rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const long long sr = list ? (long long)list[r] : r;
    const float v = (list || !mask || mask[sr]) ? dx[sr * d + (i - r * d)] : 0.f;
    ds[i] = v * (keep ? keep[sr] : 1.f);
  }
}
// 16-byte forms of the two kernels above (d % 4 == 0, aligned pointers, 32-bit element counts)
__global__ void __launch_bounds__(256) k_fill_fwd4(const float* __restrict__ s, const uint8_t* __restrict__ mask,
                                                   const int* __restrict__ pos, const float* __restrict__ keep,
                                                   const float* __restrict__ token, const float* __restrict__ pe, uint32_t n_tok,
                                                   uint32_t n4, uint32_t d4, float4* __restrict__ x) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / d4, c = (i - r * d4) * 4;
    float4 v;
    if (!mask || mask[r]) {
      v = *reinterpret_cast<const float4*>(s + (pos ? (long long)pos[r] : (long long)r) * (4 * d4) + c);
      if (keep) { const float kf = keep[r]; v.x *= kf; v.y *= kf; v.z *= kf; v.w *= kf; }
    } else {
      v = *reinterpret_cast<const float4*>(token + c);
    }
    const float4 e = *reinterpret_cast<const float4*>(pe + (long long)(r % n_tok) * (4 * d4) + c);
    x[i] = make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w);
  }
}
__global__ void __launch_bounds__(256) k_fill_bwd4(const float* __restrict__ dx, const uint8_t* __restrict__ mask,
                                                   const int* __restrict__ list, const float* __restrict__ keep, uint32_t n4,
                                                   uint32_t d4, float4* __restrict__ ds) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const uint32_t r = i / d4, c = (i - r * d4) * 4;
    const long long sr = list ? (long long)list[r] : (long long)r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (list || !mask || mask[sr]) v = *reinterpret_cast<const float4*>(dx + sr * (4 * d4) + c);
    if (keep) { const float kf = keep[sr]; v.x *= kf; v.y *= kf; v.z *= kf; v.w *= kf; }
    ds[i] = v;
  }
}
cudaError_t launch_fill_fwd(const float* s, const uint8_t* mask, const int* pos, const float* keep, const float* token,
                            const float* pe, int n_tok, long long rows, int d, float* x, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (d % 4 == 0 && fits_u32(rows * d) && al16(s) && al16(token) && al16(pe) && al16(x))
    k_fill_fwd4<<<ew_grid(rows * d / 4), 256, 0, st>>>(s, mask, pos, keep, token, pe, (uint32_t)n_tok, (uint32_t)(rows * d / 4),
                                                       (uint32_t)(d / 4), (float4*)x);
  else
    k_fill_fwd<<<ew_grid(rows * d), 256, 0, st>>>(s, mask, pos, keep, token, pe, n_tok, rows, d, x);
  return cudaGetLastError();
}
cudaError_t launch_fill_bwd(const float* dx, const uint8_t* mask, const int* list, const float* keep, long long rows, int d,
                            float* ds, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  if (d % 4 == 0 && fits_u32(rows * d) && al16(dx) && al16(ds))
    k_fill_bwd4<<<ew_grid(rows * d / 4), 256, 0, st>>>(dx, mask, list, keep, (uint32_t)(rows * d / 4), (uint32_t)(d / 4), (float4*)ds);
  else
    k_fill_bwd<<<ew_grid(rows * d), 256, 0, st>>>(dx, mask, list, keep, rows, d, ds);
  return cudaGetLastError();
}
// pos[list[i]] = i (inverse of the gather list) and dst[i] = src[list[i]] (per-frame stochastic-depth factors -> compact order)
__global__ void k_invert_list(const int* __restrict__ list, int n, int* __restrict__ pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pos[list[i]] = i;
}
__global__ void k_gather_f32(const float* __restrict__ src, const int* __restrict__ list, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[list[i]];
}
cudaError_t launch_invert_list(const int* list, int n, int* pos, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_invert_list<<<(n + 255) / 256, 256, 0, st>>>(list, n, pos);
  return cudaGetLastError();
}
cudaError_t launch_gather_f32(const float* src, const int* list, int n, float* dst, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_gather_f32<<<(n + 255) / 256, 256, 0, st>>>(src, list, n, dst);
  return cudaGetLastError();
}

// =================================================================================================
// Loss (train.py:467-491, losses_3d.py:13-14) and its gradient.
//   gt is root-centred on the fly: gt[b,n,j] - gt[b,n,root]; central gt = centred gt[:, n_tok/2].
//   loss = wc * sum ||gt_c - pred_c|| / (BS*J) + ws * sum ||gt - pred|| / (BS*N*J), BS = config BATCH_SIZE.
//   dpred = -w * (gt - pred) / ||gt - pred||   (NaN at a zero residual, like tf.norm — documented, not "fixed").
// One thread per (sample, token-or-central, joint); per-block partial sums go to `partials` (deterministic
// second pass in k_loss_reduce).
// =================================================================================================
__global__ void k_loss(const float* __restrict__ full, const float* __restrict__ central, const float* __restrict__ gt,
                       int B, int n_tok, int J, int root, float w_seq, float w_cen, float* __restrict__ dfull,
                       float* __restrict__ dcentral, float* __restrict__ partials) {
  const long long n_seq = full ? (long long)B * n_tok * J : 0, n_cen = (long long)B * J;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (i < n_seq + n_cen) {
    long long b, n, j;
    const float* pred;
    float* dp;
    float w;
    if (i < n_seq) {
      b = i / ((long long)n_tok * J); n = (i / J) % n_tok; j = i % J;
      pred = full + i * 3; dp = dfull + i * 3; w = w_seq;
    } else {
      const long long k = i - n_seq;
      b = k / J; n = n_tok / 2; j = k % J;
      pred = central + k * 3; dp = dcentral + k * 3; w = w_cen;
    }
    const float* g = gt + ((b * n_tok + n) * J + j) * 3;
    const float* g0 = gt + ((b * n_tok + n) * J + root) * 3;
    const float dx = (g[0] - g0[0]) - pred[0], dy = (g[1] - g0[1]) - pred[1], dz = (g[2] - g0[2]) - pred[2];
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    contrib = w * nrm;
    const float s = -w / nrm;
    dp[0] = s * dx; dp[1] = s * dy; dp[2] = s * dz;
  }
  __shared__ float sm[8];
  contrib = warp_sum_t(contrib);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sm[k];
    partials[blockIdx.x] = s;
  }
}
__global__ void k_loss_reduce(const float* __restrict__ partials, int n, float* __restrict__ loss) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = warp_sum_t(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    s = warp_sum_t(s);
    if (threadIdx.x == 0) loss[0] = s;
  }
}
int loss_blocks(int B, int n_tok, int J, bool has_full) {
  const long long n = (has_full ? (long long)B * n_tok * J : 0) + (long long)B * J;
  return (int)((n + 255) / 256);
}
cudaError_t launch_loss(const float* full, const float* central, const float* gt, int B, int n_tok, int J, int root,
                        float w_seq, float w_cen, float* dfull, float* dcentral, float* partials, float* loss,
                        cudaStream_t st) {
  const int nb = loss_blocks(B, n_tok, J, full != nullptr);
  k_loss<<<nb, 256, 0, st>>>(full, central, gt, B, n_tok, J, root, w_seq, w_cen, dfull, dcentral, partials);
  k_loss_reduce<<<1, 1024, 0, st>>>(partials, nb, loss);
  return cudaGetLastError();
}

// =================================================================================================
// Stochastic depth (vit:16-28): scale[i] = floor(u_i + keep) / keep with u_i from a counter-based hash RNG
// (splitmix64 of (seed, stream, i)); rate 0 gives scale 1.
// =================================================================================================
__global__ void k_droppath_scale(unsigned long long seed, unsigned long long stream, long long n, float keep,
                                 float* __restrict__ scale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (stream * 0x100000001B3ULL + (unsigned long long)i + 1ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);      // [0, 1)
  scale[i] = floorf(u + keep) / keep;
}
// Random token masking (net:287-311, training only, TOKEN_MASK_RATE > 0, masked value 0): keep[b, n] = 0 where
// u < rate and n != n_tok / 2 (the central token is never masked), else 1; x[row] *= keep[row].
__global__ void k_token_mask_draw(unsigned long long seed, unsigned long long stream, long long rows, int n_tok, float rate,
                                  float* __restrict__ keep) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= rows) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (stream * 0x100000001B3ULL + (unsigned long long)i + 1ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
  keep[i] = (u < rate && (int)(i % n_tok) != n_tok / 2) ? 0.f : 1.f;
}
cudaError_t launch_token_mask_draw(unsigned long long seed, unsigned long long stream, long long rows, int n_tok, float rate,
                                   float* keep, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  k_token_mask_draw<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(seed, stream, rows, n_tok, rate, keep);
  return cudaGetLastError();
}

cudaError_t launch_droppath_scale(unsigned long long seed, unsigned long long stream, long long n, float keep,
                                  float* scale, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  k_droppath_scale<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seed, stream, n, keep, scale);
  return cudaGetLastError();
}

// =================================================================================================
// Fused multi-tensor AdamW, tfa.optimizers.AdamW (DecoupledWeightDecayExtension + Keras Adam) semantics:
//   p -= wd_t * p  (NOT multiplied by lr; every variable decays);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= lr_t * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)          (train.py:403-415)
// One pass over the flat parameter buffer: 7 x 4 B per parameter of HBM traffic.  Optional EMA in the same pass:
//   ema -= (1 - d) * (ema - p)   (train.py:502-504).
// =================================================================================================
__global__ void k_adamw(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                        long long n, float wd, float alpha, float b1, float b2, float eps, float* __restrict__ ema,
                        float ema_decay) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x = pa[k];
      x -= wd * x;
      ma[k] = b1 * ma[k] + (1.f - b1) * ga[k];
      va[k] = b2 * va[k] + (1.f - b2) * ga[k] * ga[k];
      x -= alpha * ma[k] / (sqrtf(va[k]) + eps);
      pa[k] = x;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (ema) {
      float4 e = reinterpret_cast<float4*>(ema)[i];
      e.x -= (1.f - ema_decay) * (e.x - pp.x); e.y -= (1.f - ema_decay) * (e.y - pp.y);
      e.z -= (1.f - ema_decay) * (e.z - pp.z); e.w -= (1.f - ema_decay) * (e.w - pp.w);
      reinterpret_cast<float4*>(ema)[i] = e;
    }
  }
}
cudaError_t launch_adamw(float* p, float* m, float* v, const float* g, long long n, float wd, float alpha, float b1,
                         float b2, float eps, float* ema, float ema_decay, cudaStream_t st) {
  if (n % 4) return cudaErrorInvalidValue;
  k_adamw<<<148 * 8, 256, 0, st>>>(p, m, v, g, n, wd, alpha, b1, b2, eps, ema, ema_decay);
  return cudaGetLastError();
}

}  // namespace uu
