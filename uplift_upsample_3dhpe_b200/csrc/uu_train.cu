// Training step of the hot path (train.py:464-506): forward with saved activations and stochastic depth,
// MPJPE loss, hand-derived backward for every trainable tensor, and the fused AdamW/EMA update.
// All arithmetic is fp32 (the reference trains in fp32); gradients live in one flat buffer with the same
// layout as the parameters so the data-parallel exchange is a single sum all-reduce.
#include <unordered_map>
#include <unordered_set>
#include <cstring>

#include "model.cuh"
#include "train.cuh"

namespace uu {

const std::string& get_error();

struct BlkTape {   // saved activations of one transformer block
  float *x0 = nullptr, *y1 = nullptr, *qkv = nullptr, *o = nullptr, *x1 = nullptr, *y2 = nullptr, *hpre = nullptr,
        *hact = nullptr, *x2 = nullptr;
  // stochastic depth: the reference calls drop_path_layer twice per block (vit:185-190, net:131-135), each call with
  // its own tf.random.uniform draw, so the attention branch and the MLP branch are dropped independently
  float *scale = nullptr, *scale2 = nullptr;      // per-sample factors of the attention / MLP branch
  // spatial blocks of a strided-input model run on the valid frames only: the draws are made per frame of the full batch
  // (what uu_get_droppath_scale reports) and gathered into scale / scale2 in compact order
  float *scale_full = nullptr, *scale2_full = nullptr;
  float keep = 1.f;
};

struct TrainState {
  int B = 0;                                    // capacity / current batch of the arena
  int global_batch = 0, root = 6;
  float w_center = 1.f, w_seq = 1.f;
  float dpr[3] = {0.f, 0.f, 0.f};
  int droppath_mode = 0;                        // 0 = off, 1 = hash RNG
  unsigned long long seed = 0;
  std::vector<void*> pool;
  // tapes
  float *emb = nullptr;                         // not kept separately: spatial block 0 x0
  std::vector<BlkTape> sp, tp, st;
  float *sp_normed = nullptr;                   // S: spatial_norm output viewed as [R, J*ds]
  float *s4 = nullptr;                          // spatial_to_temporal output [R, d]
  float *full = nullptr, *central = nullptr, *dfull = nullptr, *dcentral = nullptr;
  std::vector<float*> hp, dhp;                  // zero-padded conv input / its gradient per strided block
  // gradient flow + scratch
  float *dx_sp = nullptr, *dx_t = nullptr;      // [R*J, ds], [R, d]
  std::vector<float*> dx_s;                     // per strided level output: [B*Lo, d]
  float *tmp1 = nullptr, *tmp2 = nullptr, *tmp_h = nullptr, *tmp_qkv = nullptr, *dS = nullptr;
  float *partials = nullptr, *loss = nullptr;
  // math mode 1: forward, dgrad and wgrad GEMMs of the temporal / strided blocks on tcgen05 kind::tf32 (fp32 data, TF32
  // products, fp32 accumulation — what TensorFlow 2.4 does by default on Ampere+ GPUs), their attention on bf16 hi + lo planes.
  float token_mask_rate = 0.f;                  // TOKEN_MASK_RATE (net:287-311; masked value 0), training only
  float* tok_keep = nullptr;                    // [R] 0 / 1 factors drawn for the current step
  int math = 0;
  int *gl_scratch = nullptr, *gl_list = nullptr, *gl_pos = nullptr, *gl_count = nullptr;   // valid-frame gather list (stride mask)
  int attn_split = 3;                           // attn_mma.cu: 2 = bf16 hi + lo planes (2^-16 per product), 3 = compensated TF32, 1 = TF32
  float* wg_scratch = nullptr;                  // split-K partial tiles of the tensor-core wgrad (wgrad_tc.cu)
  float* red_scratch = nullptr;                 // partial slabs of the deterministic two-pass reductions (train_kernels.cu)
  float* red_arena = nullptr;                   // ... of the backward pass, whose second passes are batched (RED_ARENA_FLOATS)
  struct PackedQkv { float *W = nullptr, *b = nullptr, *dW = nullptr, *db = nullptr; };
  std::unordered_map<const float*, PackedQkv> pk;   // W_q -> [W_q | W_k | W_v] (d, 3d), bias (3d) and their gradient slabs
  std::unordered_set<const float*> pk_valid;         // packed copies refreshed once per step
  std::unordered_map<const float*, float*> wt;   // W (K, N) -> W^T (N, K) copies for the forward GEMMs
  std::unordered_set<const float*> wt_valid;     // refreshed once per forward/backward call
};

void train_state_destroy(uu_model* m) {
  if (!m->train) return;
  free_pool(m->train->pool);
  delete m->train;
  m->train = nullptr;
}

static float* G(uu_model* m, const std::string& g, int i) {   // gradient slot of a tensor
  const size_t off = tensor_offset(m, g, i);
  return off == (size_t)-1 ? nullptr : m->grads + off;
}

constexpr size_t RED_ARENA_FLOATS = (size_t)48 << 20;      // 192 MB: about one backward pass of 512 windows between flushes
struct DeferGuard {                                         // the deferred-reduction mode never outlives train_fb
  ~DeferGuard() { train_reduce_defer_abort(); }
};

struct Ctx {
  uu_model* m;
  TrainState* t;
  cudaStream_t st;
};

#define UU_TL(expr) UU_CUDA(expr)

static int falloc(TrainState* t, float** p, size_t n_floats, bool zero = false) {
  void* q;
  if (dev_alloc(t->pool, &q, n_floats * sizeof(float), zero)) return 1;
  *p = (float*)q;
  return 0;
}

// out[c][r] = in[r][c]
__global__ void k_transpose_f32(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(long long)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(long long)c * rows + r] = tile[threadIdx.x][i];
  }
}

// tcgen05 kind::tf32 path for the large, regular GEMMs (math mode 1)
static bool tf32_ok(const Ctx& c, const void* A, long long lda, const void* Bt, long long ldb, int M, int N, int K,
                    const void* C, long long ldc) {
  return c.t->math == 1 && M >= 256 && N >= 64 && N % 64 == 0 && K >= 64 && K % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
         ldc % 4 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)Bt & 15) == 0 && ((uintptr_t)C & 15) == 0;
}
static int tf32_gemm(Ctx& c, const float* A, long long lda, int M, int K, const float* Bt, long long ldb, int N,
                     const Epilogue& e, float* C, long long ldc) {
  TcGemmPlan* p = nullptr;
  if (tc_gemm_plan_create_tf32(&p, A, lda, M, K, Bt, ldb, N)) return 1;
  cudaError_t err = tc_gemm_launch(p, e, C, 0, ldc, c.st);
  tc_gemm_plan_destroy(p);      // (tensor maps are kernel parameters: copied at launch)
  UU_CUDA(err);
  return 0;
}
static int weight_transposed(Ctx& c, const float* Wm, int K, int N, const float** out) {
  TrainState* t = c.t;
  auto it = t->wt.find(Wm);
  if (it == t->wt.end()) {
    float* p;
    if (falloc(t, &p, (size_t)K * N)) return 1;
    it = t->wt.emplace(Wm, p).first;
  }
  if (!t->wt_valid.count(Wm)) {
    k_transpose_f32<<<dim3((N + 31) / 32, (K + 31) / 32), dim3(32, 8), 0, c.st>>>(Wm, K, N, it->second);
    UU_CUDA(cudaGetLastError());
    t->wt_valid.insert(Wm);
  }
  *out = it->second;
  return 0;
}

// q | k | v as ONE linear layer: the three Keras tensors (d, d) are packed side by side into (d, 3d) once per step, so the
// block reads its LayerNorm output once (forward), writes dX once instead of three accumulating passes (dgrad) and reads X /
// dQKV once (wgrad); the packed gradient slab is added back onto the three tensors' gradient slots.
__global__ void k_pack_qkv(const float* __restrict__ wq, const float* __restrict__ wk, const float* __restrict__ wv,
                           const float* __restrict__ bq, const float* __restrict__ bk, const float* __restrict__ bv, int d,
                           float* __restrict__ W, float* __restrict__ b) {
  const int n = d * 3 * d;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n + 3 * d; i += gridDim.x * blockDim.x) {
    if (i < n) {
      const int r = i / (3 * d), c = i - r * 3 * d, k = c / d, cc = c - k * d;
      W[i] = (k == 0 ? wq : k == 1 ? wk : wv)[r * d + cc];
    } else {
      const int c = i - n, k = c / d, cc = c - k * d;
      b[c] = (k == 0 ? bq : k == 1 ? bk : bv)[cc];
    }
  }
}
__global__ void k_unpack_qkv_add(const float* __restrict__ dW, const float* __restrict__ db, int d, float* __restrict__ gq,
                                 float* __restrict__ gk, float* __restrict__ gv, float* __restrict__ gbq,
                                 float* __restrict__ gbk, float* __restrict__ gbv) {
  const int n = d * 3 * d;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n + (db ? 3 * d : 0); i += gridDim.x * blockDim.x) {
    if (i < n) {
      const int r = i / (3 * d), c = i - r * 3 * d, k = c / d, cc = c - k * d;
      (k == 0 ? gq : k == 1 ? gk : gv)[r * d + cc] += dW[i];
    } else {
      const int c = i - n, k = c / d, cc = c - k * d;
      (k == 0 ? gbq : k == 1 ? gbk : gbv)[cc] += db[c];
    }
  }
}
static int packed_qkv(Ctx& c, const std::string& g, int d, TrainState::PackedQkv* out) {
  TrainState* t = c.t;
  uu_model* m = c.m;
  const float* key = W(m, g, 2);
  auto it = t->pk.find(key);
  if (it == t->pk.end()) {
    TrainState::PackedQkv p;
    if (falloc(t, &p.W, (size_t)d * 3 * d) || falloc(t, &p.b, 3 * (size_t)d) || falloc(t, &p.dW, (size_t)d * 3 * d) ||
        falloc(t, &p.db, 3 * (size_t)d))
      return 1;
    it = t->pk.emplace(key, p).first;
  }
  if (!t->pk_valid.count(key)) {
    const int n = d * 3 * d + 3 * d;
    k_pack_qkv<<<std::min(148 * 4, (n + 255) / 256), 256, 0, c.st>>>(W(m, g, 2), W(m, g, 4), W(m, g, 6), W(m, g, 3), W(m, g, 5),
                                                                   W(m, g, 7), d, it->second.W, it->second.b);
    UU_CUDA(cudaGetLastError());
    t->pk_valid.insert(key);
  }
  *out = it->second;
  return 0;
}

// y = x @ W + b (forward linear), C may use a row map and a wider leading dimension
static int lin_fwd(Ctx& c, const float* A, long long lda, int M, int K, const float* Wm, int N, const float* bias,
                   float* C, long long ldc, const RowMap* cmap = nullptr) {
  if (tf32_ok(c, A, lda, Wm, K, M, N, K, C, ldc)) {
    const float* Wt;
    if (weight_transposed(c, Wm, K, N, &Wt)) return 1;
    Epilogue e;
    e.bias = bias;
    if (cmap) e.cmap = *cmap;
    return tf32_gemm(c, A, lda, M, K, Wt, K, N, e, C, ldc);
  }
  GemmGen g;
  g.A = A; g.lda = lda; g.B = Wm; g.ldb = N; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.bias = bias;
  if (cmap) g.cmap = *cmap;
  UU_TL(launch_gemm_gen(g, c.st));
  return 0;
}
// dX (+)= dY @ W^T ; dW += X^T dY ; db += colsum(dY).  X rows may be lda apart, dY rows ldy apart.
static int lin_bwd(Ctx& c, const float* X, long long ldx, const float* dY, long long ldy, int M, int K, int N,
                   const float* Wm, float* dX, long long lddx, int accumulate_dx, float* dW, float* db,
                   const RowMap* dxmap = nullptr) {
  if (dX && tf32_ok(c, dY, ldy, Wm, N, M, K, N, dX, lddx)) {
    // dX = dY . W^T: the weight matrix (K, N) is already the K-major B operand of this product
    Epilogue e;
    if (dxmap) e.cmap = *dxmap;
    if (accumulate_dx) { e.flags = EPI_RESIDUAL; e.res = dX; e.ldr = lddx; if (dxmap) e.rmap = *dxmap; }
    if (tf32_gemm(c, dY, ldy, M, N, Wm, N, K, e, dX, lddx)) return 1;
  } else if (dX && !dxmap && c.t->math == 1 && K > 64 && K % 64 != 0 &&
             tf32_ok(c, dY, ldy, Wm, N, M, K / 64 * 64, N, dX, lddx)) {
    // an output width that is not a multiple of the tensor-core column tile (the 544-wide input of spatial_to_temporal_fc):
    // the first floor(K / 64) * 64 columns on tcgen05 kind::tf32, the remaining K % 64 on the generic kernel
    const int K0 = K / 64 * 64;
    Epilogue e;
    if (accumulate_dx) { e.flags = EPI_RESIDUAL; e.res = dX; e.ldr = lddx; }
    if (tf32_gemm(c, dY, ldy, M, N, Wm, N, K0, e, dX, lddx)) return 1;
    GemmGen g;
    g.A = dY; g.lda = ldy; g.B = Wm + (long long)K0 * N; g.ldb = N; g.transB = 1; g.C = dX + K0; g.ldc = lddx; g.M = M;
    g.N = K - K0; g.K = N; g.accumulate = accumulate_dx;
    UU_TL(launch_gemm_gen(g, c.st));
  } else if (dX && !dxmap && dgrad_skinny_ok(dY, ldy, M, N, K, dX, lddx)) {
    // narrow layers of the spatial blocks: streaming kernel (train_kernels.cu); W is (K, N) = (outputs of dX, inputs dY)
    UU_TL(launch_dgrad_skinny(dY, ldy, Wm, M, N, K, dX, lddx, accumulate_dx, c.st));
  } else if (dX) {
    GemmGen g;
    g.A = dY; g.lda = ldy; g.B = Wm; g.ldb = N; g.transB = 1; g.C = dX; g.ldc = lddx; g.M = M; g.N = K; g.K = N;
    g.accumulate = accumulate_dx;
    if (dxmap) g.cmap = *dxmap;
    UU_TL(launch_gemm_gen(g, c.st));
  }
  if (!dW) return 0;                                    // input gradient only (the caller takes the weight gradients)
  if (wgrad_skinny_ok(X, ldx, dY, ldy, M, K, N) && db) {
    // narrow layers of the spatial blocks: one streaming pass gives dW and db (train_kernels.cu)
    UU_TL(launch_wgrad_skinny(X, ldx, dY, ldy, M, K, N, dW, db, c.st));
    return 0;
  }
  if (c.t->math == 1 && wgrad_tc_ok(X, ldx, dY, ldy, M, K, N)) {
    // dW += X^T dY on tcgen05 kind::tf32, both operands read MN-major straight from the tape (wgrad_tc.cu)
    if (wgrad_tc(X, ldx, dY, ldy, M, K, N, dW, 1, c.t->wg_scratch, c.m->num_sms, c.st)) return 1;
  } else {
    GemmGen g;
    g.A = X; g.lda = ldx; g.transA = 1; g.B = dY; g.ldb = ldy; g.C = dW; g.ldc = N; g.M = K; g.N = N; g.K = M;
    g.accumulate = 1;
    int bm, bn;
    gemm_gen_tile(N, &bm, &bn);
    const int tiles = ((K + bm - 1) / bm) * ((N + bn - 1) / bn);
    int splits = (148 * 4 + tiles - 1) / tiles;          // two CTAs per SM, two waves
    splits = std::max(1, std::min(splits, std::max(1, M / 512)));
    g.split_k = splits;
    UU_TL(launch_gemm_gen(g, c.st));
  }
  if (db) UU_TL(launch_colsum(dY, M, N, ldy, db, c.st));
  return 0;
}

struct BlkDims {
  int d, h, S, heads;
  long long nb;       // samples (windows, or frames for the spatial blocks)
  int act;            // 0 ReLU, 1 GELU
};

static int alloc_tape(TrainState* t, BlkTape& tp, const BlkDims& b, bool strided) {
  const size_t R = (size_t)b.nb * b.S;
  if (falloc(t, &tp.y1, R * b.d) || falloc(t, &tp.qkv, R * 3 * b.d) || falloc(t, &tp.o, R * b.d) ||
      falloc(t, &tp.x1, R * b.d) || falloc(t, &tp.y2, R * b.d) || falloc(t, &tp.scale, b.nb) || falloc(t, &tp.scale2, b.nb))
    return 1;
  if (!strided) {
    if (falloc(t, &tp.hpre, R * b.h) || falloc(t, &tp.hact, R * b.h) || falloc(t, &tp.x2, R * b.d)) return 1;
  }
  return 0;
}

// ---- attention half, shared by plain and strided blocks --------------------------------------------
static int attn_half_fwd(Ctx& c, const BlkDims& b, const std::string& g, BlkTape& tp, const uint8_t* keymask,
                         int mask_stride) {
  uu_model* m = c.m;
  const long long R = b.nb * b.S;
  const int d = b.d;
  UU_TL(launch_ln_fwd_gen(tp.x0, R, d, W(m, g, 0), W(m, g, 1), 1e-5f, tp.y1, c.st));
  TrainState::PackedQkv pq;
  if (packed_qkv(c, g, d, &pq)) return 1;
  if (lin_fwd(c, tp.y1, d, (int)R, d, pq.W, 3 * d, pq.b, tp.qkv, 3 * d)) return 1;
  if (attention_small_ok(b.S, b.heads, d / b.heads, keymask))       // spatial blocks: one warp per frame (train_kernels.cu)
    UU_TL(launch_attention_small_fwd(tp.qkv, b.nb, b.S, tp.o, c.st));
  else if (attention_mma_ok(b.nb, b.S, b.heads, d / b.heads))        // temporal / strided blocks: mma.sync TF32 (attn_mma.cu)
    UU_TL(launch_attention_mma_fwd(tp.qkv, b.nb, b.S, b.heads, d / b.heads, keymask, mask_stride, tp.o, c.t->attn_split, c.st));
  else
    UU_TL(launch_attention(tp.qkv, 0, (int)b.nb, b.S, b.heads, d / b.heads, keymask, mask_stride, tp.o, c.st));
  if (lin_fwd(c, tp.o, d, (int)R, d, W(m, g, 8), d, W(m, g, 9), c.t->tmp1, d)) return 1;
  const RowMap plain;
  UU_TL(launch_residual(tp.x0, plain, c.t->tmp1, tp.keep < 1.f ? tp.scale : nullptr, b.S, nullptr, 1, R, d, tp.x1, c.st));
  return 0;
}
// dx holds d(loss)/d(x1) on entry and d(loss)/d(x0) on exit
static int attn_half_bwd(Ctx& c, const BlkDims& b, const std::string& g, BlkTape& tp, const uint8_t* keymask,
                         int mask_stride, float* dx) {
  uu_model* m = c.m;
  TrainState* t = c.t;
  const long long R = b.nb * b.S;
  const int d = b.d;
  const float* dy = dx;                       // gradient of the branch output: dx itself unless stochastic depth scales it
  if (tp.keep < 1.f) {
    UU_TL(launch_scale_rows(dx, tp.scale, b.S, R, d, t->tmp1, c.st));
    dy = t->tmp1;
  }
  if (lin_bwd(c, tp.o, d, dy, d, (int)R, d, d, W(m, g, 8), t->tmp2, d, 0, G(m, g, 8), G(m, g, 9))) return 1;
  if (!attention_small_ok(b.S, b.heads, d / b.heads, keymask) && attention_mma_ok(b.nb, b.S, b.heads, d / b.heads))
    UU_TL(launch_attention_mma_bwd(tp.qkv, t->tmp2, b.nb, b.S, b.heads, d / b.heads, keymask, mask_stride, t->tmp_qkv,
                                   t->attn_split, c.st));
  else
    UU_TL(launch_attention_bwd(tp.qkv, t->tmp2, b.nb, b.S, b.heads, d / b.heads, keymask, mask_stride, t->tmp_qkv, c.st));
  TrainState::PackedQkv pq;
  if (packed_qkv(c, g, d, &pq)) return 1;
  if (wgrad_skinny_ok(tp.y1, d, t->tmp_qkv, 3 * d, R, d, d)) {
    // narrow (spatial) layers: packed input gradient, but the one-pass dW / db kernel per tensor (its 32 x 96 instantiation
    // holds 96 accumulators per thread and runs at half the speed of three 32 x 32 passes)
    if (lin_bwd(c, tp.y1, d, t->tmp_qkv, 3 * d, (int)R, d, 3 * d, pq.W, t->tmp2, d, 0, nullptr, nullptr)) return 1;
    for (int k = 0; k < 3; ++k)
      UU_TL(launch_wgrad_skinny(tp.y1, d, t->tmp_qkv + k * d, 3 * d, R, d, d, G(m, g, 2 + 2 * k), G(m, g, 3 + 2 * k), c.st));
  } else if (c.t->math == 1 && wgrad_tc_ok(tp.y1, d, t->tmp_qkv, 3 * d, R, d, 3 * d)) {
    // one tcgen05 wgrad over the packed (d, 3d) slab (its split-K reducer runs at once), added back onto the three tensors;
    // the bias gradients go straight to their slots (their second pass is deferred: nothing may read a packed copy of them)
    UU_CUDA(cudaMemsetAsync(pq.dW, 0, sizeof(float) * (size_t)d * 3 * d, c.st));
    if (lin_bwd(c, tp.y1, d, t->tmp_qkv, 3 * d, (int)R, d, 3 * d, pq.W, t->tmp2, d, 0, pq.dW, nullptr)) return 1;
    const int n = d * 3 * d;
    k_unpack_qkv_add<<<std::min(148 * 4, (n + 255) / 256), 256, 0, c.st>>>(pq.dW, nullptr, d, G(m, g, 2), G(m, g, 4), G(m, g, 6),
                                                                         nullptr, nullptr, nullptr);
    UU_CUDA(cudaGetLastError());
    for (int k = 0; k < 3; ++k) UU_TL(launch_colsum(t->tmp_qkv + k * d, (int)R, d, 3 * d, G(m, g, 3 + 2 * k), c.st));
  } else {
    // packed input gradient, per-tensor weight / bias gradients written straight to their slots
    if (lin_bwd(c, tp.y1, d, t->tmp_qkv, 3 * d, (int)R, d, 3 * d, pq.W, t->tmp2, d, 0, nullptr, nullptr)) return 1;
    for (int k = 0; k < 3; ++k)
      if (lin_bwd(c, tp.y1, d, t->tmp_qkv + k * d, 3 * d, (int)R, d, d, nullptr, nullptr, 0, 0, G(m, g, 2 + 2 * k),
                  G(m, g, 3 + 2 * k)))
        return 1;
  }
  UU_TL(launch_ln_bwd_gen(tp.x0, t->tmp2, R, d, W(m, g, 0), 1e-5f, dx, 1, G(m, g, 0), G(m, g, 1), c.st));
  return 0;
}

static int block_fwd(Ctx& c, const BlkDims& b, const std::string& g, BlkTape& tp, const uint8_t* keymask, int mstride) {
  uu_model* m = c.m;
  const long long R = b.nb * b.S;
  const int d = b.d, h = b.h;
  if (attn_half_fwd(c, b, g, tp, keymask, mstride)) return 1;
  UU_TL(launch_ln_fwd_gen(tp.x1, R, d, W(m, g, 10), W(m, g, 11), 1e-5f, tp.y2, c.st));
  if (lin_fwd(c, tp.y2, d, (int)R, d, W(m, g, 12), h, W(m, g, 13), tp.hpre, h)) return 1;
  UU_TL(launch_act_fwd(tp.hpre, R * h, b.act, tp.hact, c.st));
  if (lin_fwd(c, tp.hact, h, (int)R, h, W(m, g, 14), d, W(m, g, 15), c.t->tmp1, d)) return 1;
  const RowMap plain;
  UU_TL(launch_residual(tp.x1, plain, c.t->tmp1, tp.keep < 1.f ? tp.scale2 : nullptr, b.S, nullptr, 1, R, d, tp.x2, c.st));
  return 0;
}
// dx: d/d(x2) on entry, d/d(x0) on exit (in place)
static int block_bwd(Ctx& c, const BlkDims& b, const std::string& g, BlkTape& tp, const uint8_t* keymask, int mstride,
                     float* dx) {
  uu_model* m = c.m;
  TrainState* t = c.t;
  const long long R = b.nb * b.S;
  const int d = b.d, h = b.h;
  const RowMap plain;
  const float* dy = dx;
  if (tp.keep < 1.f) {
    UU_TL(launch_scale_rows(dx, tp.scale2, b.S, R, d, t->tmp1, c.st));
    dy = t->tmp1;
  }
  if (lin_bwd(c, tp.hact, h, dy, d, (int)R, h, d, W(m, g, 14), t->tmp_h, h, 0, G(m, g, 14), G(m, g, 15))) return 1;
  UU_TL(launch_act_bwd(tp.hpre, t->tmp_h, plain, h, R, h, b.act, t->tmp_h, c.st));
  if (lin_bwd(c, tp.y2, d, t->tmp_h, h, (int)R, d, h, W(m, g, 12), t->tmp2, d, 0, G(m, g, 12), G(m, g, 13))) return 1;
  UU_TL(launch_ln_bwd_gen(tp.x1, t->tmp2, R, d, W(m, g, 10), 1e-5f, dx, 1, G(m, g, 10), G(m, g, 11), c.st));
  return attn_half_bwd(c, b, g, tp, keymask, mstride, dx);
}

static int ensure_train(uu_model* m, int B) {
  const uu_spec& s = m->spec;
  if (!m->grads) {
    UU_CUDA(cudaMalloc(&m->grads, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMalloc(&m->adam_m, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMalloc(&m->adam_v, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMemset(m->grads, 0, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMemset(m->adam_m, 0, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMemset(m->adam_v, 0, sizeof(float) * m->n_alloc));
  }
  if (!m->train) m->train = new TrainState();
  TrainState* t = m->train;
  if (t->B == B) return 0;
  UU_CUDA(cudaDeviceSynchronize());
  free_pool(t->pool);
  t->wt.clear(); t->wt_valid.clear();          // the transposed-weight copies live in the same pool
  t->pk.clear(); t->pk_valid.clear();
  t->sp.assign(s.spatial_depth, BlkTape());
  t->tp.assign(s.temporal_depth, BlkTape());
  t->st.assign(s.n_strided, BlkTape());
  t->hp.assign(s.n_strided, nullptr); t->dhp.assign(s.n_strided, nullptr); t->dx_s.assign(s.n_strided, nullptr);
  const size_t N = s.n_tok, J = s.n_joints, ds = s.d_spatial, d = s.d_temporal, h = s.h_temporal;
  const size_t R = (size_t)B * N, Rs = R * J;
  float* x;
  if (falloc(t, &x, Rs * ds)) return 1;                         // embedding = x0 of spatial block 0
  for (int i = 0; i < s.spatial_depth; ++i) {
    BlkDims bd{(int)ds, s.h_spatial, (int)J, s.num_heads, (long long)R, 1};
    t->sp[i].x0 = x;
    if (alloc_tape(t, t->sp[i], bd, false)) return 1;
    x = t->sp[i].x2;
  }
  if (falloc(t, &t->sp_normed, Rs * ds) || falloc(t, &t->s4, R * d)) return 1;
  if (falloc(t, &x, R * d)) return 1;                           // temporal input = x0 of temporal block 0
  for (int i = 0; i < s.temporal_depth; ++i) {
    BlkDims bd{(int)d, (int)h, (int)N, s.num_heads, (long long)B, 0};
    t->tp[i].x0 = x;
    if (alloc_tape(t, t->tp[i], bd, false)) return 1;
    x = t->tp[i].x2;
  }
  for (int i = 0; i < s.n_strided; ++i) {
    const size_t L = m->seq_lens[i], Lo = m->seq_lens[i + 1];
    BlkDims bd{(int)d, (int)h, (int)L, s.num_heads, (long long)B, 0};
    if (falloc(t, &t->st[i].x0, (size_t)B * L * d)) return 1;   // x_in + PE_i
    if (alloc_tape(t, t->st[i], bd, true)) return 1;
    if (falloc(t, &t->st[i].x2, (size_t)B * Lo * d)) return 1;
    if (falloc(t, &t->hp[i], (size_t)B * Lo * s.strides[i] * h, true)) return 1;
    if (falloc(t, &t->dhp[i], (size_t)B * Lo * s.strides[i] * h, true)) return 1;
    if (falloc(t, &t->dx_s[i], (size_t)B * Lo * d)) return 1;
  }
  if (falloc(t, &t->full, R * 3 * J) || falloc(t, &t->central, (size_t)B * 3 * J) || falloc(t, &t->dfull, R * 3 * J) ||
      falloc(t, &t->dcentral, (size_t)B * 3 * J))
    return 1;
  if (falloc(t, &t->dx_sp, Rs * ds) || falloc(t, &t->dx_t, R * d)) return 1;
  const size_t rows_big = std::max(Rs, R);                      // scratch is shared by the d=32 and d=384 stages
  const size_t dmax = std::max(Rs * ds, R * d);
  (void)rows_big;
  if (falloc(t, &t->tmp1, dmax) || falloc(t, &t->tmp2, dmax) ||
      falloc(t, &t->tmp_h, std::max(Rs * (size_t)s.h_spatial, R * h)) ||
      falloc(t, &t->tmp_qkv, std::max(Rs * 3 * ds, R * 3 * d)) || falloc(t, &t->dS, R * J * ds))
    return 1;
  if (falloc(t, &t->partials, loss_blocks(B, (int)N, (int)J, true) + 8) || falloc(t, &t->loss, 4)) return 1;
  if (falloc(t, &t->tok_keep, R)) return 1;
  if (s.has_strided_input) {
    float* q;
    if (falloc(t, &q, (size_t)B + 2)) return 1;
    t->gl_scratch = reinterpret_cast<int*>(q);
    if (falloc(t, &q, R)) return 1;
    t->gl_list = reinterpret_cast<int*>(q);
    if (falloc(t, &q, R)) return 1;
    t->gl_pos = reinterpret_cast<int*>(q);
    if (falloc(t, &q, 4)) return 1;
    t->gl_count = reinterpret_cast<int*>(q);
    for (int i = 0; i < s.spatial_depth; ++i)
      if (falloc(t, &t->sp[i].scale_full, R) || falloc(t, &t->sp[i].scale2_full, R)) return 1;
  }
  if (falloc(t, &t->wg_scratch, wgrad_tc_scratch_bytes() / sizeof(float))) return 1;
  if (falloc(t, &t->red_scratch, train_reduce_scratch_floats() + 4096)) return 1;
  if (falloc(t, &t->red_arena, RED_ARENA_FLOATS)) return 1;
  t->B = B;
  return 0;
}

static float block_keep(const TrainState* t, int stage, int i, int depth) {
  if (t->droppath_mode == 0 || depth <= 1) return 1.f;
  const float rate = t->dpr[stage] * (float)i / (float)(depth - 1);      // np.linspace(0, dpr, depth)[i]
  return 1.f - rate;
}

// overlap_comm: issue the gradient all-reduce (uu_comm.cu) bucket by bucket as the backward pass finishes regions of the
// flat gradient buffer: [strided blocks, heads] after the strided stage, [temporal blocks] after the temporal stage,
// the rest (embeddings, positional tables, spatial blocks, spatial_to_temporal_fc) at the end.
static int train_fb(uu_model* m, const float* x2d, const uint8_t* mask, const float* gt3d, int B, long long step,
                    float* loss_out, cudaStream_t stream, bool overlap_comm = false) {
  const uu_spec& s = m->spec;
  UU_CHECK(m->train && m->train->global_batch > 0, "call uu_train_config first");
  UU_CHECK(B > 0 && x2d && gt3d && loss_out, "bad argument");
  UU_CHECK(!s.has_strided_input || mask, "this model has strided input: a stride mask is required");
  UU_CHECK(s.full_output, "training without the full-sequence head (USE_REFINE) is not supported");
  UU_CUDA(cudaSetDevice(m->device));
  if (ensure_train(m, B)) return 1;
  TrainState* t = m->train;
  train_reduce_scratch(t->red_scratch, train_reduce_scratch_floats());
  Ctx c{m, t, stream};
  const int N = s.n_tok, J = s.n_joints, ds = s.d_spatial, d = s.d_temporal, h = s.h_temporal, H = s.num_heads;
  const long long R = (long long)B * N, Rs = R * J;
  const bool use_mask = s.has_strided_input != 0;
  const RowMap plain;
  UU_CUDA(cudaMemsetAsync(m->grads, 0, sizeof(float) * m->n_alloc, stream));

  // Frames the stride mask drops never influence the loss (their spatial result is replaced by the upsampling token,
  // net:350-352) and receive a zero gradient: the spatial stage runs on the Rv valid frames only, in gather-list order.
  long long Rv = R;
  const int *glist = nullptr, *gpos = nullptr;
  if (use_mask) {
    UU_TL(launch_build_gather(mask, B, N, t->gl_scratch, t->gl_list, t->gl_count, stream));
    int nv = 0;
    UU_CUDA(cudaMemcpyAsync(&nv, t->gl_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
    UU_CUDA(cudaStreamSynchronize(stream));            // (the launch sizes of the spatial stage depend on the count)
    Rv = nv;
    glist = t->gl_list; gpos = t->gl_pos;
    UU_TL(launch_invert_list(glist, (int)Rv, t->gl_pos, stream));
  }
  const long long Rsv = Rv * J;

  // stochastic-depth factors for this step
  for (int stage = 0; stage < 3; ++stage) {
    std::vector<BlkTape>& tapes = stage == 0 ? t->sp : stage == 1 ? t->tp : t->st;
    const long long ns = stage == 0 ? R : B;
    for (size_t i = 0; i < tapes.size(); ++i) {
      tapes[i].keep = block_keep(t, stage, (int)i, (int)tapes.size());
      if (tapes[i].keep < 1.f) {      // two independent draws per block: RNG stream ids 2i (attention) and 2i + 1 (MLP)
        const bool compact = stage == 0 && glist;
        UU_TL(launch_droppath_scale(t->seed, (unsigned long long)step * 64ULL + stage * 16 + 2 * i, ns, tapes[i].keep,
                                    compact ? tapes[i].scale_full : tapes[i].scale, stream));
        UU_TL(launch_droppath_scale(t->seed, (unsigned long long)step * 64ULL + stage * 16 + 2 * i + 1, ns, tapes[i].keep,
                                    compact ? tapes[i].scale2_full : tapes[i].scale2, stream));
        if (compact) {
          UU_TL(launch_gather_f32(tapes[i].scale_full, glist, (int)Rv, tapes[i].scale, stream));
          UU_TL(launch_gather_f32(tapes[i].scale2_full, glist, (int)Rv, tapes[i].scale2, stream));
        }
      }
    }
  }

  // ================= forward =================
  UU_TL(launch_embed_fwd(x2d, use_mask ? mask : nullptr, glist, J, Rsv, ds, W(m, "keypoint_embedding", 0),
                         W(m, "keypoint_embedding", 1), W(m, "spatial_pe", 0), t->sp[0].x0, stream));
  const BlkDims sp_dims{ds, s.h_spatial, J, H, Rv, 1};
  for (int i = 0; i < s.spatial_depth; ++i) {
    const std::string g = "spatial_block_" + std::to_string(i + 1);
    BlkTape& tp = t->sp[i];
    if (spatial_block_fused_ok(J, ds, s.h_spatial, H, sp_dims.act)) {     // one launch per block (spatial_train.cu)
      const float* wl[16];
      for (int k = 0; k < 16; ++k) wl[k] = W(m, g, k);
      UU_TL(launch_spatial_block_fwd_tape(tp.x0, wl, tp.keep < 1.f ? tp.scale : nullptr, tp.keep < 1.f ? tp.scale2 : nullptr, Rv,
                                          tp.y1, tp.qkv, tp.o, tp.x1, tp.y2, tp.hpre, tp.hact, tp.x2, m->num_sms, stream));
    } else if (block_fwd(c, sp_dims, g, tp, nullptr, 0)) {
      return 1;
    }
  }
  float* sp_out = t->sp[s.spatial_depth - 1].x2;
  UU_TL(launch_ln_fwd_gen(sp_out, Rsv, ds, W(m, "spatial_norm", 0), W(m, "spatial_norm", 1), 1e-6f, t->sp_normed, stream));
  if (lin_fwd(c, t->sp_normed, J * ds, (int)Rv, J * ds, W(m, "spatial_to_temporal_fc", 0), d,
              W(m, "spatial_to_temporal_fc", 1), t->s4, d))
    return 1;
  const float* tok_keep = nullptr;
  if (t->token_mask_rate > 0.f) {   // random token masking (net:336-338) before the upsampling-token fill and the PE add
    UU_TL(launch_token_mask_draw(t->seed, (unsigned long long)step * 64ULL + 63ULL, R, N, t->token_mask_rate, t->tok_keep,
                                 stream));
    tok_keep = t->tok_keep;         // applied inside the fill (forward) and the gather of its gradient (backward)
  }
  UU_TL(launch_fill_fwd(t->s4, use_mask ? mask : nullptr, gpos, tok_keep, use_mask ? W(m, "strided_input_token_layer", 0) : nullptr,
                        W(m, "temporal_pe", 0), N, R, d, t->tp[0].x0, stream));
  const BlkDims tp_dims{d, h, N, H, (long long)B, 0};
  for (int i = 0; i < s.temporal_depth; ++i) {
    const uint8_t* km = (use_mask && i < s.first_strided_token_attention_layer) ? mask : nullptr;
    if (block_fwd(c, tp_dims, "temporal_block_" + std::to_string(i + 1), t->tp[i], km, N)) return 1;
  }
  float* xT = t->tp[s.temporal_depth - 1].x2;
  if (lin_fwd(c, xT, d, (int)R, d, W(m, "temporal_fc", 0), 3 * J, W(m, "temporal_fc", 1), t->full, 3 * J)) return 1;
  const float* x_in = xT;
  for (int i = 0; i < s.n_strided; ++i) {
    const std::string g = "strided_temporal_block_" + std::to_string(i + 1);
    const int L = m->seq_lens[i], Lo = m->seq_lens[i + 1], st_i = s.strides[i], pl = s.pad_left[i];
    const long long Rl = (long long)B * L, Ro = (long long)B * Lo;
    BlkTape& tp = t->st[i];
    const BlkDims bd{d, h, L, H, (long long)B, 0};
    UU_TL(launch_residual(x_in, plain, nullptr, nullptr, 1, W(m, "strided_temporal_pe_" + std::to_string(i + 1), 0), L, Rl,
                          d, tp.x0, stream));
    if (attn_half_fwd(c, bd, g, tp, nullptr, 0)) return 1;
    UU_TL(launch_ln_fwd_gen(tp.x1, Rl, d, W(m, g, 10), W(m, g, 11), 1e-5f, tp.y2, stream));
    RowMap cm;
    cm.rpb = L; cm.batch_rows = Lo * st_i; cm.offset = pl; cm.step = 1;
    if (tf32_ok(c, tp.y2, d, W(m, g, 12), d, (int)Rl, h, d, t->hp[i], h)) {
      // Conv1D k=1 + ReLU straight into the zero-padded layout (pad rows stay zero), tcgen05 kind::tf32
      const float* Wt;
      if (weight_transposed(c, W(m, g, 12), d, h, &Wt)) return 1;
      Epilogue e;
      e.bias = W(m, g, 13); e.flags = EPI_RELU; e.cmap = cm;
      if (tf32_gemm(c, tp.y2, d, (int)Rl, d, Wt, d, h, e, t->hp[i], h)) return 1;
    } else {
      GemmGen gg;
      gg.A = tp.y2; gg.lda = d; gg.B = W(m, g, 12); gg.ldb = h; gg.C = t->hp[i]; gg.ldc = h; gg.cmap = cm;
      gg.M = (int)Rl; gg.N = h; gg.K = d; gg.bias = W(m, g, 13); gg.relu = 1;
      UU_TL(launch_gemm_gen(gg, stream));
    }
    if (lin_fwd(c, t->hp[i], (long long)st_i * h, (int)Ro, 3 * h, W(m, g, 14), d, W(m, g, 15), t->tmp1, d)) return 1;
    RowMap idm;
    idm.rpb = Lo; idm.batch_rows = L; idm.offset = (st_i > 1 && pl == 0) ? 1 : 0; idm.step = st_i;
    UU_TL(launch_residual(tp.x1, idm, t->tmp1, tp.keep < 1.f ? tp.scale2 : nullptr, Lo, nullptr, 1, Ro, d, tp.x2, stream));
    x_in = tp.x2;
  }
  if (lin_fwd(c, x_in, d, B, d, W(m, "strided_temporal_fc", 0), 3 * J, W(m, "strided_temporal_fc", 1), t->central, 3 * J))
    return 1;

  // ================= loss =================
  const float bs = (float)t->global_batch;
  UU_TL(launch_loss(t->full, t->central, gt3d, B, N, J, t->root, t->w_seq / (bs * N * J), t->w_center / (bs * J),
                    t->dfull, t->dcentral, t->partials, t->loss, stream));
  UU_CUDA(cudaMemcpyAsync(loss_out, t->loss, sizeof(float), cudaMemcpyDeviceToDevice, stream));

  // ================= backward =================
  // second passes of the two-pass reductions are recorded and run in one launch per flush (before each gradient bucket
  // leaves, and at the end)
  DeferGuard defer_guard;
  train_reduce_defer_begin(t->red_arena, RED_ARENA_FLOATS);
  float* dx = t->dx_s[s.n_strided - 1];
  if (lin_bwd(c, x_in, d, t->dcentral, 3 * J, B, d, 3 * J, W(m, "strided_temporal_fc", 0), dx, d, 0,
              G(m, "strided_temporal_fc", 0), G(m, "strided_temporal_fc", 1)))
    return 1;
  for (int i = s.n_strided - 1; i >= 0; --i) {
    const std::string g = "strided_temporal_block_" + std::to_string(i + 1);
    const int L = m->seq_lens[i], Lo = m->seq_lens[i + 1], st_i = s.strides[i], pl = s.pad_left[i];
    const long long Rl = (long long)B * L, Ro = (long long)B * Lo;
    BlkTape& tp = t->st[i];
    const BlkDims bd{d, h, L, H, (long long)B, 0};
    float* dx_prev = i > 0 ? t->dx_s[i - 1] : t->dx_t;        // gradient w.r.t. this block's input sequence
    // z path
    const float* dy = dx;
    if (tp.keep < 1.f) {
      UU_TL(launch_scale_rows(dx, tp.scale2, Lo, Ro, d, t->tmp1, stream));
      dy = t->tmp1;
    }
    UU_CUDA(cudaMemsetAsync(t->dhp[i], 0, sizeof(float) * (size_t)B * Lo * st_i * h, stream));
    if (lin_bwd(c, t->hp[i], (long long)st_i * h, dy, d, (int)Ro, 3 * h, d, W(m, g, 14), t->dhp[i],
                (long long)st_i * h, 0, G(m, g, 14), G(m, g, 15)))
      return 1;
    RowMap cm;
    cm.rpb = L; cm.batch_rows = Lo * st_i; cm.offset = pl; cm.step = 1;
    // d(pre-activation)[r] = relu'(hp[map r]) * dhp[map r]; rows the conv never reads have zero gradient
    UU_TL(launch_act_bwd_mapped(t->hp[i], t->dhp[i], cm, h, Rl, h, t->tmp_h, stream));
    if (lin_bwd(c, tp.y2, d, t->tmp_h, h, (int)Rl, d, h, W(m, g, 12), t->tmp2, d, 0, G(m, g, 12), G(m, g, 13))) return 1;
    // dx1 = scatter(dx over the identity rows) + LN2 backward
    UU_CUDA(cudaMemsetAsync(dx_prev, 0, sizeof(float) * (size_t)Rl * d, stream));
    RowMap idm;
    idm.rpb = Lo; idm.batch_rows = L; idm.offset = (st_i > 1 && pl == 0) ? 1 : 0; idm.step = st_i;
    UU_TL(launch_scatter_add(dx, idm, Ro, d, dx_prev, stream));
    UU_TL(launch_ln_bwd_gen(tp.x1, t->tmp2, Rl, d, W(m, g, 10), 1e-5f, dx_prev, 1, G(m, g, 10), G(m, g, 11), stream));
    if (attn_half_bwd(c, bd, g, tp, nullptr, 0, dx_prev)) return 1;
    UU_TL(launch_period_sum(dx_prev, Rl, L, d, nullptr, 0, G(m, "strided_temporal_pe_" + std::to_string(i + 1), 0), stream));
    dx = dx_prev;
  }
  // dx == dx_t: gradient w.r.t. the temporal output from the strided path; add the full-sequence head
  if (lin_bwd(c, xT, d, t->dfull, 3 * J, (int)R, d, 3 * J, W(m, "temporal_fc", 0), t->dx_t, d, 1, G(m, "temporal_fc", 0),
              G(m, "temporal_fc", 1)))
    return 1;
  const size_t off_temporal = tensor_offset(m, "temporal_block_1", 0), off_strided = tensor_offset(m, "strided_temporal_block_1", 0);
  UU_TL(train_reduce_flush(stream));
  if (overlap_comm && comm_allreduce_range(m, off_strided, m->n_alloc, 0, stream)) return 1;
  for (int i = s.temporal_depth - 1; i >= 0; --i) {
    const uint8_t* km = (use_mask && i < s.first_strided_token_attention_layer) ? mask : nullptr;
    if (block_bwd(c, tp_dims, "temporal_block_" + std::to_string(i + 1), t->tp[i], km, N, t->dx_t)) return 1;
  }
  UU_TL(train_reduce_flush(stream));
  if (overlap_comm && comm_allreduce_range(m, off_temporal, off_strided, 1, stream)) return 1;
  // temporal input: x = m*s4 + (1-m)*token + PE
  UU_TL(launch_period_sum(t->dx_t, R, N, d, nullptr, 0, G(m, "temporal_pe", 0), stream));
  if (use_mask) {
    UU_TL(launch_period_sum(t->dx_t, R, 1, d, mask, 0, G(m, "strided_input_token_layer", 0), stream));
  }
  // gradient of the compact spatial_to_temporal_fc output: rows of dx_t in gather-list order (x token-masking factors)
  UU_TL(launch_fill_bwd(t->dx_t, use_mask ? mask : nullptr, glist, tok_keep, Rv, d, t->tmp1, stream));
  if (lin_bwd(c, t->sp_normed, J * ds, t->tmp1, d, (int)Rv, J * ds, d, W(m, "spatial_to_temporal_fc", 0), t->dS, J * ds, 0,
              G(m, "spatial_to_temporal_fc", 0), G(m, "spatial_to_temporal_fc", 1)))
    return 1;
  UU_TL(launch_ln_bwd_gen(sp_out, t->dS, Rsv, ds, W(m, "spatial_norm", 0), 1e-6f, t->dx_sp, 0, G(m, "spatial_norm", 0),
                          G(m, "spatial_norm", 1), stream));
  for (int i = s.spatial_depth - 1; i >= 0; --i)
    if (block_bwd(c, sp_dims, "spatial_block_" + std::to_string(i + 1), t->sp[i], nullptr, 0, t->dx_sp)) return 1;
  UU_TL(launch_colsum(t->dx_sp, (int)Rsv, ds, ds, G(m, "keypoint_embedding", 1), stream));
  UU_TL(launch_period_sum(t->dx_sp, Rsv, J, ds, nullptr, 0, G(m, "spatial_pe", 0), stream));
  UU_TL(launch_embed_wgrad(x2d, use_mask ? mask : nullptr, glist, J, t->dx_sp, Rsv, ds, G(m, "keypoint_embedding", 0), stream));
  UU_TL(train_reduce_defer_end(stream));
  if (overlap_comm) {
    if (comm_allreduce_range(m, 0, off_temporal, 2, stream)) return 1;
    if (comm_allreduce_scalar(m, loss_out, stream)) return 1;
    if (comm_join(m, stream)) return 1;
  }
  return 0;
}

}  // namespace uu

using namespace uu;

extern "C" {

int uu_train_config(uu_model* m, int global_batch, int root_keypoint, float w_center, float w_sequence,
                    const float* drop_path_rate3, int droppath_mode, uint64_t seed) {
  UU_CHECK(m && global_batch > 0, "bad argument");
  UU_CHECK(root_keypoint >= 0 && root_keypoint < m->spec.n_joints, "ROOT_KEYTPOINT out of range");
  if (!m->train) m->train = new TrainState();
  TrainState* t = m->train;
  t->global_batch = global_batch; t->root = root_keypoint; t->w_center = w_center; t->w_seq = w_sequence;
  for (int i = 0; i < 3; ++i) t->dpr[i] = drop_path_rate3 ? drop_path_rate3[i] : 0.f;
  t->droppath_mode = droppath_mode; t->seed = seed;
  return 0;
}

int uu_train_forward_backward(uu_model* m, const float* x2d, const uint8_t* mask, const float* gt3d, int B,
                              int64_t step, float* loss_dev, void* stream) {
  UU_CHECK(m, "null model");
  if (m->train) { m->train->wt_valid.clear(); m->train->pk_valid.clear(); }     // weights may have changed since the last call
  return train_fb(m, x2d, mask, gt3d, B, step, loss_dev, (cudaStream_t)stream);
}

/* One data-parallel training step through the C ABI alone (train.py:464-506): forward + backward of the local windows,
 * the bucketed NCCL sum all-reduce of the gradients overlapped with the backward pass (after uu_comm_init; a plain local
 * step without a communicator), the all-reduced loss in loss_dev, and the fused AdamW / EMA update. */
int uu_train_step(uu_model* m, const float* x2d, const uint8_t* mask, const float* gt3d, int B, int64_t step, float lr_t,
                  float wd_t, float beta1, float beta2, float epsilon, float ema_decay, float* loss_dev, void* stream) {
  UU_CHECK(m, "null model");
  if (m->train) { m->train->wt_valid.clear(); m->train->pk_valid.clear(); }
  if (train_fb(m, x2d, mask, gt3d, B, step, loss_dev, (cudaStream_t)stream, true)) return 1;
  return uu_adamw_step(m, lr_t, wd_t, beta1, beta2, epsilon, step + 1, ema_decay, stream);
}

int uu_train_set_token_masking(uu_model* m, float rate) {
  UU_CHECK(m && rate >= 0.f && rate < 1.f, "TOKEN_MASK_RATE must be in [0, 1)");
  if (!m->train) m->train = new TrainState();
  m->train->token_mask_rate = rate;
  return 0;
}

int uu_get_token_mask(uu_model* m, float* host, int64_t capacity) {
  UU_CHECK(m && m->train && host && m->train->tok_keep, "no training step has run yet");
  const long long R = (long long)m->train->B * m->spec.n_tok;
  UU_CHECK(R <= capacity, "output buffer too small");
  UU_CUDA(cudaSetDevice(m->device));
  if (m->train->token_mask_rate > 0.f) {
    UU_CUDA(cudaMemcpy(host, m->train->tok_keep, sizeof(float) * R, cudaMemcpyDeviceToHost));
  } else {
    for (long long i = 0; i < R; ++i) host[i] = 1.f;
  }
  return 0;
}

int uu_train_set_math(uu_model* m, int mode) {
  UU_CHECK(m && (mode == 0 || mode == 1), "math mode: 0 = fp32, 1 = tf32 tensor cores");
  if (!m->train) m->train = new TrainState();
  m->train->math = mode;
  // attention of the temporal / strided blocks (attn_mma.cu): the fp32 parity mode takes the compensated TF32 form (fp32-grade:
  // Adam turns relative gradient errors into +-lr steps wherever a gradient is small), the tensor-core mode the bf16 hi + lo form
  m->train->attn_split = mode == 0 ? 3 : 2;
  return 0;
}

int uu_op_wgrad_tf32(const float* X, int64_t ldx, const float* dY, int64_t ldy, int64_t R, int Kd, int Nd, float* dW,
                     int accumulate, void* stream) {
  UU_CHECK(X && dY && dW, "null argument");
  UU_CHECK(wgrad_tc_ok(X, ldx, dY, ldy, R, Kd, Nd), "shape not supported: R >= 256, Kd % 32 == 0 (>= 128), Nd % 64 == 0, pitches % 4 == 0");
  int dev = 0, sms = 148;
  UU_CUDA(cudaGetDevice(&dev));
  UU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* scratch = nullptr;
  UU_CUDA(cudaMalloc(&scratch, wgrad_tc_scratch_bytes()));
  const int rc = wgrad_tc(X, ldx, dY, ldy, R, Kd, Nd, dW, accumulate, scratch, sms, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(scratch);
  return rc;
}

int uu_optimizer_state(uu_model* m, int which, int allocate, float** dev_ptr, int64_t* n_floats) {
  UU_CHECK(m && dev_ptr && n_floats && which >= 0 && which <= 2, "uu_optimizer_state: which is 0 (m), 1 (v) or 2 (EMA)");
  UU_CUDA(cudaSetDevice(m->device));
  if (allocate && which < 2 && !m->grads) {
    float* g; int64_t n;
    if (uu_grad_buffer(m, &g, &n)) return 1;            // creates the gradient and both moment buffers, zero-filled
  }
  if (allocate && which == 2 && !m->ema) {
    UU_CUDA(cudaMalloc(&m->ema, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMemcpy(m->ema, m->params, sizeof(float) * m->n_alloc, cudaMemcpyDeviceToDevice));
  }
  *dev_ptr = which == 0 ? m->adam_m : which == 1 ? m->adam_v : m->ema;
  *n_floats = (int64_t)m->n_alloc;
  return 0;
}

int uu_grad_buffer(uu_model* m, float** dev_ptr, int64_t* n_floats) {
  UU_CHECK(m && dev_ptr && n_floats, "null argument");
  UU_CUDA(cudaSetDevice(m->device));
  if (!m->grads) {
    UU_CUDA(cudaMalloc(&m->grads, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMalloc(&m->adam_m, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMalloc(&m->adam_v, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMemset(m->grads, 0, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMemset(m->adam_m, 0, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMemset(m->adam_v, 0, sizeof(float) * m->n_alloc));
  }
  *dev_ptr = m->grads;
  *n_floats = (int64_t)m->n_alloc;
  return 0;
}

int uu_get_grad(uu_model* m, const char* group, int index, float* host, int64_t capacity) {
  UU_CHECK(m && group && host && m->grads, "no gradients yet");
  auto it = m->lookup.find({std::string(group), index});
  UU_CHECK(it != m->lookup.end(), "no such weight");
  const TensorInfo& t = m->tensors[it->second];
  UU_CHECK((int64_t)t.numel <= capacity, "output buffer too small");
  UU_CUDA(cudaSetDevice(m->device));
  UU_CUDA(cudaMemcpy(host, m->grads + t.offset, sizeof(float) * t.numel, cudaMemcpyDeviceToHost));
  return 0;
}

int uu_get_droppath_scale(uu_model* m, int stage, int block, int branch, float* host, int64_t capacity, float* keep_prob) {
  UU_CHECK(m && m->train && host && keep_prob && (branch == 0 || branch == 1), "bad argument (branch: 0 attention, 1 MLP)");
  TrainState* t = m->train;
  UU_CHECK(stage >= 0 && stage < 3, "stage must be 0 (spatial), 1 (temporal) or 2 (strided)");
  std::vector<BlkTape>& tapes = stage == 0 ? t->sp : stage == 1 ? t->tp : t->st;
  UU_CHECK(block >= 0 && block < (int)tapes.size(), "block out of range");
  const long long ns = stage == 0 ? (long long)t->B * m->spec.n_tok : t->B;
  UU_CHECK(ns <= capacity, "output buffer too small");
  *keep_prob = tapes[block].keep;
  UU_CUDA(cudaSetDevice(m->device));
  if (tapes[block].keep < 1.f) {
    const BlkTape& bt = tapes[block];
    const float* src = branch ? (bt.scale2_full ? bt.scale2_full : bt.scale2) : (bt.scale_full ? bt.scale_full : bt.scale);
    UU_CUDA(cudaMemcpy(host, src, sizeof(float) * ns, cudaMemcpyDeviceToHost));
  } else {
    for (long long i = 0; i < ns; ++i) host[i] = 1.f;
  }
  return 0;
}

int uu_adamw_step(uu_model* m, float lr_t, float wd_t, float beta1, float beta2, float epsilon, int64_t t,
                  float ema_decay, void* stream) {
  UU_CHECK(m && m->grads, "no gradients: run uu_train_forward_backward first");
  UU_CHECK(t >= 1, "Adam step t starts at 1 (iterations + 1)");
  UU_CUDA(cudaSetDevice(m->device));
  if (ema_decay >= 0.f && !m->ema) {   // EMA clone starts as a copy of the weights (train.py:396-401)
    UU_CUDA(cudaMalloc(&m->ema, sizeof(float) * m->n_alloc));
    UU_CUDA(cudaMemcpyAsync(m->ema, m->params, sizeof(float) * m->n_alloc, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  const double alpha = (double)lr_t * std::sqrt(1.0 - std::pow((double)beta2, (double)t)) /
                       (1.0 - std::pow((double)beta1, (double)t));
  UU_CUDA(launch_adamw(m->params, m->adam_m, m->adam_v, m->grads, (long long)m->n_alloc, wd_t, (float)alpha, beta1, beta2,
                       epsilon, ema_decay >= 0.f ? m->ema : nullptr, ema_decay, (cudaStream_t)stream));
  m->dirty = true;   // fused / packed inference weights are stale now
  return 0;
}

int uu_get_ema_weight(uu_model* m, const char* group, int index, float* host, int64_t capacity) {
  UU_CHECK(m && group && host && m->ema, "EMA is not enabled");
  auto it = m->lookup.find({std::string(group), index});
  UU_CHECK(it != m->lookup.end(), "no such weight");
  const TensorInfo& t = m->tensors[it->second];
  UU_CHECK((int64_t)t.numel <= capacity, "output buffer too small");
  UU_CUDA(cudaSetDevice(m->device));
  UU_CUDA(cudaMemcpy(host, m->ema + t.offset, sizeof(float) * t.numel, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
