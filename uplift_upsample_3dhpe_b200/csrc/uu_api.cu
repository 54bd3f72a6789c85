// C ABI of the uu3d library (include/uu3d.h): model object, weight inventory, forward schedule.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "model.cuh"
#include "train.cuh"

namespace uu {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const std::string& get_error() { return g_error; }

}  // namespace uu

using namespace uu;

namespace uu {

const float* W(const uu_model* m, const std::string& g, int i) {
  auto it = m->lookup.find({g, i});
  return it == m->lookup.end() ? nullptr : m->params + m->tensors[it->second].offset;
}

size_t tensor_offset(const uu_model* m, const std::string& g, int i) {
  auto it = m->lookup.find({g, i});
  return it == m->lookup.end() ? (size_t)-1 : m->tensors[it->second].offset;
}

static void add_tensor(uu_model* m, const std::string& g, int idx, std::vector<int64_t> shape) {
  TensorInfo t;
  t.group = g; t.index = idx; t.shape = shape; t.offset = m->n_alloc;
  t.numel = 1;
  for (auto s : shape) t.numel *= (size_t)s;
  m->n_params += t.numel;
  m->n_alloc += (t.numel + 3) & ~(size_t)3;     // every tensor starts 16-byte aligned (vector loads of biases)
  m->lookup[{g, idx}] = (int)m->tensors.size();
  m->tensors.push_back(t);
}

static void add_block(uu_model* m, const std::string& g, int64_t d, int64_t h, bool strided) {
  int i = 0;
  add_tensor(m, g, i++, {d}); add_tensor(m, g, i++, {d});                 // norm1 gamma, beta
  for (int k = 0; k < 4; ++k) { add_tensor(m, g, i++, {d, d}); add_tensor(m, g, i++, {d}); }   // wq wk wv proj
  add_tensor(m, g, i++, {d}); add_tensor(m, g, i++, {d});                 // norm2
  if (strided) {
    add_tensor(m, g, i++, {1, d, h}); add_tensor(m, g, i++, {h});         // Conv1D k=1
    add_tensor(m, g, i++, {3, h, d}); add_tensor(m, g, i++, {d});         // strided Conv1D k=3
  } else {
    add_tensor(m, g, i++, {d, h}); add_tensor(m, g, i++, {h});
    add_tensor(m, g, i++, {h, d}); add_tensor(m, g, i++, {d});
  }
}

// Keras .h5 order (weight_io.py:155-198)
static void build_inventory(uu_model* m) {
  const uu_spec& s = m->spec;
  const int64_t J = s.n_joints, ds = s.d_spatial, dt = s.d_temporal;
  add_tensor(m, "keypoint_embedding", 0, {2, ds});
  add_tensor(m, "keypoint_embedding", 1, {ds});
  add_tensor(m, "spatial_pe", 0, {J, ds});
  add_tensor(m, "temporal_pe", 0, {s.n_tok, dt});
  for (int i = 0; i < s.n_strided; ++i)
    add_tensor(m, "strided_temporal_pe_" + std::to_string(i + 1), 0, {m->seq_lens[i], dt});
  if (s.has_strided_input) add_tensor(m, "strided_input_token_layer", 0, {dt});
  for (int i = 0; i < s.spatial_depth; ++i) add_block(m, "spatial_block_" + std::to_string(i + 1), ds, s.h_spatial, false);
  add_tensor(m, "spatial_norm", 0, {ds});
  add_tensor(m, "spatial_norm", 1, {ds});
  add_tensor(m, "spatial_to_temporal_fc", 0, {J * ds, dt});
  add_tensor(m, "spatial_to_temporal_fc", 1, {dt});
  for (int i = 0; i < s.temporal_depth; ++i) add_block(m, "temporal_block_" + std::to_string(i + 1), dt, s.h_temporal, false);
  for (int i = 0; i < s.n_strided; ++i) add_block(m, "strided_temporal_block_" + std::to_string(i + 1), dt, s.h_temporal, true);
  if (s.full_output) {
    add_tensor(m, "temporal_fc", 0, {dt, 3 * J});
    add_tensor(m, "temporal_fc", 1, {3 * J});
  }
  add_tensor(m, "strided_temporal_fc", 0, {dt, 3 * J});
  add_tensor(m, "strided_temporal_fc", 1, {3 * J});
}

// ---- small device helpers -----------------------------------------------------------------------
__global__ void k_concat_cols(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                              int rows, int d, float* __restrict__ out) {
  // out[r][0:d]=a[r], [d:2d]=b[r], [2d:3d]=c[r]
  long long n = (long long)rows * 3 * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / (3 * d)), col = (int)(i % (3 * d));
    const float* src = col < d ? a : (col < 2 * d ? b : c);
    out[i] = src[(long long)r * d + (col % d)];
  }
}
// Wt[n][k] = bf16(W[k][n]) for n < N, zero rows up to n_pad
__global__ void k_pack_wt(const float* __restrict__ Wm, int K, int N, int n_pad, bf16* __restrict__ Wt) {
  long long total = (long long)n_pad * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i / K), k = (int)(i % K);
    Wt[i] = __float2bfloat16_rn(n < N ? Wm[(long long)k * N + n] : 0.f);
  }
}

// LayerNorm folded into a GEMM: Wt[n][k] = bf16(gamma[k] W[k][n]); csum[n] = sum_k Wt[n][k] (of the rounded values, so
// that rstd (x Wt - mean csum) is exactly the GEMM of the centred row); bias_out[n] = bias[n] + sum_k beta[k] W[k][n].
// One warp per output column n.
__global__ void k_pack_wt_ln(const float* __restrict__ Wm, int K, int N, int n_pad, const float* __restrict__ gamma,
                             const float* __restrict__ beta, const float* __restrict__ bias, bf16* __restrict__ Wt,
                             float* __restrict__ csum, float* __restrict__ bias_out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= n_pad) return;
  float cs = 0.f, bb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = n < N ? Wm[(long long)k * N + n] : 0.f;
    const bf16 q = __float2bfloat16_rn(w * gamma[k]);
    Wt[(long long)n * K + k] = q;
    cs += __bfloat162float(q);
    bb += beta[k] * w;
  }
  for (int o = 16; o; o >>= 1) {
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
  }
  if (lane == 0) {
    csum[n] = cs;
    bias_out[n] = (n < N ? bias[n] : 0.f) + bb;
  }
}

// Wt[n][k] = W[k][n] in fp32: K-major B operand of the kind::tf32 GEMM
__global__ void k_pack_wt32(const float* __restrict__ Wm, int K, int N, float* __restrict__ Wt) {
  long long total = (long long)N * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i / K), k = (int)(i % K);
    Wt[i] = Wm[(long long)k * N + n];
  }
}

int dev_alloc(std::vector<void*>& pool, void** p, size_t bytes, bool zero) {
  UU_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  pool.push_back(*p);
  if (zero) UU_CUDA(cudaMemset(*p, 0, bytes ? bytes : 16));
  return 0;
}

static int make_pack(uu_model* m, Pack& pk, const float* Wsrc, int K, int N, cudaStream_t st) {
  pk.k = K; pk.n = N; pk.n_pad = (N + 63) / 64 * 64;
  if (!pk.ptr) {
    void* p;
    if (dev_alloc(m->derived_allocs, &p, sizeof(bf16) * (size_t)pk.n_pad * K, false)) return 1;
    pk.ptr = (bf16*)p;
  }
  k_pack_wt<<<256, 256, 0, st>>>(Wsrc, K, N, pk.n_pad, pk.ptr);
  UU_CUDA(cudaGetLastError());
  return 0;
}

static int make_pack_ln(uu_model* m, Pack& pk, float*& csum, float*& bias_out, const float* Wsrc, int K, int N,
                        const float* gamma, const float* beta, const float* bias, cudaStream_t st) {
  pk.k = K; pk.n = N; pk.n_pad = (N + 63) / 64 * 64;
  if (!pk.ptr) {
    void* p;
    if (dev_alloc(m->derived_allocs, &p, sizeof(bf16) * (size_t)pk.n_pad * K, false)) return 1;
    pk.ptr = (bf16*)p;
    if (dev_alloc(m->derived_allocs, &p, sizeof(float) * pk.n_pad, false)) return 1;
    csum = (float*)p;
    if (dev_alloc(m->derived_allocs, &p, sizeof(float) * pk.n_pad, false)) return 1;
    bias_out = (float*)p;
  }
  k_pack_wt_ln<<<(pk.n_pad + 7) / 8, 256, 0, st>>>(Wsrc, K, N, pk.n_pad, gamma, beta, bias, pk.ptr, csum, bias_out);
  UU_CUDA(cudaGetLastError());
  return 0;
}

static int make_pack32(uu_model* m, float*& dst, const float* Wsrc, int K, int N, cudaStream_t st) {
  if (!dst) {
    void* p;
    if (dev_alloc(m->derived_allocs, &p, sizeof(float) * (size_t)N * K, false)) return 1;
    dst = (float*)p;
  }
  k_pack_wt32<<<256, 256, 0, st>>>(Wsrc, K, N, dst);
  UU_CUDA(cudaGetLastError());
  return 0;
}

static int setup_block(uu_model* m, BlockW& b, const std::string& g, int d, int h, bool strided, cudaStream_t st) {
  b.ln1_g = W(m, g, 0); b.ln1_b = W(m, g, 1);
  b.wp = W(m, g, 8); b.bp = W(m, g, 9);
  b.ln2_g = W(m, g, 10); b.ln2_b = W(m, g, 11);
  b.w1 = W(m, g, 12); b.b1 = W(m, g, 13); b.w2 = W(m, g, 14); b.b2 = W(m, g, 15);
  if (!b.wqkv) {
    void* p;
    if (dev_alloc(m->derived_allocs, &p, sizeof(float) * (size_t)d * 3 * d, false)) return 1;
    b.wqkv = (float*)p;
    if (dev_alloc(m->derived_allocs, &p, sizeof(float) * 3 * d, false)) return 1;
    b.bqkv = (float*)p;
  }
  k_concat_cols<<<256, 256, 0, st>>>(W(m, g, 2), W(m, g, 4), W(m, g, 6), d, d, b.wqkv);
  k_concat_cols<<<1, 256, 0, st>>>(W(m, g, 3), W(m, g, 5), W(m, g, 7), 1, d, b.bqkv);
  UU_CUDA(cudaGetLastError());
  if (make_pack(m, b.p_qkv, b.wqkv, d, 3 * d, st)) return 1;
  if (make_pack(m, b.p_proj, b.wp, d, d, st)) return 1;
  if (make_pack(m, b.p_fc1, b.w1, d, h, st)) return 1;
  if (make_pack(m, b.p_fc2, b.w2, strided ? 3 * h : h, d, st)) return 1;
  // (strided blocks: fc1 is the k=1 Conv1D, a (d, h) matrix like the dense fc1)
  if (make_pack_ln(m, b.p_qkv_ln, b.cs_qkv, b.bl_qkv, b.wqkv, d, 3 * d, b.ln1_g, b.ln1_b, b.bqkv, st)) return 1;
  if (make_pack_ln(m, b.p_fc1_ln, b.cs_fc1, b.bl_fc1, b.w1, d, h, b.ln2_g, b.ln2_b, b.b1, st)) return 1;
  if (m->precision == UU_PRECISION_TF32) {
    if (make_pack32(m, b.t_qkv, b.wqkv, d, 3 * d, st) || make_pack32(m, b.t_proj, b.wp, d, d, st) ||
        make_pack32(m, b.t_fc1, b.w1, d, h, st) || make_pack32(m, b.t_fc2, b.w2, strided ? 3 * h : h, d, st))
      return 1;
  }
  return 0;
}

// (Re)derive fused / packed weights after uu_set_weight / uu_adamw_step.  The repack kernels run on the CALLER's stream
// `st`: they read m->params, which the AdamW update of the same stream may just have written (a repack on the legacy
// stream would not be ordered behind an update enqueued on a non-blocking stream).  The host then waits for that stream,
// so a later forward on any other stream sees finished packs.
static int commit_weights(uu_model* m, cudaStream_t st) {
  if (!m->dirty) return 0;
  const uu_spec& s = m->spec;
  UU_CUDA(cudaSetDevice(m->device));
  if (!m->spatial_ptrs) {
    void* p;
    if (dev_alloc(m->derived_allocs, &p, sizeof(float*) * 16 * s.spatial_depth, false)) return 1;
    m->spatial_ptrs = (const float**)p;
    std::vector<const float*> h(16 * s.spatial_depth);
    for (int l = 0; l < s.spatial_depth; ++l)
      for (int i = 0; i < 16; ++i) h[l * 16 + i] = W(m, "spatial_block_" + std::to_string(l + 1), i);
    UU_CUDA(cudaMemcpy(p, h.data(), sizeof(float*) * h.size(), cudaMemcpyHostToDevice));
  }
  if (!m->sp_frags) {
    void* p;
    if (dev_alloc(m->derived_allocs, &m->sp_frags, spatial_tc_frag_bytes(s.spatial_depth), false)) return 1;
    if (dev_alloc(m->derived_allocs, &p, spatial_tc_param_bytes(s.spatial_depth), false)) return 1;
    m->sp_params = (float*)p;
  }
  UU_CUDA(launch_spatial_pack(m->spatial_ptrs, s.spatial_depth, W(m, "keypoint_embedding", 0),
                              W(m, "keypoint_embedding", 1), W(m, "spatial_pe", 0), W(m, "spatial_norm", 0),
                              W(m, "spatial_norm", 1), m->sp_frags, m->sp_params, st));
  m->tblocks.resize(s.temporal_depth);
  m->sblocks.resize(s.n_strided);
  for (int i = 0; i < s.temporal_depth; ++i)
    if (setup_block(m, m->tblocks[i], "temporal_block_" + std::to_string(i + 1), s.d_temporal, s.h_temporal, false, st)) return 1;
  for (int i = 0; i < s.n_strided; ++i)
    if (setup_block(m, m->sblocks[i], "strided_temporal_block_" + std::to_string(i + 1), s.d_temporal, s.h_temporal, true, st)) return 1;
  if (make_pack(m, m->p_s2t, W(m, "spatial_to_temporal_fc", 0), s.n_joints * s.d_spatial, s.d_temporal, st)) return 1;
  if (m->precision == UU_PRECISION_TF32 &&
      make_pack32(m, m->t_s2t, W(m, "spatial_to_temporal_fc", 0), s.n_joints * s.d_spatial, s.d_temporal, st))
    return 1;
  if (s.full_output && make_pack(m, m->p_head1, W(m, "temporal_fc", 0), s.d_temporal, 3 * s.n_joints, st)) return 1;
  if (make_pack(m, m->p_head2, W(m, "strided_temporal_fc", 0), s.d_temporal, 3 * s.n_joints, st)) return 1;
  UU_CUDA(cudaStreamSynchronize(st));
  m->dirty = false;
  return 0;
}

void free_pool(std::vector<void*>& pool) {
  for (void* p : pool) cudaFree(p);
  pool.clear();
}

static void drop_graphs(uu_model* m) {
  for (auto& kv : m->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  m->graphs.clear();
}

static void drop_plans(uu_model* m) {
  // (captured graphs stay valid: tensor maps are kernel parameters, copied into the graph nodes; what they point
  // at — workspace and weight packs — is only released in ensure_workspace / uu_destroy, which drop the cache)
  for (auto* p : m->plans) tc_gemm_plan_destroy(p);
  m->plans.clear();
  for (auto* p : m->mlp_plans) mlp_plan_destroy(p);
  m->mlp_plans.clear();
  m->plan_B = -1;
}

static int ensure_workspace(uu_model* m, int B) {
  if (B <= m->cap_B && m->ws_precision == m->precision) return 0;
  const uu_spec& s = m->spec;
  UU_CUDA(cudaDeviceSynchronize());
  drop_graphs(m);
  free_pool(m->ws_allocs);
  drop_plans(m);
  m->Xs.clear(); m->Hp.clear();
  m->Y = nullptr; m->P = nullptr; m->ln_stats = nullptr;
  const int cap = std::max(B, m->cap_B);
  const size_t R = (size_t)cap * s.n_tok;
  const size_t es = m->precision == UU_PRECISION_BF16 ? 2 : 4;
  const size_t dt = s.d_temporal, ht = s.h_temporal;
  void* p;
  if (dev_alloc(m->ws_allocs, &p, sizeof(int) * (cap + 1), true)) return 1; m->g_scratch = (int*)p;
  if (dev_alloc(m->ws_allocs, &p, sizeof(int) * R, true)) return 1; m->g_list = (int*)p;
  if (dev_alloc(m->ws_allocs, &p, sizeof(int) * 4, true)) return 1; m->g_count = (int*)p;
  if (dev_alloc(m->ws_allocs, &p, sizeof(int) * R, true)) return 1; m->w_src = (int*)p;
  if (dev_alloc(m->ws_allocs, &p, R, true)) return 1; m->w_mask = (uint8_t*)p;
  if (dev_alloc(m->ws_allocs, &m->S, es * R * s.n_joints * s.d_spatial, true)) return 1;
  if (dev_alloc(m->ws_allocs, &p, 4 * R * dt, true)) return 1; m->X = (float*)p;
  if (m->precision != UU_PRECISION_BF16 && dev_alloc(m->ws_allocs, &m->Y, es * R * dt, true)) return 1;   // LN output (fp32 schedule)
  if (dev_alloc(m->ws_allocs, &m->QKV, es * R * 3 * dt, true)) return 1;
  if (dev_alloc(m->ws_allocs, &m->O, es * R * dt, true)) return 1;
  if (dev_alloc(m->ws_allocs, &m->Hd, es * R * ht, true)) return 1;
  if (m->precision == UU_PRECISION_BF16 && dev_alloc(m->ws_allocs, &m->P, es * R * dt, true)) return 1;
  if (m->precision == UU_PRECISION_BF16) {
    if (dev_alloc(m->ws_allocs, &p, sizeof(float) * 2 * R * (dt / 64), true)) return 1;
    m->ln_stats = (float*)p;
  }
  for (int i = 0; i < s.n_strided; ++i) {
    const size_t Lo = m->seq_lens[i + 1];
    if (dev_alloc(m->ws_allocs, &p, 4 * (size_t)cap * Lo * dt, true)) return 1;
    m->Xs.push_back((float*)p);
    if (dev_alloc(m->ws_allocs, &p, es * (size_t)cap * Lo * s.strides[i] * ht, true)) return 1;   // pad rows stay zero
    m->Hp.push_back(p);
  }
  m->cap_B = cap;
  m->ws_precision = m->precision;
  return 0;
}

// ---- forward schedule ---------------------------------------------------------------------------
struct Fwd {
  uu_model* m;
  cudaStream_t st;
  int B;
  bool tc;             // tcgen05 GEMMs (bf16 activations) or CUDA-core fp32
  bool building;       // first run at this batch size: create TMA plans
  size_t plan_i = 0;
  int launches = 0;
  const float* wt32 = nullptr;   // tf32 schedule: fp32 W^T of the NEXT gemm() call (consumed by it)
};

static cudaEvent_t prof_event(uu_model* m) {
  if (m->ev_next == m->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    m->ev_pool.push_back(e);
  }
  return m->ev_pool[m->ev_next++];
}
// Launch wrapper: counts launches and, in profiling mode, brackets the launch with events on its stream.
#define UU_LAUNCH(f, kind, n, expr)                                                   \
  do {                                                                                \
    cudaEvent_t _e0 = nullptr, _e1 = nullptr;                                         \
    if ((f).m->profiling) {                                                           \
      _e0 = prof_event((f).m); _e1 = prof_event((f).m);                               \
      cudaEventRecord(_e0, (f).st);                                                   \
    }                                                                                 \
    UU_CUDA(expr);                                                                    \
    if ((f).m->profiling) {                                                           \
      cudaEventRecord(_e1, (f).st);                                                   \
      (f).m->ev_used.push_back({(kind), {_e0, _e1}});                                 \
    }                                                                                 \
    (f).launches += (n);                                                              \
  } while (0)

// One GEMM call site. A: activations in the workspace dtype; Wf: fp32 (K, N); pk: bf16 W^T.
static int gemm(Fwd& f, const void* A, long long lda, int M, int K, const float* Wf, const Pack& pk, int N,
                const Epilogue& epi, void* C, int c_bf16, long long ldc) {
  if (M == 0) return 0;
  if (f.tc) {
    if (f.building) {
      TcGemmPlan* p = nullptr;
      if (tc_gemm_plan_create(&p, (const bf16*)A, lda, M, K, pk.ptr, pk.n_pad, N)) return 1;
      f.m->plans.push_back(p);
    }
    UU_CHECK(f.plan_i < f.m->plans.size(), "internal: GEMM plan list out of sync");
    UU_LAUNCH(f, UU_KIND_GEMM_TC, 1, tc_gemm_launch(f.m->plans[f.plan_i++], epi, C, c_bf16, ldc, f.st));
  } else if (f.wt32 && M >= 64 && N % 64 == 0 && K % 4 == 0 && lda % 4 == 0 && ldc % 4 == 0 && !c_bf16) {
    // tf32 schedule: fp32 activations, W^T fp32, tcgen05 kind::tf32 (the plan is a pair of tensor maps: built per call)
    TcGemmPlan* p = nullptr;
    if (tc_gemm_plan_create_tf32(&p, (const float*)A, lda, M, K, f.wt32, K, N)) return 1;
    cudaError_t err = cudaSuccess;
    UU_LAUNCH(f, UU_KIND_GEMM_TC, 1, (err = tc_gemm_launch(p, epi, C, 0, ldc, f.st)));
    tc_gemm_plan_destroy(p);
    f.wt32 = nullptr;
  } else {
    UU_LAUNCH(f, UU_KIND_GEMM_F32, 1, launch_gemm_simt(A, 0, lda, Wf, M, N, K, epi, C, c_bf16, ldc, f.st));
    f.wt32 = nullptr;
  }
  return 0;
}

static int attention_block(Fwd& f, const BlockW& w, float* x, int L, const uint8_t* keymask) {
  // x += Proj(MHA(LN1(x)))   (vit:183-188 / net:129-133); LN1 is done by the caller (it may fuse the PE add)
  uu_model* m = f.m;
  const uu_spec& s = m->spec;
  const int d = s.d_temporal, R = f.B * L, bf = f.tc ? 1 : 0;
  const bool tf = m->precision == UU_PRECISION_TF32;
  Epilogue e;
  e.bias = w.bqkv;
  f.wt32 = tf ? w.t_qkv : nullptr;
  if (gemm(f, m->Y, d, R, d, w.wqkv, w.p_qkv, 3 * d, e, m->QKV, bf, 3 * d)) return 1;
  if (tf && attention_mma_ok(f.B, L, s.num_heads, d / s.num_heads)) {
    // tf32 schedule: fp32 rows, attention on mma.sync with bf16 hi + lo operand planes (attn_mma.cu; 2^-16 per product, finer
    // than the TF32 GEMMs around it) instead of the CUDA-core kernel of the fp32 schedule
    UU_LAUNCH(f, UU_KIND_ATTENTION, 1,
              launch_attention_mma_fwd((const float*)m->QKV, f.B, L, s.num_heads, d / s.num_heads, keymask, s.n_tok, (float*)m->O,
                                       2, f.st));
  } else {
    UU_LAUNCH(f, UU_KIND_ATTENTION, 1,
              launch_attention(m->QKV, bf, f.B, L, s.num_heads, d / s.num_heads, keymask, s.n_tok, m->O, f.st));
  }
  Epilogue ep;
  ep.bias = w.bp; ep.flags = EPI_RESIDUAL; ep.res = x; ep.ldr = d;
  f.wt32 = tf ? w.t_proj : nullptr;
  if (gemm(f, m->O, d, R, d, w.wp, w.p_proj, d, ep, x, 0, d)) return 1;
  return 0;
}

static int run_forward_impl(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central,
                            cudaStream_t st);
static int run_forward(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central,
                       cudaStream_t st) {
  m->ev_next = 0;
  m->ev_used.clear();
  const int rc = run_forward_impl(m, x2d, mask, B, full, central, st);
  if (rc) drop_plans(m);   // never keep a half-built plan list
  return rc;
}
// ---- bf16 schedule --------------------------------------------------------------------------------
// Every dense contraction is a tcgen05 GEMM with bf16 operands and a bf16 result; the residual stream is
// bf16-resident and touched only by the streaming residual+LayerNorm kernel (one read, one write per half
// block; adds, statistics and the normalisation are fp32 in registers).
static int tc_gemm_bf16(Fwd& f, const void* A, long long lda, int M, int K, const Pack& pk, int N, const float* bias,
                        bool relu, void* C, long long ldc, const RowMap* cmap = nullptr) {
  Epilogue e;
  e.bias = bias; e.flags = relu ? EPI_RELU : 0;
  if (cmap) e.cmap = *cmap;
  return gemm(f, A, lda, M, K, nullptr, pk, N, e, C, 1, ldc);
}

static int run_forward_bf16(Fwd& f, const float* x2d, const uint8_t* mask, float* full, float* central) {
  uu_model* m = f.m;
  const uu_spec& s = m->spec;
  cudaStream_t st = f.st;
  const int B = f.B, N = s.n_tok, J = s.n_joints, ds = s.d_spatial, d = s.d_temporal, h = s.h_temporal;
  const int R = B * N, H = s.num_heads, dh = d / H;
  const bool use_mask = s.has_strided_input != 0;
  const bool want_full = s.full_output && full;
  bf16 *QKV = (bf16*)m->QKV, *O = (bf16*)m->O, *Hd = (bf16*)m->Hd, *P = (bf16*)m->P;
  bf16* X = reinterpret_cast<bf16*>(m->X);      // residual stream, bf16-resident in this schedule (buffer sized for fp32)
  const RowMap plain;

  if (use_mask) UU_LAUNCH(f, UU_KIND_GATHER, 3, launch_build_gather(mask, B, N, m->g_scratch, m->g_list, m->g_count, st));
  if (m->n_chunks > 1) {
    // chunked input (uu_forward_host): window chunk c = list positions [scratch[c*Bc], scratch[(c+1)*Bc]) (device-side
    // prefix counts) or, without a mask, frames [c*Bc*N, (c+1)*Bc*N); each launch waits for its chunk's copy only
    for (int c = 0; c < m->n_chunks; ++c) {
      const int w0 = c * m->chunk_windows, w1 = std::min(B, (c + 1) * m->chunk_windows);
      UU_CUDA(cudaStreamWaitEvent(st, m->chunk_ev[c], 0));
      UU_LAUNCH(f, UU_KIND_SPATIAL, 1,
                launch_spatial_tc(x2d, use_mask ? m->g_list : nullptr, use_mask ? m->g_count : nullptr, (w1 - w0) * N,
                                  s.spatial_depth, m->sp_frags, m->sp_params, (bf16*)m->S, m->num_sms, st, m->cur_src,
                                  use_mask ? m->g_scratch + w0 : nullptr, use_mask ? m->g_scratch + w1 : nullptr,
                                  w0 * N, w1 * N, m->cur_flip));
    }
  } else {
    UU_LAUNCH(f, UU_KIND_SPATIAL, 1,
              launch_spatial_tc(x2d, use_mask ? m->g_list : nullptr, use_mask ? m->g_count : nullptr, R, s.spatial_depth,
                                m->sp_frags, m->sp_params, (bf16*)m->S, m->num_sms, st, m->cur_src, nullptr, nullptr, 0, -1,
                                m->cur_flip));
  }
  // Neither LayerNorm nor the residual / positional adds of the temporal and strided blocks are kernels of their own:
  //  * GEMMs that consume LN(x) (QKV, fc1) read the bf16 residual stream directly with gamma folded into their weights
  //    and finish the normalisation in the epilogue from per-row statistics (EPI_LNFOLD);
  //  * GEMMs that produce a residual update (projection, fc2) add the stream in their epilogue, write it in place and
  //    emit the statistics of the rows they wrote (EPI_RESID_BF16); the 544->384 GEMM does the same with the temporal
  //    positional table (EPI_ROWTABLE).
  const int slots = d / 64;
  UU_CHECK(d % 64 == 0 && slots <= 32, "temporal width must be a multiple of 64 (<= 2048)");
  UU_CHECK(s.temporal_depth > 0 && s.n_strided > 0, "the bf16 schedule needs at least one temporal and one strided block");
  {  // S4 + T1: 544->384 GEMM (+ bias + temporal PE, net:352) scattered to the token rows through the TMA-store epilogue;
     // upsampling token + PE on the rows without 2-D input (net:350-352)
    Epilogue e;
    e.bias = W(m, "spatial_to_temporal_fc", 1);
    e.flags = EPI_ROWTABLE; e.table = W(m, "temporal_pe", 0); e.table_period = N;
    e.stats_out = m->ln_stats; e.ln_slots = slots;
    if (use_mask) { e.c_rowidx = m->g_list; e.m_dev = m->g_count; }
    if (gemm(f, m->S, J * ds, R, J * ds, nullptr, m->p_s2t, d, e, X, 1, d)) return 1;
    if (use_mask)
      UU_LAUNCH(f, UU_KIND_TOKEN_FILL, 1,
                launch_token_fill_bx(mask, R, N, d, W(m, "strided_input_token_layer", 0), W(m, "temporal_pe", 0), X, st,
                                     m->ln_stats, slots));
  }
  auto ln_gemm = [&](bf16* x, int rows, const Pack& pk, int n_out, const float* bias_ln, const float* csum, bool relu,
                     void* out, const RowMap* cmap) {
    Epilogue e;
    e.bias = bias_ln; e.flags = EPI_LNFOLD | (relu ? EPI_RELU : 0);
    e.ln_stats = m->ln_stats; e.ln_csum = csum; e.ln_slots = slots; e.ln_inv_k = 1.f / d; e.ln_eps = 1e-5f;
    if (cmap) e.cmap = *cmap;
    return gemm(f, x, d, rows, d, nullptr, pk, n_out, e, out, 1, n_out);
  };
  auto resid_gemm = [&](const bf16* A, int K, bf16* x, int rows, const Pack& pk, const float* bias) {
    Epilogue e;
    e.bias = bias; e.flags = EPI_RESID_BF16;
    e.res_bf16 = x; e.stats_out = m->ln_stats; e.ln_slots = slots;
    return gemm(f, A, K, rows, K, nullptr, pk, d, e, x, 1, d);
  };
  // UU_ATTN5=1: the tcgen05 / TMEM / TMA attention kernel (attn_tc5.cu) instead of the mma.sync one.  Parity-identical;
  // measured 265 vs 191 us per 4096-window launch (a 71-token window fills 55 % of the M = 128 tile
  // and the softmax lands on two of the four SM sub-partitions, while the mma.sync kernel already runs at 89 % of the
  // HBM roofline), so the default stays the faster kernel.
  static const bool attn5 = [] { const char* e = getenv("UU_ATTN5"); return e && e[0] == '1'; }();
  static const bool fused_mlp_env = [] { const char* e = getenv("UU_FUSED_MLP"); return !e || e[0] != '0'; }();
  const bool fused_mlp = fused_mlp_env && d == 384 && h % 128 == 0 && h >= 128 && h <= 768 && R >= 512;
  size_t mlp_i = 0;
  for (int i = 0; i < s.temporal_depth; ++i) {   // T2 / T3
    const BlockW& w = m->tblocks[i];
    const uint8_t* km = (use_mask && i < s.first_strided_token_attention_layer) ? mask : nullptr;
    if (ln_gemm(X, R, w.p_qkv_ln, 3 * d, w.bl_qkv, w.cs_qkv, false, QKV, nullptr)) return 1;
    if (attn5 && attention_tc5_ok(N, H, dh))
      UU_LAUNCH(f, UU_KIND_ATTENTION, 1, launch_attention_tc5(QKV, B, N, km, N, O, m->num_sms, st));
    else
      UU_LAUNCH(f, UU_KIND_ATTENTION, 1, launch_attention_tc(QKV, B, N, H, dh, km, N, O, st));
    if (resid_gemm(O, d, X, R, w.p_proj, w.bp)) return 1;
    if (fused_mlp) {   // fc1 -> ReLU -> fc2 + residual in one kernel: the hidden activation never leaves the SM (mlp_tc.cuh)
      if (f.building) {
        MlpPlan* p = nullptr;
        if (mlp_plan_create(&p, X, d, R, d, h, w.p_fc1_ln.ptr, w.p_fc2.ptr)) return 1;
        m->mlp_plans.push_back(p);
      }
      UU_CHECK(mlp_i < m->mlp_plans.size(), "internal: MLP plan list out of sync");
      MlpArgs a;
      a.M = R; a.n_chunks = h / 64; a.ln_stats = m->ln_stats; a.ln_slots = slots; a.ln_inv_k = 1.f / d; a.ln_eps = 1e-5f;
      a.csum1 = w.cs_fc1; a.bias1 = w.bl_fc1;
      a.epi2.bias = w.b2; a.epi2.flags = EPI_RESID_BF16; a.epi2.res_bf16 = X; a.epi2.stats_out = m->ln_stats;
      a.epi2.ln_slots = slots;
      a.X = X; a.ldx = d;
      UU_LAUNCH(f, UU_KIND_GEMM_TC, 1, mlp_launch(m->mlp_plans[mlp_i++], a, st));
    } else {
      if (ln_gemm(X, R, w.p_fc1_ln, h, w.bl_fc1, w.cs_fc1, true, Hd, nullptr)) return 1;
      if (resid_gemm(Hd, h, X, R, w.p_fc2, w.b2)) return 1;
    }
  }
  if (want_full) {   // T4: full-sequence head on the temporal output
    Epilogue e;
    e.bias = W(m, "temporal_fc", 1);
    if (gemm(f, X, d, R, d, nullptr, m->p_head1, 3 * J, e, full, 0, 3 * J)) return 1;
  }
  // + PE of strided block 1 (net:128) and the statistics its QKV GEMM needs
  UU_LAUNCH(f, UU_KIND_LAYERNORM, 1,
            launch_residual_ln_bx(X, plain, nullptr, X, R, d, nullptr, nullptr, 0.f, W(m, "strided_temporal_pe_1", 0), N,
                                  nullptr, nullptr, st, m->ln_stats, slots));
  bf16* x_in = X;
  for (int i = 0; i < s.n_strided; ++i) {   // Q1/Q2
    const BlockW& w = m->sblocks[i];
    const int L = m->seq_lens[i], Lo = m->seq_lens[i + 1], st_i = s.strides[i], pl = s.pad_left[i];
    const int Rl = B * L, Ro = B * Lo;
    if (ln_gemm(x_in, Rl, w.p_qkv_ln, 3 * d, w.bl_qkv, w.cs_qkv, false, QKV, nullptr)) return 1;
    if (attn5 && attention_tc5_ok(L, H, dh))
      UU_LAUNCH(f, UU_KIND_ATTENTION, 1, launch_attention_tc5(QKV, B, L, nullptr, L, O, m->num_sms, st));
    else
      UU_LAUNCH(f, UU_KIND_ATTENTION, 1, launch_attention_tc(QKV, B, L, H, dh, nullptr, L, O, st));
    if (resid_gemm(O, d, x_in, Rl, w.p_proj, w.bp)) return 1;
    RowMap cm;   // Conv1D k=1 + ReLU into the zero-padded layout [B, Lo*s, h]
    cm.rpb = L; cm.batch_rows = Lo * st_i; cm.offset = pl; cm.step = 1;
    if (ln_gemm(x_in, Rl, w.p_fc1_ln, h, w.bl_fc1, w.cs_fc1, true, m->Hp[i], &cm)) return 1;
    // strided Conv1D k=3 as an implicit GEMM over contiguous 3h-wide rows, s*h apart
    if (tc_gemm_bf16(f, m->Hp[i], (long long)st_i * h, Ro, 3 * h, w.p_fc2, d, w.b2, false, P, d)) return 1;
    RowMap idm;  // identity path x[b, c0 + t*s]
    idm.rpb = Lo; idm.batch_rows = L; idm.offset = (st_i > 1 && pl == 0) ? 1 : 0; idm.step = st_i;
    bf16* x_out = reinterpret_cast<bf16*>(m->Xs[i]);
    if (i + 1 < s.n_strided) {   // x_out = x_in[identity rows] + conv + PE of the next block; statistics for its QKV GEMM
      UU_LAUNCH(f, UU_KIND_LAYERNORM, 1,
                launch_residual_ln_bx(x_in, idm, P, x_out, Ro, d, nullptr, nullptr, 0.f,
                                      W(m, "strided_temporal_pe_" + std::to_string(i + 2), 0), Lo, nullptr, nullptr, st,
                                      m->ln_stats, slots));
    } else {
      UU_LAUNCH(f, UU_KIND_LAYERNORM, 1,
                launch_residual_ln_bx(x_in, idm, P, x_out, Ro, d, nullptr, nullptr, 0.f, nullptr, 1, nullptr, O, st));
    }
    x_in = x_out;
  }
  {   // Q3
    Epilogue e;
    e.bias = W(m, "strided_temporal_fc", 1);
    if (gemm(f, O, d, B, d, nullptr, m->p_head2, 3 * J, e, central, 0, 3 * J)) return 1;
  }
  return 0;
}

static int run_forward_impl(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central,
                            cudaStream_t st) {
  const uu_spec& s = m->spec;
  UU_CHECK(B > 0, "batch must be positive");
  UU_CHECK(x2d && central, "x2d and central must not be null");
  UU_CHECK(!s.has_strided_input || mask, "this model has strided input: a stride mask is required");
  UU_CUDA(cudaSetDevice(m->device));
  if (commit_weights(m, st)) return 1;
  if (ensure_workspace(m, B)) return 1;
  Fwd f;
  f.m = m; f.st = st; f.B = B; f.tc = m->precision == UU_PRECISION_BF16;
  const int want_full = (s.full_output && full) ? 1 : 0;
  if (f.tc && (m->plan_B != B || m->plan_full != want_full)) drop_plans(m);
  f.building = f.tc && m->plans.empty();
  const int bf = f.tc ? 1 : 0;
  const int N = s.n_tok, J = s.n_joints, ds = s.d_spatial, d = s.d_temporal, h = s.h_temporal;
  const int R = B * N;
  const bool use_mask = s.has_strided_input != 0;
  if (f.tc) {
    static const int pdl_max_b = [] { const char* e = getenv("UU_PDL_MAX_B"); return e ? atoi(e) : 1024; }();
    pdl_set_auto(B <= pdl_max_b);     // programmatic dependent launch pays for small batches only (gemm_tc.cu)
    const int rc = run_forward_bf16(f, x2d, mask, full, central);
    pdl_set_auto(false);
    if (rc) return 1;
    m->plan_B = B; m->plan_full = want_full;
    m->launches = f.launches;
    return 0;
  }

  // ---- fp32 ("exact") schedule: CUDA-core kernels, fp32 activations; UU_PRECISION_TF32 runs the same schedule with the
  // large GEMMs on tcgen05 kind::tf32 (gemm() consumes f.wt32) ---------------------------------------------------------
  const bool tf = m->precision == UU_PRECISION_TF32;
  // K1a: gather list of frames that carry a 2-D pose
  if (use_mask) {
    UU_LAUNCH(f, UU_KIND_GATHER, 3, launch_build_gather(mask, B, N, m->g_scratch, m->g_list, m->g_count, st));
  }
  // K2: fused spatial transformer on the valid frames -> S (compact rows)
  SpatialParams sp;
  sp.x2d = x2d; sp.list = use_mask ? m->g_list : nullptr; sp.count = use_mask ? m->g_count : nullptr;
  sp.max_frames = R; sp.J = J; sp.depth = s.spatial_depth; sp.src = m->cur_src; sp.flip = m->cur_flip;
  sp.embed_k = W(m, "keypoint_embedding", 0); sp.embed_b = W(m, "keypoint_embedding", 1);
  sp.pe = W(m, "spatial_pe", 0); sp.blocks = m->spatial_ptrs;
  sp.norm_g = W(m, "spatial_norm", 0); sp.norm_b = W(m, "spatial_norm", 1);
  sp.out = m->S; sp.out_bf16 = bf;
  UU_LAUNCH(f, UU_KIND_SPATIAL, 1, launch_spatial_f32(sp, st));
  // S4 + T1: 544->384 GEMM, rows scattered to their token position, + bias + temporal PE
  {
    Epilogue e;
    e.bias = W(m, "spatial_to_temporal_fc", 1);
    e.flags = EPI_ROWTABLE; e.table = W(m, "temporal_pe", 0); e.table_period = N;
    if (use_mask) { e.c_rowidx = m->g_list; e.m_dev = m->g_count; }
    f.wt32 = tf ? m->t_s2t : nullptr;
    if (gemm(f, m->S, J * ds, R, J * ds, W(m, "spatial_to_temporal_fc", 0), m->p_s2t, d, e, m->X, 0, d)) return 1;
    if (use_mask) {
      UU_LAUNCH(f, UU_KIND_TOKEN_FILL, 1,
                launch_token_fill(mask, R, N, d, W(m, "strided_input_token_layer", 0), W(m, "temporal_pe", 0), m->X, st));
    }
  }
  // T2/T3: temporal transformer blocks (ReLU MLP)
  for (int i = 0; i < s.temporal_depth; ++i) {
    const BlockW& w = m->tblocks[i];
    UU_LAUNCH(f, UU_KIND_LAYERNORM, 1, launch_layernorm(m->X, R, d, w.ln1_g, w.ln1_b, 1e-5f, nullptr, 1, m->Y, bf, st));
    const uint8_t* km = (use_mask && i < s.first_strided_token_attention_layer) ? mask : nullptr;
    if (attention_block(f, w, m->X, N, km)) return 1;
    UU_LAUNCH(f, UU_KIND_LAYERNORM, 1, launch_layernorm(m->X, R, d, w.ln2_g, w.ln2_b, 1e-5f, nullptr, 1, m->Y, bf, st));
    Epilogue e1;
    e1.bias = w.b1; e1.flags = EPI_RELU;
    f.wt32 = tf ? w.t_fc1 : nullptr;
    if (gemm(f, m->Y, d, R, d, w.w1, w.p_fc1, h, e1, m->Hd, bf, h)) return 1;
    Epilogue e2;
    e2.bias = w.b2; e2.flags = EPI_RESIDUAL; e2.res = m->X; e2.ldr = d;
    f.wt32 = tf ? w.t_fc2 : nullptr;
    if (gemm(f, m->Hd, h, R, h, w.w2, w.p_fc2, d, e2, m->X, 0, d)) return 1;
  }
  // T4: full-sequence head (before the strided blocks modify X in place)
  if (s.full_output && full) {
    Epilogue e;
    e.bias = W(m, "temporal_fc", 1);
    if (gemm(f, m->X, d, R, d, W(m, "temporal_fc", 0), m->p_head1, 3 * J, e, full, 0, 3 * J)) return 1;
  }
  // Q1/Q2: strided transformer blocks (net:122-160)
  float* x_in = m->X;
  for (int i = 0; i < s.n_strided; ++i) {
    const BlockW& w = m->sblocks[i];
    const int L = m->seq_lens[i], Lo = m->seq_lens[i + 1], st_i = s.strides[i];
    const int pl = s.pad_left[i], pr = s.pad_right[i];
    const int Rl = B * L;
    // x += PE_i (written back), y = LN1(x)
    UU_LAUNCH(f, UU_KIND_LAYERNORM, 1,
              launch_layernorm(x_in, Rl, d, w.ln1_g, w.ln1_b, 1e-5f,
                               W(m, "strided_temporal_pe_" + std::to_string(i + 1), 0), L, m->Y, bf, st));
    if (attention_block(f, w, x_in, L, nullptr)) return 1;
    UU_LAUNCH(f, UU_KIND_LAYERNORM, 1, launch_layernorm(x_in, Rl, d, w.ln2_g, w.ln2_b, 1e-5f, nullptr, 1, m->Y, bf, st));
    // Conv1D k=1 + ReLU, written into the zero-padded layout [B, Lo*s, h] (rows never read by the
    // strided conv are dropped; pad rows stay zero from allocation)
    Epilogue e1;
    e1.bias = w.b1; e1.flags = EPI_RELU;
    e1.cmap.rpb = L; e1.cmap.batch_rows = Lo * st_i; e1.cmap.offset = pl; e1.cmap.step = 1;
    f.wt32 = tf ? w.t_fc1 : nullptr;
    if (gemm(f, m->Y, d, Rl, d, w.w1, w.p_fc1, h, e1, m->Hp[i], bf, h)) return 1;
    // strided Conv1D k=3 as an implicit GEMM: row (b,t) = Hp[b, t*s : t*s+3, :] is contiguous (3h values),
    // consecutive rows are s*h apart.  Residual = x[b, c0 + t*s] (MaxPool1D(pool 1, stride s) of the trimmed x).
    Epilogue e2;
    e2.bias = w.b2; e2.flags = EPI_RESIDUAL; e2.res = x_in; e2.ldr = d;
    e2.rmap.rpb = Lo; e2.rmap.batch_rows = L;
    e2.rmap.offset = (st_i > 1 && pl == 0) ? 1 : 0; e2.rmap.step = st_i;
    (void)pr;
    f.wt32 = tf ? w.t_fc2 : nullptr;
    if (gemm(f, m->Hp[i], (long long)st_i * h, B * Lo, 3 * h, w.w2, w.p_fc2, d, e2, m->Xs[i], 0, d)) return 1;
    x_in = m->Xs[i];
  }
  // Q3: central-frame head on the single remaining token
  {
    Epilogue e;
    e.bias = W(m, "strided_temporal_fc", 1);
    if (gemm(f, x_in, d, B, d, W(m, "strided_temporal_fc", 0), m->p_head2, 3 * J, e, central, 0, 3 * J)) return 1;
  }
  m->launches = f.launches;
  return 0;
}

static int check_spec(const uu_spec& s, std::vector<int>& lens) {
  UU_CHECK(s.n_tok >= 1 && s.n_tok <= 128, "n_tok must be in [1,128] (attention keeps a whole window on one SM)");
  UU_CHECK(s.n_joints == 17, "the fused spatial kernel is built for NUM_KEYPOINTS == 17");
  UU_CHECK(s.d_spatial == 32 && s.h_spatial == 64 && s.num_heads == 8,
           "the fused spatial kernel is built for SPATIAL_EMBED_DIM 32, MLP_RATIO 2, NUM_HEADS 8");
  UU_CHECK(s.d_temporal % 128 == 0 && s.d_temporal <= 1024, "TEMPORAL_EMBED_DIM must be a multiple of 128, <= 1024");
  UU_CHECK(s.d_temporal % s.num_heads == 0, "TEMPORAL_EMBED_DIM must be divisible by NUM_HEADS");
  const int dh = s.d_temporal / s.num_heads;
  UU_CHECK(dh == 16 || dh == 32 || dh == 48 || dh == 64, "temporal head_dim must be 16, 32, 48 or 64");
  UU_CHECK(s.h_temporal % 64 == 0, "temporal MLP width must be a multiple of 64");
  UU_CHECK(s.spatial_depth >= 1 && s.temporal_depth >= 1, "depths must be >= 1");
  UU_CHECK(s.spatial_depth <= 4, "the tensor-core spatial kernel keeps at most 4 blocks of weights in shared memory");
  UU_CHECK(s.n_strided >= 1 && s.n_strided <= UU_MAX_STRIDED, "1..8 strided blocks");
  lens.clear();
  lens.push_back(s.n_tok);
  int L = s.n_tok;
  for (int i = 0; i < s.n_strided; ++i) {
    const int st = s.strides[i], pl = s.pad_left[i], pr = s.pad_right[i];
    UU_CHECK(st >= 1 && pl >= 0 && pr >= 0, "bad stride / padding");
    const int Lp = L + pl + pr;
    UU_CHECK(Lp >= 3, "sequence too short for the k=3 strided conv");
    const int Lo = (Lp - 3) / st + 1;
    UU_CHECK(Lo == (L + pl + pr - 2 + st - 1) / st, "strided PE length (net:216) disagrees with the conv output length");
    // identity path length (net:137-152) must equal the conv output length
    int Lt = L;
    if (st > 1) {
      if (pl == 0) Lt -= 1;
      if (pr == 0) Lt -= 1;
      UU_CHECK((Lt - 1) / st + 1 == Lo, "identity path and strided conv lengths differ (the reference would fail too)");
    } else {
      UU_CHECK(L == Lo, "stride 1 needs paddings [1,1]");
    }
    UU_CHECK(st >= 3, "strides < 3 (overlapping conv windows) are not supported by the implicit-GEMM layout");
    lens.push_back(Lo);
    L = Lo;
  }
  UU_CHECK(L == 1, "the strided blocks must reduce the sequence to a single token");
  return 0;
}

}  // namespace uu

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

const char* uu_last_error(void) { return g_error.c_str(); }
int uu_version(void) { return 200; }

int uu_create(const uu_spec* spec, int device, uu_model** out) {
  UU_CHECK(spec && out, "null argument");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    set_error("no CUDA device available: uu3d has no CPU fallback");
    return 1;
  }
  UU_CHECK(device >= 0 && device < n_dev, "device index out of range");
  std::vector<int> lens;
  if (check_spec(*spec, lens)) return 1;
  UU_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  UU_CUDA(cudaGetDeviceProperties(&prop, device));
  UU_CHECK(prop.major == 10, "uu3d is built for sm_100a (B200) only");
  uu_model* m = new uu_model();
  m->num_sms = prop.multiProcessorCount;
  m->spec = *spec;
  m->device = device;
  m->seq_lens = lens;
  build_inventory(m);
  if (cudaMalloc(&m->params, sizeof(float) * m->n_alloc) != cudaSuccess ||
      cudaMemset(m->params, 0, sizeof(float) * m->n_alloc) != cudaSuccess ||
      cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("device allocation failed in uu_create");
    delete m;
    return 1;
  }
  *out = m;
  return 0;
}

int uu_destroy(uu_model* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  drop_graphs(m);
  drop_plans(m);
  free_pool(m->ws_allocs);
  free_pool(m->derived_allocs);
  train_state_destroy(m);
  comm_destroy(m);
  cudaFree(m->grads); cudaFree(m->adam_m); cudaFree(m->adam_v); cudaFree(m->ema);
  cudaFree(m->params);
  cudaFree(m->d_x);
  cudaFree(m->d_mask);
  cudaFree(m->d_full);
  cudaFree(m->d_central);
  cudaFree(m->d_video);
  cudaFree(m->flip_perm); cudaFree(m->tta_full); cudaFree(m->tta_central);
  cudaFree(m->d_centers);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  if (m->copy_stream) {
    cudaStreamDestroy(m->copy_stream);
    for (int c = 0; c < 8; ++c) cudaEventDestroy(m->chunk_ev[c]);
  }
  for (auto e : m->ev_pool) cudaEventDestroy(e);
  delete m;
  return 0;
}

int uu_set_precision(uu_model* m, int precision) {
  UU_CHECK(m, "null model");
  UU_CHECK(precision == UU_PRECISION_FP32 || precision == UU_PRECISION_BF16 || precision == UU_PRECISION_TF32, "unknown precision");
  if (precision != m->precision) {
    drop_graphs(m);
    if (precision == UU_PRECISION_TF32) m->dirty = true;      // the fp32 W^T copies are built on the next commit
  }
  m->precision = precision;
  return 0;
}
int uu_get_precision(const uu_model* m) { return m ? m->precision : -1; }

int uu_weight_count(const uu_model* m) { return m ? (int)m->tensors.size() : -1; }
int64_t uu_param_count(const uu_model* m) { return m ? (int64_t)m->n_params : -1; }

int uu_weight_info(const uu_model* m, int i, char* group, int group_cap, int* index_in_group, int64_t shape[4],
                   int* rank) {
  UU_CHECK(m && i >= 0 && i < (int)m->tensors.size(), "weight index out of range");
  const TensorInfo& t = m->tensors[i];
  if (group && group_cap > 0) {
    std::strncpy(group, t.group.c_str(), group_cap - 1);
    group[group_cap - 1] = 0;
  }
  if (index_in_group) *index_in_group = t.index;
  if (rank) *rank = (int)t.shape.size();
  if (shape) for (size_t k = 0; k < 4; ++k) shape[k] = k < t.shape.size() ? t.shape[k] : 1;
  return 0;
}

static const TensorInfo* find_tensor(uu_model* m, const char* group, int index) {
  auto it = m->lookup.find({std::string(group), index});
  if (it == m->lookup.end()) {
    set_error(std::string("no weight ") + group + "[" + std::to_string(index) + "] in this model");
    return nullptr;
  }
  return &m->tensors[it->second];
}

int uu_set_weight(uu_model* m, const char* group, int index, const float* host, const int64_t* shape, int rank) {
  UU_CHECK(m && group && host && shape, "null argument");
  const TensorInfo* t = find_tensor(m, group, index);
  if (!t) return 1;
  bool ok = rank == (int)t->shape.size();
  for (int k = 0; ok && k < rank; ++k) ok = shape[k] == t->shape[k];
  if (!ok) {   // weight_io.py:219-232 raises ValueError on a shape mismatch
    set_error(std::string("shape mismatch for ") + group + "[" + std::to_string(index) + "]");
    return 1;
  }
  UU_CUDA(cudaSetDevice(m->device));
  UU_CUDA(cudaMemcpy(m->params + t->offset, host, sizeof(float) * t->numel, cudaMemcpyHostToDevice));
  m->dirty = true;
  return 0;
}

int uu_get_weight(uu_model* m, const char* group, int index, float* host, int64_t capacity) {
  UU_CHECK(m && group && host, "null argument");
  const TensorInfo* t = find_tensor(m, group, index);
  if (!t) return 1;
  UU_CHECK((int64_t)t->numel <= capacity, "output buffer too small");
  UU_CUDA(cudaSetDevice(m->device));
  UU_CUDA(cudaMemcpy(host, m->params + t->offset, sizeof(float) * t->numel, cudaMemcpyDeviceToHost));
  return 0;
}

int uu_forward(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central, void* stream) {
  UU_CHECK(m, "null model");
  cudaStream_t st = (cudaStream_t)stream;
  static int use_graphs = -1;
  if (use_graphs < 0) { const char* e = getenv("UU_GRAPH"); use_graphs = (e && e[0] == '0') ? 0 : 1; }
  // graph replay needs a capturable (non-legacy) stream, the tcgen05 schedule (fixed launch sequence) and no per-launch events
  if (!use_graphs || st == nullptr || st == cudaStreamLegacy || m->precision != UU_PRECISION_BF16 || m->profiling || B <= 0)
    return run_forward(m, x2d, mask, B, full, central, st);
  UU_CUDA(cudaSetDevice(m->device));
  if (commit_weights(m, st)) return 1;               // derived weights are rebuilt outside any capture
  if (ensure_workspace(m, B)) return 1;              // (drops the cache when it reallocates)
  const uu_model::GraphKey key{B, x2d, mask, full, central, st};
  if (m->graphs.size() > 32 && !m->graphs.count(key)) drop_graphs(m);
  uu_model::GraphEntry& ge = m->graphs[key];
  if (ge.exec) {
    UU_CUDA(cudaGraphLaunch(ge.exec, st));
    return 0;
  }
  if (ge.failed || ge.seen++ == 0) return run_forward(m, x2d, mask, B, full, central, st);   // first sight: eager (builds plans)
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    ge.failed = true;
    return run_forward(m, x2d, mask, B, full, central, st);
  }
  const int launches = m->launches;
  const int rc = run_forward(m, x2d, mask, B, full, central, st);
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  m->launches = launches;
  if (rc || ce != cudaSuccess || !graph || cudaGraphInstantiate(&ge.exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    ge.exec = nullptr;
    ge.failed = true;
    if (rc) return rc;
    return run_forward(m, x2d, mask, B, full, central, st);
  }
  cudaGraphDestroy(graph);
  UU_CUDA(cudaGraphLaunch(ge.exec, st));
  return 0;
}

int uu_forward_host(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central) {
  UU_CHECK(m && x2d && central, "null argument");
  const uu_spec& s = m->spec;
  UU_CUDA(cudaSetDevice(m->device));
  if (B > m->stage_B) {
    UU_CUDA(cudaDeviceSynchronize());
    cudaFree(m->d_x); cudaFree(m->d_mask); cudaFree(m->d_full); cudaFree(m->d_central);
    m->d_x = m->d_full = m->d_central = nullptr; m->d_mask = nullptr;
    const size_t R = (size_t)B * s.n_tok;
    UU_CUDA(cudaMalloc(&m->d_x, sizeof(float) * R * s.n_joints * 2));
    UU_CUDA(cudaMalloc(&m->d_mask, R));
    UU_CUDA(cudaMalloc(&m->d_full, sizeof(float) * R * s.n_joints * 3));
    UU_CUDA(cudaMalloc(&m->d_central, sizeof(float) * (size_t)B * s.n_joints * 3));
    m->stage_B = B;
  }
  const size_t R = (size_t)B * s.n_tok;
  cudaStream_t st = m->own_stream;
  if (mask) UU_CUDA(cudaMemcpyAsync(m->d_mask, mask, R, cudaMemcpyHostToDevice, st));
  const int n_chunks = (m->precision == UU_PRECISION_BF16 && B >= 2048) ? 4 : 1;
  if (n_chunks > 1) {
    if (!m->copy_stream) {
      UU_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
      for (int c = 0; c < 8; ++c) UU_CUDA(cudaEventCreateWithFlags(&m->chunk_ev[c], cudaEventDisableTiming));
    }
    const int Bc = (B + n_chunks - 1) / n_chunks;
    const size_t per_window = (size_t)s.n_tok * s.n_joints * 2;
    for (int c = 0; c < n_chunks; ++c) {
      const size_t w0 = (size_t)c * Bc, w1 = std::min<size_t>(B, w0 + Bc);
      UU_CUDA(cudaMemcpyAsync(m->d_x + w0 * per_window, x2d + w0 * per_window, sizeof(float) * (w1 - w0) * per_window,
                              cudaMemcpyHostToDevice, m->copy_stream));
      UU_CUDA(cudaEventRecord(m->chunk_ev[c], m->copy_stream));
    }
    m->n_chunks = n_chunks; m->chunk_windows = Bc;
  } else {
    UU_CUDA(cudaMemcpyAsync(m->d_x, x2d, sizeof(float) * R * s.n_joints * 2, cudaMemcpyHostToDevice, st));
  }
  // full == NULL: the full-sequence head still runs (the reference always computes it) but stays on the device
  const int frc = run_forward(m, m->d_x, mask ? m->d_mask : nullptr, B, m->d_full, m->d_central, st);
  m->n_chunks = 0;
  if (frc) return 1;
  if (full) UU_CUDA(cudaMemcpyAsync(full, m->d_full, sizeof(float) * R * s.n_joints * 3, cudaMemcpyDeviceToHost, st));
  UU_CUDA(cudaMemcpyAsync(central, m->d_central, sizeof(float) * (size_t)B * s.n_joints * 3, cudaMemcpyDeviceToHost, st));
  UU_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- sliding windows cut on the device from one video ----------------------------------------
static int forward_video_dev(uu_model* m, const float* video, int T, const int32_t* centers, int B, int s_out, int s_in,
                             int pad_copy, float* full, float* central, cudaStream_t st) {
  const uu_spec& s = m->spec;
  UU_CHECK(video && centers && central && B > 0 && T > 0, "bad argument");
  UU_CHECK(s_out >= 1 && s_in >= s_out && s_in % s_out == 0, "MASK_STRIDE must be a multiple of SEQUENCE_STRIDE");   // :252-254
  UU_CUDA(cudaSetDevice(m->device));
  if (ensure_workspace(m, B)) return 1;
  UU_CUDA(launch_window_index(centers, B, s.n_tok, s_out, s_in, T, pad_copy, m->w_src, m->w_mask, st));
  m->cur_src = m->w_src;
  const int rc = run_forward(m, video, s.has_strided_input ? m->w_mask : nullptr, B, full, central, st);
  m->cur_src = nullptr;
  m->launches += 1;
  return rc;
}

int uu_forward_video(uu_model* m, const float* video2d, int T, const int32_t* centers, int B, int s_out, int s_in,
                     int pad_copy, float* full, float* central, void* stream) {
  UU_CHECK(m, "null model");
  return forward_video_dev(m, video2d, T, centers, B, s_out, s_in, pad_copy, full, central, (cudaStream_t)stream);
}

int uu_forward_video_host(uu_model* m, const float* video2d, int T, const int32_t* centers, int B, int s_out, int s_in,
                          int pad_copy, float* full, float* central) {
  UU_CHECK(m && video2d && centers && central && B > 0 && T > 0, "bad argument");
  const uu_spec& s = m->spec;
  UU_CUDA(cudaSetDevice(m->device));
  if (T > m->video_cap) {
    UU_CUDA(cudaDeviceSynchronize());
    cudaFree(m->d_video); m->d_video = nullptr;
    UU_CUDA(cudaMalloc(&m->d_video, sizeof(float) * (size_t)T * s.n_joints * 2));
    m->video_cap = T;
  }
  if (B > m->centers_cap) {
    UU_CUDA(cudaDeviceSynchronize());
    cudaFree(m->d_centers); m->d_centers = nullptr;
    UU_CUDA(cudaMalloc(&m->d_centers, sizeof(int) * (size_t)B));
    m->centers_cap = B;
  }
  if (B > m->stage_B) {     // output staging shared with uu_forward_host
    UU_CUDA(cudaDeviceSynchronize());
    cudaFree(m->d_x); cudaFree(m->d_mask); cudaFree(m->d_full); cudaFree(m->d_central);
    m->d_x = m->d_full = m->d_central = nullptr; m->d_mask = nullptr;
    const size_t R = (size_t)B * s.n_tok;
    UU_CUDA(cudaMalloc(&m->d_x, sizeof(float) * R * s.n_joints * 2));
    UU_CUDA(cudaMalloc(&m->d_mask, R));
    UU_CUDA(cudaMalloc(&m->d_full, sizeof(float) * R * s.n_joints * 3));
    UU_CUDA(cudaMalloc(&m->d_central, sizeof(float) * (size_t)B * s.n_joints * 3));
    m->stage_B = B;
  }
  const size_t R = (size_t)B * s.n_tok;
  cudaStream_t st = m->own_stream;
  UU_CUDA(cudaMemcpyAsync(m->d_video, video2d, sizeof(float) * (size_t)T * s.n_joints * 2, cudaMemcpyHostToDevice, st));
  UU_CUDA(cudaMemcpyAsync(m->d_centers, centers, sizeof(int) * (size_t)B, cudaMemcpyHostToDevice, st));
  if (forward_video_dev(m, m->d_video, T, m->d_centers, B, s_out, s_in, pad_copy, m->d_full, m->d_central, st)) return 1;
  if (full) UU_CUDA(cudaMemcpyAsync(full, m->d_full, sizeof(float) * R * s.n_joints * 3, cudaMemcpyDeviceToHost, st));
  UU_CUDA(cudaMemcpyAsync(central, m->d_central, sizeof(float) * (size_t)B * s.n_joints * 3, cudaMemcpyDeviceToHost, st));
  UU_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- test-time flip augmentation (eval.py:154-180) ---------------------------------------------
int uu_set_flip_order(uu_model* m, const int32_t* order, int n) {
  UU_CHECK(m && order && n == m->spec.n_joints, "flip order must list n_joints source joints");
  for (int i = 0; i < n; ++i) UU_CHECK(order[i] >= 0 && order[i] < n, "flip order entry out of range");
  UU_CUDA(cudaSetDevice(m->device));
  if (!m->flip_perm) UU_CUDA(cudaMalloc(&m->flip_perm, sizeof(int) * n));
  UU_CUDA(cudaMemcpy(m->flip_perm, order, sizeof(int) * n, cudaMemcpyHostToDevice));
  return 0;
}

static int tta_buffers(uu_model* m, int B) {
  if (B <= m->tta_cap) return 0;
  const uu_spec& s = m->spec;
  UU_CUDA(cudaDeviceSynchronize());
  cudaFree(m->tta_full); cudaFree(m->tta_central);
  m->tta_full = m->tta_central = nullptr;
  UU_CUDA(cudaMalloc(&m->tta_full, sizeof(float) * (size_t)B * s.n_tok * s.n_joints * 3));
  UU_CUDA(cudaMalloc(&m->tta_central, sizeof(float) * (size_t)B * s.n_joints * 3));
  m->tta_cap = B;
  return 0;
}

// pred = (f(x) + unflip(f(flip(x)))) / 2 for both outputs; `video` selects the fused-gather input path
static int forward_tta(uu_model* m, const float* x, const uint8_t* mask, int T, const int32_t* centers, int B, int s_out,
                       int s_in, int pad_copy, bool video, float* full, float* central, cudaStream_t st) {
  UU_CHECK(m->flip_perm, "call uu_set_flip_order before a flip-augmented forward");
  const uu_spec& s = m->spec;
  UU_CUDA(cudaSetDevice(m->device));
  if (tta_buffers(m, B)) return 1;
  const bool want_full = s.full_output && full;
  for (int pass = 0; pass < 2; ++pass) {
    m->cur_flip = pass ? m->flip_perm : nullptr;
    float* fo = pass ? (want_full ? m->tta_full : nullptr) : full;
    float* co = pass ? m->tta_central : central;
    const int rc = video ? forward_video_dev(m, x, T, centers, B, s_out, s_in, pad_copy, fo, co, st)
                         : run_forward(m, x, mask, B, fo, co, st);
    m->cur_flip = nullptr;
    if (rc) return 1;
  }
  UU_CUDA(launch_flip_average(central, m->tta_central, m->flip_perm, B, s.n_joints, st));
  if (want_full) UU_CUDA(launch_flip_average(full, m->tta_full, m->flip_perm, (long long)B * s.n_tok, s.n_joints, st));
  return 0;
}

int uu_forward_tta(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central, void* stream) {
  UU_CHECK(m && x2d && central && B > 0, "bad argument");
  return forward_tta(m, x2d, mask, 0, nullptr, B, 0, 0, 0, false, full, central, (cudaStream_t)stream);
}

int uu_forward_video_tta(uu_model* m, const float* video2d, int T, const int32_t* centers, int B, int s_out, int s_in,
                         int pad_copy, float* full, float* central, void* stream) {
  UU_CHECK(m && video2d && centers && central && B > 0, "bad argument");
  return forward_tta(m, video2d, nullptr, T, centers, B, s_out, s_in, pad_copy, true, full, central, (cudaStream_t)stream);
}

int uu_op_keyframe_interp(const float* pred, const int32_t* frame_indices, int n, int keyframe_stride, int values_per_frame,
                          float* out, void* stream) {
  UU_CHECK(pred && frame_indices && out && n >= 0 && keyframe_stride >= 1 && values_per_frame >= 1, "bad argument");
  UU_CUDA(launch_keyframe_interp(pred, frame_indices, n, keyframe_stride, values_per_frame, out, (cudaStream_t)stream));
  return 0;
}

int uu_op_world_to_cam_and_2d(const float* seq3d, const float* cams, int B, int points_per_sample, float* cam3d, float* p2d,
                              void* stream) {
  UU_CHECK(seq3d && cams && B > 0 && points_per_sample > 0 && (cam3d || p2d), "bad argument");
  UU_CUDA(launch_world_to_cam_2d(seq3d, cams, (long long)B * points_per_sample, points_per_sample, cam3d, p2d,
                                 (cudaStream_t)stream));
  return 0;
}

int uu_op_pose_metrics(const float* pred, const float* gt, int n, int n_joints, int root, float* jpe, float* njpe,
                       double* result_host, void* stream) {
  UU_CHECK(pred && gt && result_host && n > 0 && n_joints >= 1 && n_joints <= 32 && root >= 0 && root < n_joints,
           "bad argument (n_joints <= 32)");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch from the stream-ordered pool: no device-wide synchronisation, re-used from call to call
  float* sums = nullptr;
  double* out = nullptr;
  UU_CUDA(cudaMallocAsync(&sums, sizeof(float) * 3 * (size_t)n + 64, st));
  cudaError_t e = cudaMallocAsync(&out, sizeof(double) * 3, st);
  if (e == cudaSuccess) e = launch_pose_metrics(pred, gt, n, n_joints, root, jpe, njpe, sums, out, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(result_host, out, sizeof(double) * 3, cudaMemcpyDeviceToHost, st);
  cudaFreeAsync(sums, st);
  if (out) cudaFreeAsync(out, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  UU_CUDA(e);
  return 0;
}

int uu_op_window_gather(const float* video2d, int T, const int32_t* centers, int B, int n_tok, int n_joints, int s_out,
                        int s_in, int pad_copy, int32_t* src, uint8_t* mask, float* x2d, void* stream) {
  UU_CHECK(centers && src && mask && B > 0 && T > 0 && n_tok > 0, "bad argument");
  UU_CHECK(s_out >= 1 && s_in >= s_out && s_in % s_out == 0, "MASK_STRIDE must be a multiple of SEQUENCE_STRIDE");
  UU_CUDA(launch_window_index(centers, B, n_tok, s_out, s_in, T, pad_copy, src, mask, (cudaStream_t)stream));
  if (x2d) {
    UU_CHECK(video2d, "video2d is required to materialise the windows");
    UU_CUDA(launch_window_copy(video2d, src, B * n_tok, n_joints, x2d, (cudaStream_t)stream));
  }
  return 0;
}

int uu_last_launch_count(const uu_model* m) { return m ? m->launches : -1; }

int uu_set_profiling(uu_model* m, int on) {
  UU_CHECK(m, "null model");
  m->profiling = on != 0;
  return 0;
}

int uu_get_profile(uu_model* m, float* ms_by_kind, int32_t* launches_by_kind, int n_kinds) {
  UU_CHECK(m && ms_by_kind && launches_by_kind && n_kinds >= UU_KIND_COUNT, "bad argument");
  for (int i = 0; i < n_kinds; ++i) { ms_by_kind[i] = 0.f; launches_by_kind[i] = 0; }
  for (auto& e : m->ev_used) {
    UU_CUDA(cudaEventSynchronize(e.second.second));
    float ms = 0.f;
    UU_CUDA(cudaEventElapsedTime(&ms, e.second.first, e.second.second));
    ms_by_kind[e.first] += ms;
    launches_by_kind[e.first] += 1;
  }
  return 0;
}

int uu_stride_mask(int n_tok, int s_out, int s_in, int64_t shift, uint8_t* mask_out) {
  UU_CHECK(mask_out && n_tok > 0 && s_out > 0 && s_in > 0, "bad argument");
  UU_CHECK(s_in >= s_out && s_in % s_out == 0, "MASK_STRIDE must be a multiple of SEQUENCE_STRIDE");   // :252-254
  for (int n = 0; n < n_tok; ++n) {
    int64_t idx = (int64_t)(n - n_tok / 2) * s_out + shift;
    int64_t r = idx % s_in;
    if (r < 0) r += s_in;   // numpy floor-mod
    mask_out[n] = r == 0;
  }
  return 0;
}

// ---- single-kernel entry points ------------------------------------------------------------------
int uu_op_build_gather(const uint8_t* mask, int B, int n_tok, int32_t* scratch, int32_t* list, int32_t* count,
                       void* stream) {
  UU_CUDA(launch_build_gather(mask, B, n_tok, scratch, list, count, (cudaStream_t)stream));
  return 0;
}
int uu_op_token_fill(const uint8_t* mask, int rows, int n_tok, int d, const float* token, const float* pe, float* x,
                     void* stream) {
  UU_CUDA(launch_token_fill(mask, rows, n_tok, d, token, pe, x, (cudaStream_t)stream));
  return 0;
}
int uu_op_layernorm(float* x, int rows, int d, const float* gamma, const float* beta, float eps, const float* table,
                    int period, void* y, int y_bf16, void* stream) {
  UU_CUDA(launch_layernorm(x, rows, d, gamma, beta, eps, table, period, y, y_bf16, (cudaStream_t)stream));
  return 0;
}
int uu_op_attention(const void* qkv, int is_bf16, int B, int S, int heads, int dh, const uint8_t* keep_mask,
                    int mask_stride, void* out, void* stream) {
  UU_CUDA(launch_attention(qkv, is_bf16, B, S, heads, dh, keep_mask, mask_stride, out, (cudaStream_t)stream));
  return 0;
}
/* T3 on tcgen05 / TMEM / TMA (attn_tc5.cu): bf16 q | k | v rows (B * S, 1152) -> bf16 (B * S, 384), 8 heads of 48. */
int uu_op_attention_tc5(const void* qkv, int B, int S, const uint8_t* keep_mask, int mask_stride, void* out, void* stream) {
  UU_CHECK(qkv && out && B > 0 && attention_tc5_ok(S, 8, 48), "uu_op_attention_tc5: 1 <= S <= 80 (8 heads of dimension 48)");
  int dev = 0, sms = 148;
  UU_CUDA(cudaGetDevice(&dev));
  UU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  UU_CUDA(launch_attention_tc5((const bf16*)qkv, B, S, keep_mask, mask_stride, (bf16*)out, sms, (cudaStream_t)stream));
  return 0;
}
/* Attention of the training step on mma.sync TF32 (attn_mma.cu), fp32 rows.  dO == NULL: forward, out (B * S, heads * dh);
 * otherwise backward, dqkv (B * S, 3 * heads * dh).  nsplit 2 = bf16 hi + lo planes, 3 = compensated TF32, 1 = plain TF32. */
int uu_op_attention_train(const float* qkv, const float* dO, int B, int S, int heads, int dh, const uint8_t* keep_mask,
                          int mask_stride, float* out, float* dqkv, int nsplit, void* stream) {
  UU_CHECK(qkv && B > 0 && attention_mma_ok(B, S, heads, dh), "uu_op_attention_train: 1 <= S <= 80, head dimension 32 / 48 / 64");
  UU_CHECK(nsplit >= 1 && nsplit <= 3, "uu_op_attention_train: nsplit is 1, 2 or 3");
  if (!dO) {
    UU_CHECK(out, "uu_op_attention_train: forward needs out");
    UU_CUDA(launch_attention_mma_fwd(qkv, B, S, heads, dh, keep_mask, mask_stride, out, nsplit, (cudaStream_t)stream));
  } else {
    UU_CHECK(dqkv, "uu_op_attention_train: backward needs dqkv");
    UU_CUDA(launch_attention_mma_bwd(qkv, dO, B, S, heads, dh, keep_mask, mask_stride, dqkv, nsplit, (cudaStream_t)stream));
  }
  return 0;
}
int uu_op_spatial(uu_model* m, const float* x2d, const uint8_t* mask, int B, void* out, int32_t* n_valid_out,
                  void* stream) {
  UU_CHECK(m && x2d && out && B > 0, "bad argument");
  const uu_spec& s = m->spec;
  cudaStream_t st = (cudaStream_t)stream;
  UU_CUDA(cudaSetDevice(m->device));
  if (commit_weights(m, st)) return 1;
  if (ensure_workspace(m, B)) return 1;
  const int R = B * s.n_tok;
  const bool use_mask = mask != nullptr;
  if (use_mask) UU_CUDA(launch_build_gather(mask, B, s.n_tok, m->g_scratch, m->g_list, m->g_count, st));
  if (m->precision == UU_PRECISION_BF16) {
    UU_CUDA(launch_spatial_tc(x2d, use_mask ? m->g_list : nullptr, use_mask ? m->g_count : nullptr, R, s.spatial_depth,
                              m->sp_frags, m->sp_params, (bf16*)out, m->num_sms, st));
  } else {
    SpatialParams sp;
    sp.x2d = x2d; sp.list = use_mask ? m->g_list : nullptr; sp.count = use_mask ? m->g_count : nullptr;
    sp.max_frames = R; sp.J = s.n_joints; sp.depth = s.spatial_depth;
    sp.embed_k = W(m, "keypoint_embedding", 0); sp.embed_b = W(m, "keypoint_embedding", 1);
    sp.pe = W(m, "spatial_pe", 0); sp.blocks = m->spatial_ptrs;
    sp.norm_g = W(m, "spatial_norm", 0); sp.norm_b = W(m, "spatial_norm", 1);
    sp.out = out; sp.out_bf16 = 0;
    UU_CUDA(launch_spatial_f32(sp, st));
  }
  int n = R;
  if (use_mask) {
    UU_CUDA(cudaMemcpyAsync(&n, m->g_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  UU_CUDA(cudaStreamSynchronize(st));
  if (n_valid_out) *n_valid_out = n;
  return 0;
}
int uu_op_gemm_f32(const float* A, int64_t lda, const float* Wm, int M, int N, int K, const float* bias, int flags,
                   const float* res, int64_t ldr, float* C, int64_t ldc, void* stream) {
  Epilogue e;
  e.bias = bias; e.flags = flags & (EPI_RELU | EPI_RESIDUAL); e.res = res; e.ldr = ldr;
  UU_CUDA(launch_gemm_simt(A, 0, lda, Wm, M, N, K, e, C, 0, ldc, (cudaStream_t)stream));
  return 0;
}
int uu_op_gemm_bf16(const void* A, int64_t lda, int M, int K, const void* Wt, int N_pad, int N, const float* bias,
                    int flags, const float* res, int64_t ldr, void* C, int c_bf16, int64_t ldc, void* stream) {
  Epilogue e;
  e.bias = bias; e.flags = flags & (EPI_RELU | EPI_RESIDUAL); e.res = res; e.ldr = ldr;
  TcGemmPlan* p = nullptr;
  if (tc_gemm_plan_create(&p, (const bf16*)A, lda, M, K, (const bf16*)Wt, N_pad, N)) return 1;
  cudaError_t err = tc_gemm_launch(p, e, C, c_bf16, ldc, (cudaStream_t)stream);
  tc_gemm_plan_destroy(p);
  UU_CUDA(err);
  return 0;
}

/* tcgen05 kind::tf32 GEMM (training math mode 1): C (M, N) fp32 = A (M, K) fp32 . Bt^T (+ bias) (ReLU) (+ res). */
int uu_op_gemm_tf32(const float* A, int64_t lda, int M, int K, const float* Bt, int64_t ldb, int N, const float* bias,
                    int flags, const float* res, int64_t ldr, float* C, int64_t ldc, void* stream) {
  Epilogue e;
  e.bias = bias; e.flags = flags & (EPI_RELU | EPI_RESIDUAL); e.res = res; e.ldr = ldr;
  TcGemmPlan* p = nullptr;
  if (tc_gemm_plan_create_tf32(&p, A, lda, M, K, Bt, ldb, N)) return 1;
  cudaError_t err = tc_gemm_launch(p, e, C, 0, ldc, (cudaStream_t)stream);
  tc_gemm_plan_destroy(p);
  UU_CUDA(err);
  return 0;
}

// ---- the folded-LayerNorm / residual epilogues in isolation (parity tests) ------------------------------------------
namespace {
struct TmpPool {
  std::vector<void*> v;
  ~TmpPool() { for (void* p : v) cudaFree(p); }
};
}  // namespace

int uu_op_ln_gemm_bf16(const void* x, int rows, int d, const float* gamma, const float* beta, float eps, const float* Wm,
                       const float* bias, int N, int relu, void* out, void* stream) {
  UU_CHECK(rows > 0 && d % 128 == 0 && d <= 1024 && N % 64 == 0, "uu_op_ln_gemm_bf16: d % 128 == 0, N % 64 == 0 required");
  cudaStream_t st = (cudaStream_t)stream;
  const int slots = d / 64;
  TmpPool tp;
  void *wt, *cs, *bl, *stats;
  if (dev_alloc(tp.v, &wt, sizeof(bf16) * (size_t)N * d, false) || dev_alloc(tp.v, &cs, sizeof(float) * N, false) ||
      dev_alloc(tp.v, &bl, sizeof(float) * N, false) || dev_alloc(tp.v, &stats, sizeof(float) * 2 * (size_t)rows * slots, false))
    return 1;
  k_pack_wt_ln<<<(N + 7) / 8, 256, 0, st>>>(Wm, d, N, N, gamma, beta, bias, (bf16*)wt, (float*)cs, (float*)bl);
  UU_CUDA(cudaGetLastError());
  const RowMap plain;
  UU_CUDA(launch_residual_ln_bx((const bf16*)x, plain, nullptr, nullptr, rows, d, nullptr, nullptr, 0.f, nullptr, 1, nullptr,
                                nullptr, st, (float*)stats, slots));
  Epilogue e;
  e.bias = (const float*)bl; e.flags = EPI_LNFOLD | (relu ? EPI_RELU : 0);
  e.ln_stats = (const float*)stats; e.ln_csum = (const float*)cs; e.ln_slots = slots; e.ln_inv_k = 1.f / d; e.ln_eps = eps;
  TcGemmPlan* p = nullptr;
  if (tc_gemm_plan_create(&p, (const bf16*)x, d, rows, d, (const bf16*)wt, N, N)) return 1;
  cudaError_t err = tc_gemm_launch(p, e, out, 1, N, st);
  tc_gemm_plan_destroy(p);
  UU_CUDA(err);
  UU_CUDA(cudaStreamSynchronize(st));     // temporaries are released on return
  return 0;
}

/* x (bf16 [rows, 384], in place) += fc2(ReLU(fc1(LN(x; gamma, beta, eps)))) through the fused MLP kernel (mlp_tc.cuh);
 * W1 (384, h), W2 (h, 384) fp32 on the device; stats_out (optional) [rows, 6, 2] statistics of the result. */
int uu_op_mlp_bf16(void* x, int rows, int d, int h, const float* gamma, const float* beta, float eps, const float* W1,
                   const float* b1, const float* W2, const float* b2, float* stats_out, void* stream) {
  UU_CHECK(rows > 0 && d == 384 && h % 128 == 0 && h >= 128 && h <= 768, "uu_op_mlp_bf16: d == 384, h % 128 == 0, 128 <= h <= 768 required");
  cudaStream_t st = (cudaStream_t)stream;
  const int slots = d / 64;
  TmpPool tp;
  void *w1t, *w2t, *cs, *bl, *stats;
  if (dev_alloc(tp.v, &w1t, sizeof(bf16) * (size_t)h * d, false) || dev_alloc(tp.v, &w2t, sizeof(bf16) * (size_t)h * d, false) ||
      dev_alloc(tp.v, &cs, sizeof(float) * h, false) || dev_alloc(tp.v, &bl, sizeof(float) * h, false) ||
      dev_alloc(tp.v, &stats, sizeof(float) * 2 * (size_t)rows * slots, false))
    return 1;
  k_pack_wt_ln<<<(h + 7) / 8, 256, 0, st>>>(W1, d, h, h, gamma, beta, b1, (bf16*)w1t, (float*)cs, (float*)bl);
  k_pack_wt<<<256, 256, 0, st>>>(W2, h, d, d, (bf16*)w2t);
  UU_CUDA(cudaGetLastError());
  const RowMap plain;
  UU_CUDA(launch_residual_ln_bx((const bf16*)x, plain, nullptr, nullptr, rows, d, nullptr, nullptr, 0.f, nullptr, 1, nullptr,
                                nullptr, st, (float*)stats, slots));
  MlpPlan* p = nullptr;
  if (mlp_plan_create(&p, (const bf16*)x, d, rows, d, h, (const bf16*)w1t, (const bf16*)w2t)) return 1;
  MlpArgs a;
  a.M = rows; a.n_chunks = h / 64; a.ln_stats = (const float*)stats; a.ln_slots = slots; a.ln_inv_k = 1.f / d; a.ln_eps = eps;
  a.csum1 = (const float*)cs; a.bias1 = (const float*)bl;
  a.epi2.bias = b2; a.epi2.flags = EPI_RESID_BF16; a.epi2.res_bf16 = (const bf16*)x; a.epi2.stats_out = stats_out;
  a.epi2.ln_slots = slots;
  a.X = (bf16*)x; a.ldx = d;
  cudaError_t err = mlp_launch(p, a, st);
  mlp_plan_destroy(p);
  UU_CUDA(err);
  UU_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int uu_op_resid_gemm_bf16(const void* A, int rows, int K, const float* Wm, const float* bias, int d, void* x, int resid,
                          const float* table, int period, float* stats_out, void* stream) {
  UU_CHECK(rows > 0 && K % 8 == 0 && d % 64 == 0 && d / 64 <= 32, "uu_op_resid_gemm_bf16: K % 8 == 0, d % 64 == 0 required");
  UU_CHECK(resid || table, "uu_op_resid_gemm_bf16: residual mode or a positional table");
  cudaStream_t st = (cudaStream_t)stream;
  TmpPool tp;
  void* wt;
  if (dev_alloc(tp.v, &wt, sizeof(bf16) * (size_t)d * K, false)) return 1;
  k_pack_wt<<<256, 256, 0, st>>>(Wm, K, d, d, (bf16*)wt);
  UU_CUDA(cudaGetLastError());
  Epilogue e;
  e.bias = bias; e.stats_out = stats_out; e.ln_slots = d / 64;
  if (resid) { e.flags = EPI_RESID_BF16; e.res_bf16 = (const bf16*)x; }
  else { e.flags = EPI_ROWTABLE; e.table = table; e.table_period = period; }
  TcGemmPlan* p = nullptr;
  if (tc_gemm_plan_create(&p, (const bf16*)A, K, rows, K, (const bf16*)wt, d, d)) return 1;
  cudaError_t err = tc_gemm_launch(p, e, x, 1, d, st);
  tc_gemm_plan_destroy(p);
  UU_CUDA(err);
  UU_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
