// Shared declarations of the uu3d CUDA library (sm_100a only).
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace uu {

typedef __nv_bfloat16 bf16;

// ---- error plumbing: never throw across the C ABI -------------------------------------------
void set_error(const std::string& msg);
#define UU_CUDA(expr)                                                                            \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ::uu::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +       \
                      __FILE__ + ":" + std::to_string(__LINE__));                                \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)
#define UU_CHECK(cond, msg)                                                                      \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      ::uu::set_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +                   \
                      std::to_string(__LINE__));                                                 \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

// ---- row addressing shared by every GEMM flavour --------------------------------------------
// Logical GEMM row r -> physical row of a (batch, position) matrix:
//   pos = (r % rpb) * step + offset ; row = (r / rpb) * batch_rows + pos ; dropped if pos >= batch_rows.
// plain matrix:           rpb = INT_MAX
// zero-padded conv input: rpb = L,  batch_rows = L_out*s, offset = pad_left, step = 1   (write side)
// strided identity path:  rpb = L_out, batch_rows = L, offset = c0, step = s            (read side)
struct RowMap {
  int rpb = 0x7fffffff;
  int batch_rows = 0x7fffffff;
  int offset = 0;
  int step = 1;
};
__host__ __device__ inline long long map_row(const RowMap& m, int r) {
  if (m.rpb == 0x7fffffff) return r;
  int b = r / m.rpb;
  int pos = (r - b * m.rpb) * m.step + m.offset;
  if (pos >= m.batch_rows) return -1;
  return (long long)b * m.batch_rows + pos;
}

enum EpiFlags : int {
  EPI_RELU = 1,       // max(., 0) after bias
  EPI_RESIDUAL = 2,   // += Res[map_row(rmap, r)][col]
  EPI_ROWTABLE = 4,   // += table[(crow % table_period)][col]   (temporal positional encoding)
  // tcgen05 TMA-store epilogue only (bf16 schedule, temporal blocks):
  EPI_LNFOLD = 8,     // the GEMM consumed the raw residual stream with gamma folded into W:
                      //   out = rstd[r] * (acc - mean[r] * csum[col]) + bias'[col]   ( == LN(x) W + b )
  EPI_RESID_BF16 = 16 // out = acc + bias + res_bf16[r][col], rounded to bf16; optional row statistics of the result
};

// Programmatic dependent launch (PDL): a kernel launched with the attribute may become resident while the previous
// kernel of the stream drains; it runs its prologue (barrier init, TMEM allocation, descriptor prefetch) and then
// blocks in pdl_wait() until the previous grid has completed and its writes are visible.  pdl_trigger() at the start
// of a kernel lets ITS successor do the same.  Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();            // UU_PDL=0 / 1 forces; default: what the running schedule asked for (pdl_set_auto)
void pdl_set_auto(bool on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// Epilogue description shared by the SIMT and the tcgen05 GEMM.
struct Epilogue {
  const float* bias = nullptr;       // [N] or null
  int flags = 0;
  const float* res = nullptr;        // fp32 residual source (may alias the fp32 output)
  long long ldr = 0;
  RowMap rmap;
  const float* table = nullptr;      // [table_period, N]
  int table_period = 1;
  const int* c_rowidx = nullptr;     // optional scatter list: physical output row of logical row r
  const int* m_dev = nullptr;        // optional device-side row count (<= M)
  RowMap cmap;                       // output row mapping (ignored when c_rowidx is set)
  // LayerNorm folded into the GEMMs.  Row statistics travel as per-row partials [rows][ln_slots][2] =
  // (sum, sum of squares) over 64-column slots of the bf16 residual stream, summed in slot order by the consumer.
  const float* ln_stats = nullptr;   // EPI_LNFOLD: partials of the A rows
  const float* ln_csum = nullptr;    // EPI_LNFOLD: [N] column sums of the folded bf16 weight
  int ln_slots = 0;
  float ln_inv_k = 0.f, ln_eps = 0.f;
  const bf16* res_bf16 = nullptr;    // EPI_RESID_BF16: residual rows, same row index and pitch as the output
  float* stats_out = nullptr;        // EPI_RESID_BF16: partials of the rows written, [rows][N / 64][2] (optional)
};

// ---- kernel launchers (kernels_f32.cu) --------------------------------------------------------
// All launchers are asynchronous on `st` and return cudaGetLastError().

// Valid-frame gather list from the stride mask: list[0..count) = b*n_tok+n of frames with mask != 0,
// ascending.  count_out[0] = count.  `scratch` holds B+1 ints.
cudaError_t launch_build_gather(const uint8_t* mask, int B, int n_tok, int* scratch, int* list, int* count_out,
                                cudaStream_t st);

struct SpatialParams {
  const float* x2d;          // (B*n_tok, J, 2)
  const int* list;           // gather list (frame ids) or null = all frames
  const int* src = nullptr;  // optional: token id -> row of x2d (video frame; -1 = zeros), fused window gather
  const int* flip = nullptr; // optional: flip augmentation, joint j reads source joint flip[j] with x negated
  const int* count;          // device count of valid frames (null with list == null)
  int max_frames;            // B*n_tok
  int J, depth;
  const float* embed_k;      // (2, 32)
  const float* embed_b;      // (32)
  const float* pe;           // (J, 32)
  const float* const* blocks;  // device array [depth][16] of tensor pointers, file order
  const float* norm_g;       // spatial_norm
  const float* norm_b;
  void* out;                 // (n_valid, J*32) compact, fp32 or bf16
  int out_bf16;
};
cudaError_t launch_spatial_f32(const SpatialParams& p, cudaStream_t st);

// Rows without 2-D input: x[row] = token + pe[row % n_tok]   (net:350-352 for masked frames)
cudaError_t launch_token_fill(const uint8_t* mask, int rows, int n_tok, int d, const float* token, const float* pe,
                              float* x, cudaStream_t st);

// y = LN(x (+ table[row % period], written back to x when table != null)); d % 128 == 0, d <= 1024
cudaError_t launch_layernorm(float* x, int rows, int d, const float* gamma, const float* beta, float eps,
                             const float* table, int period, void* y, int y_bf16, cudaStream_t st);

// bf16-resident residual stream variants (X stored as bf16; arithmetic fp32).  upd / xdst / y / xcast optional.
cudaError_t launch_token_fill_bx(const uint8_t* mask, int rows, int n_tok, int d, const float* token, const float* pe,
                                 bf16* x, cudaStream_t st, float* stats = nullptr, int slots = 0);
cudaError_t launch_residual_ln_bx(const bf16* xsrc, const RowMap& smap, const bf16* upd, bf16* xdst, int rows, int d,
                                  const float* gamma, const float* beta, float eps, const float* table, int period,
                                  bf16* y, bf16* xcast, cudaStream_t st, float* stats = nullptr, int slots = 0);

// softmax(q k^T / sqrt(dh) + keymask * -1e9) v per (window, head); qkv rows = [q | k | v] (3*d)
cudaError_t launch_attention(const void* qkv, int is_bf16, int B, int S, int heads, int dh, const uint8_t* mask,
                             int mask_stride, void* out, cudaStream_t st);

// C = epi(A[M,K] * W[K,N]); A fp32 or bf16 row-major with leading dimension lda; W fp32 (in,out).
cudaError_t launch_gemm_simt(const void* A, int a_bf16, long long lda, const float* W, int M, int N, int K,
                             const Epilogue& epi, void* C, int c_bf16, long long ldc, cudaStream_t st);

// ---- tensor-core attention (attention_tc.cu): bf16 q|k|v rows in, bf16 merged heads out ---------
cudaError_t launch_attention_tc(const bf16* qkv, int B, int S, int heads, int dh, const uint8_t* mask, int mask_stride,
                                bf16* out, cudaStream_t st);

// ---- tcgen05 / TMEM attention fed by TMA (attn_tc5.cu): 8 heads of dimension 48, S <= 80 tokens per window ----
bool attention_tc5_ok(int S, int heads, int dh);
cudaError_t launch_attention_tc5(const bf16* qkv, int B, int S, const uint8_t* mask, int mask_stride, bf16* out, int num_sms,
                                 cudaStream_t st);

// ---- tensor-core spatial transformer (spatial_tc.cu) -------------------------------------------
size_t spatial_tc_frag_bytes(int depth);
size_t spatial_tc_param_bytes(int depth);
cudaError_t launch_spatial_pack(const float* const* blocks, int depth, const float* embed_k, const float* embed_b,
                                const float* pe, const float* norm_g, const float* norm_b, void* frags, float* params,
                                cudaStream_t s);
cudaError_t launch_spatial_tc(const float* x2d, const int* list, const int* count, int max_frames, int depth,
                              const void* frags, const float* params, bf16* out, int num_sms, cudaStream_t s,
                              const int* src = nullptr, const int* range_lo = nullptr, const int* range_hi = nullptr,
                              int lo = 0, int hi = -1, const int* flip = nullptr);

// ---- evaluation glue (kernels_f32.cu): flip-augmentation average, key-frame interpolation ------------------------
cudaError_t launch_flip_average(float* a, const float* b, const int* perm, long long n_poses, int J, cudaStream_t st);
cudaError_t launch_keyframe_interp(const float* pred, const int* fidx, int n, int stride, int V, float* out, cudaStream_t st);

// ---- sliding windows of one video (kernels_f32.cu): source-frame table + globally aligned stride mask ----
cudaError_t launch_window_index(const int* centers, int B, int n_tok, int s_out, int s_in, int T, int pad_copy, int* src,
                                uint8_t* mask, cudaStream_t st);
cudaError_t launch_window_copy(const float* video, const int* src, int n_tokens, int J, float* x, cudaStream_t st);

// ---- tcgen05 GEMM (gemm_tc.cu) ----------------------------------------------------------------
struct TcGemmPlan;   // holds the TMA tensor maps of one GEMM call site
int tc_gemm_plan_create(TcGemmPlan** out, const bf16* A, long long lda, int M, int K, const bf16* Wt, int N_pad,
                        int N);
// world -> camera -> 2-D (uplifiting_dataset.py:669-761): x3d (B * points_per_sample, 3), cams (B, 18)
cudaError_t launch_world_to_cam_2d(const float* x3d, const float* cams, long long n_points, int points_per_sample,
                                   float* cam3d, float* p2d, cudaStream_t st);
// MPJPE / N-MPJPE (metrics.py:13-81): pred (n, J, 3), gt (n, J, 4 = x, y, z, valid); out[3] = mpjpe, nmpjpe, valid count
cudaError_t launch_pose_metrics(const float* pred, const float* gt, int n, int J, int root, float* jpe, float* njpe,
                                float* sums, double* out, cudaStream_t st);
int tc_gemm_plan_create_tf32(TcGemmPlan** out, const float* A, long long lda, int M, int K, const float* Bt, long long ldb,
                             int N);
void tc_gemm_plan_destroy(TcGemmPlan* p);
cudaError_t tc_gemm_launch(TcGemmPlan* p, const Epilogue& epi, void* C, int c_bf16, long long ldc,
                           cudaStream_t st);


// ---- fused temporal MLP (mlp_tc.cuh, compiled into gemm_tc.cu): x += fc2(ReLU(fc1(LN2(x)))) in one tcgen05 kernel ----
struct MlpPlan;
struct MlpArgs {
  int M;                   // rows of X
  int n_chunks;            // hidden width / 64
  const float* ln_stats;   // [M][slots][2] partial sums of the rows of X (EPI_LNFOLD convention)
  int ln_slots;
  float ln_inv_k, ln_eps;
  const float* csum1;      // [h] column sums of bf16(gamma (.) W1)
  const float* bias1;      // [h] b1 + beta W1
  Epilogue epi2;           // bias = b2, res_bf16 = X, stats_out, ln_slots, flags = EPI_RESID_BF16
  bf16* X;
  long long ldx;
};

int mlp_plan_create(MlpPlan** out, const bf16* X, long long ldx, int M, int d, int h, const bf16* W1t, const bf16* W2t);
void mlp_plan_destroy(MlpPlan* p);
cudaError_t mlp_launch(const MlpPlan* p, const MlpArgs& args, cudaStream_t st);

}  // namespace uu
