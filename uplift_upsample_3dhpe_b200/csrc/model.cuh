// Internal model object shared by the inference (uu_api.cu) and training (uu_train.cu) translation units.
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/uu3d.h"
#include "common.cuh"

namespace uu {

// ------------------------------------------------------------------------------------------------
struct TensorInfo {
  std::string group;
  int index;
  std::vector<int64_t> shape;
  size_t offset;   // into the flat fp32 parameter buffer (elements)
  size_t numel;
};

struct Pack {          // bf16 W^T [n_pad, k] of a (k, n) fp32 matrix
  bf16* ptr = nullptr;
  int n = 0, n_pad = 0, k = 0;
};

struct BlockW {        // one temporal / strided block
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  float* wqkv = nullptr;   // fused fp32 [d, 3d]
  float* bqkv = nullptr;   // [3d]
  // w2: fc2 (h, d) or strided conv (3, h, d) == [3h, d]
  const float *wp = nullptr, *bp = nullptr, *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
  Pack p_qkv, p_proj, p_fc1, p_fc2;
  float *t_qkv = nullptr, *t_proj = nullptr, *t_fc1 = nullptr, *t_fc2 = nullptr;   // fp32 W^T (N, K): B operand of the tf32 schedule
  // temporal blocks, bf16 schedule: LayerNorm folded into the consuming GEMM (EPI_LNFOLD).
  //   p_*_ln = bf16((gamma (.) W)^T), cs_* = column sums of that bf16 matrix, bl_* = b + beta W
  Pack p_qkv_ln, p_fc1_ln;
  float *cs_qkv = nullptr, *bl_qkv = nullptr, *cs_fc1 = nullptr, *bl_fc1 = nullptr;
};

}  // namespace uu

using namespace uu;

namespace uu { struct TrainState; }

struct uu_model {
  uu_spec spec;
  int device = 0;
  int precision = UU_PRECISION_FP32;
  std::vector<TensorInfo> tensors;
  std::map<std::pair<std::string, int>, int> lookup;
  std::vector<int> seq_lens;
  float* params = nullptr;
  size_t n_params = 0;     // exact parameter count
  size_t n_alloc = 0;      // floats in the flat buffer (tensors padded to 16 bytes)
  bool dirty = true;

  // derived weights
  std::vector<BlockW> tblocks, sblocks;
  const float** spatial_ptrs = nullptr;    // device array [depth][16]
  void* sp_frags = nullptr;                // tensor-core spatial kernel: B-fragment image + fp32 params
  float* sp_params = nullptr;
  int num_sms = 148;
  Pack p_s2t, p_head1, p_head2;
  float* t_s2t = nullptr;                  // fp32 W^T of spatial_to_temporal_fc (tf32 schedule)
  std::vector<void*> derived_allocs;

  // workspace (sized for cap_B windows in the current precision)
  int cap_B = 0;
  int ws_precision = -1;
  int *g_scratch = nullptr, *g_list = nullptr, *g_count = nullptr;
  int* w_src = nullptr;          // sliding-window source-frame table [cap_B * n_tok] (uu_forward_video)
  uint8_t* w_mask = nullptr;     // globally aligned stride mask built on the device [cap_B * n_tok]
  const int* cur_src = nullptr;  // non-null while a video forward is being scheduled
  int* flip_perm = nullptr;      // device copy of AUGM_FLIP_KEYPOINT_ORDER (uu_set_flip_order)
  const int* cur_flip = nullptr; // non-null while the flipped half of a test-time-augmented forward is scheduled
  float *tta_full = nullptr, *tta_central = nullptr;   // outputs of the flipped pass
  int tta_cap = 0;
  float* d_video = nullptr;      // device staging for uu_forward_video_host
  int* d_centers = nullptr;
  int video_cap = 0, centers_cap = 0;
  void *S = nullptr, *Y = nullptr, *QKV = nullptr, *O = nullptr, *Hd = nullptr, *P = nullptr;
  float* ln_stats = nullptr;   // bf16 schedule: LayerNorm row partials [R][d / 64][2] (see Epilogue::ln_stats)
  float* X = nullptr;
  std::vector<float*> Xs;      // strided stream after block i: [cap_B * seq_lens[i+1], d]
  std::vector<void*> Hp;       // zero-padded conv inputs: [cap_B * Lo*s, h]
  std::vector<void*> ws_allocs;
  // device staging for uu_forward_host
  float *d_x = nullptr, *d_full = nullptr, *d_central = nullptr;
  uint8_t* d_mask = nullptr;
  int stage_B = 0;
  cudaStream_t own_stream = nullptr;
  // uu_forward_host: the input copy is split into chunks on a second stream; the spatial kernel of chunk c waits
  // only for its own chunk, so the rest of the copy overlaps spatial compute
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t chunk_ev[8] = {};
  int n_chunks = 0;              // > 1 while such a forward is being scheduled
  int chunk_windows = 0;

  // tcgen05 plans for the current batch size
  int plan_B = -1;
  int plan_full = -1;
  std::vector<TcGemmPlan*> plans;
  std::vector<MlpPlan*> mlp_plans;      // fused temporal MLP call sites (mlp_tc.cuh)
  int launches = 0;

  // CUDA-graph cache of the device-pointer forward (uu_forward): key = (batch, buffers, stream); the first call with a
  // key runs eagerly, the second is captured, later ones replay (kernel-to-kernel launch gaps: -3 % at 4096 windows,
  // -8 % at 512).  Dropped whenever plans / workspace / precision change.
  struct GraphKey {
    int B; const void *x, *mask, *full, *central; cudaStream_t st;
    bool operator<(const GraphKey& o) const {
      return std::tie(B, x, mask, full, central, st) < std::tie(o.B, o.x, o.mask, o.full, o.central, o.st);
    }
  };
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int seen = 0; bool failed = false; };
  std::map<GraphKey, GraphEntry> graphs;

  // ---- training state (uu_train.cu) ----
  float* grads = nullptr;        // flat, same layout as params
  float *adam_m = nullptr, *adam_v = nullptr, *ema = nullptr;
  struct TrainState* train = nullptr;

  // ---- data-parallel training (uu_comm.cu): NCCL communicator owned by the model, a side stream for the bucketed
  // gradient all-reduce and the events that order it against the backward pass
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_ev_ready[4] = {}, comm_ev_done = nullptr;

  // optional per-kernel-kind timing (uu_set_profiling): CUDA events around every launch
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_used;   // (kind, (start, stop))
  size_t ev_next = 0;
};

namespace uu {
// helpers defined in uu_api.cu
const float* W(const uu_model* m, const std::string& g, int i);
size_t tensor_offset(const uu_model* m, const std::string& g, int i);   // (size_t)-1 when absent
int dev_alloc(std::vector<void*>& pool, void** p, size_t bytes, bool zero);
void free_pool(std::vector<void*>& pool);
void train_state_destroy(uu_model* m);
// uu_comm.cu: sum all-reduce of grads[lo, hi) on the model's communication stream, ordered after everything enqueued on
// `main` so far (no-op without a communicator); comm_join makes `main` wait for every bucket issued since the last join
int comm_allreduce_range(uu_model* m, size_t lo, size_t hi, int bucket, cudaStream_t main);
int comm_allreduce_scalar(uu_model* m, float* dev_value, cudaStream_t main);
int comm_join(uu_model* m, cudaStream_t main);
void comm_destroy(uu_model* m);
}  // namespace uu
