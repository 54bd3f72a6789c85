// Declarations of the training-step kernels (train_kernels.cu).
#pragma once
#include "common.cuh"

namespace uu {

struct GemmGen {
  const float* A = nullptr; long long lda = 0; int transA = 0; RowMap amap;   // amap only when !transA
  const float* B = nullptr; long long ldb = 0; int transB = 0;
  float* C = nullptr; long long ldc = 0; RowMap cmap;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;   // forward only
  int relu = 0;                  // forward only
  int accumulate = 0;            // C += ...
  int split_k = 1;               // > 1: split-K partial slabs, reduced in split order (requires accumulate, plain C)
  float* partial = nullptr;      // set by launch_gemm_gen
};
// scratch for the deterministic two-pass reductions of this translation unit (set before launching a step)
void train_reduce_scratch(float* p, size_t n_floats);
size_t train_reduce_scratch_floats();
// deferred second passes (the backward pass of a training step): producers take fresh regions of `arena`, their reductions are
// recorded and run by train_reduce_flush in ONE launch; train_reduce_defer_end flushes and returns to the immediate form
void train_reduce_defer_begin(float* arena, size_t arena_floats);
cudaError_t train_reduce_flush(cudaStream_t st);
cudaError_t train_reduce_defer_end(cudaStream_t st);
void train_reduce_defer_abort();      // error paths: drop what was recorded, back to the immediate form
cudaError_t launch_gemm_gen(const GemmGen& g, cudaStream_t st);
void gemm_gen_tile(int N, int* bm, int* bn);
bool wgrad_skinny_ok(const float* X, long long ldx, const float* dY, long long ldy, long long rows, int K, int N);
cudaError_t launch_wgrad_skinny(const float* X, long long ldx, const float* dY, long long ldy, long long rows, int K, int N,
                                float* dW, float* db, cudaStream_t st);      // tile shape launch_gemm_gen picks for an N-wide output
bool dgrad_skinny_ok(const float* dY, long long ldy, long long rows, int KI, int NO, const float* dX, long long ldx);
cudaError_t launch_dgrad_skinny(const float* dY, long long ldy, const float* Wm, long long rows, int KI, int NO, float* dX,
                                long long ldx, int accumulate, cudaStream_t st);      // dX (+)= dY W^T, W (NO, KI)
cudaError_t launch_colsum(const float* Y, int M, int N, long long ld, float* out, cudaStream_t st);
cudaError_t launch_period_sum(const float* X, long long rows, int period, int d, const uint8_t* rowmask, int want,
                              float* out, cudaStream_t st);
cudaError_t launch_ln_fwd_gen(const float* x, long long rows, int d, const float* gamma, const float* beta, float eps,
                              float* y, cudaStream_t st);
cudaError_t launch_ln_bwd_gen(const float* x, const float* dy, long long rows, int d, const float* gamma, float eps,
                              float* dx, int accumulate, float* dgamma, float* dbeta, cudaStream_t st);
bool attention_small_ok(int S, int heads, int dh, const uint8_t* mask);
cudaError_t launch_attention_small_fwd(const float* qkv, long long frames, int S, float* out, cudaStream_t st);
cudaError_t launch_attention_bwd(const float* qkv, const float* dO, long long B, int S, int heads, int dh,
                                 const uint8_t* mask, int mask_stride, float* dqkv, cudaStream_t st);
// tensor-core attention of the training step (attn_mma.cu): S <= 80, head dimension 32 / 48 / 64; nsplit 3 = compensated
// TF32 (fp32-grade), 1 = plain TF32
bool attention_mma_ok(long long B, int S, int heads, int dh);
cudaError_t launch_attention_mma_fwd(const float* qkv, long long B, int S, int heads, int dh, const uint8_t* mask,
                                     int mask_stride, float* out, int nsplit, cudaStream_t st);
cudaError_t launch_attention_mma_bwd(const float* qkv, const float* dO, long long B, int S, int heads, int dh,
                                     const uint8_t* mask, int mask_stride, float* dqkv, int nsplit, cudaStream_t st);
// one spatial transformer block forward with its tape in a single launch (spatial_train.cu): 17 joints, d = 32, hidden 64,
// 8 heads, GELU
bool spatial_block_fused_ok(int J, int d, int h, int heads, int act);
cudaError_t launch_spatial_block_fwd_tape(const float* x0, const float* const* weights, const float* scale, const float* scale2,
                                          long long frames, float* y1, float* qkv, float* o, float* x1, float* y2, float* hpre,
                                          float* hact, float* x2, int num_sms, cudaStream_t st);
cudaError_t launch_act_fwd(const float* pre, long long n, int act, float* out, cudaStream_t st);
cudaError_t launch_act_bwd(const float* pre, const float* dout, const RowMap& dmap, long long ldd, long long rows,
                           int cols, int act, float* dpre, cudaStream_t st);
cudaError_t launch_act_bwd_mapped(const float* hp, const float* dhp, const RowMap& map, long long ld, long long rows,
                                  int cols, float* dpre, cudaStream_t st);
cudaError_t launch_residual(const float* base, const RowMap& bmap, const float* y, const float* scale,
                            int rows_per_sample, const float* table, int period, long long rows, int d, float* out,
                            cudaStream_t st);
cudaError_t launch_token_mask_draw(unsigned long long seed, unsigned long long stream, long long rows, int n_tok, float rate,
                                   float* keep, cudaStream_t st);
cudaError_t launch_scale_rows(const float* src, const float* scale, int rps, long long rows, int d, float* dst,
                              cudaStream_t st);
cudaError_t launch_scatter_add(const float* src, const RowMap& dmap, long long rows, int d, float* dst, cudaStream_t st);
cudaError_t launch_embed_fwd(const float* x2d, const uint8_t* mask, const int* list, int J, long long rows, int d,
                             const float* Wk, const float* b, const float* pe, float* out, cudaStream_t st);
cudaError_t launch_embed_wgrad(const float* x2d, const uint8_t* mask, const int* list, int J, const float* de, long long rows,
                               int d, float* dW, cudaStream_t st);
cudaError_t launch_fill_fwd(const float* s, const uint8_t* mask, const int* pos, const float* keep, const float* token,
                            const float* pe, int n_tok, long long rows, int d, float* x, cudaStream_t st);
cudaError_t launch_fill_bwd(const float* dx, const uint8_t* mask, const int* list, const float* keep, long long rows, int d,
                            float* ds, cudaStream_t st);
cudaError_t launch_invert_list(const int* list, int n, int* pos, cudaStream_t st);
cudaError_t launch_gather_f32(const float* src, const int* list, int n, float* dst, cudaStream_t st);
int loss_blocks(int B, int n_tok, int J, bool has_full);
cudaError_t launch_loss(const float* full, const float* central, const float* gt, int B, int n_tok, int J, int root,
                        float w_seq, float w_cen, float* dfull, float* dcentral, float* partials, float* loss,
                        cudaStream_t st);
cudaError_t launch_droppath_scale(unsigned long long seed, unsigned long long stream, long long n, float keep,
                                  float* scale, cudaStream_t st);
cudaError_t launch_adamw(float* p, float* m, float* v, const float* g, long long n, float wd, float alpha, float b1,
                         float b2, float eps, float* ema, float ema_decay, cudaStream_t st);


// ---- tcgen05 kind::tf32 weight gradients (wgrad_tc.cu): dW [Kd, Nd] (+)= X^T dY, deterministic split-K ------------
size_t wgrad_tc_scratch_bytes();
bool wgrad_tc_ok(const float* X, long long ldx, const float* dY, long long ldy, long long R, int Kd, int Nd);
int wgrad_tc(const float* X, long long ldx, const float* dY, long long ldy, long long R, int Kd, int Nd, float* dW,
             int accumulate, float* scratch, int num_sms, cudaStream_t st);

}  // namespace uu
