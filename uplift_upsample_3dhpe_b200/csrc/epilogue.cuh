// Fused GEMM epilogue shared by the SIMT (fp32) and tcgen05 (bf16) GEMMs:
//   C[crow][col] = act(acc + bias[col]) (+ Res[rrow][col]) (+ table[crow % period][col])
// crow comes from a scatter list (compact valid-frame rows -> token rows) or a RowMap
// (zero-padded conv input layout); rrow from a RowMap (strided identity path, net:146-152).
#pragma once
#include "common.cuh"

namespace uu {

struct EpiRow {
  long long crow;   // physical output row, -1 = dropped
  long long rrow;   // physical residual row (valid when EPI_RESIDUAL)
  int trow;         // row of the periodic table
};

__device__ __forceinline__ EpiRow epi_row(const Epilogue& e, int r) {
  EpiRow o;
  o.crow = e.c_rowidx ? (long long)e.c_rowidx[r] : map_row(e.cmap, r);
  o.rrow = (e.flags & EPI_RESIDUAL) ? map_row(e.rmap, r) : 0;
  o.trow = (e.flags & EPI_ROWTABLE) && o.crow >= 0 ? (int)(o.crow % e.table_period) : 0;
  return o;
}

__device__ __forceinline__ float epi_value(const Epilogue& e, const EpiRow& row, float acc, int col, int N) {
  float v = acc;
  if (e.bias) v += e.bias[col];
  if (e.flags & EPI_RELU) v = fmaxf(v, 0.f);
  if (e.flags & EPI_RESIDUAL) v += e.res[row.rrow * e.ldr + col];
  if (e.flags & EPI_ROWTABLE) v += e.table[(long long)row.trow * N + col];
  return v;
}

__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

}  // namespace uu
