// Development harness: layout experiments for the MN-major tcgen05 wgrad kernel.  nvcc -DWG_DEBUG ... && run on the GPU.
#define WG_DEBUG 1
#include <vector>
#include <cstdio>
#include <string>
#include "../../uplift_upsample_3dhpe_b200/csrc/wgrad_tc.cu"
namespace uu { void set_error(const std::string& m) { printf("ERR: %s\n", m.c_str()); } }
using namespace uu;

__global__ void k_dump(const __grid_constant__ CUtensorMap map, int row, int grp, float* out, int bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_3d(smem, &map, &bar, 0, row, grp); }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main() {
  const int R = 512, Kd = 384, Nd = 384;
  std::vector<float> X(R * Kd), Y(R * Nd), W(Kd * Nd);
  float *dX, *dY, *dW, *scr, *dump;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dY, Y.size() * 4); cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&scr, wgrad_tc_scratch_bytes()); cudaMalloc(&dump, 64 << 10);
  // E2: where does X[r][c] land in the stage?
  for (int i = 0; i < R * Kd; ++i) X[i] = (float)((i / Kd) * 1000 + (i % Kd));    // value = row*1000 + col
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap mx;
  if (encode_3d_f32(&mx, dX, Kd, R, Kd, WG_BKR, 4)) return 1;
  cudaFuncSetAttribute(k_dump, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 << 10);
  k_dump<<<1, 128, 18 << 10>>>(mx, 0, 0, dump, 16384);
  std::vector<float> h(4096);
  cudaMemcpy(h.data(), dump, 16384, cudaMemcpyDeviceToHost);
  printf("dump err=%s\n", cudaGetErrorString(cudaGetLastError()));
  for (int i = 0; i < 4096; i += 4) if (i < 80 || (i % 1024) < 8) printf("  smem float %d: %.0f %.0f %.0f %.0f\n", i, h[i], h[i + 1], h[i + 2], h[i + 3]);
  // dense random data, descriptor / instruction-descriptor variants
  {
    const int R2 = 256;
    srand(1);
    for (auto& v : X) v = (float)((rand() % 17) - 8);
    for (auto& v : Y) v = (float)((rand() % 13) - 6);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice);
    std::vector<double> want((size_t)Kd * Nd, 0.0);
    for (int r = 0; r < R2; ++r)
      for (int c = 0; c < Kd; ++c) { const double x = X[r * Kd + c]; for (int n = 0; n < Nd; ++n) want[(size_t)c * Nd + n] += x * Y[r * Nd + n]; }
    const uint32_t variants[][3] = {{4096, 512, 0}, {4096, 1024, 0}, {512, 4096, 0}, {4096, 256, 0}};
    for (auto& v : variants) {
      cudaMemcpyToSymbol(g_wg_lbo, &v[0], 4); cudaMemcpyToSymbol(g_wg_sbo, &v[1], 4); cudaMemcpyToSymbol(g_wg_idesc_xor, &v[2], 4);
      cudaMemset(scr, 0xff, 4 << 20);
      if (wgrad_tc(dX, Kd, dY, Nd, R2, Kd, Nd, dW, 0, scr, 148, 0)) return 1;
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(W.data(), dW, W.size() * 4, cudaMemcpyDeviceToHost);
      double sum = 0, err = 0; int nz = 0;
      for (int i = 0; i < Kd * Nd; ++i) { sum += fabs(W[i]); err = fmax(err, fabs(W[i] - want[i])); nz += W[i] != 0.f; }
      printf("lbo=%u sbo=%u idesc^=%x: %s  sum|W|=%g nnz=%d maxerr=%g  W[0][0..3]=%g %g %g %g want %g %g %g %g\n", v[0], v[1], v[2],
             cudaGetErrorString(e), sum, nz, err, W[0], W[1], W[2], W[3], want[0], want[1], want[2], want[3]);
    }
  }
  return 0;
}
