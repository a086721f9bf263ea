// Times potrf_diag_kernel in isolation (dev aid).  nvcc -arch=sm_100a -O3 -I../../admm_project_b200/csrc
#include <cstdio>
#include <vector>
#include <cmath>
#define CHOL_PROFILE 1
#include "chol.cuh"
namespace admmb200 { void set_error(const char*, ...) {} }
using namespace admmb200;
int main() {
  const int n = 8192, nb = 128;
  std::vector<double> h((size_t)nb * n, 0.0);
  for (int c = 0; c < nb; ++c) for (int r = 0; r < nb; ++r) h[r + (size_t)c * n] = (r == c) ? nb + 1.0 : 1.0 / (1 + abs(r - c));
  double *A, *X; int* fail;
  cudaMalloc(&A, (size_t)nb * n * 8); cudaMalloc(&X, (size_t)nb * n * 8); cudaMalloc(&fail, 4);
  cudaMemset(fail, 0, 4);
  cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_DIAG_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(A, h.data(), (size_t)nb * n * 8, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    potrf_diag_kernel<<<1, CHOL_DIAG_THREADS, CHOL_DIAG_SMEM>>>(A, n, nb, X, n, fail, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("diag kernel: %.1f us (%s)\n", ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  // back-to-back 10 launches
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) potrf_diag_kernel<<<1, CHOL_DIAG_THREADS, CHOL_DIAG_SMEM>>>(A, n, nb, X, n, fail, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("10 launches: %.1f us each\n", ms * 100);
  long long pr[8]; cudaMemcpyFromSymbol(pr, chol_prof, sizeof(pr));
  printf("cycles: load %lld  chol %lld  storeL %lld  inv %lld  storeX %lld\n", pr[1]-pr[0], pr[2]-pr[1], pr[3]-pr[2], pr[4]-pr[3], pr[5]-pr[4]);
  int f; cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost); printf("fail=%d\n", f);
  return 0;
}
