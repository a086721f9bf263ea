// tile_read_probe.cu -- how much of the HBM stream survives when a column-major m x n FP64 matrix is
// read in row tiles of R rows x ALL n columns (R*8 contiguous bytes per column, column stride m*8)?
// Decides whether a single-pass A = D iteration (tile in shared memory: D*x, prox, D'*r from one
// read of D) is worth building: the tile height is bounded by shared memory (R*n*8 <= ~200 KB).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tile_read_probe tile_read_probe.cu && ./tile_read_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int R>
__global__ void __launch_bounds__(256) tile_read(const double* __restrict__ D, int64_t m, int64_t n, double* out) {
  // persistent CTAs; tile t covers rows [t*R, t*R+R); thread -> (row pair, column lane)
  constexpr int RP = R / 2;            // double2 per column segment
  constexpr int CL = 256 / RP;         // columns in flight per sweep
  const int rp = threadIdx.x % RP, cl = threadIdx.x / RP;
  double acc = 0.0;
  const int64_t ntiles = m / R;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const double* base = D + t * R + 2 * rp;
    for (int64_t j = cl; j < n; j += 4 * CL) {
      double2 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = (j + k * CL < n) ? *reinterpret_cast<const double2*>(base + (j + k * CL) * m) : make_double2(0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += v[k].x + v[k].y;
    }
  }
  if (acc == 123.456) out[0] = acc;
}

template <int R>
float run(const double* D, int64_t m, int64_t n, double* out, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  tile_read<R><<<grid, 256>>>(D, m, n, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) tile_read<R><<<grid, 256>>>(D, m, n, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  struct Shape { int64_t m, n; } shapes[] = {{60000 / 512 * 512, 784}, {1 << 20, 1024}};
  for (auto s : shapes) {
    double *D, *out;
    cudaMalloc(&D, s.m * s.n * 8); cudaMalloc(&out, 8);
    cudaMemset(D, 0, s.m * s.n * 8);
    const double gb = s.m * s.n * 8 / 1e9;
    for (int cps : {2, 4, 8}) {
      printf("m=%lld n=%lld ctas/sm=%d :", (long long)s.m, (long long)s.n, cps);
      printf(" R=8 %.0f", gb / run<8>(D, s.m, s.n, out, cps) * 1e3);
      printf(" R=16 %.0f", gb / run<16>(D, s.m, s.n, out, cps) * 1e3);
      printf(" R=32 %.0f", gb / run<32>(D, s.m, s.n, out, cps) * 1e3);
      printf(" R=64 %.0f", gb / run<64>(D, s.m, s.n, out, cps) * 1e3);
      printf(" R=128 %.0f", gb / run<128>(D, s.m, s.n, out, cps) * 1e3);
      printf(" R=512 %.0f GB/s\n", gb / run<512>(D, s.m, s.n, out, cps) * 1e3);
    }
    cudaFree(D); cudaFree(out);
  }
  return 0;
}
