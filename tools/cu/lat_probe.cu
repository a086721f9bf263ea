// latency probes for the building blocks of the diag kernel (dev aid)
#include <cstdio>
#include <cuda_runtime.h>
__device__ long long prof[16];
__global__ void __launch_bounds__(512,1) probe(double* out, int nb) {
  extern __shared__ double S[];
  const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;
  for (int i = tid; i < 128*132; i += 512) S[i] = 1.0 / (1 + i % 97);
  __syncthreads();
  long long t0 = clock64();
  for (int j = 0; j < nb; ++j) __syncthreads();
  long long t1 = clock64();
  double acc = tid;
  for (int j = 0; j < nb; ++j) { acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2); __syncthreads(); }
  long long t2 = clock64();
  // dot only (no barrier)
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int j = 0; j < nb; ++j) {
    if (r >= j) {
      const double* pr = S + r * 132 + q; const double* pj = S + j * 132 + q;
      int k = 0;
      for (; k + 12 + q < j; k += 16) { a0 = fma(pr[k], pj[k], a0); a1 = fma(pr[k+4], pj[k+4], a1); a2 = fma(pr[k+8], pj[k+8], a2); a3 = fma(pr[k+12], pj[k+12], a3); }
      for (; k + q < j; k += 4) a0 = fma(pr[k], pj[k], a0);
    }
  }
  long long t3 = clock64();
  // rsqrt chain
  double d = 2.0 + tid;
  for (int j = 0; j < nb; ++j) d = rsqrt(d) + 1.5;
  long long t4 = clock64();
  // dependent STS -> barrier -> LDS chain
  double v = tid;
  for (int j = 0; j < nb; ++j) { S[128*132 + (tid & 127)] = v; __syncthreads(); v += S[128*132 + ((tid + 1) & 127)]; __syncthreads(); }
  long long t5 = clock64();
  // dependent DFMA chain
  double f = tid;
  for (int j = 0; j < nb; ++j) { f = fma(f, 1.0000001, 0.5); f = fma(f, 1.0000001, 0.5); f = fma(f, 1.0000001, 0.5); f = fma(f, 1.0000001, 0.5); }
  long long t6 = clock64();
  // dependent LDS chain
  int idx = tid & 127;
  for (int j = 0; j < nb; ++j) { idx = ((int)S[idx] + idx + 1) & 127; }
  long long t7 = clock64();
  out[tid] = acc + a0 + a1 + a2 + a3 + d + v + f + idx;
  if (tid == 0) { prof[0]=t1-t0; prof[1]=t2-t1; prof[2]=t3-t2; prof[3]=t4-t3; prof[4]=t5-t4; prof[5]=t6-t5; prof[6]=t7-t6; }
}
int main() {
  double* out; cudaMalloc(&out, 512*8);
  int smem = (128*132+256)*8;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1,512,smem>>>(out, 128); cudaDeviceSynchronize();
  probe<<<1,512,smem>>>(out, 128); cudaDeviceSynchronize();
  long long p[16]; cudaMemcpyFromSymbol(p, prof, sizeof(p));
  printf("per step (128 steps): barrier %.0f | 2xshfl64+barrier %.0f | dot(avg) %.0f | rsqrt+add %.0f | STS-bar-LDS-bar %.0f | 4 dep DFMA %.0f | dep LDS+cvt %.0f  (%s)\n",
         p[0]/128., p[1]/128., p[2]/128., p[3]/128., p[4]/128., p[5]/128., p[6]/128., cudaGetErrorString(cudaGetLastError()));
  return 0;
}
