// Probe: can a 128-thread x 32-register block start on an SM that already holds three 160-thread x 128-register CTAs
// with 62 KB of dynamic shared memory each (the fused kernel's footprint)?  Prints when the small kernel finishes
// relative to the hog kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 coresidency.cu -o coresidency
#include <cstdio>
#include <cuda_runtime.h>

template<int NREG>
__global__ void __launch_bounds__(160, 3) hog(long long cycles, double* out) {
  extern __shared__ double sm[];
  double acc[56];
#pragma unroll
  for(int i = 0; i < 56; i++) acc[i] = threadIdx.x + i;
  const long long t0 = clock64();
  while(clock64() - t0 < cycles) {
#pragma unroll
    for(int i = 0; i < 56; i++) acc[i] = fma(acc[i], 1.0000001, 0.5);
  }
  double s = 0;
#pragma unroll
  for(int i = 0; i < 56; i++) s += acc[i];
  sm[threadIdx.x] = s;
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

__global__ void __launch_bounds__(128, 16) small(double* out, int n) {
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i * 0.5;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int smem = 62464;
  cudaFuncSetAttribute(hog<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, hog<128>);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hog<128>, 160, smem);
  printf("hog: %d regs, occupancy %d CTAs/SM; SMs %d\n", fa.numRegs, occ, p.multiProcessorCount);
  cudaFuncGetAttributes(&fa, small);
  printf("small: %d regs\n", fa.numRegs);
  double *a, *b;
  cudaMalloc(&a, 8 * 160 * 1024);
  cudaMalloc(&b, 8 << 20);
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  int lo, hi;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, hi);
  cudaEvent_t h0, h1, k0, k1;
  cudaEventCreate(&h0), cudaEventCreate(&h1), cudaEventCreate(&k0), cudaEventCreate(&k1);
  const int smems[3] = {62464, 57344, 49152};
  const int blocks[3] = {128, 64, 32};
  for(int si = 0; si < 3; si++)
    for(int bi = 0; bi < 3; bi++) {
      const int sm_ = smems[si], bt = blocks[bi];
      cudaEventRecord(h0, s1);
      hog<128><<<p.multiProcessorCount * 3, 160, sm_, s1>>>(10000000ll, a); // ~5 ms
      cudaEventRecord(h1, s1);
      cudaEventRecord(k0, s2);
      small<<<592, bt, 0, s2>>>(b, 1 << 20);
      cudaEventRecord(k1, s2);
      cudaDeviceSynchronize();
      float th, tk1;
      cudaEventElapsedTime(&th, h0, h1);
      cudaEventElapsedTime(&tk1, h0, k1);
      printf("hog 3 CTAs/SM, %d B smem each; small block %d threads: finished %.3f ms of %.2f -> %s\n", sm_, bt, tk1, th,
             tk1 < 0.8 * th ? "CO-RESIDENT" : "waited");
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
