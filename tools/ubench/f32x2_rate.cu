// Issue-rate microbenchmark for the packed FP32 instructions (FFMA2 / FADD2) against scalar FFMA on sm_100a:
// N independent accumulator chains per thread, `warps` warps per SM, clock64 around a long unrolled loop.
// Prints warp-instructions per clock per SM and the equivalent scalar-FMA lanes per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o audio_b200/build/f32x2_rate tools/ubench/f32x2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr int CH = 8, ITERS = 512;

template <int MODE>
__global__ void rate(float* out, long long* cycles, float s0, float s1) {
  float2 seed = make_float2(threadIdx.x * 1e-3f + s0, s1);
  uint64_t acc[CH];
  float accs[2 * CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { acc[i] = pk(seed.x + i, seed.y - i); accs[2 * i] = seed.x + i; accs[2 * i + 1] = seed.y - i; }
  const uint64_t m = pk(s0, s1), a = pk(s1, s0);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) {  // scalar FFMA, 2 per chain slot (same flops as one FFMA2)
        accs[2 * i] = fma1(accs[2 * i], s0, s1);
        accs[2 * i + 1] = fma1(accs[2 * i + 1], s1, s0);
      } else if (MODE == 1) {  // FFMA2, three pair operands
        acc[i] = fma2(acc[i], m, a);
      } else if (MODE == 2) {  // FFMA2 with a scalar-broadcast operand
        acc[i] = fma2(acc[i], pk(s0, s0), a);
      } else {  // FADD2
        acc[i] = add2(acc[i], m);
      }
    }
  }
  const long long t1 = clock64();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    r += lo + hi + accs[2 * i] + accs[2 * i + 1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  float* out;
  long long* cyc;
  const int blocks = 148;
  cudaMalloc(&out, sizeof(float) * blocks * warps * 32);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  rate<MODE><<<blocks, warps * 32>>>(out, cyc, 1.0001f, 0.9999f);
  rate<MODE><<<blocks, warps * 32>>>(out, cyc, 1.0001f, 0.9999f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += (double)h[i];
  avg /= blocks;
  const double inst = (double)ITERS * CH * (MODE == 0 ? 2 : 1) * warps;  // warp-instructions per SM
  printf("%-22s warps=%2d  cycles=%9.0f  warp-inst/clk/SM=%5.2f  fp32-lane-ops/clk/SM=%6.1f\n", name, warps, avg, inst / avg,
         inst / avg * 32 * (MODE == 0 ? 1 : 2));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int warps : {4, 8, 16, 32}) {
    run<0>("FFMA (scalar)", warps);
    run<1>("FFMA2 pair,pair,pair", warps);
    run<2>("FFMA2 pair,splat,pair", warps);
    run<3>("FADD2", warps);
  }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
