// Micro-benchmark of the packed 16-bit integer instructions the SGM kernels are built from (sm_100a).
// Measures sustained warp-instructions per clock per SM for independent-chain instruction streams.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench tools/ubench_int16x2.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define CHAINS 8

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
  if (OP == 0) return __vminu2(a, b);
  if (OP == 1) return __viaddmin_u16x2(a, b, c);
  if (OP == 2) return __vimin3_u16x2(a, b, c);
  if (OP == 3) return __byte_perm(a, b, 0x5432);
  if (OP == 4) return __vadd2(a, b);
  if (OP == 5) return a + b;
  if (OP == 6) return __shfl_up_sync(0xffffffffu, a, 1) ^ b;   // SHFL + LOP3
  if (OP == 7) return __reduce_min_sync(0xffffffffu, a) + b;   // CREDUX + IADD
  if (OP == 8) return __vimax_s16x2_relu(a, b);
  if (OP == 9) return min(a, b);
  if (OP == 10) return __vsub2(a, b);
  if (OP == 11) return a * 3u + b;                              // IMAD (fma pipe)
  if (OP == 12) return (a & b) | c;                             // LOP3
  return a;
}

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) v[i] = seed * (threadIdx.x + 1) + i * 0x10001u;
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = op<OP>(v[i], v[(i + 1) % CHAINS], v[(i + 3) % CHAINS]);
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// two-op mix: one ALU-pipe op and one that may go to the FMA pipe
template <int OPA, int OPB>
__global__ void kmix(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t v[CHAINS], w[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { v[i] = seed * (threadIdx.x + 1) + i * 0x10001u; w[i] = v[i] ^ 0x1234u; }
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { v[i] = op<OPA>(v[i], v[(i + 1) % CHAINS], w[(i + 3) % CHAINS]); w[i] = op<OPB>(w[i], w[(i + 1) % CHAINS], v[(i + 5) % CHAINS]); }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s ^= v[i] ^ w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename K>
void run(const char* name, K kern, int instrPerIter) {
  uint32_t* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int threads = 1024;  // 32 warps per SM, 8 per SMSP
  kern<<<148, threads>>>(out, 12345u, cyc);
  kern<<<148, threads>>>(out, 12345u, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double warpInstr = (double)ITER * CHAINS * instrPerIter * (threads / 32);
  printf("%-34s %8.3f warp-instr/clk/SM  (%lld cycles)\n", name, warpInstr / (double)h, h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run("VIMNMX.U16x2 (__vminu2)", k<0>, 1);
  run("VIADDMNMX.U16x2 (__viaddmin)", k<1>, 1);
  run("VIMNMX3.U16x2 (__vimin3)", k<2>, 1);
  run("PRMT (__byte_perm)", k<3>, 1);
  run("VIADD.16x2 (__vadd2)", k<4>, 1);
  run("IADD (a+b)", k<5>, 1);
  run("SHFL.UP + LOP3", k<6>, 2);
  run("CREDUX.MIN + IADD", k<7>, 2);
  run("VIMNMX.S16x2.RELU", k<8>, 1);
  run("IMNMX.U32 (min)", k<9>, 1);
  run("__vsub2", k<10>, 1);
  run("IMAD", k<11>, 1);
  run("LOP3", k<12>, 1);
  run("mix VMIN16x2 + IMAD", kmix<0, 11>, 2);
  run("mix VMIN16x2 + IADD", kmix<0, 5>, 2);
  run("mix VMIN16x2 + PRMT", kmix<0, 3>, 2);
  run("mix VIADDMNMX + IMAD", kmix<1, 11>, 2);
  run("mix VMIN16x2 + VIADD.16x2", kmix<0, 4>, 2);
  return 0;
}
