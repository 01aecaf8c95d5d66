// Do MUFU.EX2 and F2FP.BF16.PACK_AB share an issue pipe on sm_100?  Times per-SMSP warp-instruction costs of each alone
// and interleaved 4:2 (the ratio of the attention softmax: 64 exponentials and 32 packs per 64 scores).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (MODE & 1) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    }
    if (MODE & 2) {
      uint32_t p0, p1;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(a), "f"(b));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(c), "f"(d));
      acc ^= p0 ^ p1;
    }
    if (MODE & 4) {   // truncating pack on the integer pipe
      uint32_t p0, p1;
      asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(p0) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
      asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(p1) : "r"(__float_as_uint(c)), "r"(__float_as_uint(d)));
      acc ^= p0 ^ p1;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(acc);
}

template <int MODE>
void run(const char* name, float* out) {
  const int iters = 20000, warps = 16;
  k<MODE><<<148, warps * 32>>>(out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // 4 warps per SMSP, iterations back to back: clocks per iteration per SMSP at ~1.9 GHz
  printf("%-40s %.3f ms  -> %.1f ns per iteration per warp-slot (4 warps/SMSP: x4 per SMSP)\n", name, ms, ms * 1e6 / iters);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 16 * 32 * 4);
  run<1>("4 x MUFU.EX2", out);
  run<2>("2 x F2FP pack", out);
  run<3>("4 x MUFU.EX2 + 2 x F2FP pack", out);
  run<4>("2 x PRMT truncating pack", out);
  run<5>("4 x MUFU.EX2 + 2 x PRMT pack", out);
  return 0;
}
