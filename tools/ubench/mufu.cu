// Micro-benchmark: per-SMSP issue rates of MUFU.EX2 / MUFU.RCP / MUFU.TANH / FFMA with 1 or 2 warps per SMSP (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + 0.001f * (threadIdx.x + 32 * i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(seed), "f"(0.5f));
      if (OP == 5) {   // F2FP.BF16.F32.PACK_AB
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(seed));
        v[i] = __uint_as_float(r | 0x3f000000u);
      }
      if (OP == 6) {   // two EX2 then one pack (the softmax inner loop's mix)
        float a, b; unsigned r;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(v[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(seed));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
        v[i] = __uint_as_float(r & 0x3fffffffu);
      }
      if (OP == 7) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(seed));
      if (OP == 4) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(seed), "f"(0.5f));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(seed), "f"(0.25f));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(seed), "f"(0.125f));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int threads, int blocks) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, sizeof(float) * threads * blocks); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<OP><<<blocks, threads>>>(out, cyc, iters, 0.9f);
  k<OP><<<blocks, threads>>>(out, cyc, iters, 0.9f);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / (double(iters) * 8 * (OP == 4 ? 4 : OP == 6 ? 3 : 1));
  printf("%-14s threads/CTA=%3d blocks=%3d : %.2f clk per warp-instruction (per warp), %.2f clk per SMSP-instruction\n", name,
         threads, blocks, per, per / ((threads / 32 + 3) / 4));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {32, 128, 256}) {
    for (int blocks : {1, 148}) {
      run<0>("MUFU.EX2", threads, blocks);
      run<1>("MUFU.RCP", threads, blocks);
      run<2>("MUFU.TANH", threads, blocks);
      run<3>("FFMA", threads, blocks);
      run<4>("EX2+3FFMA", threads, blocks);
      run<5>("F2FP.BF16x2", threads, blocks);
      run<6>("2EX2+F2FP", threads, blocks);
      run<7>("FMNMX", threads, blocks);
    }
  }
  return 0;
}
