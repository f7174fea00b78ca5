// Dependent-issue latency microbenchmark (sm_100a): one warp per SM, a single dependent chain of each op.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_latency tools/ubench_latency.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
enum Op { EX2, RCP, FADD, FMUL, FHFMA, F2FP, FMNMX, HMMA_ACC, HMMA_ACC3, HMMA_A, HMMA_TF32_ACC, LDS, NOPS };
const char* names[] = {"MUFU.EX2", "MUFU.RCP", "FADD", "FMUL", "FHFMA", "F2FP.F16", "FMNMX", "HMMA.16816 (acc chain)", "HMMA.16816 x3 accs round-robin",
                       "HMMA.16816 (D -> A chain)", "HMMA.1688.TF32 (acc chain)", "LDS.32 (pointer chase)"};
template <int OP>
__global__ void k(float* out, long long* clk, float seed) {
  __shared__ uint32_t sm[256];
  for (int i = threadIdx.x; i < 256; i += 32) sm[i] = (uint32_t)__cvta_generic_to_shared(&sm[(i + 33) & 255]);
  __syncwarp();
  float x = seed + threadIdx.x * 1e-4f;
  uint32_t u = __float_as_uint(x), ptr = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x]);
  float a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  uint32_t A[4] = {u, u, u, u};
  long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < ITERS; ++it) {
    if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
    if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x));
    if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(x) : "f"(seed));
    if (OP == FMUL) asm volatile("mul.f32 %0, %0, %1;" : "+f"(x) : "f"(seed));
    if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(seed));
    if (OP == FHFMA) asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; fma.rn.f32.f16 %0, lo, hi, %0;}" : "+f"(x) : "r"(u));
    if (OP == F2FP) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(x)); x = __uint_as_float(u); }
    if (OP == HMMA_ACC)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(a0[0]), "+f"(a0[1]), "+f"(a0[2]), "+f"(a0[3]) : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(u), "r"(u));
    if (OP == HMMA_ACC3) {
      float* a = (it % 3 == 0) ? a0 : (it % 3 == 1) ? a1 : a2;
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]) : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(u), "r"(u));
    }
    if (OP == HMMA_A) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                   : "=f"(a0[0]), "=f"(a0[1]), "=f"(a0[2]), "=f"(a0[3]) : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(u), "r"(u), "f"(0.f));
      A[0] = __float_as_uint(a0[0]); A[1] = __float_as_uint(a0[1]); A[2] = __float_as_uint(a0[2]); A[3] = __float_as_uint(a0[3]);
    }
    if (OP == HMMA_TF32_ACC)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(a0[0]), "+f"(a0[1]), "+f"(a0[2]), "+f"(a0[3]) : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(u), "r"(u));
    if (OP == LDS) asm volatile("ld.shared.u32 %0, [%0];" : "+r"(ptr));
  }
  long long t1 = clock64();
  out[blockIdx.x * 32 + threadIdx.x] = x + __uint_as_float(u) + a0[0] + a0[1] + a0[2] + a0[3] + a1[0] + a2[0] + __uint_as_float(A[0]) + __uint_as_float(ptr);
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP>
void run() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&clk, 148 * 8);
  k<OP><<<148, 32>>>(out, clk, 0.37f);
  k<OP><<<148, 32>>>(out, clk, 0.37f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, 148 * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-34s %.2f clk per dependent op\n", names[OP], avg / ITERS);
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<EX2>(); run<RCP>(); run<FADD>(); run<FMUL>(); run<FHFMA>(); run<F2FP>(); run<FMNMX>(); run<HMMA_ACC>(); run<HMMA_ACC3>(); run<HMMA_A>(); run<HMMA_TF32_ACC>(); run<LDS>();
  return 0;
}
