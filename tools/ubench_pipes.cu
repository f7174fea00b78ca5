// Pipe-throughput microbenchmark for the instruction mix of the fused score kernel (sm_100a).
// Prints warp-instructions per clock per SM for each op, measured with all SMs busy (148 x 4 CTAs x 256 threads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_pipes.cu && /tmp/ubench
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 2048, CHAINS = 8;

enum Op { EX2, RCP, TANH, F2FP, F2FP_BF16, SPLIT_F16, SPLIT_BF16T, HADD2F32, PRMT, LOP, FMNMX, FFMA, FADD, HMMA_F16, HMMA_BF16, HMMA_TF32, LDS32, MIX_SIG, MIX_MUFU_HMMA, MIX_MUFU_FFMA4, FHFMA_OP, HMMA_F16_K8, MIX_HMMA_FFMA4, MIX_HMMA_FFMA8, NOPS };
const char* names[] = {"MUFU.EX2", "MUFU.RCP", "MUFU.TANH", "F2FP.F16.F32.PACK", "F2FP.BF16.F32.PACK", "split f16 (6 ops)", "split bf16 trunc (6 ops)", "HADD2.F32 (h->f)", "PRMT", "LOP3", "FMNMX",
                       "FFMA", "FADD", "HMMA.16816.F32 f16", "HMMA.16816.F32 bf16", "HMMA.1688.F32.TF32", "LDS.32", "EX2+FADD+RCP", "EX2+HMMA", "EX2+4xFFMA", "FHFMA (f32 += f16*f16)", "HMMA.1688.F32 f16", "HMMA+4xFFMA", "HMMA+8xFFMA"};

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, long long* clk, float seed) {
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = seed * i;
  __syncthreads();
  float x[CHAINS];
  uint32_t u[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) { x[c] = seed + c * 0.01f + threadIdx.x * 1e-4f; u[c] = __float_as_uint(x[c]); }
  float acc[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
      if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
      if (OP == TANH) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[c]));
      if (OP == F2FP) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(u[c]) : "f"(x[c])); x[c] = __uint_as_float(u[c]); }
      if (OP == F2FP_BF16) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u[c]) : "f"(x[c])); x[c] = __uint_as_float(u[c]); }
      if (OP == SPLIT_F16) {
        const float e0 = x[c], e1 = x[(c + 1) % CHAINS];
        uint32_t hi, lo;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(e1), "f"(e0));
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(e1 - f.y), "f"(e0 - f.x));
        u[c] ^= hi; x[c] = __uint_as_float(lo);
      }
      if (OP == SPLIT_BF16T) {
        const float e0 = x[c], e1 = x[(c + 1) % CHAINS];
        uint32_t hi, lo;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi) : "r"(__float_as_uint(e0)), "r"(__float_as_uint(e1)));
        const float l0 = e0 - __uint_as_float(__float_as_uint(e0) & 0xffff0000u);
        const float l1 = e1 - __uint_as_float(__float_as_uint(e1) & 0xffff0000u);
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(lo) : "r"(__float_as_uint(l0)), "r"(__float_as_uint(l1)));
        u[c] ^= hi; x[c] = __uint_as_float(lo);
      }
      if (OP == HADD2F32) { asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(x[c]) : "r"(u[c])); u[c] = __float_as_uint(x[c]); }
      if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[c]) : "r"(u[(c + 1) % CHAINS]));
      if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, 0xffff0000, 0x6a;" : "+r"(u[c]) : "r"(u[(c + 1) % CHAINS]));
      if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(x[(c + 1) % CHAINS]));
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[c]) : "f"(seed));
      if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(seed));
      if (OP == HMMA_F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
      if (OP == HMMA_BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
      if (OP == HMMA_TF32)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
      if (OP == LDS32) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[c]) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(u[c] >> 7) & 1023]))); u[c] += __float_as_uint(x[c]); }
      if (OP == MIX_MUFU_HMMA) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
      }
      if (OP == MIX_MUFU_FFMA4) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[c][0]) : "f"(seed));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[c][1]) : "f"(seed));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[c][2]) : "f"(seed));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(acc[c][3]) : "f"(seed));
      }
      if (OP == FHFMA_OP) asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; fma.rn.f32.f16 %0, lo, hi, %0;}" : "+f"(x[c]) : "r"(u[c]));
      if (OP == HMMA_F16_K8)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(u[0]), "r"(u[1]), "r"(u[4]));
      if (OP == MIX_HMMA_FFMA4 || OP == MIX_HMMA_FFMA8) {
        // does the legacy tensor path overlap FMA-pipe issue?  (HMMA alone: 5 clk per sub-partition)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                     : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[c]) : "f"(seed));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[(c + 1) % CHAINS]) : "f"(seed));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[(c + 2) % CHAINS]) : "f"(seed));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[(c + 3) % CHAINS]) : "f"(seed));
        if (OP == MIX_HMMA_FFMA8) {
          asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[c]) : "f"(seed));
          asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[(c + 1) % CHAINS]) : "f"(seed));
          asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[(c + 2) % CHAINS]) : "f"(seed));
          asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[(c + 3) % CHAINS]) : "f"(seed));
        }
      }
      if (OP == MIX_SIG) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
        asm volatile("add.f32 %0, %0, 1.0;" : "+f"(x[c]));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c] + __uint_as_float(u[c]) + acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  out[blockIdx.x * 256 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, int ctas_per_sm) {
  const int grid = sms * ctas_per_sm;
  float* out; long long* clk;
  cudaMalloc(&out, grid * 256 * 4);
  cudaMalloc(&clk, grid * 8);
  k<OP><<<grid, 256>>>(out, clk, 0.37f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<grid, 256>>>(out, clk, 0.37f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[grid];
  cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  const double per_op = (OP == MIX_SIG) ? 3.0 : (OP == MIX_MUFU_HMMA) ? 2.0 : (OP == MIX_MUFU_FFMA4 || OP == MIX_HMMA_FFMA4) ? 5.0 : (OP == MIX_HMMA_FFMA8) ? 9.0 : (OP == SPLIT_F16 || OP == SPLIT_BF16T) ? 6.0 : 1.0;
  const double winst_per_cta = 8.0 * ITERS * CHAINS * per_op;     // 8 warps
  // per-SM rate from the in-kernel clocks (each SM runs ctas_per_sm CTAs concurrently for ~avg clocks)
  printf("%-22s ctas/SM=%d  %.3f warp-inst/clk/SM  (%.1f lanes/clk/SM)  [%.3f ms, %.0f clk]\n", names[OP], ctas_per_sm,
         winst_per_cta * ctas_per_sm / avg, 32.0 * winst_per_cta * ctas_per_sm / avg, ms, avg);
  cudaFree(out); cudaFree(clk); delete[] h;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  for (int c : {2, 4}) {
    run<EX2>(sms, c); run<RCP>(sms, c); run<TANH>(sms, c); run<MIX_SIG>(sms, c); run<F2FP>(sms, c); run<F2FP_BF16>(sms, c); run<SPLIT_F16>(sms, c); run<SPLIT_BF16T>(sms, c); run<HADD2F32>(sms, c);
    run<PRMT>(sms, c); run<LOP>(sms, c); run<FMNMX>(sms, c); run<FFMA>(sms, c); run<FADD>(sms, c);
    run<HMMA_F16>(sms, c); run<HMMA_BF16>(sms, c); run<HMMA_TF32>(sms, c); run<LDS32>(sms, c); run<MIX_MUFU_HMMA>(sms, c); run<MIX_MUFU_FFMA4>(sms, c); run<FHFMA_OP>(sms, c); run<HMMA_F16_K8>(sms, c); run<MIX_HMMA_FFMA4>(sms, c); run<MIX_HMMA_FFMA8>(sms, c);
  }
  return 0;
}
