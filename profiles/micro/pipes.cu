// Micro-benchmark: per-SM-sub-partition issue rate of the instructions the likelihood epilogue is made of.
// One CTA per SM, WARPS warps per sub-partition, 8 independent dependency chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 2048
template <int OP>
__global__ void k(float* out, long long* clk, float seed) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 1e-3f + i;
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(x[i]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = fmaf(x[i], 1.0001f, 0.5f);
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 2) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i])); x[i] += 1.5f; }   // + FADD (keeps ptxas from folding rcp(rcp))
      if (OP == 3) asm volatile("{.reg .b32 t; cvt.rn.f16x2.f32 t, %0, %0; mov.b32 %0, t;}" : "+f"(x[i]));
      if (OP == 4) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; cvt.f32.f16 %0, lo;}" : "+f"(x[i]));
      if (OP == 5) { if (it & 1) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(seed)); else asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(-seed)); }
      if (OP == 6) u[i] = (u[i] ^ u[(i + 1) & 7]) & 0xffffe000u;
      if (OP == 7) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 8) {  // MUFU + 3 FMA mix: does MUFU co-issue under FMA load?
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        x[i] = fmaf(x[i], 1.0001f, 0.5f); x[i] = fmaf(x[i], 0.999f, 0.25f); x[i] = fmaf(x[i], 1.0002f, 0.125f);
      }
      if (OP == 9) {  // F2FP + 3 FMA
        asm volatile("{.reg .b32 t; cvt.rn.f16x2.f32 t, %0, %0; mov.b32 %0, t;}" : "+f"(x[i]));
        x[i] = fmaf(x[i], 1.0001f, 0.5f); x[i] = fmaf(x[i], 0.999f, 0.25f); x[i] = fmaf(x[i], 1.0002f, 0.125f);
      }
      if (OP == 10) {  // MUFU + F2FP: same pipe?
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("{.reg .b32 t; cvt.rn.f16x2.f32 t, %0, %0; mov.b32 %0, t;}" : "+f"(x[i]));
      }
      if (OP == 11) {  // MUFU + min (ALU)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("min.f32 %0, %0, 30.0;" : "+f"(x[i]));
      }
      if (OP == 12) {  // cvt f16->f32 + FMA
        asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; cvt.f32.f16 %0, lo;}" : "+f"(x[i]));
        x[i] = fmaf(x[i], 1.0001f, 0.5f);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int OP>
void run(const char* name, int nop, int warps_per_smsp) {
  float* out; long long* clk;
  const int threads = 128 * warps_per_smsp;
  cudaMalloc(&out, 148 * threads * 4); cudaMalloc(&clk, 8);
  k<OP><<<148, threads>>>(out, clk, 1.0f);
  k<OP><<<148, threads>>>(out, clk, 1.0f);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double inst = (double)ITERS * 8 * nop * warps_per_smsp;   // warp-instructions per sub-partition
  printf("%-28s warps/SMSP %d: %.3f warp-inst/clk/SMSP  (%.2f clk per warp-inst)\n", name, warps_per_smsp, inst / h, h / inst);
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {1, 4}) {
    run<0>("FFMA", 1, w); run<1>("MUFU.EX2", 1, w); run<2>("MUFU.RCP + FADD", 2, w); run<7>("MUFU.LG2", 1, w);
    run<3>("F2FP.F16.F32.PACK_AB", 1, w); run<4>("HADD2.F32 (f16->f32)", 1, w); run<5>("FMNMX", 1, w); run<6>("LOP3", 1, w);
    run<8>("EX2 + 3 FFMA", 4, w); run<9>("F2FP + 3 FFMA", 4, w); run<10>("EX2 + F2FP", 2, w); run<11>("EX2 + FMNMX", 2, w);
    run<12>("HADD2.F32 + FFMA", 2, w);
  }
  return 0;
}
