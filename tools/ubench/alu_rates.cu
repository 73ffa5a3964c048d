// Development microbenchmark: issue rate of the epilogue's candidate instructions on one SM sub-partition
// (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rates alu_rates.cu).  One block of `warps` warps per SM;
// every thread runs ITER iterations of 8 independent chains of one operation; reports cycles per warp-instruction per
// sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 4096;
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8];
  int ia[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; ia[i] = __float_as_int(a[i]); }
  float b = seed * 1.0001f, c = seed * 0.9999f;
  int ib = __float_as_int(b), ic = __float_as_int(c);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fminf(a[i], b);                       // FMNMX
      if (OP == 1) a[i] = fminf(a[i], fminf(b, c));             // (b,c invariant) -> FMNMX
      if (OP == 2) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));  // FMNMX3
      if (OP == 3) ia[i] = min(ia[i], ib);                      // VIMNMX
      if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(ia[i]) : "r"(ib), "r"(ic));  // LOP3
      if (OP == 5) a[i] = fmaf(a[i], b, c);                     // FFMA
      if (OP == 6) { asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %0, %1, p;}" : "+f"(a[i]) : "f"(b)); }  // FSETP + FSEL
      if (OP == 7) a[i] = a[i] + b;                             // FADD
    }
    b += 1e-7f; ib += 1; // keep the loop from collapsing
  }
  long long t1 = clock64();
  float s = 0; int si = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += a[i]; si += ia[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + si;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024 * 8);
  const char* names[] = {"FMNMX", "FMNMX (invariant pair)", "FMNMX3 (min.f32 3-input)", "VIMNMX", "LOP3", "FFMA", "FSETP+SELP", "FADD"};
  for (int warps : {4, 8, 16}) {
    printf("warps per SM = %d (%d per sub-partition)\n", warps, warps / 4);
    for (int op = 0; op < 8; ++op) {
      auto launch = [&](int o) {
        switch (o) {
          case 0: k<0><<<148, warps * 32>>>(out, cyc, 1.5f); break; case 1: k<1><<<148, warps * 32>>>(out, cyc, 1.5f); break;
          case 2: k<2><<<148, warps * 32>>>(out, cyc, 1.5f); break; case 3: k<3><<<148, warps * 32>>>(out, cyc, 1.5f); break;
          case 4: k<4><<<148, warps * 32>>>(out, cyc, 1.5f); break; case 5: k<5><<<148, warps * 32>>>(out, cyc, 1.5f); break;
          case 6: k<6><<<148, warps * 32>>>(out, cyc, 1.5f); break; case 7: k<7><<<148, warps * 32>>>(out, cyc, 1.5f); break;
        }
      };
      launch(op); cudaDeviceSynchronize(); launch(op); cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      const double winstr_per_smsp = (double)ITER * 8 * (warps / 4.0) * (op == 6 ? 2 : 1);
      printf("  %-28s %8.0f cycles  -> %.2f cycles per warp-instruction per sub-partition\n", names[op], avg, avg / winstr_per_smsp);
    }
  }
  return 0;
}
