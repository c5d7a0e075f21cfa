// Issue-rate microbenchmark for the epilogue's candidate instruction forms (sm_100a):
// FFMA 3-register, FFMA with an immediate addend, FFMA2 (packed fp32x2) with an
// immediate addend, FMUL2, FFMA.SAT. Prints cycles per warp-instruction per SMSP
// with 1, 2 and 4 resident warps per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 64
#define ITER 256

template <int MODE>
__global__ void k(float* out, long long* cycles, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
        x7 = x0 + 7;
  uint64_t p0, p1, p2, p3, p4, p5, p6, p7, pa;
  asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
  p4 = p0; p5 = p1; p6 = p2; p7 = p3;
  asm("mov.b64 %0, {%1,%2};" : "=l"(pa) : "f"(a), "f"(b));
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int r = 0; r < REP / 8; ++r) {
      if (MODE == 0) {        // FFMA, three registers
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
      } else if (MODE == 1) { // FFMA, immediate addend
        x0 = fmaf(x0, a, 0.123f); x1 = fmaf(x1, a, 0.123f); x2 = fmaf(x2, a, 0.123f); x3 = fmaf(x3, a, 0.123f);
        x4 = fmaf(x4, a, 0.123f); x5 = fmaf(x5, a, 0.123f); x6 = fmaf(x6, a, 0.123f); x7 = fmaf(x7, a, 0.123f);
      } else if (MODE == 2) { // FFMA2, immediate addend
#define F2(p) asm volatile("{.reg .b64 c; mov.b64 c, {0f3DFBE76D, 0f3DFBE76D}; fma.rn.f32x2 %0, %0, %1, c;}" : "+l"(p) : "l"(pa));
        F2(p0) F2(p1) F2(p2) F2(p3) F2(p4) F2(p5) F2(p6) F2(p7)
      } else if (MODE == 3) { // FMUL2
#define M2(p) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(pa));
        M2(p0) M2(p1) M2(p2) M2(p3) M2(p4) M2(p5) M2(p6) M2(p7)
      } else if (MODE == 4) { // FFMA.SAT
#define S1(x) asm volatile("fma.rn.sat.f32 %0, %0, %1, 0f3F000000;" : "+f"(x) : "f"(a));
        S1(x0) S1(x1) S1(x2) S1(x3) S1(x4) S1(x5) S1(x6) S1(x7)
      } else if (MODE == 5) { // FMUL, two registers
        x0 *= a; x1 *= a; x2 *= a; x3 *= a; x4 *= a; x5 *= a; x6 *= a; x7 *= a;
      } else if (MODE == 6) { // FFMA2, three 64-bit registers
#define G2(p) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(pa));
        G2(p0) G2(p1) G2(p2) G2(p3) G2(p4) G2(p5) G2(p6) G2(p7)
      }
    }
  }
  const long long t1 = clock64();
  float lo, hi, acc = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0 ^ p1 ^ p2 ^ p3 ^ p4 ^ p5 ^ p6 ^ p7));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + lo + hi;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 << 20); cudaMalloc(&cyc, 8);
  printf("%-28s", name);
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 32 * 4 * warps_per_smsp;
    k<MODE><<<1, threads>>>(out, cyc, 1.0001f, 0.5f);
    k<MODE><<<1, threads>>>(out, cyc, 1.0001f, 0.5f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / ((double)ITER * REP * warps_per_smsp);
    printf("  %dw/smsp: %.2f cyc/warp-instr", warps_per_smsp, per);
  }
  printf("\n");
}

int main() {
  run<0>("FFMA r,r,r");
  run<1>("FFMA r,r,imm");
  run<5>("FMUL r,r");
  run<4>("FFMA.SAT r,r,imm");
  run<2>("FFMA2 rr,rr,imm");
  run<6>("FFMA2 rr,rr,rr");
  run<3>("FMUL2 rr,rr");
  return cudaGetLastError() != cudaSuccess;
}
