// Measures per-SM issue throughput of the instruction classes the exact modular MAC can be built from
// (IMAD.WIDE.U32, IMAD 32-bit, DFMA, DADD, I2F.F64.U32, FFMA) on the current GPU.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pipe_bench.cu -o build/pipe_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
constexpr int ITERS = 4096, CH = 8;

template <int OP>
__global__ void k(u64* out, u32 seed) {
  u32 a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u;
  u64 acc[CH]; double d[CH]; float f[CH]; u32 w[CH];
  for (int c = 0; c < CH; ++c) { acc[c] = c + a; d[c] = (double)(c + (a & 0xff)); f[c] = (float)c; w[c] = c + b; }
  double da = (double)(a & 0xffff), db = (double)(b & 0xffff);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
      if (OP == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(w[c]) : "r"(a), "r"(b));
      if (OP == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[c]) : "d"(da), "d"(db));
      if (OP == 3) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(da));
      if (OP == 4) asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d[c]) : "r"(w[c] + i));
      if (OP == 5) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[c]) : "f"((float)da), "f"((float)db));
      if (OP == 6) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
                     asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[c]) : "d"(da), "d"(db)); }
      if (OP == 7) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(w[c]) : "r"(a + i), "r"(b));
      if (OP == 8) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[c]) : "f"(f[c] + (float)i));
      if (OP == 9) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[c]) : "f"(__uint_as_float(w[c] + i)));
                     asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(da) : "d"(db), "d"(db)); }
    }
  }
  u64 s = 0;
  for (int c = 0; c < CH; ++c) s += acc[c] + (u64)d[c] + (u64)f[c] + w[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// shared-memory read bandwidth: every thread reads VEC-byte words at consecutive addresses (conflict-free)
template <int VEC>
__global__ void ksmem(u64* out) {
  __shared__ __align__(16) u64 buf[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i;
  __syncthreads();
  u64 s = 0;
  const int words = VEC / 8;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int idx = ((threadIdx.x + (i + c) * 32) * words) & 4095;
      const unsigned addr = (unsigned)__cvta_generic_to_shared(&buf[VEC == 16 ? (idx & ~1) : idx]);
      if (VEC == 16) { u64 x, y; asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(addr)); s += x ^ y; }
      else { u64 x; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(x) : "r"(addr)); s += x; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int VEC>
void run_smem(const char* name, int sms) {
  u64* out; cudaMalloc(&out, sizeof(u64) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  ksmem<VEC><<<sms * 2, 1024>>>(out);
  cudaEventRecord(e0);
  ksmem<VEC><<<sms * 2, 1024>>>(out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double bytes = (double)sms * 2 * 1024 * ITERS * CH * VEC;
  printf("%-28s %8.3f ms  %7.1f B/clk/SM\n", name, ms, bytes / (ms * 1e-3) / (clk * 1e3) / sms);
  cudaFree(out);
}
template <int OP>
void run(const char* name, int sms, double mult) {
  u64* out; cudaMalloc(&out, sizeof(u64) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<sms * 2, 1024>>>(out, 1);
  cudaEventRecord(e0);
  k<OP><<<sms * 2, 1024>>>(out, 2);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = (double)sms * 2 * 1024 * ITERS * CH * mult;
  printf("%-28s %8.3f ms  %7.1f ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000);
  cudaFree(out);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>("mad.wide.u32 (IMAD.WIDE)", sms, 1);
  run<1>("mad.lo.u32 (IMAD)", sms, 1);
  run<7>("mul.hi.u32 (IMAD.HI)", sms, 1);
  run<2>("fma.f64 (DFMA)", sms, 1);
  run<3>("add.f64 (DADD)", sms, 1);
  run<4>("cvt.f64.u32 (I2F)", sms, 1);
  run<5>("fma.f32 (FFMA)", sms, 1);
  run<6>("IMAD.WIDE + DFMA interleaved", sms, 2);
  run<8>("cvt.f64.f32 (F2F)", sms, 1);
  run<9>("F2F + DFMA interleaved", sms, 2);
  run_smem<8>("LDS.64 conflict-free", sms);
  run_smem<16>("LDS.128 conflict-free", sms);
  return 0;
}
