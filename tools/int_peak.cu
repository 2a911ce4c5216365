// Microbenchmark: per-SM issue rate of VABSDIFF4.U8.ACC (packed 4-byte SAD), IADD3, IMAD, LDS.32
// on sm_100a.  Register-only loops; result = lane-ops / clk / SM and chip-wide ops/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

template<int MODE>
__global__ void __launch_bounds__(256) kern(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8], acc[8];
  uint32_t b = seed * 2654435761u + threadIdx.x;
  #pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = b * (i + 3) + blockIdx.x; acc[i] = i; }
  __shared__ uint32_t sm[1024];
  if (MODE == 3) { for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * seed; __syncthreads(); }
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int u = 0; u < 8; u++) {
      #pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(acc[i]) : "r"(a[i]), "r"(b));
        if (MODE == 1) asm volatile("add.u32 %0,%0,%1;" : "+r"(acc[i]) : "r"(a[i]));
        if (MODE == 2) asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(acc[i]) : "r"(a[i]), "r"(b));
        if (MODE == 3) { uint32_t v; asm volatile("ld.shared.u32 %0,[%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x + 32 * i + u) & 1023]))); acc[i] ^= v; }
        if (MODE == 5) asm volatile("dp4a.u32.s32 %0,%1,%2,%0;" : "+r"(acc[i]) : "r"(a[i]), "r"(b));
        if (MODE == 6) asm volatile("mul.hi.s32 %0,%0,%1;" : "+r"(acc[i]) : "r"(a[i]));
        if (MODE == 7) asm volatile("prmt.b32 %0,%0,%1,%2;" : "+r"(acc[i]) : "r"(a[i]), "r"(b & 0x7777));
        if (MODE == 8) asm volatile("min.s32 %0,%0,%1;" : "+r"(acc[i]) : "r"(a[i]));
        if (MODE == 9) asm volatile("shf.r.clamp.b32 %0,%0,%1,%2;" : "+r"(acc[i]) : "r"(a[i]), "r"(b & 31));
        if (MODE == 10) { // 1 IMAD : 1 ALU(min) alternating on independent chains: do the two pipes overlap?
          if (i & 1) asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(acc[i]) : "r"(a[i]), "r"(b));
          else asm volatile("min.s32 %0,%0,%1;" : "+r"(acc[i]) : "r"(a[i]));
        }
        if (MODE == 4) { // mix: 4 VABSDIFF4 : 1 IMAD-ish
          asm volatile("vabsdiff4.u32.u32.u32.add %0,%1,%2,%0;" : "+r"(acc[i]) : "r"(a[i]), "r"(b));
          if ((i & 3) == 0) asm volatile("mad.lo.u32 %0,%1,%2,%0;" : "+r"(a[i]) : "r"(a[i]), "r"(b));
        }
      }
    }
  }
  uint32_t r = 0;
  #pragma unroll
  for (int i = 0; i < 8; i++) r += acc[i] + a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template<int MODE> int run(const char* name, int sms, double ops_per_iter_lane) {
  uint32_t* d; CK(cudaMalloc(&d, sizeof(uint32_t) * sms * 8 * 256));
  int iters = 20000;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kern<MODE><<<sms * 8, 256>>>(d, 100, 1); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); kern<MODE><<<sms * 8, 256>>>(d, iters, r + 2); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  double lane_ops = (double)sms * 8 * 256 * iters * ops_per_iter_lane;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"op\":\"%s\",\"ms\":%.4f,\"lane_ops_per_s\":%.4e,\"lane_ops_per_clk_per_sm_at_max_clock\":%.2f}\n",
         name, best, lane_ops / (best * 1e-3), lane_ops / (best * 1e-3) / sms / (clk * 1e3));
  cudaFree(d); return 0;
}
int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"sms\":%d,\"max_clock_khz\":%d}\n", sms, clk);
  if (run<0>("vabsdiff4_acc", sms, 64)) return 1;
  if (run<1>("iadd", sms, 64)) return 1;
  if (run<2>("imad", sms, 64)) return 1;
  if (run<3>("lds32", sms, 64)) return 1;
  if (run<4>("vabsdiff4+imad(4:1)", sms, 64)) return 1;
  if (run<5>("dp4a", sms, 64)) return 1;
  if (run<6>("mul.hi.s32", sms, 64)) return 1;
  if (run<7>("prmt", sms, 64)) return 1;
  if (run<8>("min.s32", sms, 64)) return 1;
  if (run<9>("shf.r", sms, 64)) return 1;
  if (run<10>("imad+min(1:1)", sms, 64)) return 1;
  return 0;
}
