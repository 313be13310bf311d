// Microbenchmark: issue throughput of the instructions the packed DP kernel is made of, alone and in
// pairs, to find out which share an execution pipe on sm_100a (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Prints warp-instructions per cycle per SM (max 4 = one per SMSP per cycle).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t d; asm volatile("prmt.b32 %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }
__device__ __forceinline__ uint32_t imad1(uint32_t a, uint32_t c) { uint32_t d; asm volatile("mad.lo.u32 %0,%1,1,%2;" : "=r"(d) : "r"(a), "r"(c)); return d; }
__device__ __forceinline__ uint32_t hset(uint32_t a, uint32_t b) { return __hgt2_mask(*(__half2*)&a, *(__half2*)&b); }
__device__ __forceinline__ uint32_t hsetbf(uint32_t a, uint32_t b) { __half2 r = __hgt2(*(__half2*)&a, *(__half2*)&b); return *(uint32_t*)&r; }
__device__ __forceinline__ uint32_t hfma(uint32_t a, uint32_t b, uint32_t c) { __half2 r = __hfma2(*(__half2*)&a, *(__half2*)&b, *(__half2*)&c); return *(uint32_t*)&r; }

enum { OP_HMNMX, OP_HADD, OP_VIMNMX, OP_VIADDMNMX, OP_VIADD16, OP_PRMT, OP_LOP3, OP_IMAD, OP_HSET, OP_HSETBF, OP_HFMA, OP_IADD, OP_VIMNMX3, OP_NONE };
template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x, uint32_t y, uint32_t z) {
  if (OP == OP_HMNMX) { __half2 r = __hmax2(*(__half2*)&x, *(__half2*)&y); return *(uint32_t*)&r; }
  if (OP == OP_HADD) { __half2 r = __hsub2(*(__half2*)&x, *(__half2*)&y); return *(uint32_t*)&r; }
  if (OP == OP_VIMNMX) return __vmaxs2(x, y);
  if (OP == OP_VIADDMNMX) return __viaddmax_s16x2(x, y, z);
  if (OP == OP_VIADD16) return __vadd2(x, y);
  if (OP == OP_PRMT) return prmt(x, y, 0x5410);
  if (OP == OP_LOP3) return (x & y) ^ z;
  if (OP == OP_IMAD) return imad1(x, y);
  if (OP == OP_HSET) return hset(x, y);
  if (OP == OP_HSETBF) return hsetbf(x, y);
  if (OP == OP_HFMA) return hfma(x, y, z);
  if (OP == OP_IADD) return x + y + z + 0x10001;
  if (OP == OP_VIMNMX3) return __vimax3_s16x2(x, y, z);
  return x;
}

template <int A, int B>
__global__ void bench(uint32_t* out, long long* cyc, int iters) {
  uint32_t a[8], b[8];
  for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 7 + k + 0x30003000; b[k] = threadIdx.x * 13 + k + 0x31003100; }
  const uint32_t y = out[0] | 0x30013001, z = out[1] | 0x2fff2fff;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a[k] = op<A>(a[k], y, z);
      if (B != OP_NONE) b[k] = op<B>(b[k], z, y);
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int k = 0; k < 8; ++k) s += a[k] ^ b[k];
  out[2 + blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Instruction mixes of one packed cell of the forward traceback kernel, on independent chains.
template <int MIX>
__global__ void mix(uint32_t* out, long long* cyc, int iters) {
  uint32_t r[20];
  for (int k = 0; k < 20; ++k) r[k] = threadIdx.x * 7 + k + 0x30003000;
  const uint32_t y = out[0] | 0x30013001, z = out[1] | 0x2fff2fff;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    // common part: 4 VIMNMX, 1 VIADDMNMX, 3 PRMT, 3 IMAD
    r[0] = op<OP_VIMNMX>(r[0], y, z); r[1] = op<OP_VIMNMX>(r[1], z, y); r[2] = op<OP_VIMNMX>(r[2], y, z); r[3] = op<OP_VIMNMX>(r[3], z, y);
    r[4] = op<OP_VIADDMNMX>(r[4], y, z);
    r[5] = op<OP_PRMT>(r[5], y, z); r[6] = op<OP_PRMT>(r[6], z, y); r[7] = op<OP_PRMT>(r[7], y, z);
    r[8] = op<OP_IMAD>(r[8], y, z); r[9] = op<OP_IMAD>(r[9], z, y); r[10] = op<OP_IMAD>(r[10], y, z);
    if (MIX == 1) {  // flags: 4 HSET2 (mask) + 4 LOP3
      r[11] = op<OP_HSET>(r[11], y, z); r[12] = op<OP_HSET>(r[12], z, y); r[13] = op<OP_HSET>(r[13], y, z); r[14] = op<OP_HSET>(r[14], z, y);
      r[15] = op<OP_LOP3>(r[15], r[11], z); r[16] = op<OP_LOP3>(r[16], r[12], z); r[17] = op<OP_LOP3>(r[17], r[13], z); r[18] = op<OP_LOP3>(r[18], r[14], z);
    } else if (MIX == 2) {  // flags: 4 HSET2.BF + 4 HFMA2
      r[11] = op<OP_HSETBF>(r[11], y, z); r[12] = op<OP_HSETBF>(r[12], z, y); r[13] = op<OP_HSETBF>(r[13], y, z); r[14] = op<OP_HSETBF>(r[14], z, y);
      r[15] = op<OP_HFMA>(r[15], y, r[11]); r[16] = op<OP_HFMA>(r[16], y, r[12]); r[17] = op<OP_HFMA>(r[17], y, r[13]); r[18] = op<OP_HFMA>(r[18], y, r[14]);
    } else if (MIX == 3) {  // flags via 4 VIADD.16x2 differences + 2 PRMT + 2 LOP3 (sign gather)
      r[11] = op<OP_VIADD16>(r[11], y, z); r[12] = op<OP_VIADD16>(r[12], z, y); r[13] = op<OP_VIADD16>(r[13], y, z); r[14] = op<OP_VIADD16>(r[14], z, y);
      r[15] = op<OP_PRMT>(r[11], r[12], z); r[16] = op<OP_PRMT>(r[13], r[14], z);
      r[17] = op<OP_LOP3>(r[17], r[15], z); r[18] = op<OP_LOP3>(r[18], r[16], z);
    } else if (MIX == 4) {  // no flags (score-only like)
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int k = 0; k < 20; ++k) s += r[k];
  out[2 + blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MIX>
void runmix(const char* name, int ninst, uint32_t* d_out, long long* d_cyc) {
  const int iters = 2000, threads = 512, blocks = 148;  // 16 warps per SM
  mix<MIX><<<blocks, threads>>>(d_out, d_cyc, iters);
  mix<MIX><<<blocks, threads>>>(d_out, d_cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += (double)h[i];
  avg /= blocks;
  printf("%-44s %6.2f cycles per cell per SMSP (%d inst)\n", name, avg / iters / (threads / 32 / 4), ninst);
}

template <int A, int B>
double run(const char* name, uint32_t* d_out, long long* d_cyc) {
  const int iters = 2000, threads = 1024, blocks = 148;  // 32 warps per SM, one CTA per SM
  bench<A, B><<<blocks, threads>>>(d_out, d_cyc, iters);
  bench<A, B><<<blocks, threads>>>(d_out, d_cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += (double)h[i];
  avg /= blocks;
  const double ninst = (double)iters * 8 * (B == OP_NONE ? 1 : 2) * (threads / 32);
  const double ipc = ninst / avg;
  printf("%-28s %6.2f warp-inst/clk/SM\n", name, ipc);
  return ipc;
}

int main() {
  uint32_t* d_out; long long* d_cyc;
  cudaMalloc(&d_out, (2 + 148 * 1024) * 4); cudaMemset(d_out, 0, (2 + 148 * 1024) * 4);
  cudaMalloc(&d_cyc, 148 * 8);
#define ONE(X) run<X, OP_NONE>(#X, d_out, d_cyc)
#define TWO(X, Y) run<X, Y>(#X " + " #Y, d_out, d_cyc)
  ONE(OP_HMNMX); ONE(OP_HADD); TWO(OP_HMNMX, OP_VIMNMX); TWO(OP_HMNMX, OP_PRMT); TWO(OP_HMNMX, OP_HADD); TWO(OP_HMNMX, OP_IMAD); TWO(OP_HMNMX, OP_LOP3); TWO(OP_HADD, OP_IMAD); TWO(OP_HADD, OP_VIMNMX);
  ONE(OP_VIMNMX); ONE(OP_VIADDMNMX); ONE(OP_VIMNMX3); ONE(OP_VIADD16); ONE(OP_PRMT); ONE(OP_LOP3); ONE(OP_IADD); ONE(OP_IMAD); ONE(OP_HSET); ONE(OP_HSETBF); ONE(OP_HFMA);
  TWO(OP_VIMNMX, OP_LOP3); TWO(OP_VIMNMX, OP_PRMT); TWO(OP_VIMNMX, OP_VIADD16); TWO(OP_VIMNMX, OP_IMAD); TWO(OP_VIMNMX, OP_HSET); TWO(OP_VIMNMX, OP_HSETBF);
  TWO(OP_VIMNMX, OP_HFMA); TWO(OP_IMAD, OP_HSET); TWO(OP_IMAD, OP_HFMA); TWO(OP_HSET, OP_HFMA); TWO(OP_LOP3, OP_HSET); TWO(OP_IMAD, OP_VIADD16); TWO(OP_LOP3, OP_IADD);
  TWO(OP_VIMNMX, OP_VIADDMNMX); TWO(OP_PRMT, OP_IMAD); TWO(OP_PRMT, OP_HSET); TWO(OP_PRMT, OP_VIADD16); TWO(OP_PRMT, OP_HFMA); TWO(OP_PRMT, OP_VIADDMNMX); TWO(OP_HSET, OP_VIADD16); TWO(OP_HSET, OP_VIADDMNMX); TWO(OP_LOP3, OP_PRMT); TWO(OP_LOP3, OP_VIADD16); TWO(OP_LOP3, OP_IMAD); TWO(OP_LOP3, OP_HFMA); TWO(OP_VIADD16, OP_HFMA); TWO(OP_VIADD16, OP_VIADDMNMX);
  runmix<4>("mix: common (4vimnmx,1viaddmnmx,3prmt,3imad)", 11, d_out, d_cyc);
  runmix<1>("mix: + 4 HSET2 + 4 LOP3", 19, d_out, d_cyc);
  runmix<2>("mix: + 4 HSET2.BF + 4 HFMA2", 19, d_out, d_cyc);
  runmix<3>("mix: + 4 VIADD16 + 2 PRMT + 2 LOP3", 19, d_out, d_cyc);
  return 0;
}
