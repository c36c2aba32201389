// tools/microbench_gather4.cu -- can the TMA engine serve 4-byte random gathers?
// x is viewed as a 2-D tensor [n/4 rows][4 floats]; cp.async.bulk.tensor.2d
// .tile::gather4 fetches four 16-byte rows per instruction into shared memory.
// Measures gathers/clk/SM for (a) TMA gather4 alone, (b) LSU gathers alone,
// (c) both at once (is the TMA path additive to the 1 req/clk LSU path?).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <loops/util/tma.hxx>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned hsh(unsigned a) { a *= 2654435761u; a ^= a >> 15; a *= 2246822519u; a ^= a >> 13; return a; }

__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               :: "r"(loops::tma::smem_addr(dst)), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(loops::tma::smem_addr(bar)) : "memory");
}

// MODE 1: TMA only, 2: LSU only, 3: both (each thread does TPT tma gather4 + 4*TPT lsu gathers per round)
template <int MODE, int THREADS, int ISSUERS>
__global__ void __launch_bounds__(THREADS) k(const __grid_constant__ CUtensorMap map, const float* __restrict__ x, unsigned mask, int rounds, float* out) {
  __shared__ __align__(128) float buf[2][ISSUERS][32];  // TMA tensor destinations need 128-byte alignment
  __shared__ unsigned long long bars[2];
  const int t = threadIdx.x;
  if (t == 0) { loops::tma::barrier_init((uint64_t*)&bars[0], 1); loops::tma::barrier_init((uint64_t*)&bars[1], 1); }
  __syncthreads();
  float acc = 0.f;
  unsigned seed = (blockIdx.x * THREADS + t) * 977u;
  auto issue = [&](int b, int r) {
    if (MODE & 1) {
      if (t == 0) loops::tma::barrier_arrive_expect_tx((uint64_t*)&bars[b], ISSUERS * 64);
      __syncthreads();
      if (t < ISSUERS) {
        unsigned h = hsh(seed + r * 131u);
        gather4(&buf[b][t][0], &map, 0, (h & mask) >> 2, (hsh(h) & mask) >> 2, (hsh(h + 1) & mask) >> 2, (hsh(h + 2) & mask) >> 2, (uint64_t*)&bars[b]);
      }
    }
  };
  issue(0, 0);
  for (int r = 0; r < rounds; ++r) {
    const int b = r & 1;
    if (r + 1 < rounds) issue(b ^ 1, r + 1);
    if (MODE & 2) {
      float s[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] = __ldg(x + (hsh(seed + r * 8191u + u * 7u) & mask));
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += s[u];
    }
    if (MODE & 1) {
      loops::tma::barrier_wait((uint64_t*)&bars[b], (r >> 1) & 1);
      if (t < ISSUERS) acc += buf[b][t][0] + buf[b][t][5] + buf[b][t][10] + buf[b][t][15];
      __syncthreads();
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t N = 1u << 20;   // 4 MB of floats
  float *x, *out; CK(cudaMalloc(&x, N * 4)); CK(cudaMalloc(&out, 64)); CK(cudaMemset(x, 0, N * 4));
  typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap map;
  cuuint64_t gdim[2] = {4, N / 4}; cuuint64_t gstr[1] = {16}; cuuint32_t box[2] = {4, 1}; cuuint32_t estr[2] = {1, 1};
  CUresult r = ((enc_t)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("cuTensorMapEncodeTiled rc=%d\n", (int)r);
  if (r != CUDA_SUCCESS) return 2;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int rounds = 2000;
  auto run = [&](auto kern, const char* name, int ctas, int threads, int issuers, int mode) {
    kern<<<sms * ctas, threads>>>(map, x, (unsigned)(N - 1), 20, out);
    cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaEventRecord(a); kern<<<sms * ctas, threads>>>(map, x, (unsigned)(N - 1), rounds, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double tma = (mode & 1) ? double(sms) * ctas * issuers * 4.0 * rounds : 0, lsu = (mode & 2) ? double(sms) * ctas * threads * 8.0 * rounds : 0;
    printf("%-34s %8.1f us  tma %.3f + lsu %.3f = %.3f gathers/clk/SM (@1.9GHz)\n", name, ms * 1e3, tma / ms / 1e6 / sms / 1.9, lsu / ms / 1e6 / sms / 1.9, (tma + lsu) / ms / 1e6 / sms / 1.9);
  };
  run(k<2, 256, 32>, "LSU only, 4 CTAs x256", 4, 256, 32, 2);
  run(k<1, 256, 32>, "TMA only, 32 issuers, 4 CTAs", 4, 256, 32, 1);
  run(k<1, 256, 128>, "TMA only, 128 issuers, 4 CTAs", 4, 256, 128, 1);
  run(k<1, 256, 128>, "TMA only, 128 issuers, 6 CTAs", 6, 256, 128, 1);
  run(k<3, 256, 32>, "both, 32 issuers, 4 CTAs", 4, 256, 32, 3);
  run(k<3, 256, 128>, "both, 128 issuers, 4 CTAs", 4, 256, 128, 3);
  run(k<3, 256, 8>, "both, 8 issuers, 4 CTAs", 4, 256, 8, 3);
  return 0;
}
