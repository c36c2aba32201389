// tools/microbench.cu -- B200 ceilings that bound the SpMV design (not product
// code): streaming reads (LDG.128 vs TMA bulk), and 4-byte random gathers from
// an L2-resident table (what x[col] is). Build: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -Iinclude tools/microbench.cu -o tools/microbench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <loops/util/tma.hxx>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_stream_ldg(const float4* __restrict__ p, size_t n4, float* out) {
  float acc = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
    acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
  }
  for (; i < n4; i += stride) { float4 a = __ldg(p + i); acc += a.x + a.y + a.z + a.w; }
  if (acc == 123.456f) out[0] = acc;
}

template <int CHUNK_BYTES, int STAGES>
__global__ void k_stream_tma(const char* __restrict__ p, size_t nchunks, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long bars[STAGES];
  const int t = threadIdx.x;
  if (t == 0) for (int s = 0; s < STAGES; ++s) loops::tma::barrier_init((uint64_t*)&bars[s], 1);
  __syncthreads();
  size_t next = blockIdx.x;
  if (t == 0) {
    for (int s = 0; s < STAGES && next < nchunks; ++s, next += gridDim.x) {
      loops::tma::barrier_arrive_expect_tx((uint64_t*)&bars[s], CHUNK_BYTES);
      loops::tma::bulk_g2s(smem + (size_t)s * CHUNK_BYTES, p + next * CHUNK_BYTES, CHUNK_BYTES, (uint64_t*)&bars[s]);
    }
  }
  float acc = 0.f;
  int k = 0;
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++k) {
    const int s = k % STAGES;
    loops::tma::barrier_wait((uint64_t*)&bars[s], (k / STAGES) & 1);
    const float4* v = (const float4*)(smem + (size_t)s * CHUNK_BYTES);
    for (int i = t; i < CHUNK_BYTES / 16; i += blockDim.x) { float4 a = v[i]; acc += a.x + a.y + a.z + a.w; }
    __syncthreads();
    if (t == 0 && next < nchunks) {
      loops::tma::fence_proxy_async();
      loops::tma::barrier_arrive_expect_tx((uint64_t*)&bars[s], CHUNK_BYTES);
      loops::tma::bulk_g2s(smem + (size_t)s * CHUNK_BYTES, p + next * CHUNK_BYTES, CHUNK_BYTES, (uint64_t*)&bars[s]);
      next += gridDim.x;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

// gather with a streamed index array (what SpMV does): 16 gathers in flight
__global__ void k_gather_idx(const int4* __restrict__ idx4, size_t n4, const float* __restrict__ x, float* out) {
  float acc = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    int4 a = __ldg(idx4 + i), b = __ldg(idx4 + i + stride), c = __ldg(idx4 + i + 2 * stride), d = __ldg(idx4 + i + 3 * stride);
    float s0 = __ldg(x + a.x), s1 = __ldg(x + a.y), s2 = __ldg(x + a.z), s3 = __ldg(x + a.w);
    float s4 = __ldg(x + b.x), s5 = __ldg(x + b.y), s6 = __ldg(x + b.z), s7 = __ldg(x + b.w);
    float s8 = __ldg(x + c.x), s9 = __ldg(x + c.y), sa = __ldg(x + c.z), sb = __ldg(x + c.w);
    float sc = __ldg(x + d.x), sd = __ldg(x + d.y), se = __ldg(x + d.z), sf = __ldg(x + d.w);
    acc += s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7 + s8 + s9 + sa + sb + sc + sd + se + sf;
  }
  if (acc == 123.456f) out[0] = acc;
}

// pure gather: indices from a hash (no index traffic)
__global__ void k_gather_hash(size_t n, unsigned mask, const float* __restrict__ x, float* out) {
  float acc = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += 8 * stride) {
    float s[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      unsigned h = (unsigned)(i + u * stride) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      s[u] = __ldg(x + (h & mask));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += s[u];
  }
  if (acc == 123.456f) out[0] = acc;
}

// gather throughput as a function of the shared-memory carve-out (L1 = 228 KB - smem)
// and of the load's cache operator. MODE 0: ld.global.nc (default), 1: .L1::no_allocate,
// 2: ld.global.cg (L2 only), 3: ld.global.L1::evict_first
template <int MODE>
__device__ __forceinline__ float gload(const float* p) {
  float v;
  if (MODE == 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 2) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else asm volatile("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <int MODE>
__global__ void k_gather_smem(size_t n, unsigned mask, const float* __restrict__ x, float* out) {
  extern __shared__ float dummy[];
  float acc = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += 16 * stride) {
    float s[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      unsigned h = (unsigned)(i + u * stride) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      s[u] = gload<MODE>(x + (h & mask));
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) acc += s[u];
  }
  if (acc == 123.456f) { out[0] = acc; dummy[threadIdx.x] = acc; }
}

template <typename F> float time_ms(F f, int reps = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t NNZ = 1u << 25;             // 32M elements
  const size_t BYTES = NNZ * 8;            // 256 MiB stream (values+indices worth)
  char* buf; float* out; int* idx; float* x;
  CK(cudaMalloc(&buf, BYTES)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&idx, NNZ * 4)); CK(cudaMalloc(&x, 64u << 20));
  CK(cudaMemset(buf, 0, BYTES)); CK(cudaMemset(x, 0, 64u << 20));
  printf("SMs %d\n", sms);
  for (int occ : {2, 4, 8}) {
    float ms = time_ms([&] { k_stream_ldg<<<sms * occ, 512>>>((const float4*)buf, BYTES / 16, out); });
    printf("stream LDG.128   grid %4d x512 : %.1f us  %.0f GB/s\n", sms * occ, ms * 1e3, BYTES / ms / 1e6);
  }
  {
    auto k = k_stream_tma<16384, 4>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384);
    for (int occ : {1, 2, 3}) {
      float ms = time_ms([&] { k<<<sms * occ, 256, 4 * 16384>>>(buf, BYTES / 16384, out); });
      printf("stream TMA 16KBx4 grid %4d x256 : %.1f us  %.0f GB/s\n", sms * occ, ms * 1e3, BYTES / ms / 1e6);
    }
    auto k2 = k_stream_tma<32768, 2>;
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768);
    for (int occ : {1, 2, 3}) {
      float ms = time_ms([&] { k2<<<sms * occ, 256, 2 * 32768>>>(buf, BYTES / 32768, out); });
      printf("stream TMA 32KBx2 grid %4d x256 : %.1f us  %.0f GB/s\n", sms * occ, ms * 1e3, BYTES / ms / 1e6);
    }
  }
  for (unsigned tbl_log : {20u, 22u, 24u}) {   // table of 2^k floats: 4 MB, 16 MB, 64 MB
    std::vector<int> h(NNZ);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < NNZ; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s & ((1u << tbl_log) - 1)); }
    CK(cudaMemcpy(idx, h.data(), NNZ * 4, cudaMemcpyHostToDevice));
    for (int occ : {4, 8}) {
      float ms = time_ms([&] { k_gather_idx<<<sms * occ, 256>>>((const int4*)idx, NNZ / 4, x, out); });
      printf("gather idx-stream table %3u MB grid %4d x256: %.1f us  %.1f Ggather/s  (idx %.0f GB/s + sectors %.0f GB/s)\n",
             (4u << tbl_log) >> 20, sms * occ, ms * 1e3, NNZ / ms / 1e6, NNZ * 4 / ms / 1e6, NNZ * 32.0 / ms / 1e6);
    }
    float ms = time_ms([&] { k_gather_hash<<<sms * 8, 256>>>(NNZ, (1u << tbl_log) - 1, x, out); });
    printf("gather hash       table %3u MB grid %4d x256: %.1f us  %.1f Ggather/s  (sectors %.0f GB/s)\n",
           (4u << tbl_log) >> 20, sms * 8, ms * 1e3, NNZ / ms / 1e6, NNZ * 32.0 / ms / 1e6);
  }
  printf("-- gather rate vs shared-memory carve-out (1 CTA/SM x 512 threads, 16 gathers in flight per thread, 4 MB table)\n");
  for (int smem_kb : {0, 32, 64, 100, 132, 164, 196, 220}) {
    auto run = [&](auto kern, const char* name) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
      float ms = time_ms([&] { kern<<<sms, 512, smem_kb * 1024>>>(NNZ, (1u << 20) - 1, x, out); });
      printf("  smem %3d KB %-22s: %.1f us  %.1f Ggather/s (%.2f /clk/SM @1.9GHz)\n", smem_kb, name, ms * 1e3, NNZ / ms / 1e6, NNZ / ms / 1e6 / sms / 1.9);
    };
    run(k_gather_smem<0>, "ld.global.nc");
    run(k_gather_smem<1>, "nc.L1::no_allocate");
    run(k_gather_smem<2>, "ld.global.cg");
  }
  printf("-- fine sweep: ld.global.nc gather rate vs dynamic smem per CTA (1 CTA/SM x 512 thr) and (2 CTAs/SM x 256 thr)\n");
  for (int smem_kb = 0; smem_kb <= 224; smem_kb += 8) {
    auto kern = k_gather_smem<0>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    float ms1 = time_ms([&] { kern<<<sms, 512, smem_kb * 1024>>>(NNZ, (1u << 20) - 1, x, out); }, 5);
    float ms2 = -1.f;
    if (smem_kb <= 112) ms2 = time_ms([&] { kern<<<2 * sms, 256, smem_kb * 1024>>>(NNZ, (1u << 20) - 1, x, out); }, 5);
    printf("  smem/CTA %3d KB: 1 CTA/SM %.2f /clk/SM   2 CTAs/SM (total %3d KB) %.2f /clk/SM\n", smem_kb,
           NNZ / ms1 / 1e6 / sms / 1.9, 2 * smem_kb, ms2 > 0 ? NNZ / ms2 / 1e6 / sms / 1.9 : 0.0);
  }
  return 0;
}
