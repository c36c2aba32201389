// tools/microbench_tex.cu -- which SM path serves 4-byte random gathers fastest on B200
// (not product code). The merge-path SpMV is bound by the L1TEX wavefront rate: one
// 32-byte sector request per x[col]. This measures, for a 4 MB and a 64 MB table:
//   ldg      : ld.global.nc through the LSU pipe          (what the kernels do today)
//   tex      : tex1Dfetch<float> on a linear texture object (TEX pipe front end)
//   mix      : half the gathers on each pipe
//   ldg+lds  : LSU gathers with K conflict-free shared-memory loads per gather, to see
//              whether shared-memory wavefronts and global tag lookups share one budget
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench_tex.cu -o tools/microbench_tex
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned hash_u(unsigned v) {
  v *= 2654435761u; v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
  return v;
}

// MODE 0 ldg, 1 tex, 2 mix (even u -> ldg, odd u -> tex)
template <int MODE, int U>
__global__ void k_gather(size_t n, unsigned mask, const float* __restrict__ x, cudaTextureObject_t tx, float* out) {
  float acc = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += U * stride) {
    float s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned h = hash_u((unsigned)(i + u * stride)) & mask;
      const bool use_tex = MODE == 1 || (MODE == 2 && (u & 1));
      s[u] = use_tex ? tex1Dfetch<float>(tx, int(h)) : __ldg(x + h);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += s[u];
  }
  if (acc == 123.456f) out[0] = acc;
}

// LSU gathers + LDS_PER16 conflict-free shared loads per 16 gathers
template <int LDS_PER16>
__global__ void k_gather_lds(size_t n, unsigned mask, const float* __restrict__ x, float* out) {
  __shared__ float sh[2048];
  for (int k = threadIdx.x; k < 2048; k += blockDim.x) sh[k] = float(k);
  __syncthreads();
  float acc = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int rot = threadIdx.x;
  for (; i < n; i += 16 * stride) {
    float s[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) s[u] = __ldg(x + (hash_u((unsigned)(i + u * stride)) & mask));
#pragma unroll
    for (int k = 0; k < LDS_PER16; ++k) { acc += sh[(rot + 32 * k) & 2047]; }
    rot += 7 * 32;
#pragma unroll
    for (int u = 0; u < 16; ++u) acc += s[u];
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F> float time_ms(F f, int reps = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz / 1e6;
  const size_t N = 1u << 25;
  float *x, *out;
  CK(cudaMalloc(&x, 64u << 20)); CK(cudaMalloc(&out, 64)); CK(cudaMemset(x, 0, 64u << 20));
  printf("SMs %d  clock %.3f GHz\n", sms, ghz);
  for (unsigned tbl_log : {20u, 24u}) {
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x;
    rd.res.linear.desc = cudaCreateChannelDesc<float>(); rd.res.linear.sizeInBytes = size_t(4) << tbl_log;
    cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tx = 0;
    CK(cudaCreateTextureObject(&tx, &rd, &td, nullptr));
    const unsigned mask = (1u << tbl_log) - 1;
    auto report = [&](const char* name, float ms) {
      printf("table %3u MB %-28s: %7.1f us  %6.1f Ggather/s  %.2f /clk/SM\n", (4u << tbl_log) >> 20, name, ms * 1e3,
             N / ms / 1e6, N / ms / 1e6 / sms / ghz);
    };
    for (int occ : {4, 8}) {
      char nm[64];
      snprintf(nm, 64, "ldg  x16 grid %dx256", sms * occ);
      report(nm, time_ms([&] { k_gather<0, 16><<<sms * occ, 256>>>(N, mask, x, tx, out); }));
      snprintf(nm, 64, "tex  x16 grid %dx256", sms * occ);
      report(nm, time_ms([&] { k_gather<1, 16><<<sms * occ, 256>>>(N, mask, x, tx, out); }));
      snprintf(nm, 64, "mix  x16 grid %dx256", sms * occ);
      report(nm, time_ms([&] { k_gather<2, 16><<<sms * occ, 256>>>(N, mask, x, tx, out); }));
    }
    report("tex  x8  grid 8/SM x256", time_ms([&] { k_gather<1, 8><<<sms * 8, 256>>>(N, mask, x, tx, out); }));
    report("ldg+lds 0/16", time_ms([&] { k_gather_lds<0><<<sms * 8, 256>>>(N, mask, x, out); }));
    report("ldg+lds 2/16", time_ms([&] { k_gather_lds<2><<<sms * 8, 256>>>(N, mask, x, out); }));
    report("ldg+lds 4/16", time_ms([&] { k_gather_lds<4><<<sms * 8, 256>>>(N, mask, x, out); }));
    report("ldg+lds 8/16", time_ms([&] { k_gather_lds<8><<<sms * 8, 256>>>(N, mask, x, out); }));
    report("ldg+lds 16/16", time_ms([&] { k_gather_lds<16><<<sms * 8, 256>>>(N, mask, x, out); }));
    cudaDestroyTextureObject(tx);
  }
  return 0;
}
