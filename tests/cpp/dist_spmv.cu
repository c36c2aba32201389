// tests/cpp/dist_spmv.cu -- BASELINE configs[4] from a plain C++ host: the multi-GPU step of
// include/loopsb.h (loopsb_dist_*) without Python, PyTorch or MPI. One host thread per GPU (the
// ABI's "one process or host thread per GPU"); rank 0 makes the NCCL id and the other threads read
// it from memory -- where a multi-process launcher would broadcast the 128 bytes. Every rank owns a
// contiguous row range of a random matrix with GLOBAL column ids and its slice of x, runs
// y_shard = A_shard * allgather(x_shard) several times and compares y with a host SpMV bit for bit
// (values k/8, x integers: every sum is exact, so the order of the adds does not matter).
//
//   dist_spmv [--world N] [--groups a,b,...]      N in {1,2,4,8} <= visible GPUs; default 1
// Prints "OK <what>" / "FAIL <what>" lines; exit code = number of failures.
#include <loopsb.h>

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace {

struct host_csr {
  int rows = 0, cols = 0;
  std::vector<int> off, idx;
  std::vector<float> val;
};

host_csr make_matrix(int rows, int cols, unsigned seed) {
  host_csr m;
  m.rows = rows; m.cols = cols;
  m.off.assign(size_t(rows) + 1, 0);
  std::mt19937 rng(seed);
  for (int r = 0; r < rows; ++r) {
    int deg = (r % 13 == 0) ? 0 : (r == 5 ? std::min(cols, 1500) : int(rng() % 33));
    const int step = std::max(1, cols / std::max(deg, 1));
    for (int k = 0; k < deg && k * step < cols; ++k) {       // unique, ascending columns
      m.idx.push_back(k * step + int(rng() % unsigned(step)));
      m.val.push_back(float(1 + rng() % 16) / 8.0f);
    }
    m.off[size_t(r) + 1] = int(m.idx.size());
  }
  return m;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      std::printf("FAIL rank %d: %s -> %s\n", rank, #call, cudaGetErrorString(e_));           \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)
#define LB(call)                                                                              \
  do {                                                                                        \
    int rc_ = (call);                                                                         \
    if (rc_ != LOOPSB_OK) {                                                                   \
      std::printf("FAIL rank %d: %s -> %s (%s)\n", rank, #call, loopsb_status_string(rc_),    \
                  loopsb_last_error());                                                       \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

int run_rank(int world, int rank, const char* id, const host_csr& A, const std::vector<float>& x,
             const std::vector<float>& y_ref, const std::vector<int>& groups) {
  CK(cudaSetDevice(rank));
  const int r0 = int((long long)A.rows * rank / world), r1 = int((long long)A.rows * (rank + 1) / world);
  const int local_rows = r1 - r0, a0 = A.off[size_t(r0)], a1 = A.off[size_t(r1)];
  const long long local_nnz = a1 - a0;
  std::vector<int> off(static_cast<size_t>(local_rows) + 1, 0);
  for (int r = 0; r <= local_rows; ++r) off[size_t(r)] = A.off[size_t(r0 + r)] - a0;   // rebased; columns stay global
  const int xs = A.cols / world;
  int *d_off = nullptr, *d_idx = nullptr;
  float *d_val = nullptr, *d_x = nullptr, *d_y = nullptr;
  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));
  CK(cudaMalloc(&d_off, off.size() * 4));
  CK(cudaMalloc(&d_idx, size_t(std::max<long long>(local_nnz, 1)) * 4));
  CK(cudaMalloc(&d_val, size_t(std::max<long long>(local_nnz, 1)) * 4));
  CK(cudaMalloc(&d_x, size_t(xs) * 4));
  CK(cudaMalloc(&d_y, size_t(std::max(local_rows, 1)) * 4));
  CK(cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, A.idx.data() + a0, size_t(local_nnz) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_val, A.val.data() + a0, size_t(local_nnz) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_x, x.data() + size_t(rank) * xs, size_t(xs) * 4, cudaMemcpyHostToDevice));

  loopsb_dist_t* d = nullptr;
  LB(loopsb_dist_create(&d, world > 1 ? id : nullptr, world, rank, local_rows, A.cols, local_nnz, d_off, d_idx, d_val,
                        groups.empty() ? nullptr : groups.data(), int32_t(groups.size()), stream));
  loopsb_dist_info_t info;
  LB(loopsb_dist_info(d, &info));
  int bad = 0;
  if (info.world != world || info.rank != rank || info.local_rows != local_rows || info.local_nnz != local_nnz ||
      info.num_blocks != int(groups.size()) + 1) {
    std::printf("FAIL rank %d: loopsb_dist_info disagrees with what was passed in\n", rank);
    ++bad;
  }
  std::vector<float> y(static_cast<size_t>(local_rows), 0.0f);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemsetAsync(d_y, 0xff, size_t(std::max(local_rows, 1)) * 4, stream));   // NaN pattern: y must be overwritten
    LB(loopsb_dist_spmv(d, d_x, d_y, stream));
    CK(cudaMemcpyAsync(y.data(), d_y, y.size() * 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (std::memcmp(y.data(), y_ref.data() + r0, y.size() * 4) != 0) {
      std::printf("FAIL rank %d: y shard differs from the host SpMV (step %d)\n", rank, rep);
      ++bad;
      break;
    }
  }
  const float* x_full = nullptr;
  LB(loopsb_dist_x_full(d, &x_full));
  std::vector<float> xg(static_cast<size_t>(A.cols), 0.0f);
  CK(cudaMemcpy(xg.data(), x_full, xg.size() * 4, cudaMemcpyDeviceToHost));
  if (std::memcmp(xg.data(), x.data(), xg.size() * 4) != 0) {
    std::printf("FAIL rank %d: the gathered x differs from x\n", rank);
    ++bad;
  }
  LB(loopsb_dist_destroy(d));
  cudaFree(d_off); cudaFree(d_idx); cudaFree(d_val); cudaFree(d_x); cudaFree(d_y);
  cudaStreamDestroy(stream);
  if (!bad)
    std::printf("OK rank %d/%d: %d rows, %lld nonzeros, %d column block(s), transport %d, NCCL %d\n", rank, world,
                local_rows, local_nnz, info.num_blocks, info.transport, info.nccl_version);
  return bad;
}

}  // namespace

int main(int argc, char** argv) {
  int world = 1;
  std::vector<int> groups;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--world") && i + 1 < argc) world = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--groups") && i + 1 < argc) {
      std::string s = argv[++i];
      for (size_t p = 0; p < s.size();) {
        size_t q = s.find(',', p);
        if (q == std::string::npos) q = s.size();
        groups.push_back(std::atoi(s.substr(p, q - p).c_str()));
        p = q + 1;
      }
    }
  }
  int ngpu = 0;
  if (cudaGetDeviceCount(&ngpu) != cudaSuccess || ngpu < 1) {
    std::printf("FAIL no CUDA device (loops-b200 has no CPU fallback)\n");
    return 1;
  }
  if (!(world == 1 || world == 2 || world == 4 || world == 8) || world > ngpu) {
    std::printf("FAIL --world %d needs that many visible GPUs (have %d) and must be 1, 2, 4 or 8\n", world, ngpu);
    return 1;
  }
  const int rows = 4096 * world + 37, cols = 4096 * world;     // rows need not divide evenly, x shards must
  const host_csr A = make_matrix(rows, cols, 7);
  std::vector<float> x(static_cast<size_t>(cols), 0.0f), y_ref(static_cast<size_t>(rows), 0.0f);
  std::mt19937 rng(11);
  for (auto& v : x) v = float(1 + rng() % 10);
  for (int r = 0; r < rows; ++r) {
    float s = 0.0f;
    for (int a = A.off[size_t(r)]; a < A.off[size_t(r) + 1]; ++a) s += A.val[size_t(a)] * x[size_t(A.idx[size_t(a)])];
    y_ref[size_t(r)] = s;
  }
  char id[LOOPSB_DIST_ID_BYTES] = {0};
  if (world > 1) {
    cudaSetDevice(0);
    const int rc = loopsb_dist_unique_id(id);
    if (rc != LOOPSB_OK) {
      std::printf("FAIL loopsb_dist_unique_id: %s (%s)\n", loopsb_status_string(rc), loopsb_last_error());
      return 1;
    }
  }
  std::atomic<int> failures(0);
  std::vector<std::thread> ranks;
  for (int r = 0; r < world; ++r)
    ranks.emplace_back([&, r]() { failures += run_rank(world, r, id, A, x, y_ref, groups); });
  for (auto& t : ranks) t.join();
  std::printf("failures: %d\n", failures.load());
  return failures.load();
}
