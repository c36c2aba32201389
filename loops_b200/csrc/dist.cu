// loops_b200/csrc/dist.cu -- row-partitioned multi-GPU SpMV behind the C ABI
// (include/loopsb.h, loopsb_dist_*; SURVEY.md section 8e, BASELINE configs[4]).
//
// The reference is single-GPU (memory.hxx:24 only mentions multi-GPU as a TODO); the
// partitioning is fixed by the north star: rank r owns a contiguous row range with
// GLOBAL column ids, and the only thing that crosses GPUs is the dense x, all-gathered
// over NVLink with NCCL. One process (or host thread) per GPU; NCCL is loaded at run
// time (dlopen) so libloopsb200.so itself does not depend on it.
//
// Two ways to run a step  y_shard = A_shard * allgather(x_shard):
//   * groups == none : ONE ncclAllGather on the caller's stream, then one merge-path
//     SpMV over the caller's CSR arrays (nothing is copied);
//   * groups given   : the all-gather is issued as ring-shifted NCCL send/recv phases on
//     a side stream -- phase g brings the chunks of ranks r+k, k in group g -- and the
//     shard is held as column blocks (loopsb_csr_split_columns_*): block 0 = the rank's
//     own columns, block g = the columns that arrive in phase g. Block g's SpMV
//     (y += A_g x, same merge-path kernel) starts as soon as phase g has landed, so
//     every phase after the first hides behind the SpMV of the block before it, and
//     each block gathers from a slice of x small enough to stay L2-resident.
// y is bit-identical between the two on exactly representable inputs (the blocks only
// regroup a row's adds by column range).
//
// Transport of the phased all-gather. NCCL's send/recv kernels hold SMs while they run:
// measured on 8 B200s they slowed the concurrent SpMV blocks 2x (and moved 56 MB in 360 us).
// So the chunks are PULLED by the copy engines instead: every rank keeps a double-buffered
// copy of its x shard plus a few flag words in a CUDA-IPC region all ranks map; a step is
// nothing but stream-ordered operations --
//     S   : wait(peers have pulled my buffer of step k-2) ; x_shard -> stage[k&1] ;
//           write ready[k&1][me] = k into every peer's region
//     side: per chunk  wait(ready[k&1][peer] >= k) ; cudaMemcpyAsync(peer stage -> x_full) ;
//           write pulled[k&1][me] = k into the peer's region
// (cuStreamWaitValue32 / cuStreamWriteValue32, no kernels, no host barrier, no SM), with
// the copies of a phase spread over two side streams. NCCL stays the rendezvous (handle
// exchange at create) and the fallback transport (LOOPSB_DIST_TRANSPORT=nccl, or when the
// IPC mapping is refused).
#include "common.cuh"

#include <cuda.h>      // types of the stream memory operations only (entry points come from the runtime)
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>

using namespace loopsb;

namespace {

// ---- the slice of NCCL's ABI used here (nccl.h, stable since 2.x) ---------------
typedef struct ncclComm* nccl_comm_t;
struct nccl_unique_id { char internal[128]; };
constexpr int kNcclFloat32 = 7;   // ncclFloat32 in ncclDataType_t

struct nccl_api {
  void* handle = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

const nccl_api* nccl() {
  static nccl_api api;
  static bool tried = false;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("LOOPSB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);   // a copy the process already loaded (torch's) is re-used by soname
    if (h) break;
  }
  if (!h) {
    set_error("NCCL not found (dlopen libnccl.so.2: %s); set LOOPSB_NCCL_LIB", dlerror());
    return nullptr;
  }
  auto sym = [&](const char* n) { return dlsym(h, n); };
  api.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
  api.GetUniqueId = reinterpret_cast<int (*)(nccl_unique_id*)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(nccl_comm_t*, int, nccl_unique_id, int)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<int (*)(nccl_comm_t)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t)>(sym("ncclAllGather"));
  api.Send = reinterpret_cast<int (*)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<int (*)(void*, size_t, int, int, nccl_comm_t, cudaStream_t)>(sym("ncclRecv"));
  api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.Send || !api.Recv ||
      !api.GroupStart || !api.GroupEnd) {
    set_error("libnccl is missing a required symbol");
    dlclose(h);
    return nullptr;
  }
  api.handle = h;
  return &api;
}

#define NCCL_TRY(expr)                                                                   \
  do {                                                                                   \
    const int r__ = (expr);                                                              \
    if (r__ != 0) {                                                                      \
      const nccl_api* n__ = nccl();                                                      \
      set_error("%s failed: %s", #expr, (n__ && n__->GetErrorString) ? n__->GetErrorString(r__) : "nccl error"); \
      return LOOPSB_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

// y = ((y + p1) + p2) + ... in block order: every column block after the first writes its
// partial row sums to its own buffer (a read-modify-write of y inside the SpMV kernel put a
// dependent global load behind the gather queue for every row: +20 % per block), and this
// streaming pass folds them in a fixed order -- deterministic, 4 bytes per row and block.
struct part_ptrs { const float* p[8]; };
__global__ void __launch_bounds__(256) combine_parts_kernel(float* __restrict__ y, part_ptrs parts, int nparts, int rows) {
  const int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    float acc = y[r];
    for (int q = 0; q < nparts; ++q) acc = __fadd_rn(acc, __ldg(parts.p[q] + r));
    y[r] = acc;
  }
}

struct col_block {
  loopsb_plan_t* plan = nullptr;
  const int32_t* offsets = nullptr;    // [rows + 1]
  const int32_t* indices = nullptr;    // global column ids
  const float* values = nullptr;
  int64_t nnz = 0;
  std::vector<int> shifts;             // ring shifts k whose chunks (rank r + k) this block reads
  float* y_part = nullptr;             // blocks >= 1 write their partial y here (rows floats, owned)
  cudaEvent_t landed = nullptr;        // recorded on the side stream when its chunks are in x_full
  cudaEvent_t k_begin = nullptr, k_end = nullptr;   // breakdown probes
};

}  // namespace

struct loopsb_dist {
  int world = 1, rank = 0, device = 0;
  int rows = 0, cols = 0, chunk_cols = 0;
  int64_t nnz = 0;
  nccl_comm_t comm = nullptr;
  float* x_full = nullptr;
  cudaStream_t side = nullptr;         // the all-gather phases run here when the shard is split
  cudaEvent_t start = nullptr, comm_begin = nullptr, comm_end = nullptr;
  std::vector<col_block> blocks;       // 1 block = no split (borrowed arrays)
  // owned storage of the split copy
  int32_t* blk_offsets = nullptr;
  int32_t* blk_indices = nullptr;
  float* blk_values = nullptr;
  bool probing = false;
  long long bytes = 0;
  // copy-engine transport (see the header comment)
  struct p2p_state {
    bool on = false;
    char* region = nullptr;              // [stage 0 | stage 1 | flags]
    size_t stage_bytes = 0, flags_off = 0;
    std::vector<char*> peer;             // every rank's region as this process sees it (peer[rank] = region)
    std::vector<char> opened;            // 1 = mapped with cudaIpcOpenMemHandle (closed on destroy)
    static constexpr int kPull = 7;      // copy streams the pulls are spread over (one per peer up to 8 ranks:
                                         // an 8 MB pull is ~30 us of mostly latency, 7 in flight fill the link)
    cudaStream_t side[kPull] = {};
    cudaStream_t pub = nullptr;          // publishes my shard: stage copy + ready flags
    cudaEvent_t joined[kPull] = {};
    cudaEvent_t published = nullptr;
    uint32_t step = 0;
    CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*batch)(CUstream, unsigned int, CUstreamBatchMemOpParams*, unsigned int) = nullptr;
  } p2p;
  std::vector<cudaEvent_t> landed_more;  // per block, per extra pull stream (kPull - 1 each)
  // a step is ~70 driver calls (flag operations, copies, event edges, launches); with 8 ranks the
  // host could not issue them as fast as the GPU ran them (0.50 ms per step, whatever the phases).
  // The whole step is therefore captured ONCE per (x_shard, y_shard, buffer parity) into a CUDA
  // graph and replayed with one cudaGraphLaunch.
  struct graph_entry { const float* x; float* y; uint32_t q; cudaGraphExec_t exec; };
  std::vector<graph_entry> graphs;
  bool use_graphs = true;
};

namespace {
void free_dist(loopsb_dist* d) {
  if (!d) return;
  for (col_block& b : d->blocks) {
    if (b.plan) loopsb_plan_destroy(b.plan);
    if (b.y_part) cudaFree(b.y_part);
    if (b.landed) cudaEventDestroy(b.landed);
    if (b.k_begin) cudaEventDestroy(b.k_begin);
    if (b.k_end) cudaEventDestroy(b.k_end);
  }
  if (d->p2p.region) {
    // nobody may still be reading my staged shard: wait (bounded) until every peer has
    // acknowledged the pull of the last step it took part in, then unmap and free
    cudaDeviceSynchronize();
    if (d->p2p.on && d->p2p.step > 0) {
      const size_t off = d->p2p.flags_off + size_t(2 * d->world) * 4;
      std::vector<uint32_t> acks(size_t(2 * d->world));
      for (int spin = 0; spin < 2000; ++spin) {
        if (cudaMemcpy(acks.data(), d->p2p.region + off, acks.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        bool all = true;
        for (int i = 0; i < 2 * d->world; ++i)
          if (i % d->world != d->rank && acks[size_t(i)] != 1u) all = false;
        if (all) break;
        usleep(1000);
      }
    }
    for (int p = 0; p < int(d->p2p.peer.size()); ++p)
      if (d->p2p.opened[size_t(p)]) cudaIpcCloseMemHandle(d->p2p.peer[size_t(p)]);
    cudaFree(d->p2p.region);
  }
  for (cudaStream_t st : d->p2p.side) if (st) cudaStreamDestroy(st);
  if (d->p2p.pub) cudaStreamDestroy(d->p2p.pub);
  for (cudaEvent_t e : d->p2p.joined) if (e) cudaEventDestroy(e);
  if (d->p2p.published) cudaEventDestroy(d->p2p.published);
  for (cudaEvent_t e : d->landed_more) if (e) cudaEventDestroy(e);
  for (auto& g : d->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (d->comm && nccl()) nccl()->CommDestroy(d->comm);
  if (d->x_full) cudaFree(d->x_full);
  if (d->blk_offsets) cudaFree(d->blk_offsets);
  if (d->blk_indices) cudaFree(d->blk_indices);
  if (d->blk_values) cudaFree(d->blk_values);
  if (d->side) cudaStreamDestroy(d->side);
  if (d->start) cudaEventDestroy(d->start);
  if (d->comm_begin) cudaEventDestroy(d->comm_begin);
  if (d->comm_end) cudaEventDestroy(d->comm_end);
  (void)cudaGetLastError();
  delete d;
}
}  // namespace

namespace {
// Collective: allocate the IPC region, exchange handles through NCCL, map the peers.
// Any failure leaves d->p2p.on == false (the NCCL phases are used instead) -- but every rank
// must reach the same verdict, so the outcome is agreed on with one more tiny all-gather.
int setup_p2p(loopsb_dist* d, cudaStream_t s) {
  const nccl_api* n = nccl();
  auto& P = d->p2p;
  const int W = d->world;
  bool ok = true;
  cudaDriverEntryPointQueryResult qr;
  void* fw = nullptr; void* fr = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fw, cudaEnableDefault, &qr) != cudaSuccess || !fw ||
      cudaGetDriverEntryPoint("cuStreamWriteValue32", &fr, cudaEnableDefault, &qr) != cudaSuccess || !fr) {
    (void)cudaGetLastError();
    ok = false;
  }
  void* fb = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamBatchMemOp", &fb, cudaEnableDefault, &qr) != cudaSuccess) { (void)cudaGetLastError(); fb = nullptr; }
  P.batch = reinterpret_cast<CUresult (*)(CUstream, unsigned int, CUstreamBatchMemOpParams*, unsigned int)>(fb);
  P.wait32 = reinterpret_cast<CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int)>(fw);
  P.write32 = reinterpret_cast<CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int)>(fr);
  P.stage_bytes = (size_t(d->chunk_cols) * sizeof(float) + 255) & ~size_t(255);
  P.flags_off = 2 * P.stage_bytes;
  const size_t region_bytes = P.flags_off + size_t(4 * W) * 4 + 256;
  if (ok && cudaMalloc(&P.region, region_bytes) != cudaSuccess) { (void)cudaGetLastError(); P.region = nullptr; ok = false; }
  if (ok) {
    std::vector<uint32_t> init(size_t(4 * W), 0u);
    for (int i = 2 * W; i < 4 * W; ++i) init[size_t(i)] = 1u;       // pulled[q][p] = 1: both buffers are free
    if (cudaMemsetAsync(P.region, 0, region_bytes, s) != cudaSuccess ||
        cudaMemcpyAsync(P.region + P.flags_off, init.data(), init.size() * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) ok = false;
  }
  // record exchanged per rank: {ipc handle 64 B, pid, device, raw pointer, ok}
  struct rec { cudaIpcMemHandle_t h; long long pid; long long dev; unsigned long long ptr; long long ok; };
  static_assert(sizeof(rec) == 96, "exchange record layout");
  rec mine{};
  if (ok && cudaIpcGetMemHandle(&mine.h, P.region) != cudaSuccess) { (void)cudaGetLastError(); ok = false; }
  mine.pid = (long long)getpid(); mine.dev = d->device; mine.ptr = (unsigned long long)(uintptr_t)P.region; mine.ok = ok;
  rec* dbuf = nullptr;
  std::vector<rec> all(size_t(W), rec{});
  if (cudaMalloc(&dbuf, sizeof(rec) * (size_t(W) + 1)) != cudaSuccess) { (void)cudaGetLastError(); return LOOPSB_ERR_ALLOC; }
  auto exchange = [&](rec& r) -> bool {     // all-gather of one record per rank (bytes as ncclInt8 = 0)
    if (cudaMemcpyAsync(dbuf + W, &r, sizeof(rec), cudaMemcpyHostToDevice, s) != cudaSuccess) return false;
    if (n->AllGather(dbuf + W, dbuf, sizeof(rec), 0, d->comm, s) != 0) return false;
    if (cudaMemcpyAsync(all.data(), dbuf, sizeof(rec) * W, cudaMemcpyDeviceToHost, s) != cudaSuccess) return false;
    return cudaStreamSynchronize(s) == cudaSuccess;
  };
  if (!exchange(mine)) { cudaFree(dbuf); set_error("handle exchange over NCCL failed"); return LOOPSB_ERR_CUDA; }
  for (int p = 0; p < W; ++p) ok = ok && all[size_t(p)].ok != 0;
  P.peer.assign(size_t(W), nullptr);
  P.opened.assign(size_t(W), 0);
  if (ok) {
    for (int p = 0; p < W && ok; ++p) {
      if (p == d->rank) { P.peer[size_t(p)] = P.region; continue; }
      const rec& r = all[size_t(p)];
      if (r.pid == mine.pid) {              // ranks as threads of one process: plain peer access
        int can = 0;
        cudaDeviceCanAccessPeer(&can, d->device, int(r.dev));
        if (!can) { ok = false; break; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(int(r.dev), 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
        (void)cudaGetLastError();
        P.peer[size_t(p)] = reinterpret_cast<char*>(uintptr_t(r.ptr));
      } else {
        void* m = nullptr;
        if (cudaIpcOpenMemHandle(&m, r.h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); ok = false; break; }
        P.peer[size_t(p)] = static_cast<char*>(m);
        P.opened[size_t(p)] = 1;
      }
    }
  }
  // second round: did everybody map everybody?
  mine.ok = ok;
  if (!exchange(mine)) { cudaFree(dbuf); set_error("handle exchange over NCCL failed"); return LOOPSB_ERR_CUDA; }
  for (int p = 0; p < W; ++p) ok = ok && all[size_t(p)].ok != 0;
  cudaFree(dbuf);
  if (ok) {
    for (cudaStream_t& st : P.side)
      if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) ok = false;
    if (cudaStreamCreateWithFlags(&P.pub, cudaStreamNonBlocking) != cudaSuccess) ok = false;
    for (cudaEvent_t& e : P.joined)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) ok = false;
    if (cudaEventCreateWithFlags(&P.published, cudaEventDisableTiming) != cudaSuccess) ok = false;
    d->landed_more.assign(d->blocks.size() * (P.kPull - 1), nullptr);
    for (cudaEvent_t& e : d->landed_more)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) ok = false;
    (void)cudaGetLastError();
  }
  P.on = ok;
  if (ok) d->bytes += (long long)region_bytes;
  return LOOPSB_OK;
}

#define CU_TRY(expr)                                                               \
  do {                                                                             \
    const CUresult r__ = (expr);                                                   \
    if (r__ != CUDA_SUCCESS) {                                                     \
      set_error("%s failed with CUresult %d", #expr, int(r__));                    \
      return LOOPSB_ERR_CUDA;                                                      \
    }                                                                              \
  } while (0)

// One step's x exchange with the copy engines. Nothing here sits in front of block 0 on the
// caller's stream `s` except the copy of its own chunk: the flag traffic (a posted write over
// NVLink per peer) runs on the publish stream, the pulls on kPull copy streams. On return `s`
// may run block g once it has waited for blocks[g].landed and the block's landed_more events.
int exchange_p2p(loopsb_dist* d, const float* x_shard, cudaStream_t s, uint32_t q) {
  auto& P = d->p2p;
  constexpr int K = loopsb_dist::p2p_state::kPull;
  const int W = d->world, me = d->rank;
  const size_t chunk_b = size_t(d->chunk_cols) * sizeof(float);
  // Flags are binary semaphores with constant wait / write values, so that a whole step can be
  // replayed as a CUDA graph:
  //   ready[q][src]    in the PULLER's region : 1 = rank src's stage[q] is full (src sets, puller clears)
  //   pulled[q][puller] in the OWNER's region : 1 = puller is done with my stage[q] (puller sets, owner
  //                                             clears; starts at 1)
  auto ready_at = [&](char* region, int src) { return CUdeviceptr(uintptr_t(region + P.flags_off + (size_t(q) * W + src) * 4)); };
  auto pulled_at = [&](char* region, int puller) { return CUdeviceptr(uintptr_t(region + P.flags_off + (size_t(2 * W) + size_t(q) * W + puller) * 4)); };
  auto wait_op = [](CUdeviceptr a, uint32_t v) {
    CUstreamBatchMemOpParams o; memset(&o, 0, sizeof(o));
    o.waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_32; o.waitValue.address = a; o.waitValue.value = v;
    o.waitValue.flags = CU_STREAM_WAIT_VALUE_EQ;
    return o;
  };
  auto write_op = [](CUdeviceptr a, uint32_t v) {
    CUstreamBatchMemOpParams o; memset(&o, 0, sizeof(o));
    o.writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_32; o.writeValue.address = a; o.writeValue.value = v;
    o.writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;
    return o;
  };
  auto submit = [&](cudaStream_t st, std::vector<CUstreamBatchMemOpParams>& ops) -> int {
    CUstream cst = reinterpret_cast<CUstream>(st);
    if (ops.empty()) return LOOPSB_OK;
    if (P.batch) { CU_TRY(P.batch(cst, unsigned(ops.size()), ops.data(), 0)); }
    else for (auto& o : ops) {
      if (o.operation == CU_STREAM_MEM_OP_WAIT_VALUE_32) CU_TRY(P.wait32(cst, o.waitValue.address, o.waitValue.value, o.waitValue.flags));
      else CU_TRY(P.write32(cst, o.writeValue.address, o.writeValue.value, o.writeValue.flags));
    }
    ops.clear();
    return LOOPSB_OK;
  };
  // `start`: x_shard is ready and the previous step's kernels are done with x_full
  LOOPSB_CUDA_TRY(cudaEventRecord(d->start, s));
  LOOPSB_CUDA_TRY(cudaMemcpyAsync(d->x_full + size_t(me) * d->chunk_cols, x_shard, chunk_b, cudaMemcpyDeviceToDevice, s));

  // ---- publish: every peer is done with my stage[q] -> stage the shard -> raise the peers' flags
  std::vector<CUstreamBatchMemOpParams> ops;
  LOOPSB_CUDA_TRY(cudaStreamWaitEvent(P.pub, d->start, 0));
  for (int p = 0; p < W; ++p) if (p != me) ops.push_back(wait_op(pulled_at(P.region, p), 1u));
  for (int p = 0; p < W; ++p) if (p != me) ops.push_back(write_op(pulled_at(P.region, p), 0u));
  if (int rc = submit(P.pub, ops)) return rc;
  LOOPSB_CUDA_TRY(cudaMemcpyAsync(P.region + size_t(q) * P.stage_bytes, x_shard, chunk_b, cudaMemcpyDeviceToDevice, P.pub));
  for (int sh = 1; sh < W; ++sh)             // the peer that pulls me first is told first
    ops.push_back(write_op(ready_at(P.peer[size_t((me - sh + W) % W)], me), 1u));
  if (int rc = submit(P.pub, ops)) return rc;
  LOOPSB_CUDA_TRY(cudaEventRecord(P.published, P.pub));

  // ---- pulls, ring order, spread over the copy streams
  for (cudaStream_t st : P.side) LOOPSB_CUDA_TRY(cudaStreamWaitEvent(st, d->start, 0));
  if (d->probing) LOOPSB_CUDA_TRY(cudaEventRecord(d->comm_begin, P.side[0]));
  int turn = 0;
  for (size_t g = 1; g < d->blocks.size(); ++g) {
    col_block& b = d->blocks[g];
    for (int sh : b.shifts) {
      const int p = (me + sh) % W;
      cudaStream_t st = P.side[turn++ % K];
      ops.push_back(wait_op(ready_at(P.region, p), 1u));
      ops.push_back(write_op(ready_at(P.region, p), 0u));
      if (int rc = submit(st, ops)) return rc;
      LOOPSB_CUDA_TRY(cudaMemcpyAsync(d->x_full + size_t(p) * d->chunk_cols, P.peer[size_t(p)] + size_t(q) * P.stage_bytes,
                                      chunk_b, cudaMemcpyDefault, st));
      ops.push_back(write_op(pulled_at(P.peer[size_t(p)], me), 1u));
      if (int rc = submit(st, ops)) return rc;
    }
    LOOPSB_CUDA_TRY(cudaEventRecord(b.landed, P.side[0]));
    for (int j = 1; j < K; ++j) LOOPSB_CUDA_TRY(cudaEventRecord(d->landed_more[g * (K - 1) + (j - 1)], P.side[j]));
  }
  if (d->probing) {
    for (int j = 1; j < K; ++j) {
      LOOPSB_CUDA_TRY(cudaEventRecord(P.joined[j], P.side[j]));
      LOOPSB_CUDA_TRY(cudaStreamWaitEvent(P.side[0], P.joined[j], 0));
    }
    LOOPSB_CUDA_TRY(cudaEventRecord(d->comm_end, P.side[0]));
  }
  return LOOPSB_OK;
}

// Everything one step enqueues when the shard is held as column blocks (also what is captured).
int enqueue_split_step(loopsb_dist* d, const float* x_shard, float* y_shard, cudaStream_t s, uint32_t q) {
  const nccl_api* n = nccl();
  const size_t chunk = size_t(d->chunk_cols);
  const bool probe = d->probing;
  if (d->p2p.on) {
    const int rc = exchange_p2p(d, x_shard, s, q);
    if (rc != LOOPSB_OK) return rc;
  } else {
    // own chunk in place, then the ring-shifted NCCL phases on the side stream
    if (chunk)
      LOOPSB_CUDA_TRY(cudaMemcpyAsync(d->x_full + size_t(d->rank) * chunk, x_shard, chunk * sizeof(float),
                                      cudaMemcpyDeviceToDevice, s));
    LOOPSB_CUDA_TRY(cudaEventRecord(d->start, s));
    LOOPSB_CUDA_TRY(cudaStreamWaitEvent(d->side, d->start, 0));
    if (probe) LOOPSB_CUDA_TRY(cudaEventRecord(d->comm_begin, d->side));
    for (size_t g = 1; g < d->blocks.size(); ++g) {
      col_block& b = d->blocks[g];
      NCCL_TRY(n->GroupStart());
      for (int k : b.shifts) {
        const int to = (d->rank - k + d->world) % d->world, from = (d->rank + k) % d->world;
        NCCL_TRY(n->Send(x_shard, chunk, kNcclFloat32, to, d->comm, d->side));
        NCCL_TRY(n->Recv(d->x_full + size_t(from) * chunk, chunk, kNcclFloat32, from, d->comm, d->side));
      }
      NCCL_TRY(n->GroupEnd());
      LOOPSB_CUDA_TRY(cudaEventRecord(b.landed, d->side));
    }
    if (probe) LOOPSB_CUDA_TRY(cudaEventRecord(d->comm_end, d->side));
  }
  for (size_t g = 0; g < d->blocks.size(); ++g) {
    col_block& b = d->blocks[g];
    if (g > 0) LOOPSB_CUDA_TRY(cudaStreamWaitEvent(s, b.landed, 0));
    if (g > 0 && d->p2p.on)
      for (int j = 0; j < d->p2p.kPull - 1; ++j)
        LOOPSB_CUDA_TRY(cudaStreamWaitEvent(s, d->landed_more[g * (d->p2p.kPull - 1) + j], 0));
    if (probe) LOOPSB_CUDA_TRY(cudaEventRecord(b.k_begin, s));
    const int rc = loopsb_spmv_f32(b.plan, b.values, b.indices, nullptr, d->x_full, g == 0 ? y_shard : b.y_part,
                                   d->rows, d->cols, s);
    if (rc != LOOPSB_OK) return rc;
    if (g + 1 == d->blocks.size() && d->rows > 0) {
      part_ptrs parts{};
      for (size_t qq = 1; qq < d->blocks.size(); ++qq) parts.p[qq - 1] = d->blocks[qq].y_part;
      const int grid = std::min((d->rows + 255) / 256, 148 * 8);
      combine_parts_kernel<<<grid, 256, 0, s>>>(y_shard, parts, int(d->blocks.size()) - 1, d->rows);
      LOOPSB_CUDA_TRY(cudaGetLastError());
    }
    if (probe) LOOPSB_CUDA_TRY(cudaEventRecord(b.k_end, s));
  }
  // x_shard is read by the publish stream too: in stream order it is free once this call's work is
  if (d->p2p.on) LOOPSB_CUDA_TRY(cudaStreamWaitEvent(s, d->p2p.published, 0));
  return LOOPSB_OK;
}
}  // namespace

extern "C" {

int loopsb_dist_unique_id(void* id128) {
  LOOPSB_REQUIRE(id128 != nullptr, "id buffer is null");
  const nccl_api* n = nccl();
  if (!n) return LOOPSB_ERR_UNSUPPORTED;
  nccl_unique_id id;
  NCCL_TRY(n->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return LOOPSB_OK;
}

int loopsb_dist_create(loopsb_dist_t** out, const void* id128, int32_t world, int32_t rank, int32_t local_rows,
                       int32_t num_cols, int64_t local_nnz, const int32_t* offsets, const int32_t* col_indices,
                       const float* values, const int32_t* groups, int32_t num_groups, void* stream) {
  LOOPSB_REQUIRE(out != nullptr, "null argument");
  *out = nullptr;
  LOOPSB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad world / rank");
  LOOPSB_REQUIRE(world == 1 || id128 != nullptr, "unique id is null");
  LOOPSB_REQUIRE(local_rows >= 0 && num_cols >= 0 && local_nnz >= 0 && local_nnz < (int64_t(1) << 31), "bad sizes");
  LOOPSB_REQUIRE(num_cols % world == 0, "equal x shards need num_cols % world == 0 (plain all-gather)");
  LOOPSB_REQUIRE(offsets != nullptr && (local_nnz == 0 || (col_indices && values)), "null matrix arrays");
  LOOPSB_REQUIRE(num_groups >= 0 && (num_groups == 0 || groups != nullptr), "bad groups");
  if (current_device() == nullptr) return LOOPSB_ERR_CUDA;
  int total = 0;
  for (int g = 0; g < num_groups; ++g) {
    LOOPSB_REQUIRE(groups[g] >= 1, "a group holds at least one chunk");
    total += groups[g];
  }
  LOOPSB_REQUIRE(num_groups == 0 || total == world - 1, "group sizes must add up to world - 1");
  LOOPSB_REQUIRE(num_groups + 1 <= 8 && world <= 64, "at most 8 column blocks / 64 ranks");
  cudaStream_t s = as_stream(stream);

  loopsb_dist* d = new (std::nothrow) loopsb_dist();
  if (!d) { set_error("host allocation failed"); return LOOPSB_ERR_ALLOC; }
  auto fail = [&](int code) { free_dist(d); return code; };
  d->world = world; d->rank = rank; d->rows = local_rows; d->cols = num_cols; d->nnz = local_nnz;
  d->chunk_cols = num_cols / world;
  cudaGetDevice(&d->device);

  if (world > 1) {
    const nccl_api* n = nccl();
    if (!n) return fail(LOOPSB_ERR_UNSUPPORTED);
    nccl_unique_id id;
    memcpy(&id, id128, sizeof(id));
    const int r = n->CommInitRank(&d->comm, world, id, rank);
    if (r != 0) {
      set_error("ncclCommInitRank failed: %s", n->GetErrorString ? n->GetErrorString(r) : "nccl error");
      return fail(LOOPSB_ERR_CUDA);
    }
  }
  if (cudaMalloc(&d->x_full, size_t(num_cols ? num_cols : 1) * sizeof(float)) != cudaSuccess ||
      cudaEventCreateWithFlags(&d->start, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&d->comm_begin) != cudaSuccess || cudaEventCreate(&d->comm_end) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("device allocation of the gathered x failed");
    return fail(LOOPSB_ERR_ALLOC);
  }
  d->bytes = (long long)num_cols * 4;

  const int nblocks = num_groups + 1;
  d->blocks.resize(size_t(nblocks));
  for (col_block& b : d->blocks) {
    if (cudaEventCreateWithFlags(&b.landed, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&b.k_begin) != cudaSuccess || cudaEventCreate(&b.k_end) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("event creation failed");
      return fail(LOOPSB_ERR_CUDA);
    }
  }
  loopsb_layout_t lay{};
  lay.kind = LOOPSB_LAYOUT_CSR;
  lay.num_tiles = local_rows;

  if (nblocks == 1) {
    col_block& b = d->blocks[0];
    b.offsets = offsets; b.indices = col_indices; b.values = values; b.nnz = local_nnz;
    for (int k = 1; k < world; ++k) b.shifts.push_back(k);
    lay.num_atoms = int32_t(local_nnz);
    lay.offsets = offsets;
    int rc = loopsb_plan_create(&b.plan, &lay, LOOPSB_SCHED_MERGE_PATH_FLAT, s);
    if (rc != LOOPSB_OK) return fail(rc);
  } else {
    if (cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("side stream creation failed");
      (void)cudaGetLastError();
      return fail(LOOPSB_ERR_CUDA);
    }
    // block 0 = own chunk; block g = chunks of ranks rank + k for the shifts of group g
    std::vector<int32_t> block_of_chunk(size_t(world), 0);
    block_of_chunk[size_t(rank)] = 0;
    int k = 1;
    for (int g = 0; g < num_groups; ++g)
      for (int q = 0; q < groups[g]; ++q, ++k) {
        d->blocks[size_t(g) + 1].shifts.push_back(k);
        block_of_chunk[size_t((rank + k) % world)] = g + 1;
      }
    const size_t stride = size_t(local_rows) + 1;
    std::vector<int64_t> bnnz(size_t(nblocks), 0);
    if (cudaMalloc(&d->blk_offsets, stride * nblocks * 4) != cudaSuccess ||
        cudaMalloc(&d->blk_indices, (size_t(local_nnz) + 4 * nblocks) * 4) != cudaSuccess ||
        cudaMalloc(&d->blk_values, (size_t(local_nnz) + 4 * nblocks) * 4) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("device allocation of the column-block copy failed (%lld nonzeros)", (long long)local_nnz);
      return fail(LOOPSB_ERR_ALLOC);
    }
    d->bytes += (long long)(stride * nblocks * 4) + local_nnz * 8;
    int rc = loopsb_csr_split_columns_count(local_rows, local_nnz, offsets, col_indices, d->chunk_cols, world,
                                            block_of_chunk.data(), nblocks, d->blk_offsets, bnnz.data(), s);
    if (rc != LOOPSB_OK) return fail(rc);
    rc = loopsb_csr_split_columns_fill(local_rows, local_nnz, offsets, col_indices, values, d->chunk_cols, world,
                                       block_of_chunk.data(), nblocks, d->blk_offsets, bnnz.data(), d->blk_indices,
                                       d->blk_values, s);
    if (rc != LOOPSB_OK) return fail(rc);
    int64_t base = 0;
    for (int b = 0; b < nblocks; ++b) {
      col_block& cb = d->blocks[size_t(b)];
      cb.offsets = d->blk_offsets + size_t(b) * stride;
      base = (base + 3) & ~int64_t(3);     // loopsb_csr_split_columns_fill starts every block on 16 bytes
      cb.indices = d->blk_indices + base;
      cb.values = d->blk_values + base;
      cb.nnz = bnnz[size_t(b)];
      base += cb.nnz;
      lay.num_atoms = int32_t(cb.nnz);
      lay.offsets = cb.offsets;
      rc = loopsb_plan_create(&cb.plan, &lay, LOOPSB_SCHED_MERGE_PATH_FLAT, s);
      if (rc != LOOPSB_OK) return fail(rc);
      const size_t nchunks = b == 0 ? 1 : cb.shifts.size();
      loopsb_plan_hint_x_bytes(cb.plan, int64_t(nchunks) * d->chunk_cols * int64_t(sizeof(float)));
      if (b > 0) {
        if (cudaMalloc(&cb.y_part, size_t(local_rows ? local_rows : 1) * sizeof(float)) != cudaSuccess) {
          (void)cudaGetLastError();
          set_error("device allocation of a partial y failed");
          return fail(LOOPSB_ERR_ALLOC);
        }
        d->bytes += (long long)local_rows * 4;
      }
    }
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) {
    set_error("distributed plan set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(LOOPSB_ERR_CUDA);
  }
  if (world > 1 && nblocks > 1) {
    const char* tr = getenv("LOOPSB_DIST_TRANSPORT");
    if (!(tr && strcmp(tr, "nccl") == 0)) {
      const int rc = setup_p2p(d, s);
      if (rc != LOOPSB_OK) return fail(rc);
    }
  }
  *out = d;
  return LOOPSB_OK;
}

int loopsb_dist_destroy(loopsb_dist_t* d) {
  free_dist(d);
  return LOOPSB_OK;
}

int loopsb_dist_info(const loopsb_dist_t* d, loopsb_dist_info_t* info) {
  LOOPSB_REQUIRE(d != nullptr && info != nullptr, "null argument");
  memset(info, 0, sizeof(*info));
  info->world = d->world; info->rank = d->rank;
  info->local_rows = d->rows; info->num_cols = d->cols; info->local_nnz = d->nnz;
  info->num_blocks = int32_t(d->blocks.size());
  for (size_t b = 0; b < d->blocks.size() && b < 8; ++b) info->block_nnz[b] = d->blocks[b].nnz;
  info->bytes = d->bytes;
  int v = 0;
  if (d->world > 1 && nccl() && nccl()->GetVersion) nccl()->GetVersion(&v);
  info->nccl_version = v;
  info->transport = d->blocks.size() == 1 ? 0 : (d->p2p.on ? 2 : 1);
  info->graphs_cached = int32_t(d->graphs.size());
  return LOOPSB_OK;
}

int loopsb_dist_x_full(const loopsb_dist_t* d, const float** x_full) {
  LOOPSB_REQUIRE(d != nullptr && x_full != nullptr, "null argument");
  *x_full = d->x_full;
  return LOOPSB_OK;
}

int loopsb_dist_probe(loopsb_dist_t* d, int32_t enable) {
  LOOPSB_REQUIRE(d != nullptr, "null argument");
  d->probing = enable != 0;
  return LOOPSB_OK;
}

int loopsb_dist_probe_read(loopsb_dist_t* d, float* comm_ms, float* kernel_ms, float* block_ms, int32_t capacity) {
  LOOPSB_REQUIRE(d != nullptr, "null argument");
  LOOPSB_CUDA_TRY(cudaDeviceSynchronize());
  float c = 0.0f, k = 0.0f;
  if (d->world > 1) LOOPSB_CUDA_TRY(cudaEventElapsedTime(&c, d->comm_begin, d->comm_end));
  for (size_t b = 0; b < d->blocks.size(); ++b) {
    float ms = 0.0f;
    LOOPSB_CUDA_TRY(cudaEventElapsedTime(&ms, d->blocks[b].k_begin, d->blocks[b].k_end));
    k += ms;
    if (block_ms && int32_t(b) < capacity) block_ms[b] = ms;
  }
  if (comm_ms) *comm_ms = c;
  if (kernel_ms) *kernel_ms = k;
  return LOOPSB_OK;
}

int loopsb_dist_spmv(loopsb_dist_t* d, const float* x_shard, float* y_shard, void* stream) {
  LOOPSB_REQUIRE(d != nullptr, "null argument");
  LOOPSB_REQUIRE(d->cols == 0 || x_shard != nullptr, "x shard is null");
  LOOPSB_REQUIRE(d->rows == 0 || y_shard != nullptr, "y shard is null");
  cudaStream_t s = as_stream(stream);
  const nccl_api* n = d->world > 1 ? nccl() : nullptr;
  if (d->world > 1 && !n) return LOOPSB_ERR_UNSUPPORTED;
  const size_t chunk = size_t(d->chunk_cols);
  const bool probe = d->probing;

  if (d->blocks.size() == 1) {
    col_block& b = d->blocks[0];
    if (d->world > 1) {
      if (probe) LOOPSB_CUDA_TRY(cudaEventRecord(d->comm_begin, s));
      NCCL_TRY(n->AllGather(x_shard, d->x_full, chunk, kNcclFloat32, d->comm, s));
      if (probe) LOOPSB_CUDA_TRY(cudaEventRecord(d->comm_end, s));
    } else if (chunk) {
      LOOPSB_CUDA_TRY(cudaMemcpyAsync(d->x_full, x_shard, chunk * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    if (probe) LOOPSB_CUDA_TRY(cudaEventRecord(b.k_begin, s));
    int rc = loopsb_spmv_f32(b.plan, b.values, b.indices, nullptr, d->x_full, y_shard, d->rows, d->cols, s);
    if (rc != LOOPSB_OK) return rc;
    if (probe) LOOPSB_CUDA_TRY(cudaEventRecord(b.k_end, s));
    return LOOPSB_OK;
  }

  if (d->p2p.on) {
    const uint32_t q = (++d->p2p.step) & 1u;
    // replaying the step as one CUDA graph is opt-in: measured on 8 B200s it did not beat the
    // ~70 individual calls (0.52 vs 0.50 ms per step), the step is bound by the chunks' arrival
    static const bool no_graphs = getenv("LOOPSB_DIST_GRAPH") == nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    if (!d->use_graphs || no_graphs || probe || s == nullptr || cap != cudaStreamCaptureStatusNone)
      return enqueue_split_step(d, x_shard, y_shard, s, q);
    for (auto& g : d->graphs)
      if (g.x == x_shard && g.y == y_shard && g.q == q) {
        LOOPSB_CUDA_TRY(cudaGraphLaunch(g.exec, s));
        return LOOPSB_OK;
      }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      (void)cudaGetLastError();
      d->use_graphs = false;
      return enqueue_split_step(d, x_shard, y_shard, s, q);
    }
    const int rc = enqueue_split_step(d, x_shard, y_shard, s, q);
    const cudaError_t ee = cudaStreamEndCapture(s, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc != LOOPSB_OK || ee != cudaSuccess || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      // capture refused (old driver, unsupported node): run this and every later step directly
      (void)cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      d->use_graphs = false;
      return enqueue_split_step(d, x_shard, y_shard, s, q);
    }
    cudaGraphDestroy(graph);
    if (d->graphs.size() >= 16) {              // callers cycling through many buffers: drop the oldest
      cudaGraphExecDestroy(d->graphs.front().exec);
      d->graphs.erase(d->graphs.begin());
    }
    d->graphs.push_back({x_shard, y_shard, q, exec});
    LOOPSB_CUDA_TRY(cudaGraphLaunch(exec, s));
    return LOOPSB_OK;
  }
  return enqueue_split_step(d, x_shard, y_shard, s, 0);
}

}  // extern "C"
