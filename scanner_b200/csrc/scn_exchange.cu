// scn_exchange.cu -- exchange of the per-retune-step detection records between the GPUs of one box over
// NVLink peer memory ("only the small per-step records are gathered").
//
// The reference has no counterpart (one process, one FFT, process.cpp:316-331); this is the collective a sweep
// sharded by retune step needs.  An NCCL all-gather does the job (scanner_b200/sweep.py, csrc/host/sweepProcessor)
// but costs a rendezvous per batch: its kernel needs an SM while the persistent fused kernel of the next batch owns
// every register file, and every rank waits for the slowest one once per batch.  Here every rank instead owns a
// WINDOW in its own HBM that all peers can write:
//
//   window = data[kSlots][world][n_steps * rec_words]  +  flags[kSlots][world]
//
// publish(seq): one small kernel stores this rank's partial records into slot (seq mod kSlots), row `rank`, of
//               EVERY peer's window (plain st.global on peer-mapped addresses: NVLink writes, fire and forget),
//               then __threadfence_system() and a flag store of `seq`.  No rank waits for any other.
// merge(seq):   one small kernel polls the `world` flags of its OWN window (local HBM/L2) until they reach seq and
//               folds the rows by sum / OR (same rule as merge_records_kernel).  Issued one batch late
//               (publish(i), then merge(i-1)) it never spins in practice and ranks run up to a batch apart.
// Contract that makes kSlots = 4 race free: every rank merges EVERY sequence number, and issues merge(s) before
// publish(s + 2) on the same stream.  (A peer's publish(s + 4) reuses the slot of s; it comes after that peer's
// merge(s + 2), which waited for this rank's publish(s + 2), which stream order puts after this rank's merge(s).)
// scn_exchange_merge refuses sequence numbers that break the rule.
//
// Windows are cudaMalloc'ed by the library; other PROCESSES map them with CUDA IPC handles
// (scn_exchange_handle / scn_exchange_connect_ipc -- the handles travel over whatever the host already has,
// torch.distributed in bench.py), other DEVICES OF THE SAME PROCESS with cudaDeviceEnablePeerAccess
// (scn_exchange_connect_local).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scanner_b200.h"

namespace scn {
int api_fail(int code, const char* fmt, ...);   // scn_api.cu: records the thread-local error text
}

namespace {

constexpr uint32_t kSlots = 4;
constexpr uint32_t kMaxWorld = 64;
constexpr long long kSpinLimitCycles = 20000000000ll;   // ~10 s at 2 GHz: a peer that never publishes is an error, not a hang

}  // namespace

struct scn_exchange {
  int device = 0;
  uint32_t rank = 0, world = 1, n_steps = 0, rec_words = 0, rec_total = 0;
  uint32_t* window = nullptr;                 // own window (cudaMalloc)
  size_t window_bytes = 0;
  std::vector<uint32_t*> peer;                // peer[r] = rank r's window as seen from this device (peer[rank] == window)
  std::vector<bool> ipc_opened;
  uint32_t** d_peer = nullptr;                // device copy of peer[]
  uint32_t* d_error = nullptr;                // set by a merge that gave up waiting
  uint32_t* d_scratch = nullptr;              // staging of the host-pointer forms
  uint32_t* d_scratch2 = nullptr;
  uint64_t seq = 0;                           // last published sequence number
  bool connected = false;
};

namespace {

__host__ __device__ inline size_t data_offset(uint32_t slot, uint32_t row, uint32_t world, uint32_t rec_total) {
  return (size_t(slot) * world + row) * rec_total;
}
__host__ __device__ inline size_t flag_offset(uint32_t slot, uint32_t row, uint32_t world, uint32_t rec_total) {
  return size_t(kSlots) * world * rec_total + size_t(slot) * world + row;
}

// One CTA's copy of a record table into a (peer) window: 8-byte stores when the table length allows (NVLink stores are
// fire-and-forget, so the instruction count is what matters: cfg4's table is 137 KB per rank).
__device__ __forceinline__ void copy_records(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t n) {
  if ((n & 1u) == 0u && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 7u) == 0u) {
    uint2* d2 = reinterpret_cast<uint2*>(dst);
    const uint2* s2 = reinterpret_cast<const uint2*>(src);
    for (uint32_t x = threadIdx.x; x < n / 2; x += blockDim.x) d2[x] = s2[x];
  } else {
    for (uint32_t x = threadIdx.x; x < n; x += blockDim.x) dst[x] = src[x];
  }
}

// grid = world CTAs: CTA d copies this rank's records into peer d's window and then raises its flag there.
__global__ void __launch_bounds__(256)
publish_records_kernel(const uint32_t* __restrict__ records, uint32_t* const* __restrict__ peers, uint32_t rank,
                       uint32_t world, uint32_t rec_total, uint32_t slot, uint32_t seq) {
  uint32_t* win = peers[blockIdx.x];
  copy_records(win + data_offset(slot, rank, world, rec_total), records, rec_total);
  __syncthreads();                            // the CTA's stores happen-before thread 0's fence (bar.sync), and a
  if (threadIdx.x == 0) {                     // system-scope fence is cumulative: ONE fence per CTA, not 256
    __threadfence_system();
    volatile uint32_t* flag = win + flag_offset(slot, rank, world, rec_total);
    *flag = seq;
  }
}

// Waits (bounded) until every rank's row of `slot` carries sequence number `seq`, then merges the rows:
// words 0, 1 of a record add, the mask words OR (scn_records.cu).  All reads of the window bypass L1.
__global__ void __launch_bounds__(256)
merge_published_kernel(const uint32_t* __restrict__ window, uint32_t world, uint32_t rec_total, uint32_t rec_words,
                       uint32_t slot, uint32_t seq, uint32_t* __restrict__ out, uint32_t* __restrict__ error) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    const volatile uint32_t* flag = window + flag_offset(slot, threadIdx.x, world, rec_total);
    const long long t0 = clock64();
    // sequence numbers are compared modulo 2^32 (a flag is never more than kSlots batches away from seq)
    while (int32_t(*flag - seq) < 0) {
      if (clock64() - t0 > kSpinLimitCycles) { s_ok = 0; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (!s_ok) {
    if (threadIdx.x == 0) atomicExch(error, seq ? seq : 1u);
    return;
  }
  __threadfence_system();
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < rec_total; x += gridDim.x * blockDim.x) {
    const bool is_sum = (x % rec_words) < 2;
    uint32_t v = 0;
    for (uint32_t r = 0; r < world; r++) {
      const uint32_t y = __ldcv(window + data_offset(slot, r, world, rec_total) + x);
      v = is_sum ? v + y : (v | y);
    }
    out[x] = v;
  }
}

// publish(seq) and merge(seq - 1) in ONE launch: CTAs [0, world) publish, the rest merge the previous batch.
__global__ void __launch_bounds__(512)
exchange_step_kernel(const uint32_t* __restrict__ records, uint32_t* const* __restrict__ peers, uint32_t rank,
                     uint32_t world, uint32_t rec_total, uint32_t rec_words, uint32_t slot, uint32_t seq,
                     const uint32_t* __restrict__ window, uint32_t prev_slot, uint32_t prev_seq,
                     uint32_t* __restrict__ out, uint32_t* __restrict__ error) {
  if (blockIdx.x < world) {
    uint32_t* win = peers[blockIdx.x];
    copy_records(win + data_offset(slot, rank, world, rec_total), records, rec_total);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      volatile uint32_t* flag = win + flag_offset(slot, rank, world, rec_total);
      *flag = seq;
    }
    return;
  }
  if (prev_seq == 0) return;
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    const volatile uint32_t* flag = window + flag_offset(prev_slot, threadIdx.x, world, rec_total);
    const long long t0 = clock64();
    while (int32_t(*flag - prev_seq) < 0) {
      if (clock64() - t0 > kSpinLimitCycles) { s_ok = 0; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (!s_ok) {
    if (threadIdx.x == 0) atomicExch(error, prev_seq);
    return;
  }
  __threadfence_system();
  const uint32_t mb = blockIdx.x - world, nmb = gridDim.x - world;
  for (uint32_t x = mb * blockDim.x + threadIdx.x; x < rec_total; x += nmb * blockDim.x) {
    const bool is_sum = (x % rec_words) < 2;
    uint32_t v = 0;
    for (uint32_t r = 0; r < world; r++) {
      const uint32_t y = __ldcv(window + data_offset(prev_slot, r, world, rec_total) + x);
      v = is_sum ? v + y : (v | y);
    }
    out[x] = v;
  }
}

#define SCN_XCUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return scn::api_fail(SCN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                           __FILE__, __LINE__);                                                 \
  } while (0)

int upload_peers(scn_exchange* x) {
  SCN_XCUDA(cudaSetDevice(x->device));
  if (!x->d_peer) SCN_XCUDA(cudaMalloc(&x->d_peer, sizeof(uint32_t*) * x->world));
  SCN_XCUDA(cudaMemcpy(x->d_peer, x->peer.data(), sizeof(uint32_t*) * x->world, cudaMemcpyHostToDevice));
  x->connected = true;
  return SCN_OK;
}

}  // namespace

extern "C" {

SCN_API int scn_exchange_create(int device, uint32_t rank, uint32_t world, uint32_t n_steps, uint32_t record_words,
                                scn_exchange** out) {
  if (!out) return scn::api_fail(SCN_ERR_INVALID, "exchange_create: out is NULL");
  *out = nullptr;
  if (world == 0 || world > kMaxWorld || rank >= world || n_steps == 0 || record_words < 3)
    return scn::api_fail(SCN_ERR_INVALID, "exchange_create: bad arguments (rank %u of %u, %u steps, %u words)", rank,
                         world, n_steps, record_words);
  SCN_XCUDA(cudaSetDevice(device));
  scn_exchange* x = new scn_exchange();
  x->device = device;
  x->rank = rank; x->world = world; x->n_steps = n_steps; x->rec_words = record_words;
  x->rec_total = n_steps * record_words;
  x->window_bytes = sizeof(uint32_t) * (size_t(kSlots) * world * x->rec_total + size_t(kSlots) * world);
  x->peer.assign(world, nullptr);
  x->ipc_opened.assign(world, false);
  cudaError_t e = cudaMalloc(&x->window, x->window_bytes);
  if (e == cudaSuccess) e = cudaMemset(x->window, 0, x->window_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&x->d_error, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(x->d_error, 0, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    scn_exchange_destroy(x);
    return scn::api_fail(SCN_ERR_CUDA, "exchange_create: %s", cudaGetErrorString(e));
  }
  x->peer[rank] = x->window;
  if (world == 1) {
    int rc = upload_peers(x);
    if (rc != SCN_OK) { scn_exchange_destroy(x); return rc; }
  }
  *out = x;
  return SCN_OK;
}

SCN_API int scn_exchange_handle(scn_exchange* x, unsigned char* handle) {
  if (!x || !handle) return scn::api_fail(SCN_ERR_INVALID, "exchange_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == SCN_IPC_HANDLE_BYTES, "SCN_IPC_HANDLE_BYTES must match CUDA");
  SCN_XCUDA(cudaSetDevice(x->device));
  cudaIpcMemHandle_t h;
  SCN_XCUDA(cudaIpcGetMemHandle(&h, x->window));
  std::memcpy(handle, &h, sizeof(h));
  return SCN_OK;
}

SCN_API int scn_exchange_connect_ipc(scn_exchange* x, const unsigned char* handles) {
  if (!x || !handles) return scn::api_fail(SCN_ERR_INVALID, "exchange_connect_ipc: NULL argument");
  SCN_XCUDA(cudaSetDevice(x->device));
  for (uint32_t r = 0; r < x->world; r++) {
    if (r == x->rank || x->peer[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + size_t(r) * SCN_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    SCN_XCUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[r] = static_cast<uint32_t*>(p);
    x->ipc_opened[r] = true;
  }
  return upload_peers(x);
}

SCN_API int scn_exchange_connect_local(scn_exchange* const* all, uint32_t world) {
  if (!all || world == 0) return scn::api_fail(SCN_ERR_INVALID, "exchange_connect_local: bad arguments");
  for (uint32_t a = 0; a < world; a++)
    if (!all[a] || all[a]->world != world || all[a]->rank != a)
      return scn::api_fail(SCN_ERR_INVALID, "exchange_connect_local: entry %u is not rank %u of %u", a, a, world);
  for (uint32_t a = 0; a < world; a++) {
    scn_exchange* x = all[a];
    SCN_XCUDA(cudaSetDevice(x->device));
    for (uint32_t b = 0; b < world; b++) {
      if (b == a) continue;
      if (all[b]->device != x->device) {
        int can = 0;
        SCN_XCUDA(cudaDeviceCanAccessPeer(&can, x->device, all[b]->device));
        if (!can)
          return scn::api_fail(SCN_ERR_CUDA, "device %d cannot access device %d as a peer", x->device, all[b]->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(all[b]->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess)
          return scn::api_fail(SCN_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", all[b]->device, cudaGetErrorString(e));
      }
      x->peer[b] = all[b]->window;
    }
    int rc = upload_peers(x);
    if (rc != SCN_OK) return rc;
  }
  return SCN_OK;
}

SCN_API int scn_exchange_publish(scn_exchange* x, const uint32_t* d_records, void* stream, uint64_t* seq_out) {
  if (!x || !d_records) return scn::api_fail(SCN_ERR_INVALID, "exchange_publish: NULL argument");
  if (!x->connected) return scn::api_fail(SCN_ERR_INVALID, "exchange_publish: peers are not connected");
  SCN_XCUDA(cudaSetDevice(x->device));
  const uint64_t seq = ++x->seq;
  publish_records_kernel<<<x->world, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_records, x->d_peer, x->rank, x->world, x->rec_total, uint32_t(seq % kSlots), uint32_t(seq));
  SCN_XCUDA(cudaGetLastError());
  if (seq_out) *seq_out = seq;
  return SCN_OK;
}

SCN_API int scn_exchange_step(scn_exchange* x, const uint32_t* d_records, uint32_t* d_merged_previous, void* stream,
                              uint64_t* seq_out) {
  if (!x || !d_records || !d_merged_previous) return scn::api_fail(SCN_ERR_INVALID, "exchange_step: NULL argument");
  if (!x->connected) return scn::api_fail(SCN_ERR_INVALID, "exchange_step: peers are not connected");
  SCN_XCUDA(cudaSetDevice(x->device));
  const uint64_t seq = ++x->seq;
  uint32_t mgrid = (x->rec_total + 511) / 512;
  if (mgrid > 32) mgrid = 32;
  exchange_step_kernel<<<x->world + mgrid, 512, 0, static_cast<cudaStream_t>(stream)>>>(
      d_records, x->d_peer, x->rank, x->world, x->rec_total, x->rec_words, uint32_t(seq % kSlots), uint32_t(seq),
      x->window, uint32_t((seq - 1) % kSlots), uint32_t(seq - 1), d_merged_previous, x->d_error);
  SCN_XCUDA(cudaGetLastError());
  if (seq_out) *seq_out = seq;
  return SCN_OK;
}

SCN_API int scn_exchange_merge(scn_exchange* x, uint64_t seq, uint32_t* d_merged, void* stream) {
  if (!x || !d_merged || seq == 0) return scn::api_fail(SCN_ERR_INVALID, "exchange_merge: bad arguments");
  if (seq > x->seq || x->seq - seq > 1)
    return scn::api_fail(SCN_ERR_INVALID, "exchange_merge: sequence %llu is not mergeable (last published %llu): "
                         "merge(s) must be issued before publish(s + 2)", (unsigned long long)seq,
                         (unsigned long long)x->seq);
  SCN_XCUDA(cudaSetDevice(x->device));
  uint32_t grid = (x->rec_total + 255) / 256;
  if (grid > 32) grid = 32;
  merge_published_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x->window, x->world, x->rec_total, x->rec_words, uint32_t(seq % kSlots), uint32_t(seq), d_merged, x->d_error);
  SCN_XCUDA(cudaGetLastError());
  return SCN_OK;
}

// Host-pointer forms for callers that keep their partial records on the host (csrc/host/sweepProcessor.cpp):
// upload + publish, and merge + download (synchronous).
SCN_API int scn_exchange_publish_host(scn_exchange* x, const uint32_t* host_records, uint64_t* seq_out) {
  if (!x || !host_records) return scn::api_fail(SCN_ERR_INVALID, "exchange_publish_host: NULL argument");
  SCN_XCUDA(cudaSetDevice(x->device));
  if (!x->d_scratch) SCN_XCUDA(cudaMalloc(&x->d_scratch, sizeof(uint32_t) * x->rec_total));
  SCN_XCUDA(cudaMemcpy(x->d_scratch, host_records, sizeof(uint32_t) * x->rec_total, cudaMemcpyHostToDevice));
  int rc = scn_exchange_publish(x, x->d_scratch, nullptr, seq_out);
  if (rc != SCN_OK) return rc;
  SCN_XCUDA(cudaStreamSynchronize(nullptr));       // the scratch is reused by the next call
  return SCN_OK;
}

SCN_API int scn_exchange_merge_host(scn_exchange* x, uint64_t seq, uint32_t* host_merged) {
  if (!x || !host_merged) return scn::api_fail(SCN_ERR_INVALID, "exchange_merge_host: NULL argument");
  SCN_XCUDA(cudaSetDevice(x->device));
  if (!x->d_scratch2) SCN_XCUDA(cudaMalloc(&x->d_scratch2, sizeof(uint32_t) * x->rec_total));
  int rc = scn_exchange_merge(x, seq, x->d_scratch2, nullptr);
  if (rc != SCN_OK) return rc;
  SCN_XCUDA(cudaMemcpy(host_merged, x->d_scratch2, sizeof(uint32_t) * x->rec_total, cudaMemcpyDeviceToHost));
  uint32_t bad = 0;
  rc = scn_exchange_status(x, &bad);
  if (rc != SCN_OK) return rc;
  if (bad) return scn::api_fail(SCN_ERR_CUDA, "exchange_merge_host: gave up waiting for a peer to publish sequence %u", bad);
  return SCN_OK;
}

SCN_API int scn_exchange_status(scn_exchange* x, uint32_t* timed_out_seq) {
  if (!x || !timed_out_seq) return scn::api_fail(SCN_ERR_INVALID, "exchange_status: NULL argument");
  SCN_XCUDA(cudaSetDevice(x->device));
  SCN_XCUDA(cudaMemcpy(timed_out_seq, x->d_error, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return SCN_OK;
}

SCN_API uint32_t scn_exchange_slots(void) { return kSlots; }

SCN_API int scn_exchange_destroy(scn_exchange* x) {
  if (!x) return SCN_OK;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (uint32_t r = 0; r < x->world; r++)
    if (r < x->ipc_opened.size() && x->ipc_opened[r] && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
  if (x->d_peer) cudaFree(x->d_peer);
  if (x->d_scratch) cudaFree(x->d_scratch);
  if (x->d_scratch2) cudaFree(x->d_scratch2);
  if (x->d_error) cudaFree(x->d_error);
  if (x->window) cudaFree(x->window);
  delete x;
  return SCN_OK;
}

}  // extern "C"
