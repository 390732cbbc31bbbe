// scn_nccl.cu -- in-process, multi-device NCCL gather of the per-retune-step records (SURVEY.md section 8e:
// "single process, ncclCommInitAll over 1/2/4/8 devices, one all-gather of the small records per sweep").
//
// Used by the C++ host (csrc/host/sweepProcessor.cpp), which replaces the reference's two worker threads on one FFT
// (process.cpp:316-331) with one worker per GPU.  Every device contributes its partial record table
// [n_steps][record_words]; one grouped ncclAllGather over NVLink puts all partial tables on every device and the
// merge kernel of scn_records.cu folds them (sum of words 0,1; OR of the mask words).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, so a process
// that already carries its own NCCL (PyTorch bundles one) keeps using that copy and a host without NCCL can still
// use everything else (the NVLink peer-memory exchange of scn_exchange.cu needs no NCCL at all).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <vector>

#include "../../include/scanner_b200.h"

namespace scn {
int api_fail(int code, const char* fmt, ...);
cudaError_t launch_merge(const uint32_t* parts, uint32_t n_parts, uint32_t n_steps, uint32_t rec_words,
                         uint32_t* out, cudaStream_t stream);
}

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetErrorString; }
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy this process already has, if any
    if (!a.lib) a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) {
      a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(a.lib, "ncclCommInitAll"));
      a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
      a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.lib, "ncclAllGather"));
      a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(a.lib, "ncclGroupStart"));
      a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(a.lib, "ncclGroupEnd"));
      a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
    }
    return a;
  }();
  return api;
}

#define SCN_NCUDA(expr)                                                                                    \
  do {                                                                                                     \
    cudaError_t e_ = (expr);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      return scn::api_fail(SCN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define SCN_NCCL(expr)                                                                                     \
  do {                                                                                                     \
    ncclResult_t r_ = (expr);                                                                              \
    if (r_ != ncclSuccess)                                                                                 \
      return scn::api_fail(SCN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

}  // namespace

struct scn_gather {
  std::vector<int> devices;
  uint32_t n_steps = 0, rec_words = 0;
  size_t rec_total = 0;
  std::vector<ncclComm_t> comms;
  std::vector<cudaStream_t> streams;
  std::vector<uint32_t*> d_part, d_all, d_merged;
  std::vector<uint32_t> h_check;
};

extern "C" {

SCN_API int scn_nccl_gather_create(const int* devices, uint32_t n_devices, uint32_t n_steps, uint32_t record_words,
                                   scn_gather** out) {
  if (!out) return scn::api_fail(SCN_ERR_INVALID, "nccl_gather_create: out is NULL");
  *out = nullptr;
  if (!devices || n_devices == 0 || n_steps == 0 || record_words < 3)
    return scn::api_fail(SCN_ERR_INVALID, "nccl_gather_create: bad arguments");
  if (!nccl().ok()) return scn::api_fail(SCN_ERR_CUDA, "NCCL is not available (dlopen libnccl.so.2: %s)", dlerror());
  scn_gather* g = new scn_gather();
  g->devices.assign(devices, devices + n_devices);
  g->n_steps = n_steps;
  g->rec_words = record_words;
  g->rec_total = size_t(n_steps) * record_words;
  g->comms.assign(n_devices, nullptr);
  g->streams.assign(n_devices, nullptr);
  g->d_part.assign(n_devices, nullptr);
  g->d_all.assign(n_devices, nullptr);
  g->d_merged.assign(n_devices, nullptr);
  auto bail = [&](int rc) { scn_nccl_gather_destroy(g); return rc; };
  ncclResult_t r = nccl().CommInitAll(g->comms.data(), int(n_devices), g->devices.data());
  if (r != ncclSuccess) {
    for (auto& c : g->comms) c = nullptr;
    return bail(scn::api_fail(SCN_ERR_CUDA, "ncclCommInitAll over %u devices failed: %s", n_devices, nccl().GetErrorString(r)));
  }
  for (uint32_t d = 0; d < n_devices; d++) {
    cudaError_t e = cudaSetDevice(g->devices[d]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->streams[d], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_part[d], sizeof(uint32_t) * g->rec_total);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_all[d], sizeof(uint32_t) * g->rec_total * n_devices);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_merged[d], sizeof(uint32_t) * g->rec_total);
    if (e != cudaSuccess) return bail(scn::api_fail(SCN_ERR_CUDA, "nccl_gather_create: device %d: %s", g->devices[d], cudaGetErrorString(e)));
  }
  *out = g;
  return SCN_OK;
}

SCN_API int scn_nccl_gather_merge_host(scn_gather* g, const uint32_t* const* host_partials, uint32_t* host_merged) {
  if (!g || !host_partials || !host_merged) return scn::api_fail(SCN_ERR_INVALID, "nccl_gather_merge_host: NULL argument");
  const uint32_t G = uint32_t(g->devices.size());
  const size_t bytes = sizeof(uint32_t) * g->rec_total;
  for (uint32_t d = 0; d < G; d++) {
    if (!host_partials[d]) return scn::api_fail(SCN_ERR_INVALID, "nccl_gather_merge_host: partial table %u is NULL", d);
    SCN_NCUDA(cudaSetDevice(g->devices[d]));
    SCN_NCUDA(cudaMemcpyAsync(g->d_part[d], host_partials[d], bytes, cudaMemcpyHostToDevice, g->streams[d]));
  }
  SCN_NCCL(nccl().GroupStart());
  for (uint32_t d = 0; d < G; d++) {
    ncclResult_t r = nccl().AllGather(g->d_part[d], g->d_all[d], g->rec_total, ncclUint32, g->comms[d], g->streams[d]);
    if (r != ncclSuccess) {
      nccl().GroupEnd();
      return scn::api_fail(SCN_ERR_CUDA, "ncclAllGather on device %d failed: %s", g->devices[d], nccl().GetErrorString(r));
    }
  }
  SCN_NCCL(nccl().GroupEnd());
  for (uint32_t d = 0; d < G; d++) {
    SCN_NCUDA(cudaSetDevice(g->devices[d]));
    SCN_NCUDA(scn::launch_merge(g->d_all[d], G, g->n_steps, g->rec_words, g->d_merged[d], g->streams[d]));
  }
  // every device now holds the merged table; device 0's copy is returned and the others are checked against it
  g->h_check.resize(g->rec_total);
  for (uint32_t d = 0; d < G; d++) {
    SCN_NCUDA(cudaSetDevice(g->devices[d]));
    SCN_NCUDA(cudaMemcpyAsync(d == 0 ? host_merged : g->h_check.data(), g->d_merged[d], bytes, cudaMemcpyDeviceToHost,
                              g->streams[d]));
    SCN_NCUDA(cudaStreamSynchronize(g->streams[d]));
    if (d > 0 && std::memcmp(host_merged, g->h_check.data(), bytes) != 0)
      return scn::api_fail(SCN_ERR_CUDA, "nccl_gather_merge_host: device %d disagrees with device %d after the all-gather",
                           g->devices[d], g->devices[0]);
  }
  return SCN_OK;
}

SCN_API int scn_nccl_gather_destroy(scn_gather* g) {
  if (!g) return SCN_OK;
  for (size_t d = 0; d < g->devices.size(); d++) {
    cudaSetDevice(g->devices[d]);
    if (g->streams[d]) { cudaStreamSynchronize(g->streams[d]); cudaStreamDestroy(g->streams[d]); }
    if (g->d_part[d]) cudaFree(g->d_part[d]);
    if (g->d_all[d]) cudaFree(g->d_all[d]);
    if (g->d_merged[d]) cudaFree(g->d_merged[d]);
    if (g->comms[d] && nccl().ok()) nccl().CommDestroy(g->comms[d]);
  }
  delete g;
  return SCN_OK;
}

}  // extern "C"
