// common.cuh — internals shared by the sm_100a engine's translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <map>
#include <unordered_map>
#include <vector>

#include "../../include/dfgpu.h"

namespace dfgpu {

// ---------------------------------------------------------------------------------------------
// errors: C++ exception inside the library, converted to (code, thread-local message) at the ABI.
// ---------------------------------------------------------------------------------------------
struct Error {
  int code;
  std::string msg;
};
[[noreturn]] inline void fail(int code, const std::string& m) { throw Error{code, m}; }

void set_last_error(const std::string& m);

#define DF_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      ::dfgpu::fail(_e == cudaErrorMemoryAllocation ? DFGPU_ERR_OOM : DFGPU_ERR_CUDA,              \
                    std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                        std::to_string(__LINE__) + " (" #expr ")");                                \
  } while (0)

template <class Fn>
int guarded(Fn&& fn) {
  try {
    fn();
    return DFGPU_OK;
  } catch (const Error& e) {
    set_last_error(e.msg);
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(std::string("internal: ") + e.what());
    return DFGPU_ERR_INTERNAL;
  }
}

const char* dtype_name(int dt);
int dtype_width(int dt);  // bytes; 0 for bool/utf8
inline bool is_signed_int(int dt) { return dt >= DFGPU_INT8 && dt <= DFGPU_INT64; }
inline bool is_unsigned_int(int dt) { return dt >= DFGPU_UINT8 && dt <= DFGPU_UINT64; }
inline bool is_int(int dt) { return dt >= DFGPU_INT8 && dt <= DFGPU_UINT64; }
inline bool is_float(int dt) { return dt == DFGPU_FLOAT32 || dt == DFGPU_FLOAT64; }
inline bool is_numeric(int dt) { return dt >= DFGPU_INT8 && dt <= DFGPU_FLOAT64; }

}  // namespace dfgpu

// ---------------------------------------------------------------------------------------------
// opaque handle definitions
// ---------------------------------------------------------------------------------------------
struct dfgpu_ctx {
  int device = 0;
  int sm_count = 148;
  size_t device_mem_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  int64_t launches = 0;
  bool force_direct_kernel = false;  // DFGPU_FP_KERNEL=direct: bypass the TMA pipeline (A/B testing)
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // device scratch shared by operators (error flags, counters); 64 x u64
  unsigned long long* d_scratch = nullptr;
  unsigned long long* h_scratch = nullptr;  // pinned mirror
  // pinned staging ring for uploads from pageable memory
  void* stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  // per-kernel profiling (dfgpu_profile_*): ring of event pairs around the dominant kernels
  static constexpr int kProfRing = 32;
  bool prof_on = false;
  cudaEvent_t prof_ev[kProfRing][2] = {};
  bool prof_pending[kProfRing] = {};
  int prof_next = 0;
  double prof_ms = 0.0;
  int64_t prof_n = 0;
  int prof_begin();          // returns ring slot or -1
  void prof_end(int slot);
  void prof_drain();
  // streamed host->host operators: copy-in / copy-out streams and a cache of pinned result buffers
  cudaStream_t stream_in = nullptr, stream_out = nullptr;
  struct HostBlock { void* p; size_t bytes; bool used; };
  std::vector<HostBlock> host_blocks;
  void* host_alloc(size_t bytes);   // pinned, cached
  void host_release(void* p);
  // multi-GPU
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;

  // kernels whose per-device function attributes (dynamic shared memory limit) were set through this
  // ctx: the attribute belongs to the device, so it is tracked per ctx and not per process
  std::vector<const void*> configured_kernels;
  bool first_use(const void* kernel) {
    for (const void* k : configured_kernels)
      if (k == kernel) return false;
    configured_kernels.push_back(kernel);
    return true;
  }

  // device memory: see api.cu
  static constexpr size_t kBigBlock = 256u << 10;
  std::map<size_t, std::vector<void*>> big_free;  // size class -> cached blocks
  std::unordered_map<void*, size_t> big_live;     // block -> size class
  size_t big_cached_bytes = 0;
  void release_cached();
  void* alloc(size_t bytes);
  void free(void* p);
  void use();  // cudaSetDevice(device)
};

struct DevColumn {
  int dtype = 0;
  void* values = nullptr;        // device
  size_t values_bytes = 0;
  uint8_t* validity = nullptr;   // device, bit 0 = row 0 (re-based to offset 0), or null
  int32_t* offsets = nullptr;    // utf8: device i32 offsets (re-based view keeps original values)
  int64_t null_count = 0;
};

namespace dfgpu {
// collectives over the ctx's NCCL communicator (api.cu); counts and offsets in u64 words, all on ctx->stream
void comm_allgather_u64(dfgpu_ctx* ctx, const unsigned long long* send, unsigned long long* recv, size_t count);
void comm_exchange_v(dfgpu_ctx* ctx, const unsigned long long* send, const size_t* send_off, const size_t* send_cnt,
                     unsigned long long* recv, const size_t* recv_off, const size_t* recv_cnt);
void comm_allgather_v(dfgpu_ctx* ctx, const unsigned long long* send, unsigned long long* recv, const size_t* off, const size_t* cnt);
void comm_allgather_bytes_v(dfgpu_ctx* ctx, const void* send, void* recv, const size_t* off, const size_t* cnt);
void comm_allreduce_aggs(dfgpu_ctx* ctx, int naggs, const int* funcs, const int* mtypes, unsigned long long* d_vals, unsigned long long* d_nonnull,
                         unsigned long long* d_rows);
// one Utf8 column on the device (arrow 0.12 BinaryArray): the unit of the multi-source string gather
struct Utf8Source {
  const int* off;
  const unsigned char* bytes;
};
constexpr int UTF8_SRC_SHIFT = 40;  // gather index = (source << 40) | row
}  // namespace dfgpu

struct dfgpu_batch {
  dfgpu_ctx* ctx = nullptr;
  int64_t nrows = 0;
  std::vector<DevColumn> cols;
  bool owns = true;  // false: a view into buffers owned elsewhere (chunks of dfgpu_aggregate_update_host)
  ~dfgpu_batch();  // returns the column buffers to the ctx pool
};

struct dfgpu_result {
  dfgpu_ctx* ctx = nullptr;
  int64_t nrows = 0;
  std::vector<DevColumn> cols;
  bool on_host = false;  // columns live in pinned host memory (dfgpu_filter_project_host)
  ~dfgpu_result();
};
