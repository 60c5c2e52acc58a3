#include "common.h"

#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <set>
#include <unordered_map>
#include <utility>
#include <vector>

namespace hoigen {

static thread_local char g_err[512] = "";
static std::atomic<int> g_num_sms[64];   // per device ordinal, filled by hoigen_init

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  const int n = g_num_sms[dev].load(std::memory_order_relaxed);
  return n > 0 ? n : 148;
}

int set_max_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  HOIGEN_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({kernel, dev})) return HOIGEN_OK;
  HOIGEN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({kernel, dev});
  return HOIGEN_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int resolve_driver() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return -1;
  }
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return 0;
}

struct TmapKey {
  uint64_t v[10];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 10; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return size_t(h);
  }
};
// CUtensorMap must stay at a stable 64-byte aligned address: heap-allocate each entry.
static std::unordered_map<TmapKey, CUtensorMap*, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

static const CUtensorMap* get_tmap(int rank, const void* ptr, const uint64_t* dims, const uint64_t* strides,
                                   const uint32_t* box) {
  if (resolve_driver() != 0) return nullptr;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.v[0] = reinterpret_cast<uint64_t>(ptr);
  key.v[1] = uint64_t(rank);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[5 + i] = (uint64_t(box[i]) << 40) ^ (i > 0 ? strides[i - 1] : 0);
  }
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return it->second;
  // Bound the cache for callers that keep passing fresh buffers: retire a full generation, free it one generation later
  // (descriptors are copied into kernel parameters at launch, and no single call creates thousands of them, so a
  // pointer handed out earlier in the same C-ABI call stays valid).
  static std::vector<CUtensorMap*> retired;
  if (g_tmaps.size() >= 8192) {
    for (CUtensorMap* old : retired) free(old);
    retired.clear();
    for (auto& kv : g_tmaps) retired.push_back(kv.second);
    g_tmaps.clear();
  }
  CUtensorMap* m = nullptr;
  if (posix_memalign(reinterpret_cast<void**>(&m), 64, sizeof(CUtensorMap)) != 0) {
    set_error("posix_memalign failed");
    return nullptr;
  }
  cuuint64_t gdims[3];
  cuuint64_t gstrides[2];
  cuuint32_t gbox[3];
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    if (i > 0) gstrides[i - 1] = strides[i - 1];
  }
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cuuint32_t(rank), const_cast<void*>(ptr), gdims,
                        gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d ptr %p dims [%llu,%llu,%llu] stride1 %llu box [%u,%u,%u]",
              int(r), rank, ptr, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides[0] : 0),
              box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    free(m);
    return nullptr;
  }
  g_tmaps.emplace(key, m);
  return m;
}

const CUtensorMap* get_tmap_2d_bf16(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                                    uint32_t box0, uint32_t box1) {
  uint64_t dims[2] = {dim0, dim1};
  uint64_t strides[1] = {stride1_bytes};
  uint32_t box[2] = {box0, box1};
  return get_tmap(2, ptr, dims, strides, box);
}

const CUtensorMap* get_tmap_3d_bf16(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                                    uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0,
                                    uint32_t box1, uint32_t box2) {
  uint64_t dims[3] = {dim0, dim1, dim2};
  uint64_t strides[2] = {stride1_bytes, stride2_bytes};
  uint32_t box[3] = {box0, box1, box2};
  return get_tmap(3, ptr, dims, strides, box);
}

// ---------------------------------------------------------------------------------------------------
// launch counter + optional per-launch CUDA-event profiler
// ---------------------------------------------------------------------------------------------------
struct ProfRec {
  const char* tag;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static const int kMaxProf = 8192;
static ProfRec g_prof[kMaxProf];
static int g_prof_n = 0;
static int g_prof_events = 0;  // events created so far (reused across resets)
static std::atomic<bool> g_prof_on{false};
static std::atomic<long long> g_launches{0};
static std::mutex g_prof_mu;   // callers may launch from several host threads

KernelScope::KernelScope(const char* tag, cudaStream_t stream, double flops, double bytes) : slot_(-1), stream_(stream) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (g_prof_n >= kMaxProf) return;
  ProfRec& r = g_prof[g_prof_n];
  if (g_prof_n >= g_prof_events) {
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    g_prof_events = g_prof_n + 1;
  }
  r.tag = tag; r.flops = flops; r.bytes = bytes;
  cudaEventRecord(r.e0, stream);
  slot_ = g_prof_n++;
}
KernelScope::~KernelScope() {
  if (slot_ >= 0) cudaEventRecord(g_prof[slot_].e1, stream_);
}

StreamKWorkspace get_streamk_workspace(cudaStream_t stream, size_t slot_bytes, int num_flags) {
  struct Entry { StreamKWorkspace ws; size_t bytes; int flags; };
  static std::mutex mu;
  static std::unordered_map<cudaStream_t, Entry> table;
  std::lock_guard<std::mutex> lock(mu);
  auto it = table.find(stream);
  if (it != table.end() && it->second.bytes >= slot_bytes && it->second.flags >= num_flags) return it->second.ws;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    set_error("stream-K workspace must be created before stream capture starts (run the path once, uncaptured, first)");
    return StreamKWorkspace{nullptr, nullptr};
  }
  if (it != table.end()) {   // grow: the old launches on this stream must be done with the old buffers
    cudaStreamSynchronize(stream);
    cudaFree(it->second.ws.slots);
    cudaFree(it->second.ws.flags);
    table.erase(it);
  }
  Entry e{{nullptr, nullptr}, slot_bytes, num_flags};
  cudaError_t err = cudaMalloc(reinterpret_cast<void**>(&e.ws.slots), slot_bytes);
  if (err == cudaSuccess) err = cudaMalloc(reinterpret_cast<void**>(&e.ws.flags), size_t(num_flags) * sizeof(int));
  if (err == cudaSuccess) err = cudaMemset(e.ws.flags, 0, size_t(num_flags) * sizeof(int));
  if (err != cudaSuccess) {
    set_error("stream-K workspace allocation failed: %s", cudaGetErrorString(err));
    cudaFree(e.ws.slots);
    cudaFree(e.ws.flags);
    return StreamKWorkspace{nullptr, nullptr};
  }
  table.emplace(stream, e);
  return e.ws;
}

}  // namespace hoigen

extern "C" {

long long hoigen_launch_count(void) { return hoigen::g_launches.load(); }

int hoigen_profile_enable(int on) {
  hoigen::g_prof_on = on != 0;
  return HOIGEN_OK;
}

int hoigen_profile_reset(void) {
  std::lock_guard<std::mutex> lock(hoigen::g_prof_mu);
  hoigen::g_prof_n = 0;
  hoigen::g_launches = 0;
  return HOIGEN_OK;
}

// Writes one line per recorded launch: "tag ms flops bytes start_ms\n". Synchronises the device. Returns bytes written.
long long hoigen_profile_read(char* buf, long long cap) {
  using namespace hoigen;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  long long off = 0;
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f, t0 = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof[i].e0, g_prof[i].e1) != cudaSuccess) ms = -1.f;
    if (cudaEventElapsedTime(&t0, g_prof[0].e0, g_prof[i].e0) != cudaSuccess) t0 = -1.f;
    int n = snprintf(buf + off, size_t(cap - off), "%s %.6f %.0f %.0f %.6f\n", g_prof[i].tag, ms, g_prof[i].flops, g_prof[i].bytes, t0);
    if (n < 0 || off + n >= cap) break;
    off += n;
  }
  return off;
}

int hoigen_abi_version(void) { return HOIGEN_ABI_VERSION; }

const char* hoigen_last_error(void) { return hoigen::g_err; }

int hoigen_init(int device) {
  cudaDeviceProp prop;
  HOIGEN_CHECK_CUDA(cudaSetDevice(device));
  HOIGEN_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    hoigen::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                      prop.minor);
    return HOIGEN_ERR_ARCH;
  }
  if (device >= 0 && device < 64) hoigen::g_num_sms[device] = prop.multiProcessorCount;
  if (hoigen::resolve_driver() != 0) return HOIGEN_ERR_CUDA;
  return HOIGEN_OK;
}

}  // extern "C"
