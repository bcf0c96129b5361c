#include <stdlib.h>

#include "common.cuh"

#include <cstring>
#include <list>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace vq {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

struct TmapKey {
  const void* base;
  uint64_t inner, outer, ld;
  uint32_t bi, bo;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && bi == o.bi &&
           bo == o.bo;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner);
    mix(k.outer);
    mix(k.ld);
    mix((uint64_t(k.bi) << 32) | k.bo);
    return static_cast<size_t>(h);
  }
};

// Bounded LRU cache: descriptors are keyed by every argument (base pointer included), so a caller that allocates fresh
// buffers every step would otherwise grow the cache without limit.  Entries are copied out under the lock, so eviction
// can never invalidate a descriptor a launch is about to use.
constexpr size_t kTmapCacheCapacity = 2048;

bool get_tmap_2d_bf16(const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                      uint32_t box_outer, CUtensorMap* out) {
  struct Entry { TmapKey key; CUtensorMap map; };
  static std::mutex mu;
  static std::list<Entry> lru;   // front = most recently used
  static std::unordered_map<TmapKey, std::list<Entry>::iterator, TmapKeyHash> index;
  TmapKey key{base, inner, outer, ld, box_inner, box_outer};
  std::lock_guard<std::mutex> lock(mu);
  auto it = index.find(key);
  if (it != index.end()) {
    lru.splice(lru.begin(), lru, it->second);
    *out = it->second->map;
    return true;
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return false;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * 2) & 15u) != 0 || box_inner != 64 ||
      box_outer == 0 || box_outer > 256 || inner == 0 || outer == 0) {
    set_error("tensor map: base/pitch must be 16-byte aligned, box = {64, 1..256}, extents > 0");
    return false;
  }
  Entry e;
  e.key = key;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * 2};  // bytes, dims 1..rank-1
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
    return false;
  }
  lru.push_front(e);
  index[key] = lru.begin();
  if (lru.size() > kTmapCacheCapacity) {
    index.erase(lru.back().key);
    lru.pop_back();
  }
  *out = e.map;
  return true;
}

cudaError_t splitk_workspace(cudaStream_t stream, size_t bytes, size_t n_flags, void** ws, unsigned** flags) {
  struct Slot { void* ws = nullptr; size_t bytes = 0; unsigned* flags = nullptr; size_t n_flags = 0; };
  static std::mutex mu;
  static std::unordered_map<uint64_t, Slot> slots;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t key = (reinterpret_cast<uint64_t>(stream) << 6) ^ (uint64_t)(dev & 63);
  std::lock_guard<std::mutex> lock(mu);
  Slot& s = slots[key];
  if (s.bytes < bytes) {
    if (s.ws) {
      if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
      cudaFree(s.ws);
      s.ws = nullptr; s.bytes = 0;
    }
    const size_t want = bytes < (size_t(20) << 20) ? (size_t(20) << 20) : bytes;
    if ((e = cudaMalloc(&s.ws, want)) != cudaSuccess) return e;
    s.bytes = want;
  }
  if (s.n_flags < n_flags) {
    if (s.flags) {
      if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
      cudaFree(s.flags);
      s.flags = nullptr; s.n_flags = 0;
    }
    const size_t want = n_flags < 1024 ? 1024 : n_flags;
    if ((e = cudaMalloc(reinterpret_cast<void**>(&s.flags), want * sizeof(unsigned))) != cudaSuccess) return e;
    if ((e = cudaMemset(s.flags, 0, want * sizeof(unsigned))) != cudaSuccess) return e;
    s.n_flags = want;
  }
  *ws = s.ws;
  *flags = s.flags;
  return cudaSuccess;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("VQ_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// per-device caches (a process may touch more than one GPU; function attributes and the SM count are per device)
constexpr int kMaxDevices = 64;

int num_sms() {
  static int n[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) v = 0;
    n[dev] = v;
  }
  return n[dev];
}

cudaError_t ensure_dynamic_smem(const void* func, int bytes, bool max_carveout) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, int> done;   // (device, function) -> bytes configured
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t key = (reinterpret_cast<uint64_t>(func) << 6) ^ (uint64_t)(dev & 63);
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  if (max_carveout) {
    e = cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  done[key] = bytes;
  return cudaSuccess;
}

// ------------------------------------------------------------------ launch accounting / profiling
struct ProfRecord {
  int kind;
  const char* tag;
  int m, n, k;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static long long g_launches = 0;

static cudaEvent_t take_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

LaunchScope::LaunchScope(int kind, const char* tag, int m, int n, int k, double flops, double bytes,
                         cudaStream_t stream)
    : slot_(-1), stream_(stream) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  ++g_launches;
  if (!g_prof_on) return;
  ProfRecord r{kind, tag, m, n, k, flops, bytes, take_event(), take_event()};
  cudaEventRecord(r.e0, stream);
  slot_ = (int)g_prof.size();
  g_prof.push_back(r);
}
LaunchScope::~LaunchScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  cudaEventRecord(g_prof[slot_].e1, stream_);
}

}  // namespace vq

extern "C" {
int vq_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(vq::g_prof_mu);
  for (auto& r : vq::g_prof) {
    vq::g_event_pool.push_back(r.e0);
    vq::g_event_pool.push_back(r.e1);
  }
  vq::g_prof.clear();
  vq::g_prof_on = on != 0;
  return VQ_OK;
}
long long vq_launch_count(void) {
  std::lock_guard<std::mutex> lock(vq::g_prof_mu);
  return vq::g_launches;
}
// Text table of the recorded launches aggregated by (kind, tag, m, n, k):
//   kind tag m n k launches total_ms total_flops total_bytes
// Synchronises on the recorded events.  Returns the number of bytes written (0 if buf too small).
long long vq_profile_dump(char* buf, long long cap) {
  std::lock_guard<std::mutex> lock(vq::g_prof_mu);
  struct Agg { int kind; const char* tag; int m, n, k; long long cnt; double ms, flops, bytes; };
  std::vector<Agg> aggs;
  for (auto& r : vq::g_prof) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
    Agg* a = nullptr;
    for (auto& x : aggs)
      if (x.kind == r.kind && x.tag == r.tag && x.m == r.m && x.n == r.n && x.k == r.k) { a = &x; break; }
    if (!a) {
      aggs.push_back(Agg{r.kind, r.tag, r.m, r.n, r.k, 0, 0, 0, 0});
      a = &aggs.back();
    }
    a->cnt += 1; a->ms += ms; a->flops += r.flops; a->bytes += r.bytes;
  }
  std::string out;
  char line[256];
  for (auto& a : aggs) {
    snprintf(line, sizeof(line), "%d %s %d %d %d %lld %.6f %.6e %.6e\n", a.kind, a.tag, a.m, a.n, a.k, a.cnt, a.ms,
             a.flops, a.bytes);
    out += line;
  }
  if ((long long)out.size() + 1 > cap) return 0;
  memcpy(buf, out.c_str(), out.size() + 1);
  return (long long)out.size();
}

const char* vq_last_error(void) { return vq::g_last_error.c_str(); }
int vq_version(void) { return 1; }
int vq_num_sms(void) { return vq::num_sms(); }
}
