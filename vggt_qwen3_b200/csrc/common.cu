#include "common.cuh"

#include <cstring>
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

const CUtensorMap* get_tmap_2d_bf16(const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                                    uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, std::unique_ptr<CUtensorMap>, TmapKeyHash> cache;
  TmapKey key{base, inner, outer, ld, box_inner, box_outer};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second.get();

  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return nullptr;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * 2) & 15u) != 0 || box_inner != 64 ||
      box_outer == 0 || box_outer > 256 || inner == 0 || outer == 0) {
    set_error("tensor map: base/pitch must be 16-byte aligned, box = {64, 1..256}, extents > 0");
    return nullptr;
  }
  std::unique_ptr<CUtensorMap> tm(new CUtensorMap);
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * 2};  // bytes, dims 1..rank-1
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm.get(), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
    return nullptr;
  }
  const CUtensorMap* out = tm.get();
  cache.emplace(key, std::move(tm));
  return out;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
  }
  return n;
}

}  // namespace vq

extern "C" {
const char* vq_last_error(void) { return vq::g_last_error.c_str(); }
int vq_version(void) { return 1; }
int vq_num_sms(void) { return vq::num_sms(); }
}
