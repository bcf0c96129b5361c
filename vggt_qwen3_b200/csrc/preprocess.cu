// f2: the image preprocessing that feeds the visual-encode path, on the GPU from uint8 pixels.
//
// Reference: src/dataio/collate_multiview.py:12-19 — transforms.Resize(image_size, BICUBIC) on a PIL image (i.e.
// PIL.Image.resize: separable antialiased convolution, bicubic a = -0.5, 22-bit fixed-point taps, uint8 rounding after
// each pass, horizontal pass first), transforms.CenterCrop(image_size), transforms.ToTensor().  The reference runs it
// per view on the host in the training process (single-threaded PIL) and ships fp32 pixels; here the host ships the
// uint8 views (4x fewer bytes over PCIe) and two HBM-bound kernels reproduce the arithmetic bit for bit
// (integer accumulation, same taps), computing only the rows / columns that survive the centre crop.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace vq {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow Resample.c PRECISION_BITS

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole-image box, restricted to the output indices
// [first, first + count).  table: count x (2 + ksize) int32 = {xmin, n, taps...}.
static int precompute_taps(int in_size, int out_size, int first, int count, std::vector<int>& table) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  table.assign((size_t)count * (2 + ksize), 0);
  std::vector<double> k(ksize);
  for (int i = 0; i < count; ++i) {
    const int xx = first + i;
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    int* row = table.data() + (size_t)i * (2 + ksize);
    row[0] = xmin;
    row[1] = xmax;
    for (int x = 0; x < xmax; ++x) {
      double v = k[x];
      if (ww != 0.0) v /= ww;
      row[2 + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
  }
  return ksize;
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> kPrecisionBits;   // arithmetic shift, as Pillow's lookup table indexes it
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: mid[f, y, x, c] (x in the cropped column range) from in[f, y, :, c]; one thread per output byte
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ mid, const int* __restrict__ taps, int ksize,
                long long total, int W, int size) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % 3);
  const long long p = i / 3;
  const int x = (int)(p % size);
  const long long fy = p / size;   // f * H + y
  const int* t = taps + (long long)x * (2 + ksize);
  const int x0 = t[0], n = t[1];
  const uint8_t* src = in + (fy * W + x0) * 3 + c;
  int acc = 1 << (kPrecisionBits - 1);
  for (int k = 0; k < n; ++k) acc += t[2 + k] * (int)src[3 * k];
  mid[i] = (uint8_t)clip8(acc);
}

// vertical pass + ToTensor: out[f, c, y, x] = u8 / 255 with u8 from mid[f, :, x, c]; one thread per output element,
// x fastest so that the f32 stores of a warp are contiguous
__global__ void __launch_bounds__(256)
resize_v_totensor_kernel(const uint8_t* __restrict__ mid, float* __restrict__ out, const int* __restrict__ taps,
                         int ksize, long long total, int H, int size) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % size);
  const long long r = i / size;
  const int y = (int)(r % size);
  const long long fc = r / size;   // f * 3 + c
  const int c = (int)(fc % 3);
  const long long f = fc / 3;
  const int* t = taps + (long long)y * (2 + ksize);
  const int y0 = t[0], n = t[1];
  const uint8_t* src = mid + ((f * H + y0) * size + x) * 3 + c;
  int acc = 1 << (kPrecisionBits - 1);
  for (int k = 0; k < n; ++k) acc += t[2 + k] * (int)src[(long long)k * size * 3];
  out[i] = __fdiv_rn((float)clip8(acc), 255.0f);
}

struct PreGeom {
  int nh, nw, top, left, kh, kv;   // resized size, crop offsets, tap counts
  size_t taps_h_bytes, taps_v_bytes, mid_bytes;
};

// torchvision _compute_resized_output_size (int size, no max_size) + center_crop offsets
static bool pre_geometry(int frames, int H, int W, int size, PreGeom& g) {
  if (frames <= 0 || H <= 0 || W <= 0 || size <= 0) return false;
  const int shrt = W <= H ? W : H, lng = W <= H ? H : W;
  const int new_long = (int)((double)size * lng / shrt);
  g.nh = W <= H ? new_long : size;
  g.nw = W <= H ? size : new_long;
  g.top = (int)nearbyint((g.nh - size) / 2.0);    // python round(): half to even, the default FP rounding mode
  g.left = (int)nearbyint((g.nw - size) / 2.0);
  auto ksize = [](int in_size, int out_size) {
    const double scale = (double)in_size / out_size;
    return (int)ceil(2.0 * (scale < 1.0 ? 1.0 : scale)) * 2 + 1;
  };
  g.kh = ksize(W, g.nw);
  g.kv = ksize(H, g.nh);
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  g.taps_h_bytes = up((size_t)size * (2 + g.kh) * sizeof(int));
  g.taps_v_bytes = up((size_t)size * (2 + g.kv) * sizeof(int));
  g.mid_bytes = up((size_t)frames * H * size * 3);
  return true;
}

}  // namespace vq

using namespace vq;

extern "C" size_t vq_preprocess_workspace_bytes(int32_t frames, int32_t H, int32_t W, int32_t image_size) {
  PreGeom g;
  if (!pre_geometry(frames, H, W, image_size, g)) {
    set_error("invalid argument: preprocess geometry");
    return 0;
  }
  return g.taps_h_bytes + g.taps_v_bytes + g.mid_bytes;
}

// Host-only helpers (no GPU needed): the geometry and the fixed-point taps the kernels use, for inspection / tests.
extern "C" int vq_preprocess_geometry(int32_t H, int32_t W, int32_t image_size, int32_t* out6) {
  VQ_CHECK_ARG(out6 != nullptr, "null pointer");
  PreGeom g;
  VQ_CHECK_ARG(pre_geometry(1, H, W, image_size, g), "preprocess geometry");
  out6[0] = g.nh; out6[1] = g.nw; out6[2] = g.top; out6[3] = g.left; out6[4] = g.kh; out6[5] = g.kv;
  return VQ_OK;
}

extern "C" int vq_preprocess_taps(int32_t in_size, int32_t out_size, int32_t first, int32_t count, int32_t* table,
                                  int64_t capacity) {
  VQ_CHECK_ARG(table && in_size > 0 && out_size > 0 && first >= 0 && count > 0 && first + count <= out_size,
               "preprocess taps arguments");
  std::vector<int> t;
  const int ksize = precompute_taps(in_size, out_size, first, count, t);
  VQ_CHECK_ARG((int64_t)t.size() <= capacity, "table capacity");
  for (size_t i = 0; i < t.size(); ++i) table[i] = t[i];
  return ksize;   // > 0: taps per output index; each row is {first input index, tap count, taps...}
}

extern "C" int vq_preprocess_views(const uint8_t* views_hwc, float* out_chw, int32_t frames, int32_t H, int32_t W,
                                   int32_t image_size, void* workspace, size_t workspace_bytes, vq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VQ_CHECK_ARG(views_hwc && out_chw && workspace, "null pointer");
  PreGeom g;
  VQ_CHECK_ARG(pre_geometry(frames, H, W, image_size, g), "preprocess geometry");
  VQ_CHECK_ARG(workspace_bytes >= g.taps_h_bytes + g.taps_v_bytes + g.mid_bytes, "workspace too small");
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) % 16) == 0, "workspace alignment");
  std::vector<int> th, tv;
  precompute_taps(W, g.nw, g.left, image_size, th);
  precompute_taps(H, g.nh, g.top, image_size, tv);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int* d_th = reinterpret_cast<int*>(ws);
  int* d_tv = reinterpret_cast<int*>(ws + g.taps_h_bytes);
  uint8_t* mid = ws + g.taps_h_bytes + g.taps_v_bytes;
  // pageable host memory: cudaMemcpyAsync stages the bytes before it returns, so the vectors may die with this call
  VQ_CHECK_CUDA(cudaMemcpyAsync(d_th, th.data(), th.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  VQ_CHECK_CUDA(cudaMemcpyAsync(d_tv, tv.data(), tv.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  {
    const long long total = (long long)frames * H * image_size * 3;
    LaunchScope scope(kKindRow, "resize_h", frames, H, W, 0, (double)frames * H * ((double)W + image_size) * 3, stream);
    resize_h_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(views_hwc, mid, d_th, g.kh, total, W, image_size);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  {
    const long long total = (long long)frames * 3 * image_size * image_size;
    LaunchScope scope(kKindRow, "resize_v_totensor", frames, image_size, image_size, 0,
                      (double)frames * image_size * 3.0 * (H + 4.0 * image_size), stream);
    resize_v_totensor_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(mid, out_chw, d_tv, g.kv, total, H,
                                                                                   image_size);
  }
  VQ_CHECK_CUDA(cudaGetLastError());
  return VQ_OK;
}
