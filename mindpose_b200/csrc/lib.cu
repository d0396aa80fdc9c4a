// lib.cu -- library plumbing of libposecodec: errors, device query, and the
// host-buffer front end (pc_ctx / pc_topdown_decode_host).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace pc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  cudaGetLastError();  // clear the sticky-less error state
  return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? PC_ERR_NO_DEVICE
                                                                    : PC_ERR_CUDA;
}

int sm_count_cached() {
  // per-device cache; devices do not change while the process lives
  static int counts[64];
  static std::once_flag once;
  std::call_once(once, [] {
    for (int& c : counts) c = 0;
  });
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (counts[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    counts[dev] = v;
  }
  return counts[dev];
}

// A pool keeps freed memory only if its release threshold says so, and raising the threshold
// of the device's DEFAULT pool would change how every other cudaMallocAsync user of the
// process returns memory to the driver: the library has its own.
cudaError_t scratch_pool(cudaMemPool_t* out) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool;
    e = cudaMemPoolCreate(&pool, &props);
    if (e != cudaSuccess) return e;
    uint64_t keep = UINT64_MAX;  // of this pool only: a call must not cost a driver allocation
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) {
      cudaMemPoolDestroy(pool);
      return e;
    }
    pools[dev] = pool;
  }
  *out = pools[dev];
  return cudaSuccess;
}

}  // namespace pc

using namespace pc;

extern "C" int pc_version(void) { return PC_VERSION; }

extern "C" const char* pc_last_error(void) { return g_err; }

extern "C" int pc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  PC_REQUIRE(device >= 0 && device < count, PC_ERR_NO_DEVICE,
             "pc_device_info: device %d not present (%d visible)", device, count);
  int sms = 0, maj = 0, mnr = 0;
  PC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  PC_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, device));
  PC_CUDA(cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, device));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = mnr;
  return PC_OK;
}

// ---------------------------------------------------------------------------
// Host-buffer front end.  Crops are cut into chunks; chunk c uses slot c % 2
// (its own stream and device buffers), so the host->device copy of chunk c+1
// runs while the kernel of chunk c is decoding.  Results are copied back on the
// slot's stream; one synchronisation per stream at the end.
// ---------------------------------------------------------------------------
struct pc_ctx {
  int device;
  int64_t scratch_bytes;  // per slot, for heatmap (+ flipped) planes
  cudaStream_t stream[2];
  unsigned char* d_maps[2];
  float* d_small[2];  // center | scale | score | preds | boxes per chunk
  int64_t small_floats;
  int64_t last_h2d_bytes, last_d2h_bytes;  // of the last *_host call
  // host-side tables of pc_topdown_affine_host (offsets, sizes, fetch jobs of one chunk)
  int64_t* h_off;
  int32_t* h_hw;
  int4* h_jobs;
  int64_t h_cap;
};

static const int64_t kMaxChunkCrops = 1 << 16;

extern "C" int pc_ctx_create(int device, int64_t scratch_bytes, pc_ctx** out) {
  PC_REQUIRE(out != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_ctx_create: out is NULL");
  *out = nullptr;
  PC_REQUIRE(scratch_bytes >= (1 << 20), PC_ERR_INVALID_ARGUMENT,
             "pc_ctx_create: scratch_bytes must be >= 1 MiB");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  PC_REQUIRE(device >= 0 && device < count, PC_ERR_NO_DEVICE,
             "pc_ctx_create: device %d not present (%d visible)", device, count);
  PC_CUDA(cudaSetDevice(device));
  pc_ctx* c = new pc_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->scratch_bytes = scratch_bytes / 2;
  // per crop: center 2 + scale 2 + score 1 + boxes 6, plus preds K*3 (K <= 64)
  c->small_floats = kMaxChunkCrops * (11 + 3 * PC_MAX_JOINTS);
  for (int s = 0; s < 2; ++s) {
    cudaError_t e1 = cudaStreamCreateWithFlags(&c->stream[s], cudaStreamNonBlocking);
    cudaError_t e2 = cudaMalloc((void**)&c->d_maps[s], (size_t)c->scratch_bytes);
    cudaError_t e3 = cudaMalloc((void**)&c->d_small[s], sizeof(float) * c->small_floats);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
      cudaError_t bad = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
      pc_ctx_destroy(c);
      return cuda_fail(bad, "pc_ctx_create allocation");
    }
  }
  *out = c;
  return PC_OK;
}

extern "C" int pc_ctx_destroy(pc_ctx* c) {
  if (!c) return PC_OK;
  cudaSetDevice(c->device);
  for (int s = 0; s < 2; ++s) {
    if (c->stream[s]) {
      cudaStreamSynchronize(c->stream[s]);
      cudaStreamDestroy(c->stream[s]);
    }
    if (c->d_maps[s]) cudaFree(c->d_maps[s]);
    if (c->d_small[s]) cudaFree(c->d_small[s]);
  }
  free(c->h_off);
  free(c->h_hw);
  free(c->h_jobs);
  delete c;
  return PC_OK;
}

extern "C" int pc_ctx_last_transfer_bytes(const pc_ctx* c, int64_t* h2d_bytes,
                                          int64_t* d2h_bytes) {
  PC_REQUIRE(c != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_ctx_last_transfer_bytes: ctx is NULL");
  if (h2d_bytes) *h2d_bytes = c->last_h2d_bytes;
  if (d2h_bytes) *d2h_bytes = c->last_d2h_bytes;
  return PC_OK;
}

extern "C" int pc_topdown_decode_host(pc_ctx* c, const float* h_heatmap, const float* h_flipped,
                                      const float* h_center, const float* h_scale,
                                      const float* h_score, float* h_all_preds,
                                      float* h_all_boxes, const pc_topdown_decode_params* p,
                                      int64_t n) {
  PC_REQUIRE(c != nullptr && p != nullptr, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_decode_host: ctx / params is NULL");
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode_host: n < 0");
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS && p->height >= 1 &&
                 p->width >= 1,
             PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode_host: bad num_joints / map size");
  if (n == 0) return PC_OK;
  PC_REQUIRE(h_heatmap && h_center && h_scale && h_score && h_all_preds && h_all_boxes &&
                 (!p->flip_test || h_flipped),
             PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode_host: NULL host pointer");
  PC_CUDA(cudaSetDevice(c->device));
  const int K = p->num_joints;
  const int64_t plane = (int64_t)K * p->height * p->width;  // floats per crop
  const int64_t crop_bytes = plane * 4 * (p->flip_test ? 2 : 1);
  int64_t chunk = c->scratch_bytes / crop_bytes;
  if (chunk > kMaxChunkCrops) chunk = kMaxChunkCrops;
  // at least ~4 chunks so copies and kernels overlap
  if (chunk > (n + 3) / 4) chunk = (n + 3) / 4;
  PC_REQUIRE(chunk >= 1, PC_ERR_UNSUPPORTED,
             "pc_topdown_decode_host: one crop (%lld bytes) exceeds the context scratch",
             (long long)crop_bytes);
  c->last_h2d_bytes = n * (crop_bytes + 5 * (int64_t)sizeof(float));
  c->last_d2h_bytes = n * (K * 3 + 6) * (int64_t)sizeof(float);
  // One chunk: every failure returns from the lambda only, so that the two streams are
  // always drained below before the caller gets its host buffers back.
  auto run_chunk = [&](int64_t i0, int slot) -> int {
    const int64_t m = (n - i0 < chunk) ? n - i0 : chunk;
    cudaStream_t st = c->stream[slot];
    float* d_hm = reinterpret_cast<float*>(c->d_maps[slot]);
    float* d_fl = p->flip_test ? d_hm + m * plane : nullptr;
    float* d_center = c->d_small[slot];
    float* d_scale = d_center + 2 * m;
    float* d_score = d_scale + 2 * m;
    float* d_boxes = d_score + m;
    float* d_preds = d_boxes + 6 * m;
    PC_CUDA(cudaMemcpyAsync(d_hm, h_heatmap + i0 * plane, sizeof(float) * m * plane,
                            cudaMemcpyHostToDevice, st));
    if (p->flip_test)
      PC_CUDA(cudaMemcpyAsync(d_fl, h_flipped + i0 * plane, sizeof(float) * m * plane,
                              cudaMemcpyHostToDevice, st));
    PC_CUDA(cudaMemcpyAsync(d_center, h_center + 2 * i0, sizeof(float) * 2 * m,
                            cudaMemcpyHostToDevice, st));
    PC_CUDA(cudaMemcpyAsync(d_scale, h_scale + 2 * i0, sizeof(float) * 2 * m,
                            cudaMemcpyHostToDevice, st));
    PC_CUDA(cudaMemcpyAsync(d_score, h_score + i0, sizeof(float) * m, cudaMemcpyHostToDevice,
                            st));
    const int rc =
        pc_topdown_decode(d_hm, d_fl, d_center, d_scale, d_score, d_preds, d_boxes, p, m, st);
    if (rc != PC_OK) return rc;
    PC_CUDA(cudaMemcpyAsync(h_all_preds + i0 * K * 3, d_preds, sizeof(float) * m * K * 3,
                            cudaMemcpyDeviceToHost, st));
    PC_CUDA(cudaMemcpyAsync(h_all_boxes + i0 * 6, d_boxes, sizeof(float) * m * 6,
                            cudaMemcpyDeviceToHost, st));
    return PC_OK;
  };
  int rc = PC_OK;
  int slot = 0;
  for (int64_t i0 = 0; i0 < n && rc == PC_OK; i0 += chunk, slot ^= 1) rc = run_chunk(i0, slot);
  // single exit: copies and kernels already enqueued may still touch the caller's buffers
  cudaError_t e0 = cudaStreamSynchronize(c->stream[0]);
  cudaError_t e1 = cudaStreamSynchronize(c->stream[1]);
  if (rc != PC_OK) return rc;
  if (e0 != cudaSuccess) return cuda_fail(e0, "cudaStreamSynchronize");
  if (e1 != cudaSuccess) return cuda_fail(e1, "cudaStreamSynchronize");
  return PC_OK;
}

// The rectangle [x0, x1) x [y0, y1) of source pixels one crop can sample, clipped to the
// image.  The destination rectangle maps to a rotated rectangle around the box centre
// (get_affine_transform: a similarity of ratio scale_w * pixel_std / image_w; get_warp_matrix:
// an anisotropic one of half sizes scale * pixel_std / 2), whose bounding box is taken and
// padded: 4 pixels cover the second bilinear tap, the fixed-point rounding of the warp
// (< 1/16 pixel) and the float32 / float64 differences to the device's matrix (< 1e-3
// pixel).  Returns false when the box is not finite or is degenerate (the caller then
// uploads the image).  tests/test_upload_rect.py checks it against the oracle warp.
static bool crop_source_rect(const float* box, float rot_deg, const pc_affine_host_params* p,
                             int* x0, int* y0, int* x1, int* y1) {
  double w = box[2], h = box[3];
  const double cx = (double)box[0] + w * 0.5, cy = (double)box[1] + h * 0.5;
  const double aspect = (double)p->image_w / (double)p->image_h;
  if (w > aspect * h)
    h = w / aspect;
  else if (w < aspect * h)
    w = h * aspect;
  const double sw = fabs(w * p->scale_padding), sh = fabs(h * p->scale_padding);
  // a degenerate box gives a singular matrix, whose "inverse" (all zeros in cv::warpAffine)
  // samples pixel (0, 0) for the whole crop: not a rectangle around the box
  if (!(sw >= 1.0 && sh >= 1.0)) return false;
  // the device computes centre and scale in float32: beyond 1e5 pixels its rounding (ulp
  // 0.008 at 1e5) is no longer small against the padding
  if (!(fabs(cx) <= 1e5 && fabs(cy) <= 1e5 && sw <= 1e5 && sh <= 1e5)) return false;
  // half sizes of the sampled rectangle along the crop's own axes, in source pixels
  const double a = 0.5 * sw;
  const double b = p->use_udp ? 0.5 * sh : 0.5 * sw * (double)p->image_h / (double)p->image_w;
  const double rad = (double)rot_deg * (3.141592653589793 / 180.0);
  const double cs = fabs(cos(rad)), sn = fabs(sin(rad));
  const double hx = cs * a + sn * b, hy = sn * a + cs * b;
  const double pad = 4.0;
  const double fx0 = floor(cx - hx - pad), fx1 = ceil(cx + hx + pad) + 1.0;
  const double fy0 = floor(cy - hy - pad), fy1 = ceil(cy + hy + pad) + 1.0;
  if (!(fx0 == fx0 && fx1 == fx1 && fy0 == fy0 && fy1 == fy1)) return false;  // NaN
  if (fabs(fx0) > 1e9 || fabs(fx1) > 1e9 || fabs(fy0) > 1e9 || fabs(fy1) > 1e9) return false;
  *x0 = (int)fmax(fx0, 0.0);
  *y0 = (int)fmax(fy0, 0.0);
  *x1 = (int)fmin(fx1, (double)p->src_w);
  *y1 = (int)fmin(fy1, (double)p->src_h);
  return true;
}

extern "C" int pc_crop_source_rect(const float* h_box, float rot,
                                   const pc_affine_host_params* p, int32_t* h_rect) {
  PC_REQUIRE(h_box && p && h_rect, PC_ERR_INVALID_ARGUMENT, "pc_crop_source_rect: NULL pointer");
  PC_REQUIRE(p->src_h >= 1 && p->src_w >= 1 && p->image_w >= 1 && p->image_h >= 1,
             PC_ERR_INVALID_ARGUMENT, "pc_crop_source_rect: bad sizes");
  int x0, y0, x1, y1;
  PC_REQUIRE(crop_source_rect(h_box, rot, p, &x0, &y0, &x1, &y1), PC_ERR_UNSUPPORTED,
             "pc_crop_source_rect: the box is not finite or is degenerate");
  h_rect[0] = x0;
  h_rect[1] = y0;
  h_rect[2] = x1;
  h_rect[3] = y1;
  return PC_OK;
}

// PC_UPLOAD_ROI_KERNEL: the rectangles of a chunk fetched over PCIe by the GPU itself.
// job = (first byte of a row's span, first row, span bytes (multiple of 16), rows) per crop;
// kUploadParts CTAs per crop, a warp per row, a lane per 16 bytes.  Nothing is computed: the
// point is one launch per chunk instead of one copy-engine launch per crop, with enough
// 16-byte reads in flight (148 SMs x 64 warps x 512 B) to keep the link busy.
constexpr int kUploadParts = 4;
constexpr int kUploadWarps = 8;
__global__ void __launch_bounds__(kUploadWarps * 32)
    upload_rects_kernel(const uint8_t* __restrict__ h_src, uint8_t* __restrict__ d_src,
                        const int4* __restrict__ jobs, int64_t src_bytes, int pitch) {
  const int64_t crop = blockIdx.x / kUploadParts;
  const int part = blockIdx.x - (int)(crop * kUploadParts);
  const int4 j = jobs[crop];
  const int nvec = j.z >> 4, pv = pitch >> 4;
  const int64_t base = crop * src_bytes + (int64_t)j.y * pitch + j.x;
  const uint4* s = reinterpret_cast<const uint4*>(h_src + base);
  uint4* d = reinterpret_cast<uint4*>(d_src + base);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = part * kUploadWarps + warp; r < j.w; r += kUploadParts * kUploadWarps) {
    const uint4* sr = s + (int64_t)r * pv;
    uint4* dr = d + (int64_t)r * pv;
    for (int v = lane; v < nvec; v += 32) dr[v] = __ldcs(sr + v);
  }
}

// Device-side address of a page-locked host range, or NULL if [p, p + bytes) is not
// page-locked memory this device can read.
static const uint8_t* mapped_host_range(const uint8_t* p, int64_t bytes) {
  cudaPointerAttributes a0, a1;
  if (cudaPointerGetAttributes(&a0, p) != cudaSuccess ||
      cudaPointerGetAttributes(&a1, p + bytes - 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer)
    return nullptr;
  return static_cast<const uint8_t*>(a0.devicePointer);
}

extern "C" int pc_topdown_affine_host(pc_ctx* c, const uint8_t* h_images, const float* h_boxes,
                                      const float* h_rot, uint8_t* h_crops, float* h_center,
                                      float* h_scale, const pc_affine_host_params* p,
                                      int64_t n) {
  PC_REQUIRE(c != nullptr && p != nullptr, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_affine_host: ctx / params is NULL");
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_topdown_affine_host: n < 0");
  PC_REQUIRE(p->src_h >= 1 && p->src_w >= 1 && p->channels >= 1 && p->channels <= 4 &&
                 p->image_w >= 1 && p->image_h >= 1,
             PC_ERR_INVALID_ARGUMENT, "pc_topdown_affine_host: bad sizes");
  PC_REQUIRE(p->upload >= PC_UPLOAD_FULL && p->upload <= PC_UPLOAD_ROI_KERNEL,
             PC_ERR_INVALID_ARGUMENT, "pc_topdown_affine_host: upload must be a PC_UPLOAD_* value");
  if (n == 0) return PC_OK;
  PC_REQUIRE(h_images && h_boxes && h_crops, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_affine_host: NULL host pointer");
  PC_CUDA(cudaSetDevice(c->device));
  const int64_t src_bytes = (int64_t)p->src_h * p->src_w * p->channels;
  const int64_t dst_bytes = (int64_t)p->image_h * p->image_w * p->channels;
  // per crop in the maps scratch: source image, crop, then (16-B aligned) offsets,
  // sizes and the two matrices
  const int64_t per_crop = ((src_bytes + dst_bytes + 15) / 16) * 16 + 16 + 8 + 8 + 48 + 48;
  int64_t chunk = (c->scratch_bytes - 64) / per_crop;
  if (chunk > kMaxChunkCrops) chunk = kMaxChunkCrops;
  if (chunk > (n + 3) / 4) chunk = (n + 3) / 4;
  PC_REQUIRE(chunk >= 1, PC_ERR_UNSUPPORTED,
             "pc_topdown_affine_host: one crop (%lld bytes) exceeds the context scratch",
             (long long)per_crop);
  pc_box_params bp = {p->image_w, p->image_h, p->pixel_std, p->scale_padding};
  pc_affine_params ap = {p->image_w, p->image_h, p->pixel_std, p->use_udp};
  pc_warp_params wp = {p->image_w, p->image_h, p->channels};
  // offsets / sizes are the same for every chunk: build once on the host (tables owned by
  // the context, freed by pc_ctx_destroy)
  if (c->h_cap < chunk) {
    free(c->h_off);
    free(c->h_hw);
    free(c->h_jobs);
    c->h_off = (int64_t*)malloc(sizeof(int64_t) * chunk);
    c->h_hw = (int32_t*)malloc(sizeof(int32_t) * 2 * chunk);
    c->h_jobs = (int4*)malloc(sizeof(int4) * chunk);
    c->h_cap = (c->h_off && c->h_hw && c->h_jobs) ? chunk : 0;
  }
  int64_t* const h_off = c->h_off;
  int32_t* const h_hw = c->h_hw;
  int4* const h_jobs = c->h_jobs;
  PC_REQUIRE(c->h_cap >= chunk, PC_ERR_CUDA, "pc_topdown_affine_host: out of host memory");
  // the fetch kernel needs the device mapping of a page-locked source and 16-byte geometry
  const int64_t pitch = (int64_t)p->src_w * p->channels;
  int upload = p->upload;
  const uint8_t* m_images = nullptr;
  if (upload == PC_UPLOAD_ROI_KERNEL) {
    m_images = mapped_host_range(h_images, n * src_bytes);
    if (!m_images || (pitch & 15) || (src_bytes & 15) ||
        (reinterpret_cast<uintptr_t>(m_images) & 15))
      upload = PC_UPLOAD_ROI;
  }
  for (int64_t i = 0; i < chunk; ++i) {
    h_off[i] = i * src_bytes;
    h_hw[2 * i] = p->src_h;
    h_hw[2 * i + 1] = p->src_w;
  }
  int64_t h2d = n * (int64_t)(16 + (h_rot ? 4 : 0) + 8 + 8);
  c->last_h2d_bytes = 0;
  c->last_d2h_bytes = n * (dst_bytes + (h_center ? 8 : 0) + (h_scale ? 8 : 0));
  // One chunk (failures return from the lambda only: the streams are drained below).  The
  // pageable h_jobs / h_off / h_hw tables may be rewritten for the next chunk at once:
  // cudaMemcpyAsync from pageable memory returns after staging the source.
  auto run_chunk = [&](int64_t i0, int slot) -> int {
    const int64_t m = (n - i0 < chunk) ? n - i0 : chunk;
    cudaStream_t st = c->stream[slot];
    unsigned char* base = c->d_maps[slot];
    uint8_t* d_src = base;
    uint8_t* d_dst = base + ((m * src_bytes + 15) / 16) * 16;
    unsigned char* q = d_dst + ((m * dst_bytes + 15) / 16) * 16;
    int4* d_jobs = (int4*)q;
    q += 16 * m;
    int64_t* d_off = (int64_t*)q;
    q += 8 * m;
    double* d_fwd = (double*)q;
    q += 48 * m;
    double* d_inv = (double*)q;
    q += 48 * m;
    int32_t* d_hw = (int32_t*)q;
    float* d_boxes = c->d_small[slot];
    float* d_center = d_boxes + 4 * m;
    float* d_scale = d_center + 2 * m;
    float* d_rot = d_scale + 2 * m;
    if (upload == PC_UPLOAD_FULL) {
      PC_CUDA(cudaMemcpyAsync(d_src, h_images + i0 * src_bytes, (size_t)(m * src_bytes),
                              cudaMemcpyHostToDevice, st));
      h2d += m * src_bytes;
    } else {
      // only the rectangle each crop samples, into the same full-image layout (the bytes
      // around it keep whatever the scratch held: the warp never reads them)
      for (int64_t i = 0; i < m; ++i) {
        int x0 = 0, y0 = 0, x1 = p->src_w, y1 = p->src_h;
        if (!crop_source_rect(h_boxes + 4 * (i0 + i), h_rot ? h_rot[i0 + i] : 0.f, p, &x0, &y0,
                              &x1, &y1)) {
          x0 = y0 = 0;
          x1 = p->src_w;
          y1 = p->src_h;
        }
        if (x1 <= x0 || y1 <= y0) x0 = x1 = y0 = y1 = 0;  // the crop is all border
        int64_t b0 = (int64_t)x0 * p->channels, b1 = (int64_t)x1 * p->channels;
        const int64_t rows = y1 - y0;
        if (upload == PC_UPLOAD_ROI_KERNEL) {
          b0 &= ~(int64_t)63;
          b1 = (b1 + 63) & ~(int64_t)63;
          if (b1 > pitch) b1 = pitch;
          h_jobs[i] = make_int4((int)b0, y0, (int)(b1 - b0), (int)rows);
        } else if (rows > 0) {
          const int64_t at = (int64_t)y0 * pitch + b0;
          const uint8_t* h_at = h_images + (i0 + i) * src_bytes + at;
          if (b1 - b0 == pitch)
            PC_CUDA(cudaMemcpyAsync(d_src + i * src_bytes + at, h_at, (size_t)(rows * pitch),
                                    cudaMemcpyHostToDevice, st));
          else
            PC_CUDA(cudaMemcpy2DAsync(d_src + i * src_bytes + at, (size_t)pitch, h_at,
                                      (size_t)pitch, (size_t)(b1 - b0), (size_t)rows,
                                      cudaMemcpyHostToDevice, st));
        }
        h2d += rows * (b1 - b0);
      }
      if (upload == PC_UPLOAD_ROI_KERNEL) {
        PC_CUDA(cudaMemcpyAsync(d_jobs, h_jobs, sizeof(int4) * m, cudaMemcpyHostToDevice, st));
        h2d += 16 * m;
        upload_rects_kernel<<<(unsigned)(m * kUploadParts), kUploadWarps * 32, 0, st>>>(
            m_images + i0 * src_bytes, d_src, d_jobs, src_bytes, (int)pitch);
        PC_CUDA(cudaGetLastError());
      }
    }
    PC_CUDA(cudaMemcpyAsync(d_boxes, h_boxes + 4 * i0, sizeof(float) * 4 * m,
                            cudaMemcpyHostToDevice, st));
    if (h_rot)
      PC_CUDA(cudaMemcpyAsync(d_rot, h_rot + i0, sizeof(float) * m, cudaMemcpyHostToDevice, st));
    PC_CUDA(cudaMemcpyAsync(d_off, h_off, sizeof(int64_t) * m, cudaMemcpyHostToDevice, st));
    PC_CUDA(cudaMemcpyAsync(d_hw, h_hw, sizeof(int32_t) * 2 * m, cudaMemcpyHostToDevice, st));
    int rc = pc_box_to_center_scale(d_boxes, d_center, d_scale, &bp, m, st);
    if (rc == PC_OK)
      rc = pc_affine_matrices(d_center, d_scale, h_rot ? d_rot : nullptr, d_fwd, d_inv, &ap, m,
                              st);
    if (rc == PC_OK) rc = pc_warp_affine_u8(d_src, d_off, d_hw, d_inv, d_dst, &wp, m, st);
    if (rc != PC_OK) return rc;
    PC_CUDA(cudaMemcpyAsync(h_crops + i0 * dst_bytes, d_dst, (size_t)(m * dst_bytes),
                            cudaMemcpyDeviceToHost, st));
    if (h_center)
      PC_CUDA(cudaMemcpyAsync(h_center + 2 * i0, d_center, sizeof(float) * 2 * m,
                              cudaMemcpyDeviceToHost, st));
    if (h_scale)
      PC_CUDA(cudaMemcpyAsync(h_scale + 2 * i0, d_scale, sizeof(float) * 2 * m,
                              cudaMemcpyDeviceToHost, st));
    return PC_OK;
  };
  int rc = PC_OK;
  int slot = 0;
  for (int64_t i0 = 0; i0 < n && rc == PC_OK; i0 += chunk, slot ^= 1) rc = run_chunk(i0, slot);
  c->last_h2d_bytes = h2d;
  // single exit: copies and kernels already enqueued may still touch the caller's buffers
  cudaError_t e0 = cudaStreamSynchronize(c->stream[0]);
  cudaError_t e1 = cudaStreamSynchronize(c->stream[1]);
  if (rc != PC_OK) return rc;
  if (e0 != cudaSuccess) return cuda_fail(e0, "cudaStreamSynchronize");
  if (e1 != cudaSuccess) return cuda_fail(e1, "cudaStreamSynchronize");
  return PC_OK;
}
