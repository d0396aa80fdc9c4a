// warp_affine.cu -- crop geometry and OpenCV-exact affine crop warp for sm_100a.
//
// Replaces, for batches of crops:
//   TopDownBoxToCenterScale._xywh2cs   (data/transform/topdown_transform.py:131-154)
//   get_affine_transform               (data/transform/utils.py:44-98)
//   get_warp_matrix                    (data/transform/utils.py:158-190)
//   cv2.warpAffine(..., INTER_LINEAR)  (topdown_transform.py:217-222 / :248-253)
//   keypoint half of TopDownAffine     (topdown_transform.py:224-231 / :255-259)
//
// The warp reproduces OpenCV's integer pipeline (third-party, not under the
// reference tree): fp64 inverse matrix, 10-bit fixed-point coordinates reduced
// to 1/32 pixel, 15-bit bilinear weights, constant border 0.  With a = fx/32
// and b = fy/32 the four float32 weight products are exact multiples of 2^-10,
// so rint(w * 32768) = 32 * (32-fx or fx) * (32-fy or fy) and
//   dst = (sum(w * p) + 16384) >> 15 = (sum((..)(..) * p) + 512) >> 10.
#include <math.h>

#include "common.cuh"

namespace pc {

// ---------------------------------------------------------------- A1 -------
__global__ void box_to_center_scale_kernel(const float* __restrict__ boxes,
                                           float* __restrict__ center,
                                           float* __restrict__ scale, double aspect,
                                           float pixel_std, float scale_padding, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = boxes[4 * i], y = boxes[4 * i + 1];
  const float wf = boxes[4 * i + 2], hf = boxes[4 * i + 3];
  // centre: float32 box values, x + w * 0.5 evaluated in float32
  center[2 * i] = __fadd_rn(x, __fmul_rn(wf, 0.5f));
  center[2 * i + 1] = __fadd_rn(y, __fmul_rn(hf, 0.5f));
  // aspect fix-up: the float32 box value is compared against a float64 ratio;
  // the adjusted side becomes float64, the untouched side stays float32, and
  // each is divided by pixel_std in its own precision before the float32 store.
  const double ah = __dmul_rn(aspect, (double)hf);
  float sw, sh;
  if ((double)wf > ah) {
    sw = __fdiv_rn(wf, pixel_std);
    sh = (float)__ddiv_rn(__ddiv_rn((double)wf, aspect), (double)pixel_std);
  } else if ((double)wf < ah) {
    sw = (float)__ddiv_rn(__dmul_rn((double)hf, aspect), (double)pixel_std);
    sh = __fdiv_rn(hf, pixel_std);
  } else {
    sw = __fdiv_rn(wf, pixel_std);
    sh = __fdiv_rn(hf, pixel_std);
  }
  scale[2 * i] = __fmul_rn(sw, scale_padding);
  scale[2 * i + 1] = __fmul_rn(sh, scale_padding);
}

// ------------------------------------------------------------ A2 / A3 ------
__device__ __forceinline__ void invert_affine_cv(const double* m, double* o) {
  // cv::warpAffine's own inversion, op for op
  double d = __dsub_rn(__dmul_rn(m[0], m[4]), __dmul_rn(m[1], m[3]));
  d = d != 0.0 ? __ddiv_rn(1.0, d) : 0.0;
  const double a11 = __dmul_rn(m[4], d), a22 = __dmul_rn(m[0], d);
  const double i00 = a11;
  const double i01 = __dmul_rn(m[1], -d);
  const double i10 = __dmul_rn(m[3], -d);
  const double i11 = a22;
  const double b1 = __dsub_rn(__dmul_rn(-i00, m[2]), __dmul_rn(i01, m[5]));
  const double b2 = __dsub_rn(__dmul_rn(-i10, m[2]), __dmul_rn(i11, m[5]));
  o[0] = i00;
  o[1] = i01;
  o[2] = b1;
  o[3] = i10;
  o[4] = i11;
  o[5] = b2;
}

// cv2.getAffineTransform, op for op: the 6x6 system [x y 1 0 0 0; 0 0 0 x y 1] m = d
// solved by OpenCV's in-house LU (partial pivoting, no fused multiply-add).  The
// crop translation often makes the warp's fixed-point rounding an exact tie (the
// inverse translation is a float32 value, so b * 1024 has only ~6 fractional
// bits), and a last-bit difference in the matrix then moves every pixel of the
// crop; reproducing the solver bit for bit is what makes the warp bit-exact.
__device__ void solve_three_points(const double s[3][2], const double d[3][2], double* m) {
  double A[6][6], b[6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) A[i][j] = 0.0;
  for (int i = 0; i < 3; ++i) {
    A[2 * i][0] = A[2 * i + 1][3] = s[i][0];
    A[2 * i][1] = A[2 * i + 1][4] = s[i][1];
    A[2 * i][2] = A[2 * i + 1][5] = 1.0;
    b[2 * i] = d[i][0];
    b[2 * i + 1] = d[i][1];
  }
  for (int i = 0; i < 6; ++i) {
    int k = i;
    for (int j = i + 1; j < 6; ++j)
      if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
    if (k != i) {
      for (int j = i; j < 6; ++j) {
        const double t = A[i][j];
        A[i][j] = A[k][j];
        A[k][j] = t;
      }
      const double t = b[i];
      b[i] = b[k];
      b[k] = t;
    }
    const double dd = __ddiv_rn(-1.0, A[i][i]);
    for (int j = i + 1; j < 6; ++j) {
      const double alpha = __dmul_rn(A[j][i], dd);
      for (int c = i + 1; c < 6; ++c) A[j][c] = __dadd_rn(A[j][c], __dmul_rn(alpha, A[i][c]));
      b[j] = __dadd_rn(b[j], __dmul_rn(alpha, b[i]));
    }
  }
  for (int i = 5; i >= 0; --i) {
    double acc = b[i];
    for (int c = i + 1; c < 6; ++c) acc = __dsub_rn(acc, __dmul_rn(A[i][c], b[c]));
    b[i] = __ddiv_rn(acc, A[i][i]);
  }
  for (int c = 0; c < 6; ++c) m[c] = b[c];
}

__global__ void affine_matrices_kernel(const float* __restrict__ center,
                                       const float* __restrict__ scale,
                                       const float* __restrict__ rot, double* __restrict__ fwd,
                                       double* __restrict__ inv, int image_w, int image_h,
                                       float pixel_std, int use_udp, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float cx = center[2 * i], cy = center[2 * i + 1];
  const float sx = scale[2 * i], sy = scale[2 * i + 1];
  const double r = rot ? (double)rot[i] : 0.0;
  double m[6];
  if (!use_udp) {
    // get_affine_transform: points are STORED in float32 (utils.py:83-91)
    const float sw = __fmul_rn(sx, pixel_std);  // scale_tmp[0], float32
    const double rad = 3.141592653589793 * r / 180.0;
    const double sn = sin(rad), cs = cos(rad);
    const double pyv = (double)__fmul_rn(sw, -0.5f);
    const double dirx = 0.0 * cs - pyv * sn;
    const double diry = 0.0 * sn + pyv * cs;
    float s[3][2], d[3][2];
    // scale_tmp * shift with shift = (0, 0) adds +0.0 in float32: no effect
    s[0][0] = cx;
    s[0][1] = cy;
    s[1][0] = (float)((double)cx + dirx);
    s[1][1] = (float)((double)cy + diry);
    s[2][0] = __fadd_rn(s[1][0], -__fsub_rn(s[0][1], s[1][1]));
    s[2][1] = __fadd_rn(s[1][1], __fsub_rn(s[0][0], s[1][0]));
    const double dw = (double)image_w, dh = (double)image_h;
    d[0][0] = (float)(dw * 0.5);
    d[0][1] = (float)(dh * 0.5);
    d[1][0] = (float)(dw * 0.5 + 0.0);
    d[1][1] = (float)(dh * 0.5 + dw * -0.5);
    d[2][0] = __fadd_rn(d[1][0], -__fsub_rn(d[0][1], d[1][1]));
    d[2][1] = __fadd_rn(d[1][1], __fsub_rn(d[0][0], d[1][0]));
    double sd[3][2], dd[3][2];
    for (int p = 0; p < 3; ++p)
      for (int c = 0; c < 2; ++c) {
        sd[p][c] = (double)s[p][c];
        dd[p][c] = (double)d[p][c];
      }
    solve_three_points(sd, dd, m);
  } else {
    // get_warp_matrix(rot, center * 2, image_size - 1, scale * pixel_std), float32 result
    const double theta = r * (3.141592653589793 / 180.0);
    const double in_w = (double)__fmul_rn(cx, 2.0f), in_h = (double)__fmul_rn(cy, 2.0f);
    const double tw = (double)__fmul_rn(sx, pixel_std), th = (double)__fmul_rn(sy, pixel_std);
    const double kx = ((double)image_w - 1.0) / tw, ky = ((double)image_h - 1.0) / th;
    const double sn = sin(theta), cs = cos(theta);
    m[0] = (double)(float)(cs * kx);
    m[1] = (double)(float)(-sn * kx);
    m[2] = (double)(float)(kx * (__dadd_rn(__dadd_rn(-0.5 * in_w * cs, 0.5 * in_h * sn),
                                          0.5 * tw)));
    m[3] = (double)(float)(sn * ky);
    m[4] = (double)(float)(cs * ky);
    m[5] = (double)(float)(ky * (__dadd_rn(__dsub_rn(-0.5 * in_w * sn, 0.5 * in_h * cs),
                                          0.5 * th)));
  }
  if (fwd)
    for (int c = 0; c < 6; ++c) fwd[6 * i + c] = m[c];
  if (inv) {
    double o[6];
    invert_affine_cv(m, o);
    for (int c = 0; c < 6; ++c) inv[6 * i + c] = o[c];
  }
}

// cv2.getAffineTransform(src, dst) for caller-built point triples (float32, as the reference
// stores them: utils.py:81-96), then cv::warpAffine's inversion.
__global__ void affine_from_points_kernel(const float* __restrict__ src_pts,
                                          const float* __restrict__ dst_pts,
                                          double* __restrict__ fwd, double* __restrict__ inv,
                                          int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sd[3][2], dd[3][2], m[6];
  for (int p = 0; p < 3; ++p)
    for (int c = 0; c < 2; ++c) {
      sd[p][c] = (double)src_pts[6 * i + 2 * p + c];
      dd[p][c] = (double)dst_pts[6 * i + 2 * p + c];
    }
  solve_three_points(sd, dd, m);
  if (fwd)
    for (int c = 0; c < 6; ++c) fwd[6 * i + c] = m[c];
  if (inv) {
    double o[6];
    invert_affine_cv(m, o);
    for (int c = 0; c < 6; ++c) inv[6 * i + c] = o[c];
  }
}

__global__ void invert_affine_kernel(const double* __restrict__ fwd, double* __restrict__ inv,
                                     int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m[6], o[6];
  for (int c = 0; c < 6; ++c) m[c] = fwd[6 * i + c];
  invert_affine_cv(m, o);
  for (int c = 0; c < 6; ++c) inv[6 * i + c] = o[c];
}

// keypoints through the forward matrix
__global__ void affine_joints_kernel(float* __restrict__ kps, const double* __restrict__ fwd,
                                     int K, int use_udp, int64_t total) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t i = t / K;
  float* kp = kps + t * 3;
  const double* m = fwd + 6 * i;
  const float x = kp[0], y = kp[1];
  if (!use_udp) {
    // np.array(M) @ [x, y, 1.0] in float64, only when visibility > 0
    if (!(kp[2] > 0.f)) return;
    const double xd = (double)x, yd = (double)y;
    kp[0] = (float)__dadd_rn(__dadd_rn(__dmul_rn(m[0], xd), __dmul_rn(m[1], yd)), m[2]);
    kp[1] = (float)__dadd_rn(__dadd_rn(__dmul_rn(m[3], xd), __dmul_rn(m[4], yd)), m[5]);
  } else {
    // np.dot([x, y, 1] float32, M.T float32) runs in BLAS sgemm: a k-ordered
    // chain of float32 fused multiply-adds (verified against numpy/OpenBLAS)
    const float m0 = (float)m[0], m1 = (float)m[1], m2 = (float)m[2];
    const float m3 = (float)m[3], m4 = (float)m[4], m5 = (float)m[5];
    kp[0] = __fmaf_rn(1.0f, m2, __fmaf_rn(y, m1, __fmul_rn(x, m0)));
    kp[1] = __fmaf_rn(1.0f, m5, __fmaf_rn(y, m4, __fmul_rn(x, m3)));
  }
}

// ---------------------------------------------------------------- A4 -------
constexpr int kWarpThreads = 256;
constexpr int kWarpTileRows = 8;
constexpr int kWarpMaxDstW = 1024;

template <int C>
__global__ void __launch_bounds__(kWarpThreads)
    warp_affine_u8_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ src_off,
                          const int32_t* __restrict__ src_hw, const double* __restrict__ inv,
                          uint8_t* __restrict__ dst, int dst_w, int dst_h, int tiles_per_crop,
                          FastDiv div_w) {
  __shared__ int s_adelta[kWarpMaxDstW];
  __shared__ int s_bdelta[kWarpMaxDstW];
  __shared__ int s_x0[kWarpTileRows];
  __shared__ int s_y0[kWarpTileRows];
  extern __shared__ __align__(16) uint8_t s_out[];  // [rows][dst_w * C]

  const int64_t crop = blockIdx.x / tiles_per_crop;
  const int tile = blockIdx.x - (int)(crop * tiles_per_crop);
  const int row0 = tile * kWarpTileRows;
  const int rows = min(kWarpTileRows, dst_h - row0);
  const double* m = inv + 6 * crop;
  const double m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5];

  for (int x = threadIdx.x; x < dst_w; x += blockDim.x) {
    // saturate_cast<int>(M * x * AB_SCALE): round half to even, saturating
    s_adelta[x] = __double2int_rn(__dmul_rn(__dmul_rn(m00, (double)x), 1024.0));
    s_bdelta[x] = __double2int_rn(__dmul_rn(__dmul_rn(m10, (double)x), 1024.0));
  }
  if (threadIdx.x < rows) {
    const double y = (double)(row0 + threadIdx.x);
    s_x0[threadIdx.x] =
        __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m01, y), m02), 1024.0)) + 16;
    s_y0[threadIdx.x] =
        __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m11, y), m12), 1024.0)) + 16;
  }
  __syncthreads();

  const int hs = src_hw[2 * crop], ws = src_hw[2 * crop + 1];
  const uint8_t* img = src + src_off[crop];
  const int npix = rows * dst_w;
  for (int p = threadIdx.x; p < npix; p += blockDim.x) {
    const int ry = (int)fdiv((uint32_t)p, div_w);
    const int x = p - ry * dst_w;
    const int X = (s_x0[ry] + s_adelta[x]) >> 5;
    const int Y = (s_y0[ry] + s_bdelta[x]) >> 5;
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx));  // saturate_cast<short>
    sy = max(-32768, min(32767, sy));
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy);
    const int w10 = (32 - fx) * fy, w11 = fx * fy;
    uint8_t* o = s_out + (size_t)p * C;
    if ((unsigned)sx < (unsigned)(ws - 1) && (unsigned)sy < (unsigned)(hs - 1)) {
      const uint8_t* r0 = img + ((int64_t)sy * ws + sx) * C;
      const uint8_t* r1 = r0 + (int64_t)ws * C;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int acc = w00 * r0[c] + w01 * r0[C + c] + w10 * r1[c] + w11 * r1[C + c];
        o[c] = (uint8_t)((acc + 512) >> 10);
      }
    } else if (sx >= ws || sx + 1 < 0 || sy >= hs || sy + 1 < 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) o[c] = 0;
    } else {
      const bool x0in = sx >= 0 && sx < ws, x1in = sx + 1 >= 0 && sx + 1 < ws;
      const bool y0in = sy >= 0 && sy < hs, y1in = sy + 1 >= 0 && sy + 1 < hs;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        int acc = 0;
        if (y0in && x0in) acc += w00 * img[((int64_t)sy * ws + sx) * C + c];
        if (y0in && x1in) acc += w01 * img[((int64_t)sy * ws + sx + 1) * C + c];
        if (y1in && x0in) acc += w10 * img[((int64_t)(sy + 1) * ws + sx) * C + c];
        if (y1in && x1in) acc += w11 * img[((int64_t)(sy + 1) * ws + sx + 1) * C + c];
        o[c] = (uint8_t)((acc + 512) >> 10);
      }
    }
  }
  __syncthreads();

  // the tile is one contiguous span of the destination crop: coalesced copy-out
  const size_t tile_bytes = (size_t)npix * C;
  uint8_t* out = dst + ((size_t)crop * dst_h + row0) * dst_w * C;
  if ((((uintptr_t)out) & 15) == 0) {
    const int nvec = (int)(tile_bytes >> 4);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x)
      reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(s_out)[i];
    for (int i = (nvec << 4) + threadIdx.x; i < (int)tile_bytes; i += blockDim.x)
      out[i] = s_out[i];
  } else {
    for (int i = threadIdx.x; i < (int)tile_bytes; i += blockDim.x) out[i] = s_out[i];
  }
}


// ---- 3-channel fast path ----------------------------------------------------
// The generic kernel above issues 12 byte loads per output pixel and is bound by
// load-instruction issue, not by HBM.  For interleaved RGB the two horizontally
// adjacent source pixels are 6 contiguous bytes: they are fetched as two (three
// when the run starts at byte 3 of a word) ALIGNED 32-bit words per source row and
// realigned with a funnel shift, each thread produces four consecutive output
// pixels (12 bytes = three packed words), and a warp's 384 output bytes are
// transposed through shared memory into 24 full 16-byte stores.
constexpr int kWarp3TileRows = 16;

struct Six {
  uint32_t lo, hi;  // bytes 0..3, bytes 4..7 of the run (6 are used)
};
// 6-byte run at byte offset `off` from the 4-byte aligned base: three aligned words,
// realigned with a funnel shift.
__device__ __forceinline__ Six load_six(const uint8_t* __restrict__ base4, uint32_t off) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(base4 + (off & ~3u));
  const uint32_t sh = (off & 3u) << 3;
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
  Six r;
  r.lo = __funnelshift_r(w0, w1, sh);
  r.hi = __funnelshift_r(w1, w2, sh);
  return r;
}

// Source row pair of one output pixel: shared by the four pixels of a thread when the
// matrix has no rotation (then Y does not depend on x).
//
// cv::remap limits sources to SHRT_MAX rows / columns, so saturate_cast<short> of the
// integer coordinates can only turn an out-of-image tap into another out-of-image tap;
// it is not reproduced.
//
// Fast path range: the aligned words of a run reach up to 3 bytes before and 6 bytes
// after it, so the run must not start at the first pixel of the image nor end in its
// last 3 pixels; [sx_lo, sx_lo + sx_span] is the range of sx that is both interior
// (all four taps inside) and safe for this row pair.
struct RowCtx {
  int sy, fy;
  uint32_t off;     // byte offset of row sy from the aligned base (valid rows only)
  int sx_lo;
  uint32_t sx_span;
};
__device__ __forceinline__ RowCtx make_row(int Y, int hs, int ws, uint32_t ws3, uint32_t delta) {
  RowCtx r;
  r.sy = Y >> 5;
  r.fy = Y & 31;
  r.off = (uint32_t)r.sy * ws3 + delta;
  const int lo = r.sy >= 1 ? 0 : 1;
  const int hi = r.sy + 2 < hs ? ws - 2 : ws - 4;
  const bool ok = (unsigned)r.sy < (unsigned)(hs - 1) && hi >= lo;
  r.sx_lo = ok ? lo : 0x40000000;
  r.sx_span = ok ? (uint32_t)(hi - lo) : 0u;
  return r;
}

// One output pixel (3 channels packed r | g << 8 | b << 16) at fixed-point source X in
// the row pair rc.  OpenCV's sum of four 15-bit weighted taps equals, exactly (integer
// arithmetic), a horizontal blend with weights (32-fx, fx) followed by a vertical blend
// with (32-fy, fy); the horizontal blends are two-way dot products (dp2a) of the weight
// pair with the channel's two bytes, paired up by a byte permute of the realigned words.
// ROWAL: the source row pitch is a multiple of 4 bytes, so the run has the same alignment
// in both rows and the lower row's words are the upper row's plus the pitch.
// Interior pixel: all four taps inside the image and the aligned words of the run inside
// the buffer (the caller has checked sx against [rc.sx_lo, rc.sx_lo + rc.sx_span]).
template <bool ROWAL>
__device__ __forceinline__ uint32_t warp_pixel3_interior(const uint8_t* __restrict__ base4,
                                                         uint32_t ws3, int X, const RowCtx rc) {
  const int sx = X >> 5;
  const int fx = X & 31, fy = rc.fy;
  {
    const uint32_t off = rc.off + (uint32_t)sx * 3u;
    Six a, b;
    if (ROWAL) {
      const uint8_t* pa = base4 + (off & ~3u);
      const uint32_t* wa = reinterpret_cast<const uint32_t*>(pa);
      const uint32_t* wb = reinterpret_cast<const uint32_t*>(pa + ws3);
      const uint32_t sh = (off & 3u) << 3;
      const uint32_t a0w = __ldg(wa), a1w = __ldg(wa + 1), a2w = __ldg(wa + 2);
      const uint32_t b0w = __ldg(wb), b1w = __ldg(wb + 1), b2w = __ldg(wb + 2);
      a.lo = __funnelshift_r(a0w, a1w, sh);
      a.hi = __funnelshift_r(a1w, a2w, sh);
      b.lo = __funnelshift_r(b0w, b1w, sh);
      b.hi = __funnelshift_r(b1w, b2w, sh);
    } else {
      a = load_six(base4, off);
      b = load_six(base4, off + ws3);
    }
    // lo = r0 g0 b0 r1, hi = g1 b1 . . : pair the bytes of each channel (byte permute) and
    // take the 16-bit x 8-bit dot products with the weight pair (32 - fx, fx)
    const uint32_t wg = (32u - (uint32_t)fx) | ((uint32_t)fx << 16);
    const uint32_t arg = __byte_perm(a.lo, a.hi, 0x4130), abb = __byte_perm(a.lo, a.hi, 0x0052);
    const uint32_t brg = __byte_perm(b.lo, b.hi, 0x4130), bbb = __byte_perm(b.lo, b.hi, 0x0052);
    // The vertical weights are folded into the weight pairs: both 16-bit halves of wg scale
    // by gy (uy) in one multiply without a carry between them (products <= 1024), so each
    // channel is two chained dot products -- the same integer sum as blending horizontally,
    // then vertically.
    const uint32_t gy = 32u - (uint32_t)fy, uy = (uint32_t)fy;
    const uint32_t wt = wg * gy, wu = wg * uy;
    const uint32_t c0 = __dp2a_lo(wu, brg, __dp2a_lo(wt, arg, 512u)) >> 10;
    const uint32_t c1 = __dp2a_hi(wu, brg, __dp2a_hi(wt, arg, 512u)) >> 10;
    const uint32_t c2 = __dp2a_lo(wu, bbb, __dp2a_lo(wt, abb, 512u)) >> 10;
    return __byte_perm(__byte_perm(c0, c1, 0x0040), c2, 0x5410);  // byte 3 = high byte of c2 = 0
  }
}

template <bool ROWAL>
__device__ __forceinline__ uint32_t warp_pixel3(const uint8_t* __restrict__ img,
                                                const uint8_t* __restrict__ base4, int hs,
                                                int ws, uint32_t ws3, int X, const RowCtx rc) {
  const int sx = X >> 5;
  const int fx = X & 31, fy = rc.fy, sy = rc.sy;
  if ((uint32_t)(sx - rc.sx_lo) <= rc.sx_span)
    return warp_pixel3_interior<ROWAL>(base4, ws3, X, rc);
  if (sx >= ws || sx + 1 < 0 || sy >= hs || sy + 1 < 0) return 0u;
  const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy);
  const int w10 = (32 - fx) * fy, w11 = fx * fy;
  const bool x0in = sx >= 0 && sx < ws, x1in = sx + 1 >= 0 && sx + 1 < ws;
  const bool y0in = sy >= 0 && sy < hs, y1in = sy + 1 >= 0 && sy + 1 < hs;
  uint32_t out = 0u;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int acc = 0;
    if (y0in && x0in) acc += w00 * img[((int64_t)sy * ws + sx) * 3 + c];
    if (y0in && x1in) acc += w01 * img[((int64_t)sy * ws + sx + 1) * 3 + c];
    if (y1in && x0in) acc += w10 * img[((int64_t)(sy + 1) * ws + sx) * 3 + c];
    if (y1in && x1in) acc += w11 * img[((int64_t)(sy + 1) * ws + sx + 1) * 3 + c];
    out |= (uint32_t)((acc + 512) >> 10) << (8 * c);
  }
  return out;
}

struct NormArgs {
  float mean[3], std[3];
  // 1 / std and whether x * rcp corrected by one residual step equals the IEEE quotient for
  // every value this channel can take (checked on the host for all 256: norm_fast_ok)
  float rcp[3];
  int32_t fast;
};

// (pixel - mean) / std, correctly rounded: the division itself, or -- when the host has
// verified it for all 256 pixel values of the channel -- q0 = x * (1 / std) plus one exact
// residual step (three instructions instead of the nine of an IEEE division).
__device__ __forceinline__ float norm_value(uint32_t px, float mean, float std, float rcp,
                                            bool fast) {
  const float x = __fsub_rn((float)px, mean);
  if (!fast) return __fdiv_rn(x, std);
  const float q0 = __fmul_rn(x, rcp);
  return __fmaf_rn(__fmaf_rn(-q0, std, x), rcp, q0);
}

// NORM = false: uint8 HWC crops; dst_w % 16 == 0 and dst 16-byte aligned, so every warp's
//   32 quads are 384 contiguous, 16-byte aligned output bytes.
// NORM = true: the step that follows the warp in the
//   reference's pipeline -- vision.Normalize(mean * 255, std * 255) and HWC2CHW
//   (mindpose/data/data_factory.py:127-138) -- fused into the store: float32 CHW planes,
//   (pixel - mean[c]) / std[c] in float32, one 16-byte store per channel and thread
//   (dst_w % 4 == 0).  The uint8 crop is never written.
// QUAD (rotation-free crops of the uint8 variant): the fixed-point column deltas are monotone
// in x, so a quad whose first and last pixel are interior is interior as a whole -- one range
// test per quad, and the four pixels' tap loads sit in one basic block.
// Shared memory of the quad kernel below (12.6 KB); the band kernel further down lends it a
// piece of its band buffer when a tile takes this path.
struct QuadSmem {
  int adelta[kWarpMaxDstW];
  int bdelta[kWarpMaxDstW];
  int x0[kWarp3TileRows];
  int y0[kWarp3TileRows];
  int4 rowa[kWarp3TileRows];  // rotation-free crops: sy, fy, sx_lo, sx_span per row
  uint32_t rowoff[kWarp3TileRows];
  uint32_t stage[16][96];     // one 384-byte transpose buffer per warp (<= 16 warps)
};

// One tile (kWarp3TileRows output rows of one crop) by any number of whole warps.
template <bool NORM, bool QUAD>
__device__ __forceinline__ void warp3_quad_tile(
    QuadSmem& sm, const uint8_t* __restrict__ src, const int64_t* __restrict__ src_off,
    const int32_t* __restrict__ src_hw, const double* __restrict__ inv, void* __restrict__ dst_any,
    int dst_w, int dst_h, int64_t crop, int tile, FastDiv div_wq, const NormArgs& norm) {
  uint8_t* dst = static_cast<uint8_t*>(dst_any);
  int* const s_adelta = sm.adelta;
  int* const s_bdelta = sm.bdelta;
  int* const s_x0 = sm.x0;
  int* const s_y0 = sm.y0;
  int4* const s_rowa = sm.rowa;
  uint32_t* const s_rowoff = sm.rowoff;
  const int nthreads = blockDim.x;

  const int row0 = tile * kWarp3TileRows;
  const int rows = min(kWarp3TileRows, dst_h - row0);
  const double* m = inv + 6 * crop;
  const double m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5];

  for (int x = threadIdx.x; x < dst_w; x += nthreads) {
    s_adelta[x] = __double2int_rn(__dmul_rn(__dmul_rn(m00, (double)x), 1024.0));
    s_bdelta[x] = __double2int_rn(__dmul_rn(__dmul_rn(m10, (double)x), 1024.0));
  }
  if (threadIdx.x < rows) {
    const double y = (double)(row0 + threadIdx.x);
    s_x0[threadIdx.x] =
        __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m01, y), m02), 1024.0)) + 16;
    s_y0[threadIdx.x] =
        __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m11, y), m12), 1024.0)) + 16;
  }
  __syncthreads();

  const int hs = src_hw[2 * crop], ws = src_hw[2 * crop + 1];
  const uint8_t* img = src + src_off[crop];
  const uint32_t delta = (uint32_t)(reinterpret_cast<uintptr_t>(img) & 3u);
  const uint8_t* base4 = img - delta;  // 4-byte aligned
  const uint32_t ws3 = (uint32_t)ws * 3u;
  const int wq = dst_w >> 2;
  const int nquads = rows * wq;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* out = dst + ((size_t)crop * dst_h + row0) * dst_w * 3;
  uint32_t* stage = sm.stage[warp];
  // the fixed-point column deltas are monotone in x: the last one being 0 means all are
  const bool axis = s_bdelta[dst_w - 1] == 0;
  const bool rowal = (ws3 & 3u) == 0;
  if (!NORM && axis) {  // Y does not depend on x: one row context per output row
    if (threadIdx.x < rows) {
      const RowCtx rc = make_row(s_y0[threadIdx.x] >> 5, hs, ws, ws3, delta);
      s_rowa[threadIdx.x] = make_int4(rc.sy, rc.fy, rc.sx_lo, (int)rc.sx_span);
      s_rowoff[threadIdx.x] = rc.off;
    }
    __syncthreads();
  }

  for (int base = warp * 32; base < nquads; base += nthreads) {
    const int t = base + lane;
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    if (t < nquads) {
      const int ry = (int)fdiv((uint32_t)t, div_wq);
      const int x = (t - ry * wq) << 2;
      const int X0 = s_x0[ry], Y0 = s_y0[ry];
      const int4 ad = *reinterpret_cast<const int4*>(&s_adelta[x]);
      // (uint8 variant only: in the float32 CHW variant the extra code path costs 15
      // registers and two resident CTAs per SM)
      if (!NORM && axis) {  // no rotation: Y does not depend on x, one row pair per quad
        const int4 rw = s_rowa[ry];
        RowCtx ra;
        ra.sy = rw.x, ra.fy = rw.y, ra.sx_lo = rw.z, ra.sx_span = (uint32_t)rw.w;
        ra.off = s_rowoff[ry];
        const int Xa = (X0 + ad.x) >> 5, Xd = (X0 + ad.w) >> 5;
        const int sx_min = min(Xa, Xd) >> 5, sx_max = max(Xa, Xd) >> 5;
        if (QUAD && (uint32_t)(sx_min - ra.sx_lo) <= ra.sx_span &&
            (uint32_t)(sx_max - ra.sx_lo) <= ra.sx_span) {
          const int Xb = (X0 + ad.y) >> 5, Xc = (X0 + ad.z) >> 5;
          if (rowal) {
            p0 = warp_pixel3_interior<true>(base4, ws3, Xa, ra);
            p1 = warp_pixel3_interior<true>(base4, ws3, Xb, ra);
            p2 = warp_pixel3_interior<true>(base4, ws3, Xc, ra);
            p3 = warp_pixel3_interior<true>(base4, ws3, Xd, ra);
          } else {
            p0 = warp_pixel3_interior<false>(base4, ws3, Xa, ra);
            p1 = warp_pixel3_interior<false>(base4, ws3, Xb, ra);
            p2 = warp_pixel3_interior<false>(base4, ws3, Xc, ra);
            p3 = warp_pixel3_interior<false>(base4, ws3, Xd, ra);
          }
        } else if (rowal) {
          p0 = warp_pixel3<true>(img, base4, hs, ws, ws3, (X0 + ad.x) >> 5, ra);
          p1 = warp_pixel3<true>(img, base4, hs, ws, ws3, (X0 + ad.y) >> 5, ra);
          p2 = warp_pixel3<true>(img, base4, hs, ws, ws3, (X0 + ad.z) >> 5, ra);
          p3 = warp_pixel3<true>(img, base4, hs, ws, ws3, (X0 + ad.w) >> 5, ra);
        } else {
          p0 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.x) >> 5, ra);
          p1 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.y) >> 5, ra);
          p2 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.z) >> 5, ra);
          p3 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.w) >> 5, ra);
        }
      } else {
        const int4 bd = *reinterpret_cast<const int4*>(&s_bdelta[x]);
        const int Ya = (Y0 + bd.x) >> 5, Yb = (Y0 + bd.y) >> 5;
        const int Yc = (Y0 + bd.z) >> 5, Yd = (Y0 + bd.w) >> 5;
        const RowCtx ra = make_row(Ya, hs, ws, ws3, delta);
        const bool same = Ya == Yb && Ya == Yc && Ya == Yd;  // small angles: one row pair
        // Two spellings of the same thing: ptxas allocates 40 registers for the float32
        // variant with the branch and 48 (not 54) for the uint8 variant with the selects.
        if constexpr (NORM) {
          if (same) {
            p0 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.x) >> 5, ra);
            p1 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.y) >> 5, ra);
            p2 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.z) >> 5, ra);
            p3 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.w) >> 5, ra);
          } else {
            p0 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.x) >> 5, ra);
            p1 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.y) >> 5,
                                    make_row(Yb, hs, ws, ws3, delta));
            p2 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.z) >> 5,
                                    make_row(Yc, hs, ws, ws3, delta));
            p3 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.w) >> 5,
                                    make_row(Yd, hs, ws, ws3, delta));
          }
        } else {
          p0 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.x) >> 5, ra);
          p1 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.y) >> 5,
                                  same ? ra : make_row(Yb, hs, ws, ws3, delta));
          p2 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.z) >> 5,
                                  same ? ra : make_row(Yc, hs, ws, ws3, delta));
          p3 = warp_pixel3<false>(img, base4, hs, ws, ws3, (X0 + ad.w) >> 5,
                                  same ? ra : make_row(Yd, hs, ws, ws3, delta));
        }
      }
    }
    if (NORM) {
      if (t < nquads) {
        const int ry = (int)fdiv((uint32_t)t, div_wq);
        const int x = (t - ry * wq) << 2;
        float* fdst = static_cast<float*>(dst_any);
        const uint32_t px[4] = {p0, p1, p2, p3};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            v[e] = __fdiv_rn(__fsub_rn((float)((px[e] >> (8 * c)) & 255u), norm.mean[c]),
                             norm.std[c]);
          st_stream_f4(fdst + (((size_t)crop * 3 + c) * dst_h + row0 + ry) * dst_w + x,
                       make_float4(v[0], v[1], v[2], v[3]));
        }
      }
      continue;
    }
    stage[3 * lane] = p0 | (p1 << 24);
    stage[3 * lane + 1] = (p1 >> 8) | (p2 << 16);
    stage[3 * lane + 2] = (p2 >> 16) | (p3 << 8);
    __syncwarp();
    const int nq = min(32, nquads - base);  // quads this warp produced
    if (lane * 4 < nq * 3) {                // nq * 12 bytes = nq * 3 words; nq % 4 == 0
      const uint4 v = reinterpret_cast<const uint4*>(stage)[lane];
      asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(out + (size_t)base * 12 +
                                                                           lane * 16),
                   "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                   : "memory");
    }
    __syncwarp();
  }
}

template <bool NORM, bool QUAD>
__global__ void __launch_bounds__(kWarpThreads)
    warp_affine_u8x3_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ src_off,
                            const int32_t* __restrict__ src_hw, const double* __restrict__ inv,
                            void* __restrict__ dst_any, int dst_w, int dst_h, int tiles_per_crop,
                            FastDiv div_wq, const NormArgs norm) {
  __shared__ __align__(16) QuadSmem sm;
  const int64_t crop = blockIdx.x / tiles_per_crop;
  const int tile = blockIdx.x - (int)(crop * tiles_per_crop);
  warp3_quad_tile<NORM, QUAD>(sm, src, src_off, src_hw, inv, dst_any, dst_w, dst_h, crop, tile,
                              div_wq, norm);
}

// ---- band kernel: rotation-free 3-channel crops out of shared memory --------------------
// The evaluation path (rot = 0) and most of the bench.  The quad kernel above gathers its taps
// from global memory through L1: 81 instructions per pixel (alignment, range tests and address
// arithmetic per tap), every sector fetched 14 times out of L1, 0.52 of the HBM roofline.
// Without rotation the source pixels of a run of output rows are a RECTANGLE of the image, so
// here
//   * the rectangle ("band": source rows sy_lo .. sy_hi, columns clo .. chi, clipped to the
//     image) is staged in shared memory by 1-D bulk copies (cp.async.bulk, one per source row,
//     issued by the lanes of warp 0, completion on an mbarrier): the copies are 16-byte
//     aligned in global memory, so a row lands at the phase (address mod 16) it has there;
//   * OpenCV's constant border costs nothing: a tap outside the image gets WEIGHT 0 (the
//     column's weight pair is masked once per tile, the row's pair once per pass) and its
//     address is clamped into the band, so whatever bytes it reads contribute exactly 0 --
//     no range test and no zero fill anywhere;
//   * a lane owns ONE output column for the whole tile: the source column, the weight pair
//     (32 - fx, fx), the word offsets of its three tap words and the two byte-permute selectors
//     that pair up the channels (r0 r1 g0 g1 | b0 b1) are per-thread constants -- the row
//     pitch is a multiple of 4 bytes (ws % 4 == 0), so the alignment of a column's 6-byte run
//     is the same in every row;
//   * per pixel: one broadcast load of the row's entry (two row addresses, two weights),
//     3 + 3 aligned 32-bit shared-memory loads, 4 byte permutes and the same six dp2a dot
//     products as the quad kernel (the identical integer sum, so the identical bytes);
//   * a warp's 32 pixels of a row (96 bytes) leave as 24 words: one shuffle and one byte
//     permute per lane (tests/test_warp_identities.py), three lanes of four store.
// Tiles that do not qualify (rotation, ws % 4 != 0, a mirrored matrix, a band of one output
// row that does not fit) run the quad path in the same launch, on a piece of the band buffer.
// One band buffer per CTA; the copies of a pass overlap the other five CTAs of the SM.  (Two
// buffers per CTA, the next pass copied while this one is computed, were measured and lost:
// 2 x 18 KB 0.469 ms, 2 x 24 KB 0.480 ms against 0.445 ms for 1 x 36 KB -- shorter passes,
// fewer CTAs; profiles/README.md, r02f.)
constexpr int kBandBytes = 36 * 1024;
// (Measured and not kept, profiles/README.md r02s: 44 KB bands / 5 CTAs 0.569 ms, 128-row
// tiles 0.514 ms, unroll 2 / 8 0.484 / 0.478 ms, and skipping the loads of a source row the
// previous output row already fetched 0.480 ms -- its uniform branches cost more than the
// loads they save -- against 0.445 ms for this form.)
constexpr int kBandUnroll = 4;
constexpr int kBandMaxThreads = 512;   // dst_w <= 512 (one column per thread)
constexpr int kBandPassRows = 16;      // output rows staged and computed at a time
constexpr int kBandMaxTileRows = 64;   // output rows per CTA (the host picks 16, 32 or 64)
struct BandRow {
  uint32_t addr_a;  // shared-memory address of the word holding source pixel (sy, clo) ...
  uint32_t addr_b;  // ... and (sy + 1, clo): the rows' 16-byte phases differ when ws3 % 16 != 0
  uint32_t gy, uy;  // 32 - fy, fy; 0 for a source row outside the image
};

__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst_smem, const void* src_gmem,
                                               uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr));
  return v;
}

// 48 registers: 6 CTAs of 192 threads per SM, as many as a 36 KB band allows.
// todo: [0] = number of entries, [1 ...] = the 16-row tiles (crop * tiles16 + tile) this kernel
// leaves to the quad kernel.
// NORM: the float32 CHW variant (row N2, see the quad kernel): (pixel - mean[c]) / std[c] to
// three planes, a warp's 32 pixels of a row as one 128-byte store per channel.
template <bool NORM>
__global__ void __maxnreg__(48)
    warp_affine_u8x3_band_kernel(const uint8_t* __restrict__ src,
                                 const int64_t* __restrict__ src_off,
                                 const int32_t* __restrict__ src_hw,
                                 const double* __restrict__ inv, void* __restrict__ dst_any,
                                 int dst_w, int dst_h, int tile_rows, int tiles_per_crop,
                                 int* __restrict__ todo, const NormArgs norm) {
  uint8_t* const dst = static_cast<uint8_t*>(dst_any);
  extern __shared__ __align__(128) uint8_t s_band[];  // kBandBytes
  __shared__ int s_x0[kBandMaxTileRows];
  __shared__ int s_y0[kBandMaxTileRows];
  __shared__ __align__(16) BandRow s_row[kBandPassRows];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t crop = blockIdx.x / tiles_per_crop;
  const int tile = blockIdx.x - (int)(crop * tiles_per_crop);
  const int row0 = tile * tile_rows;
  const int rows = min(tile_rows, dst_h - row0);
  const double* m = inv + 6 * crop;
  const double m00 = m[0], m10 = m[3];
  const int hs = src_hw[2 * crop], ws = src_hw[2 * crop + 1];
  const uint8_t* img = src + src_off[crop];
  const uint32_t ws3 = (uint32_t)ws * 3u;

  if (tid < rows) {
    const double m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5];
    const double y = (double)(row0 + tid);
    s_x0[tid] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m01, y), m02), 1024.0)) + 16;
    s_y0[tid] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m11, y), m12), 1024.0)) + 16;
  }
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  // this thread's column (blockDim.x == dst_w, a multiple of 32)
  const int ad = __double2int_rn(__dmul_rn(__dmul_rn(m00, (double)tid), 1024.0));
  const int ad_last = __double2int_rn(__dmul_rn(__dmul_rn(m00, (double)(dst_w - 1)), 1024.0));
  const int bd_last = __double2int_rn(__dmul_rn(__dmul_rn(m10, (double)(dst_w - 1)), 1024.0));
  __syncthreads();

  // ---- does the tile qualify? (block-uniform) --------------------------------------------
  // Y independent of x (the fixed-point column deltas of Y are monotone: the last one being 0
  // means all are), X0 the same in every row of the tile, sx and sy non-decreasing.
  const int X0 = s_x0[0];
  const bool row_ok = tid == 0 || tid >= rows || (s_x0[tid] == X0 && s_y0[tid] >= s_y0[tid - 1]);
  bool ok = __syncthreads_and(row_ok) != 0;
  ok = ok && bd_last == 0 && ad_last >= 0 && (ws3 & 3u) == 0 && hs >= 1 && ws >= 1;
  const int sx_lo = X0 >> 10, sx_hi = ((X0 + ad_last) >> 10) + 1;  // columns the taps touch
  const int clo = max(sx_lo, 0), chi = min(sx_hi, ws - 1);         // ... inside the image
  const bool cols_in = clo <= chi;
  // Row pitch of the band: 16 bytes in front when there are columns outside the image (a
  // clamped run may start 3 bytes before column clo), the phase of the row (<= 15), the columns
  // clo .. chi, the three tap words of the last column (<= 12 bytes from the start of its
  // run), the tail of the last 16-byte unit of the copy (<= 15).
  const int L = (sx_lo < 0 || !cols_in) ? 16 : 0;
  const int P = (L + 15 + (cols_in ? (chi - clo + 1) * 3 : 3) + 12 + 15 + 15) & ~15;
  const int nb_max = kBandBytes / P;  // source rows the buffer holds
  // nb_max >= 2: one output row needs exactly two source rows, so the passes always advance
  ok = ok && nb_max >= 2;
  if (!ok) {
    // Rotation, ws % 4 != 0, a mirrored matrix, a band that does not fit: the tile goes to the
    // quad kernel (its 16-row tiles), launched right after this kernel over the list.
    if (tid == 0) {
      const int tiles16 = (dst_h + kWarp3TileRows - 1) / kWarp3TileRows;
      const int first = row0 / kWarp3TileRows;  // tile_rows is a multiple of kWarp3TileRows
      const int cnt = (rows + kWarp3TileRows - 1) / kWarp3TileRows;
      const int at = atomicAdd(todo, cnt);
      for (int i = 0; i < cnt; ++i) todo[1 + at + i] = (int)crop * tiles16 + first + i;
    }
    return;
  }

  // ---- per-thread column constants ---------------------------------------------------------
  const int X = (X0 + ad) >> 5;
  const int sx = X >> 5, fx = X & 31;
  // weights of the two taps, 0 for a tap outside the image (constant border 0)
  const uint32_t wg = ((uint32_t)sx < (uint32_t)ws ? 32u - (uint32_t)fx : 0u) |
                      ((uint32_t)(sx + 1) < (uint32_t)ws ? (uint32_t)fx << 16 : 0u);
  // the run is read from a column inside the band, wherever the true column is: a tap that
  // moved has weight 0 (clo - 1 keeps tap 1 = column clo in place)
  const int sxc = min(max(sx, clo - 1), max(chi, clo - 1));
  const uint32_t band0 = smem_u32(s_band);
  // phase mod 4 of column clo, the same in every row (ws3 % 4 == 0); BandRow.addr_* is the
  // address of column clo rounded down to a word, so a column's run starts `run` bytes on
  const uint32_t rho = (uint32_t)((reinterpret_cast<uintptr_t>(img) + (size_t)clo * 3u) & 3u);
  const int run = (sxc - clo) * 3 + (int)rho;  // >= -3
  const uint32_t a = (uint32_t)run & 3u;       // alignment of the run
  const uint32_t colA = (uint32_t)(run - (int)a);     // word A (B = A + 4)
  const uint32_t colZ = colA + (a == 3u ? 8u : 0u);   // word C when the run starts at byte 3, else A
  const uint32_t selA = 0x4130u + 0x1111u * a;               // (A, B) -> r0 r1 g0 g1
  const uint32_t selB = ((6u + a) & 7u) | ((1u + a) << 4);   // (B, Z) -> b0 b1 . .
  // output: lane k of a group of four holds pixel k; words 0..2 of the group's 12 bytes
  const int k4 = lane & 3;
  const uint32_t selO = k4 == 0 ? 0x4210u : (k4 == 1 ? 0x5421u : 0x6542u);
  uint8_t* out = dst + ((size_t)crop * dst_h + row0) * dst_w * 3 + (size_t)(tid >> 5) * 96 +
                 (size_t)(((lane >> 2) * 3 + k4) << 2);
  const uint32_t out_pitch = (uint32_t)dst_w * 3u;
  // NORM: this column of channel 0's plane; the planes are dst_h * dst_w floats apart
  float* fout = static_cast<float*>(dst_any) + ((size_t)crop * 3 * dst_h + row0) * dst_w + tid;
  const size_t fplane = (size_t)dst_h * dst_w;
  const uint32_t row_tab = smem_u32(s_row);
  // fixed-point source rows per output row, for the pass planner (an estimate: the planner
  // checks the rows it picks)
  const int y_step = rows > 1 ? max((s_y0[rows - 1] - s_y0[0]) / (rows - 1), 1) : 1;

  struct Pass {
    int r_begin, r_end, sy_lo, sy_hi;
    bool any;  // some source pixel of the band lies inside the image
  };
  // the output rows of the pass that starts at row r: as many as one buffer holds source
  // rows for, at most kBandPassRows
  auto plan = [&](int r) {
    Pass ps;
    ps.r_begin = r;
    ps.sy_lo = s_y0[r] >> 10;
    int cnt = min(min(((nb_max - 2) << 10) / y_step + 1, kBandPassRows), rows - r);
    while (cnt > 1 && (s_y0[r + cnt - 1] >> 10) + 2 - ps.sy_lo > nb_max) --cnt;
    ps.r_end = r + cnt;
    ps.sy_hi = (s_y0[ps.r_end - 1] >> 10) + 1;
    ps.any = max(ps.sy_lo, 0) <= min(ps.sy_hi, hs - 1) && cols_in;
    return ps;
  };
  // row table and bulk copies of one pass
  auto issue = [&](const Pass& ps) {
    const uint32_t buf = band0;
    const int rlo = max(ps.sy_lo, 0), rhi = min(ps.sy_hi, hs - 1);
    if (tid < ps.r_end - ps.r_begin) {
      const int Y = s_y0[ps.r_begin + tid] >> 5;
      const int sy = Y >> 5, fy = Y & 31;
      // Band row j = r - sy_lo starts at j * P; its copy starts at L and lands at the phase
      // (address mod 16) the row has in global memory.  A source row outside the image has
      // weight 0 and reads the nearest row of the band instead.
      const int ra = ps.any ? min(max(sy, rlo), rhi) : ps.sy_lo;
      const int rb = ps.any ? min(max(sy + 1, rlo), rhi) : ps.sy_lo;
      const uintptr_t g0 = reinterpret_cast<uintptr_t>(img) + (size_t)clo * 3u;
      const uintptr_t ga = g0 + (size_t)(uint32_t)max(ra, 0) * ws3;
      const uintptr_t gb = g0 + (size_t)(uint32_t)max(rb, 0) * ws3;
      BandRow br;
      br.addr_a = buf + (uint32_t)((ra - ps.sy_lo) * P + L) + ((uint32_t)ga & 12u);
      br.addr_b = buf + (uint32_t)((rb - ps.sy_lo) * P + L) + ((uint32_t)gb & 12u);
      br.gy = (uint32_t)sy < (uint32_t)hs ? 32u - (uint32_t)fy : 0u;
      br.uy = (uint32_t)(sy + 1) < (uint32_t)hs ? (uint32_t)fy : 0u;
      s_row[tid] = br;
    }
    if (tid < 32 && ps.any) {  // one bulk copy per source row, issued by the lanes of warp 0
      const uint32_t span = (uint32_t)(chi - clo + 1) * 3u;
      uint32_t bytes = 0;
      for (int r = rlo + lane; r <= rhi; r += 32) {
        const uintptr_t g = reinterpret_cast<uintptr_t>(img) + (size_t)r * ws3 + (size_t)clo * 3u;
        bytes += (uint32_t)(((g & 15u) + span + 15u) & ~15u);
      }
      bytes = __reduce_add_sync(0xffffffffu, bytes);
      if (lane == 0) mbar_arrive_expect_tx(&s_bar, bytes);
      __syncwarp();
      for (int r = rlo + lane; r <= rhi; r += 32) {
        const uintptr_t g = reinterpret_cast<uintptr_t>(img) + (size_t)r * ws3 + (size_t)clo * 3u;
        const uint32_t n16 = (uint32_t)(((g & 15u) + span + 15u) & ~15u);
        bulk_g2s_plain(buf + (uint32_t)((r - ps.sy_lo) * P + L),
                       reinterpret_cast<const void*>(g & ~(uintptr_t)15), n16, &s_bar);
      }
    }
  };

  uint32_t phase = 0u;  // parity of the barrier's next completion
  Pass cur = plan(0);
  issue(cur);
#pragma unroll 1
  for (;;) {
    if (cur.any) {
      if (tid < 32) mbar_wait(&s_bar, phase);  // the other warps wait at the barrier below
      phase ^= 1u;
    }
    __syncthreads();

    // ---- one column, r_end - r_begin rows ---------------------------------------------------
    {
      const int cnt = cur.r_end - cur.r_begin;
      uint8_t* o = out + (size_t)cur.r_begin * out_pitch;
#pragma unroll kBandUnroll
      for (int i = 0; i < cnt; ++i) {
        const uint4 rc = lds128(row_tab + 16u * (uint32_t)i);  // addr_a, addr_b, gy, uy
        const uint32_t A0 = lds32(rc.x + colA), B0 = lds32(rc.x + colA + 4u);
        const uint32_t Z0 = lds32(rc.x + colZ);
        const uint32_t A1 = lds32(rc.y + colA), B1 = lds32(rc.y + colA + 4u);
        const uint32_t Z1 = lds32(rc.y + colZ);
        const uint32_t arg = __byte_perm(A0, B0, selA), abb = __byte_perm(B0, Z0, selB);
        const uint32_t brg = __byte_perm(A1, B1, selA), bbb = __byte_perm(B1, Z1, selB);
        const uint32_t wt = wg * rc.z, wu = wg * rc.w;  // (32 - fx, fx) x (gy, uy)
        const uint32_t c0 = __dp2a_lo(wu, brg, __dp2a_lo(wt, arg, 512u)) >> 10;
        const uint32_t c1 = __dp2a_hi(wu, brg, __dp2a_hi(wt, arg, 512u)) >> 10;
        const uint32_t c2 = __dp2a_lo(wu, bbb, __dp2a_lo(wt, abb, 512u)) >> 10;
        if (NORM) {
          float* f = fout + (size_t)(cur.r_begin + i) * dst_w;
          const bool fast = norm.fast != 0;  // uniform
          const float v0 = norm_value(c0, norm.mean[0], norm.std[0], norm.rcp[0], fast);
          const float v1 = norm_value(c1, norm.mean[1], norm.std[1], norm.rcp[1], fast);
          const float v2 = norm_value(c2, norm.mean[2], norm.std[2], norm.rcp[2], fast);
          asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(f), "f"(v0) : "memory");
          asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(f + fplane), "f"(v1) : "memory");
          asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(f + 2 * fplane), "f"(v2) : "memory");
          continue;
        }
        const uint32_t px = __byte_perm(__byte_perm(c0, c1, 0x0040), c2, 0x5410);
        const uint32_t nxt_px = __shfl_down_sync(0xffffffffu, px, 1);
        if (k4 < 3) {
          const uint32_t w = __byte_perm(px, nxt_px, selO);
          asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(o), "r"(w) : "memory");
        }
        o += out_pitch;
      }
    }
    if (cur.r_end >= rows) break;
    // the next pass overwrites the band and the row table: order this pass's generic-proxy
    // reads before the bulk copies of the async proxy
    fence_proxy_async();
    __syncthreads();
    cur = plan(cur.r_end);
    issue(cur);
  }
}

// The tiles the band kernel left: the quad kernel over a list, with a fixed grid (the list is
// empty on the evaluation path, and the launch then costs a few microseconds).
template <bool NORM>
__global__ void __launch_bounds__(kWarpThreads)
    warp_affine_u8x3_list_kernel(const uint8_t* __restrict__ src,
                                 const int64_t* __restrict__ src_off,
                                 const int32_t* __restrict__ src_hw,
                                 const double* __restrict__ inv, void* __restrict__ dst, int dst_w,
                                 int dst_h, FastDiv div_wq, const int* __restrict__ todo,
                                 const NormArgs norm) {
  __shared__ __align__(16) QuadSmem sm;
  const int count = todo[0];
  const int tiles16 = (dst_h + kWarp3TileRows - 1) / kWarp3TileRows;
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    const int t = todo[1 + i];
    const int crop = t / tiles16;
    warp3_quad_tile<NORM, !NORM>(sm, src, src_off, src_hw, inv, dst, dst_w, dst_h, crop,
                                 t - crop * tiles16, div_wq, norm);
    __syncthreads();  // the tile's tables are rebuilt for the next one
  }
}

}  // namespace pc

using namespace pc;

static inline unsigned blocks_for(int64_t n, int threads) {
  return (unsigned)((n + threads - 1) / threads);
}

extern "C" int pc_box_to_center_scale(const float* d_boxes, float* d_center, float* d_scale,
                                      const pc_box_params* p, int64_t n, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_box_to_center_scale: params is NULL");
  PC_REQUIRE(n >= 0 && p->image_w > 0 && p->image_h > 0, PC_ERR_INVALID_ARGUMENT,
             "pc_box_to_center_scale: bad n / image size");
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_boxes && d_center && d_scale, PC_ERR_INVALID_ARGUMENT,
             "pc_box_to_center_scale: NULL tensor pointer");
  const double aspect = (double)p->image_w / (double)p->image_h;
  box_to_center_scale_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      d_boxes, d_center, d_scale, aspect, p->pixel_std, p->scale_padding, n);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_affine_matrices(const float* d_center, const float* d_scale,
                                  const float* d_rot, double* d_fwd, double* d_inv,
                                  const pc_affine_params* p, int64_t n, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_affine_matrices: params is NULL");
  PC_REQUIRE(n >= 0 && p->image_w > 0 && p->image_h > 0, PC_ERR_INVALID_ARGUMENT,
             "pc_affine_matrices: bad n / image size");
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_center && d_scale && (d_fwd || d_inv), PC_ERR_INVALID_ARGUMENT,
             "pc_affine_matrices: NULL tensor pointer");
  affine_matrices_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(
      d_center, d_scale, d_rot, d_fwd, d_inv, p->image_w, p->image_h, p->pixel_std, p->use_udp,
      n);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_affine_from_points(const float* d_src_points, const float* d_dst_points,
                                     double* d_fwd, double* d_inv, int64_t n, void* stream) {
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_affine_from_points: n < 0");
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_src_points && d_dst_points && (d_fwd || d_inv), PC_ERR_INVALID_ARGUMENT,
             "pc_affine_from_points: NULL tensor pointer");
  affine_from_points_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(
      d_src_points, d_dst_points, d_fwd, d_inv, n);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_invert_affine(const double* d_fwd, double* d_inv, int64_t n, void* stream) {
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_invert_affine: n < 0");
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_fwd && d_inv, PC_ERR_INVALID_ARGUMENT, "pc_invert_affine: NULL tensor pointer");
  invert_affine_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(d_fwd, d_inv, n);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_affine_joints(float* d_keypoints, const double* d_fwd, int32_t num_joints,
                                int32_t use_udp, int64_t n, void* stream) {
  PC_REQUIRE(n >= 0 && num_joints >= 1, PC_ERR_INVALID_ARGUMENT,
             "pc_affine_joints: bad n / num_joints");
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_keypoints && d_fwd, PC_ERR_INVALID_ARGUMENT,
             "pc_affine_joints: NULL tensor pointer");
  const int64_t total = n * num_joints;
  affine_joints_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      d_keypoints, d_fwd, num_joints, use_udp, total);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

// Does q0 = x * (1 / std) corrected by one residual step give the IEEE quotient x / std for all
// 256 pixel values of every channel?  (It does for the ImageNet statistics; any other mean /
// std is checked the same way and falls back to the division when one value differs.)
static bool norm_fast_ok(NormArgs* na) {
  bool ok = true;
  for (int c = 0; c < 3; ++c) {
    const float std = na->std[c], mean = na->mean[c];
    const float rcp = 1.0f / std;
    na->rcp[c] = rcp;
    for (int v = 0; v < 256 && ok; ++v) {
      const float x = (float)v - mean;
      const float q0 = x * rcp;
      const float q = fmaf(fmaf(-q0, std, x), rcp, q0);
      const float want = x / std;
      ok = memcmp(&q, &want, sizeof(float)) == 0;
    }
  }
  return ok;
}

// The band kernel takes 3-channel crops of a width that is a multiple of 32 (one thread per
// output column, whole warps) up to 512.
static bool band_path_takes(int dst_w, int dst_h, int64_t n) {
  const int tiles16 = (dst_h + kWarp3TileRows - 1) / kWarp3TileRows;
  return dst_w % 32 == 0 && dst_w <= kBandMaxThreads && n * (int64_t)tiles16 < 0x3fffffffLL;
}

// Rotation-free tiles out of a shared-memory band; tiles that do not qualify are listed in
// stream-ordered scratch (the library's own pool) and done by the quad kernel right after.
template <bool NORM>
static int launch_band_path(const uint8_t* d_src, const int64_t* d_src_offset,
                            const int32_t* d_src_hw, const double* d_inv, void* d_dst, int dst_w,
                            int dst_h, int64_t n, const NormArgs& na, cudaStream_t st) {
  const int tiles16 = (dst_h + kWarp3TileRows - 1) / kWarp3TileRows;
  // rows per CTA: per-CTA set-up (matrix, column constants) is paid once per tile, so tiles
  // are as tall as still leaves every SM a few rounds of CTAs
  int tile_rows = kBandMaxTileRows;
  const int64_t want = (int64_t)sm_count_cached() * 6 * 3;  // three rounds of CTAs
  while (tile_rows > kWarp3TileRows && n * ((dst_h + tile_rows - 1) / tile_rows) < want)
    tile_rows >>= 1;
  const int tiles_b = (dst_h + tile_rows - 1) / tile_rows;
  cudaMemPool_t pool;
  PC_CUDA(scratch_pool(&pool));
  int* todo = nullptr;
  PC_CUDA(cudaMallocFromPoolAsync((void**)&todo, sizeof(int) * (size_t)(1 + n * tiles16), pool, st));
  cudaError_t le = cudaMemsetAsync(todo, 0, sizeof(int), st);
  if (le == cudaSuccess) {
    warp_affine_u8x3_band_kernel<NORM><<<(unsigned)(n * tiles_b), dst_w, kBandBytes, st>>>(
        d_src, d_src_offset, d_src_hw, d_inv, d_dst, dst_w, dst_h, tile_rows, tiles_b, todo, na);
    le = cudaGetLastError();
  }
  if (le == cudaSuccess) {
    int64_t lgrid = (int64_t)sm_count_cached() * 4;  // 64 registers x 256 threads: 4 CTAs per SM
    // (capped at 48 registers for 5 CTAs per SM: rotated crops 1.52 instead of 1.49 ms)
    if (lgrid > n * tiles16) lgrid = n * tiles16;
    // (A programmatic dependent launch would hide this launch under the band kernel's tail --
    // 3 us on the evaluation path -- but the quad path lives on L1 hits, and CTAs that become
    // resident next to six 36 KB bands run with the SM's shared-memory carve-out: rotated
    // crops 3.45 ms instead of 1.48, measured.)
    warp_affine_u8x3_list_kernel<NORM><<<(unsigned)lgrid, kWarpThreads, 0, st>>>(
        d_src, d_src_offset, d_src_hw, d_inv, d_dst, dst_w, dst_h,
        make_fastdiv((uint32_t)(dst_w >> 2)), todo, na);
    le = cudaGetLastError();
  }
  const cudaError_t fe = cudaFreeAsync(todo, st);  // on every path
  PC_CUDA(le);
  PC_CUDA(fe);
  return PC_OK;
}

extern "C" int pc_warp_affine_u8(const uint8_t* d_src, const int64_t* d_src_offset,
                                 const int32_t* d_src_hw, const double* d_inv, uint8_t* d_dst,
                                 const pc_warp_params* p, int64_t n, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_warp_affine_u8: params is NULL");
  PC_REQUIRE(n >= 0 && p->dst_w >= 1 && p->dst_h >= 1, PC_ERR_INVALID_ARGUMENT,
             "pc_warp_affine_u8: bad n / destination size");
  PC_REQUIRE(p->channels >= 1 && p->channels <= 4, PC_ERR_UNSUPPORTED,
             "pc_warp_affine_u8: channels %d outside [1, 4]", p->channels);
  PC_REQUIRE(p->dst_w <= kWarpMaxDstW, PC_ERR_UNSUPPORTED,
             "pc_warp_affine_u8: dst_w %d > %d", p->dst_w, kWarpMaxDstW);
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_src && d_src_offset && d_src_hw && d_inv && d_dst, PC_ERR_INVALID_ARGUMENT,
             "pc_warp_affine_u8: NULL tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->channels == 3 && p->dst_w % 16 == 0 && ((uintptr_t)d_dst & 15) == 0) {
    const int tiles3 = (p->dst_h + kWarp3TileRows - 1) / kWarp3TileRows;
    const int64_t grid3 = n * tiles3;
    PC_REQUIRE(grid3 < 0x7fffffffLL, PC_ERR_UNSUPPORTED, "pc_warp_affine_u8: batch too large");
    if (band_path_takes(p->dst_w, p->dst_h, n))
      return launch_band_path<false>(d_src, d_src_offset, d_src_hw, d_inv, d_dst, p->dst_w,
                                     p->dst_h, n, NormArgs(), st);
    warp_affine_u8x3_kernel<false, true><<<(unsigned)grid3, kWarpThreads, 0, st>>>(
        d_src, d_src_offset, d_src_hw, d_inv, d_dst, p->dst_w, p->dst_h, tiles3,
        make_fastdiv((uint32_t)(p->dst_w >> 2)), NormArgs());
    PC_CUDA(cudaGetLastError());
    return PC_OK;
  }
  const int tiles = (p->dst_h + kWarpTileRows - 1) / kWarpTileRows;
  const int64_t grid = n * tiles;
  PC_REQUIRE(grid < 0x7fffffffLL, PC_ERR_UNSUPPORTED, "pc_warp_affine_u8: batch too large");
  const size_t smem = (size_t)kWarpTileRows * p->dst_w * p->channels;
  const FastDiv dw = make_fastdiv((uint32_t)p->dst_w);
#define PC_LAUNCH_WARP(CH)                                                              \
  warp_affine_u8_kernel<CH><<<(unsigned)grid, kWarpThreads, smem, st>>>(               \
      d_src, d_src_offset, d_src_hw, d_inv, d_dst, p->dst_w, p->dst_h, tiles, dw)
  switch (p->channels) {
    case 1: PC_LAUNCH_WARP(1); break;
    case 2: PC_LAUNCH_WARP(2); break;
    case 3: PC_LAUNCH_WARP(3); break;
    default: PC_LAUNCH_WARP(4); break;
  }
#undef PC_LAUNCH_WARP
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_warp_affine_u8_norm_chw(const uint8_t* d_src, const int64_t* d_src_offset,
                                          const int32_t* d_src_hw, const double* d_inv,
                                          float* d_dst, const pc_warp_norm_params* p, int64_t n,
                                          void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_warp_affine_u8_norm_chw: params is NULL");
  PC_REQUIRE(n >= 0 && p->dst_w >= 1 && p->dst_h >= 1, PC_ERR_INVALID_ARGUMENT,
             "pc_warp_affine_u8_norm_chw: bad n / destination size");
  PC_REQUIRE(p->channels == 3, PC_ERR_UNSUPPORTED,
             "pc_warp_affine_u8_norm_chw: channels %d (only 3-channel images)", p->channels);
  PC_REQUIRE(p->dst_w % 4 == 0 && p->dst_w <= kWarpMaxDstW, PC_ERR_UNSUPPORTED,
             "pc_warp_affine_u8_norm_chw: dst_w %d must be a multiple of 4 and <= %d", p->dst_w,
             kWarpMaxDstW);
  for (int c = 0; c < 3; ++c)
    PC_REQUIRE(p->std[c] != 0.f, PC_ERR_INVALID_ARGUMENT,
               "pc_warp_affine_u8_norm_chw: std[%d] is zero", c);
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_src && d_src_offset && d_src_hw && d_inv && d_dst, PC_ERR_INVALID_ARGUMENT,
             "pc_warp_affine_u8_norm_chw: NULL tensor pointer");
  PC_REQUIRE(((uintptr_t)d_dst & 15) == 0, PC_ERR_UNSUPPORTED,
             "pc_warp_affine_u8_norm_chw: destination must be 16-byte aligned");
  const int tiles3 = (p->dst_h + kWarp3TileRows - 1) / kWarp3TileRows;
  const int64_t grid3 = n * tiles3;
  PC_REQUIRE(grid3 < 0x7fffffffLL, PC_ERR_UNSUPPORTED,
             "pc_warp_affine_u8_norm_chw: batch too large");
  NormArgs na;
  for (int c = 0; c < 3; ++c) {
    na.mean[c] = p->mean[c];
    na.std[c] = p->std[c];
  }
  na.fast = norm_fast_ok(&na) ? 1 : 0;
  if (band_path_takes(p->dst_w, p->dst_h, n))
    return launch_band_path<true>(d_src, d_src_offset, d_src_hw, d_inv, d_dst, p->dst_w, p->dst_h,
                                  n, na, (cudaStream_t)stream);
  warp_affine_u8x3_kernel<true, false><<<(unsigned)grid3, kWarpThreads, 0, (cudaStream_t)stream>>>(
      d_src, d_src_offset, d_src_hw, d_inv, d_dst, p->dst_w, p->dst_h, tiles3,
      make_fastdiv((uint32_t)(p->dst_w >> 2)), na);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
