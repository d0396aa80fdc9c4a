// rescale.cu -- bottom-up evaluation preprocessing: bilinear rescale + pad + mask (sm_100a).
//
// Replaces, for a batch of images at once, the two validation transforms of the shipped
// HigherHRNet recipe (configs/higher_hrnet/higher_hrnet_w32_ascend.yaml:49-51):
//   BottomUpRescale.transform  mindpose/data/transform/bottomup_transform.py:170-209
//       cv2.resize(image, target, interpolation=cv2.INTER_LINEAR)
//   BottomUpPad.transform      bottomup_transform.py:610-648
//       np.pad to max_image_size (zeros right / below), mask = 1 on the image, 0 on the padding
// i.e. what sits between the decoded image and the network on the bottom-up path.  One pass:
// every byte of the canvas (image, padding) and of the mask is written once, the source is
// read once (its rows come back from L1 / L2 for the second tap and the next output row).
//
// Arithmetic = OpenCV's 8-bit bilinear resize (modules/imgproc/src/resize.cpp, restated in
// oracle/resize.py and pinned there against cv2 itself):
//   column dx: fx = float((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx;
//              sx < 0 -> (0, 0); sx >= src_w - 1 -> (src_w - 1, 0);
//              a0 = rint((1 - fx) * 2048), a1 = rint(fx * 2048)                  (int16)
//   row dy:    the same without zeroing the fraction; the two rows are clamped one by one
//   horizontal r = S[sx] * a0 + S[sx + 1] * a1                                  (int32)
//   vertical   (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
//   src == 2 * dst in both directions: the INTER_AREA fast path, (a + b + c + d + 2) >> 2.
// The target size of each image is the caller's (BottomUpRescale._get_new_size is integer /
// float64 host arithmetic with Python's round(): mindpose_b200/transforms.py).
#include "common.cuh"

namespace pc {

constexpr int kRescaleThreads = 256;
constexpr int kRescaleRows = 16;  // output rows per CTA

struct AxisTap {
  int s0, s1;  // source indices of the two taps (already clamped)
  int w0, w1;  // fixed-point weights (sum 2048 up to rounding)
};

// cv2's table entry of destination index d on an axis of src_n -> dst_n samples
__device__ __forceinline__ AxisTap axis_tap(int d, double scale, int src_n, bool zero_frac) {
  float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
  const float fl = floorf(f);
  int s = (int)fl;
  f = __fsub_rn(f, fl);
  if (zero_frac) {
    if (s < 0) s = 0, f = 0.f;
    if (s >= src_n - 1) s = src_n - 1, f = 0.f;
  }
  AxisTap t;
  t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  t.s0 = min(max(s, 0), src_n - 1);
  t.s1 = min(max(s + 1, 0), src_n - 1);
  return t;
}

// One thread: four consecutive output pixels (12 bytes, three aligned words) of kRescaleRows
// rows.  The horizontal pass of the lower source row is kept for the next output row, which
// usually starts from it (scale factors near 1).
__global__ void __launch_bounds__(kRescaleThreads)
    rescale_pad_u8x3_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ src_off,
                            const int32_t* __restrict__ src_hw, const int32_t* __restrict__ dst_wh,
                            uint8_t* __restrict__ dst, uint8_t* __restrict__ mask, int CW, int CH,
                            int tiles_per_image) {
  const int64_t img = blockIdx.x / tiles_per_image;
  const int row_begin = (blockIdx.x - (int)(img * tiles_per_image)) * kRescaleRows;
  const int row_end = min(row_begin + kRescaleRows, CH);
  const int sh = src_hw[2 * img], sw = src_hw[2 * img + 1];
  int tw = dst_wh[2 * img], th = dst_wh[2 * img + 1];
  if (sh < 1 || sw < 1) tw = th = 0;  // nothing to sample: the canvas is all padding
  tw = min(tw, CW);
  th = min(th, CH);
  const uint8_t* image = src + src_off[img];
  const size_t spitch = (size_t)sw * 3;
  uint8_t* out_img = dst + (size_t)img * CH * CW * 3;
  uint8_t* out_mask = mask ? mask + (size_t)img * CH * CW : nullptr;
  const bool area = tw > 0 && sw == 2 * tw && sh == 2 * th;
  const double scale_x = tw > 0 ? __ddiv_rn(1.0, __ddiv_rn((double)tw, (double)sw)) : 1.0;
  const double scale_y = th > 0 ? __ddiv_rn(1.0, __ddiv_rn((double)th, (double)sh)) : 1.0;

  for (int q = threadIdx.x; 4 * q < CW; q += kRescaleThreads) {
    const int x0 = 4 * q;
    const int npx = min(4, CW - x0);  // (CW % 4 != 0: the last thread of a row writes bytes)
    AxisTap tx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) tx[i] = axis_tap(min(x0 + i, max(tw - 1, 0)), scale_x, sw, true);
    int prev_y1 = -1;
    int r1[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) r1[e] = 0;
    // horizontal pass of source row y for the thread's four pixels
    auto hpass = [&](int y, int (&r)[12]) {
      const uint8_t* row = image + (size_t)y * spitch;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint8_t* p0 = row + tx[i].s0 * 3;
        const uint8_t* p1 = row + tx[i].s1 * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          r[3 * i + c] = (int)__ldg(p0 + c) * tx[i].w0 + (int)__ldg(p1 + c) * tx[i].w1;
      }
    };
    for (int y = row_begin; y < row_end; ++y) {
      uint32_t px[4] = {0u, 0u, 0u, 0u};  // packed r | g << 8 | b << 16 per pixel
      if (y < th && x0 < tw) {
        if (area) {
          const uint8_t* ra = image + (size_t)(2 * y) * spitch;
          const uint8_t* rb = ra + spitch;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (x0 + i < tw) {
              const uint8_t* a = ra + (size_t)(2 * (x0 + i)) * 3;
              const uint8_t* b = rb + (size_t)(2 * (x0 + i)) * 3;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const int v = ((int)__ldg(a + c) + (int)__ldg(a + 3 + c) + (int)__ldg(b + c) +
                               (int)__ldg(b + 3 + c) + 2) >> 2;
                px[i] |= (uint32_t)v << (8 * c);
              }
            }
          }
        } else {
          const AxisTap ty = axis_tap(y, scale_y, sh, false);
          int r0[12];
          if (ty.s0 == prev_y1) {
#pragma unroll
            for (int e = 0; e < 12; ++e) r0[e] = r1[e];
          } else {
            hpass(ty.s0, r0);
          }
          if (ty.s1 == ty.s0) {
#pragma unroll
            for (int e = 0; e < 12; ++e) r1[e] = r0[e];
          } else {
            hpass(ty.s1, r1);
          }
          prev_y1 = ty.s1;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (x0 + i < tw) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const int v = (((ty.w0 * (r0[3 * i + c] >> 4)) >> 16) +
                               ((ty.w1 * (r1[3 * i + c] >> 4)) >> 16) + 2) >> 2;
                px[i] |= (uint32_t)min(max(v, 0), 255) << (8 * c);
              }
            }
          }
        }
      }
      uint8_t* o = out_img + ((size_t)y * CW + x0) * 3;
      if (npx == 4 && ((reinterpret_cast<uintptr_t>(o) & 3u) == 0)) {
        // 12 bytes = r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
        uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
        o32[0] = px[0] | (px[1] << 24);
        o32[1] = (px[1] >> 8) | (px[2] << 16);
        o32[2] = (px[2] >> 16) | (px[3] << 8);
      } else {
        for (int i = 0; i < npx; ++i)
          for (int c = 0; c < 3; ++c) o[3 * i + c] = (uint8_t)(px[i] >> (8 * c));
      }
      if (out_mask) {
        uint8_t* m = out_mask + (size_t)y * CW + x0;
        uint32_t bits = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (y < th && x0 + i < tw) bits |= 1u << (8 * i);
        if (npx == 4 && ((reinterpret_cast<uintptr_t>(m) & 3u) == 0)) {
          *reinterpret_cast<uint32_t*>(m) = bits;
        } else {
          for (int i = 0; i < npx; ++i) m[i] = (uint8_t)(bits >> (8 * i));
        }
      }
    }
  }
}

}  // namespace pc

using namespace pc;

extern "C" int pc_rescale_pad_u8(const uint8_t* d_src, const int64_t* d_src_offset,
                                 const int32_t* d_src_hw, const int32_t* d_dst_wh,
                                 uint8_t* d_dst, uint8_t* d_mask, int32_t canvas_w,
                                 int32_t canvas_h, int32_t channels, int64_t n, void* stream) {
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_rescale_pad_u8: n < 0");
  PC_REQUIRE(channels == 3, PC_ERR_UNSUPPORTED,
             "pc_rescale_pad_u8: %d channels (the pipeline's images are 3-channel uint8)",
             channels);
  PC_REQUIRE(canvas_w >= 1 && canvas_h >= 1 && canvas_w <= 16384 && canvas_h <= 16384,
             PC_ERR_INVALID_ARGUMENT, "pc_rescale_pad_u8: canvas %d x %d outside [1, 16384]",
             canvas_w, canvas_h);
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_src && d_src_offset && d_src_hw && d_dst_wh && d_dst, PC_ERR_INVALID_ARGUMENT,
             "pc_rescale_pad_u8: NULL tensor pointer");
  const int tiles = (canvas_h + kRescaleRows - 1) / kRescaleRows;
  PC_REQUIRE(n * tiles < 0x7fffffffLL, PC_ERR_UNSUPPORTED, "pc_rescale_pad_u8: batch too large");
  rescale_pad_u8x3_kernel<<<(unsigned)(n * tiles), kRescaleThreads, 0, (cudaStream_t)stream>>>(
      d_src, d_src_offset, d_src_hw, d_dst_wh, d_dst, d_mask, canvas_w, canvas_h, tiles);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
