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

constexpr int kGenericThreads = 256;
constexpr int kGenericRows = 16;  // output rows per CTA of the plain kernel
constexpr int kRescaleMaxThreads = 256;  // columns per CTA (one thread each)
constexpr int kRescaleRows = 32;         // output rows per CTA
#ifndef PC_RESCALE_AHEAD
#define PC_RESCALE_AHEAD 4
#endif
#ifndef PC_RESCALE_MINB
#define PC_RESCALE_MINB 6  // CTAs of 256 threads per SM the register budget is cut for
#endif
constexpr int kRescaleAhead = PC_RESCALE_AHEAD;  // source rows prefetched ahead of the walk

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

// The plain form, for what the column kernel below does not take (a canvas whose rows do not
// start on word boundaries, a source one pixel wide): one thread = four consecutive output
// pixels of kGenericRows rows, byte loads, the tables evaluated in place.
__global__ void __launch_bounds__(kGenericThreads)
    rescale_pad_generic_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ src_off,
                            const int32_t* __restrict__ src_hw, const int32_t* __restrict__ dst_wh,
                            uint8_t* __restrict__ dst, uint8_t* __restrict__ mask, int CW, int CH,
                            int tiles_per_image) {
  const int64_t img = blockIdx.x / tiles_per_image;
  const int row_begin = (blockIdx.x - (int)(img * tiles_per_image)) * kGenericRows;
  const int row_end = min(row_begin + kGenericRows, CH);
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

  for (int q = threadIdx.x; 4 * q < CW; q += kGenericThreads) {
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


// cv2's table entry of destination index d: first source index, fraction
__device__ __forceinline__ void axis_entry(int d, double scale, int src_n, bool zero_frac,
                                           int& s, float& f) {
  f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
  const float fl = floorf(f);
  s = (int)fl;
  f = __fsub_rn(f, fl);
  if (zero_frac) {
    if (s < 0) s = 0, f = 0.f;
    if (s >= src_n - 1) s = src_n - 1, f = 0.f;
  }
}

struct RescaleRow {
  uint32_t off0, off1;  // byte offsets of the two source rows in the image
  uint32_t b0, b1;      // vertical weights << 16; both 0 on a row below the image (-> pixel 0)
};

// (a * b) >> 32 as its own instruction: two of them and one three-input add per channel (left
// to the compiler the pair becomes a 64-bit multiply-add chain with extra register moves)
__device__ __forceinline__ uint32_t mulhi_u32(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// keeps a per-thread constant in its register (the compiler would rebuild it in the loop)
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}

// vision.Normalize(mean * 255, std * 255) of the step after the transforms
// (mindpose/data/data_factory.py:127-138), as in the fused crop warp (warp_affine.cu, row N2):
// the IEEE quotient, or x * (1 / std) plus one exact residual step when the host has verified
// that form for all 256 pixel values of every channel.
struct RescaleNorm {
  float mean[3], std[3], rcp[3];
  int32_t fast;
};
__device__ __forceinline__ float rescale_norm_value(uint32_t px, float mean, float std, float rcp,
                                                    bool fast) {
  const float x = __fsub_rn((float)px, mean);
  if (!fast) return __fdiv_rn(x, std);
  const float q0 = __fmul_rn(x, rcp);
  return __fmaf_rn(__fmaf_rn(-q0, std, x), rcp, q0);
}

// Where a thread's pixel of each row goes.  NORM = false: uint8 HWC, a warp's 32 pixels as 24
// words (lanes 4g .. 4g+3 hold the pixels of columns 4c .. 4c+3; three of them store a word
// each).  NORM = true: float32 CHW planes, Normalize fused (the uint8 canvas is never written;
// the zero padding becomes (0 - mean) / std, as Normalize makes it in the reference's pipeline).
template <bool NORM>
struct RescaleOut {
  uint8_t* o;        // !NORM: this lane's word of its group of four pixels
  float* f;          // NORM: this column of channel 0's plane
  uint8_t* m;        // mask byte of this column (dereferenced only if store_m)
  uint32_t opitch, fplane, CW, selO;
  bool store_px, store_m;
  RescaleNorm norm;

  __device__ __forceinline__ void row(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t mask_value) {
    if (NORM) {
      if (store_px) {
        const bool fast = norm.fast != 0;
        f[0] = rescale_norm_value(v0, norm.mean[0], norm.std[0], norm.rcp[0], fast);
        f[fplane] = rescale_norm_value(v1, norm.mean[1], norm.std[1], norm.rcp[1], fast);
        f[2 * (size_t)fplane] = rescale_norm_value(v2, norm.mean[2], norm.std[2], norm.rcp[2], fast);
      }
      f += CW;
    } else {
      const uint32_t px = __byte_perm(__byte_perm(v0, v1, 0x0040), v2, 0x5410);
      const uint32_t nxt = __shfl_down_sync(0xffffffffu, px, 1);
      if (store_px) *reinterpret_cast<uint32_t*>(o) = __byte_perm(px, nxt, selO);
      o += opitch;
    }
    if (store_m) *m = (uint8_t)mask_value;
    m += CW;
  }
};

// This thread's column of one tile.  AREA: the exact 2 x 2 reduction.  PITCH4: the source row
// pitch is a multiple of 4 bytes, so the alignment of the thread's window -- and with it the
// byte-permute selectors -- is the same in every row.
template <bool AREA, bool PITCH4, bool NORM>
__device__ __forceinline__ void rescale_column(const uint8_t* __restrict__ image, uint32_t win,
                                               uint32_t wpair, const RescaleRow* s_row, int rows,
                                               RescaleOut<NORM>& out, uint32_t mval) {
  uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(image) + win) & 3u;
  uint32_t selA = pinned(0x4130u + 0x1111u * a);              // (A, B) -> r0 r1 g0 g1
  uint32_t selB = pinned(((6u + a) & 7u) | ((1u + a) << 4));  // (B, Z) -> b0 b1 . .
  bool third = a == 3u;                                        // the window reaches a third word
  const uint8_t* base = image + win - a;  // word aligned (PITCH4: in every row)
  // horizontal pass of the source row at byte offset `off`: h[c] = S0[c] * w0 + S1[c] * w1
  auto hpass = [&](uint32_t off, uint32_t (&h)[3]) {
    const uint8_t* p = base + off;
    if (!PITCH4) {
      const uint32_t d = (uint32_t)reinterpret_cast<uintptr_t>(p) & 3u;  // any phase
      const uint32_t ar = a + d;   // the window starts ar (0 .. 6) bytes into the word at p - d
      p = p - d + (ar & 4u);       // ... re-aimed at the word that holds its first byte
      const uint32_t aa = ar & 3u;
      selA = 0x4130u + 0x1111u * aa;
      selB = ((6u + aa) & 7u) | ((1u + aa) << 4);
      third = aa == 3u;
    }
    const uint32_t A = __ldg(reinterpret_cast<const uint32_t*>(p));
    const uint32_t B = __ldg(reinterpret_cast<const uint32_t*>(p + 4));
    uint32_t Z = A;
    if (third) Z = __ldg(reinterpret_cast<const uint32_t*>(p + 8));
    const uint32_t rg = __byte_perm(A, B, selA), bb = __byte_perm(B, Z, selB);
    h[0] = __dp2a_lo(wpair, rg, 0u);
    h[1] = __dp2a_hi(wpair, rg, 0u);
    h[2] = __dp2a_lo(wpair, bb, 0u);
    if (!AREA) h[0] >>= 4, h[1] >>= 4, h[2] >>= 4;
  };
  // Every output row starts a source row nobody has touched yet: without help each step is
  // one DRAM round trip (measured: the kernel ran at the latency of 32 of them per tile).
  // The rows kRescaleAhead steps on are prefetched into L2 as the walk goes.
  auto prefetch_row = [&](int j) {
    const uint32_t off = s_row[j].off1;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
  };
  {
    const uint32_t off = s_row[0].off0;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
  }
  for (int j = 0; j < min(kRescaleAhead, rows); ++j) prefetch_row(j);
  uint32_t prev_off1 = 0xffffffffu;
  // one output row: hx = horizontal pass of the upper source row, hy = of the lower one.  The
  // caller alternates the two register sets, so that "the upper row is the previous lower
  // row" (scale factors near 1) costs nothing.
  auto step = [&](int i, uint32_t (&hx)[3], uint32_t (&hy)[3]) {
    const uint4 rr = *reinterpret_cast<const uint4*>(s_row + i);  // off0, off1, b0, b1
    if (i + kRescaleAhead < rows) {
      prefetch_row(i + kRescaleAhead);
      if (AREA) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + s_row[i + kRescaleAhead].off0));
    }
    if (AREA || rr.x != prev_off1) hpass(rr.x, hx);
    hpass(rr.y, hy);  // (the same row again where the two are clamped together: border rows)
    prev_off1 = rr.y;
    // (b0 * (r0 >> 4) >> 16) + (b1 * (r1 >> 4) >> 16) + 2 >> 2, in [0, 255] (the weights sum
    // to 2048 +- 1); the area path: (sum of four + 2) >> 2.  Columns right of the image have
    // weight pair 0 and rows below it b0 = b1 = 0: their pixels come out 0.
    uint32_t v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      v[c] = AREA ? ((rr.z ? hx[c] + hy[c] : 0u) + 2u) >> 2
                  : (mulhi_u32(rr.z, hx[c]) + mulhi_u32(rr.w, hy[c]) + 2u) >> 2;
    out.row(v[0], v[1], v[2], (rr.z | rr.w) ? mval : 0u);
  };
  uint32_t ha[3] = {0u, 0u, 0u}, hb[3] = {0u, 0u, 0u};
#pragma unroll 1
  for (int i = 0; i < rows; i += 2) {
    step(i, ha, hb);
    if (i + 1 < rows) step(i + 1, hb, ha);
  }
}

// One thread per output column, kRescaleRows rows per CTA, walking down: the horizontal pass of
// a source row (the two or three aligned words that hold the two taps' six bytes, two byte
// permutes, three dp2a) is kept in registers for the next output row, which usually starts
// from it.  Lanes are neighbouring columns, so a warp's loads of one source row fall into one
// or two 128-byte lines; a warp's 32 pixels (96 bytes) leave as 24 words (one shuffle, one
// byte permute).  Needs canvas rows that start on word boundaries (canvas_w % 4 == 0) and
// sources at least two pixels wide; row offsets are 32-bit (images below 4 GB).
template <bool NORM>
__global__ void __launch_bounds__(kRescaleMaxThreads, PC_RESCALE_MINB)
    rescale_pad_u8x3_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ src_off,
                            const int32_t* __restrict__ src_hw, const int32_t* __restrict__ dst_wh,
                            void* __restrict__ dst_any, uint8_t* __restrict__ mask, int CW, int CH,
                            int col_blocks, int row_tiles, const RescaleNorm norm) {
  __shared__ __align__(16) RescaleRow s_row[kRescaleRows];
  const int tid = threadIdx.x, lane = tid & 31;
  int b = blockIdx.x;
  const int cb = b % col_blocks;
  b /= col_blocks;
  const int rt = b % row_tiles;
  const int64_t img = b / row_tiles;
  const int row_begin = rt * kRescaleRows;
  const int rows = min(kRescaleRows, CH - row_begin);
  const int x = cb * (int)blockDim.x + tid;

  const int sh = src_hw[2 * img], sw = src_hw[2 * img + 1];
  int tw = dst_wh[2 * img], th = dst_wh[2 * img + 1];
  if (sh < 1 || sw < 1 || tw < 1 || th < 1) tw = th = 0;  // nothing to sample: all padding
  tw = min(tw, CW);
  th = min(th, CH);
  const uint8_t* image = src + src_off[img];
  const uint32_t spitch = (uint32_t)sw * 3u;
  const bool area = tw > 0 && sw == 2 * tw && sh == 2 * th;

  const uint32_t k4 = (uint32_t)lane & 3u;
  RescaleOut<NORM> out;
  out.opitch = (uint32_t)CW * 3u;
  out.fplane = (uint32_t)CH * (uint32_t)CW;
  out.CW = (uint32_t)CW;
  out.o = static_cast<uint8_t*>(dst_any) + ((size_t)img * CH + row_begin) * out.opitch +
          (size_t)(x >> 2) * 12 + (size_t)k4 * 4;
  out.f = static_cast<float*>(dst_any) + ((size_t)img * 3 * CH + row_begin) * CW + x;
  out.store_px = x < CW && (NORM || k4 < 3);
  out.store_m = mask != nullptr && x < CW;
  out.m = mask + ((size_t)img * CH + row_begin) * CW + x;
  out.selO = pinned(k4 == 0 ? 0x4210u : (k4 == 1 ? 0x5421u : 0x6542u));
  out.norm = norm;

  if (tw == 0 || row_begin >= th || sw == 1) {
    // (uniform) A tile of padding, or a source one pixel wide (no second pixel to make a load
    // window with: every column is source column 0 with weight 2048).
    const double scale_y = th > 0 ? __ddiv_rn(1.0, __ddiv_rn((double)th, (double)sh)) : 1.0;
    for (int i = 0; i < rows; ++i) {
      const int y = row_begin + i;
      uint32_t px = 0u;
      const bool inside = x < tw && y < th;
      if (inside) {
        int s;
        float f;
        axis_entry(y, scale_y, sh, false, s, f);
        const uint32_t b0 = (uint32_t)__float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
        const uint32_t b1 = (uint32_t)__float2int_rn(__fmul_rn(f, 2048.f));
        const uint8_t* p0 = image + (size_t)min(max(s, 0), sh - 1) * spitch;
        const uint8_t* p1 = image + (size_t)min(max(s + 1, 0), sh - 1) * spitch;
        for (int c = 0; c < 3; ++c) {
          const uint32_t r0 = ((uint32_t)p0[c] * 2048u) >> 4, r1 = ((uint32_t)p1[c] * 2048u) >> 4;
          px |= ((((b0 * r0) >> 16) + ((b1 * r1) >> 16) + 2u) >> 2) << (8 * c);
        }
      }
      out.row(px & 0xffu, (px >> 8) & 0xffu, px >> 16, inside ? 1u : 0u);
    }
    return;
  }

  // ---- the tile's row table
  if (tid < rows) {
    RescaleRow r;
    const int y = row_begin + tid;
    r.off0 = r.off1 = 0u, r.b0 = r.b1 = 0u;
    if (y < th) {
      if (area) {
        r.off0 = (uint32_t)(2 * y) * spitch;
        r.off1 = r.off0 + spitch;
        r.b0 = r.b1 = 1u;
      } else {
        const double scale_y = __ddiv_rn(1.0, __ddiv_rn((double)th, (double)sh));
        int s;
        float f;
        axis_entry(y, scale_y, sh, false, s, f);
        r.off0 = (uint32_t)min(max(s, 0), sh - 1) * spitch;
        r.off1 = (uint32_t)min(max(s + 1, 0), sh - 1) * spitch;
        r.b0 = (uint32_t)__float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f)) << 16;
        r.b1 = (uint32_t)__float2int_rn(__fmul_rn(f, 2048.f)) << 16;
      }
    }
    s_row[tid] = r;
  }

  // ---- this thread's column: byte offset of its two-pixel window in a source row, the dp2a
  // weight pair.  A column whose second tap is clamped away (weight 0, last source pixel)
  // reads the window one pixel to the left with the weights swapped, so that no load leaves
  // the row.  Columns right of the image read window 0 with weights 0.
  uint32_t win = 0u, wpair = 0u;
  if (x < tw) {
    if (area) {
      win = (uint32_t)(2 * x) * 3u;
      wpair = 0x00010001u;
    } else {
      const double scale_x = __ddiv_rn(1.0, __ddiv_rn((double)tw, (double)sw));
      int s;
      float f;
      axis_entry(x, scale_x, sw, true, s, f);
      const uint32_t w0 = (uint32_t)__float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
      const uint32_t w1 = (uint32_t)__float2int_rn(__fmul_rn(f, 2048.f));
      if (s + 1 <= sw - 1) {
        win = (uint32_t)s * 3u;
        wpair = w0 | (w1 << 16);
      } else {  // s == sw - 1, f == 0: (pixel s - 1) * 0 + (pixel s) * w0
        win = (uint32_t)(s - 1) * 3u;
        wpair = w0 << 16;
      }
    }
  }
  const uint32_t mval = x < tw ? 1u : 0u;
  __syncthreads();
  if (x - lane >= tw) {  // (warp-uniform) every column of this warp is padding
    for (int i = 0; i < rows; ++i) out.row(0u, 0u, 0u, 0u);
    return;
  }
  const bool pitch4 = (spitch & 3u) == 0u;
#define PC_RESCALE_GO(A_, P_) \
  rescale_column<A_, P_, NORM>(image, win, wpair, s_row, rows, out, mval)
  if (area) {
    if (pitch4) PC_RESCALE_GO(true, true);
    else PC_RESCALE_GO(true, false);
  } else {
    if (pitch4) PC_RESCALE_GO(false, true);
    else PC_RESCALE_GO(false, false);
  }
#undef PC_RESCALE_GO
}

}  // namespace pc

using namespace pc;

// Does q0 = x * (1 / std) corrected by one residual step give the IEEE quotient x / std for all
// 256 pixel values of every channel?  (As in warp_affine.cu.)
static bool rescale_norm_fast_ok(RescaleNorm* na) {
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

static int rescale_launch(const char* who, const uint8_t* d_src, const int64_t* d_src_offset,
                          const int32_t* d_src_hw, const int32_t* d_dst_wh, void* d_dst,
                          uint8_t* d_mask, int32_t canvas_w, int32_t canvas_h, int32_t channels,
                          const RescaleNorm* norm, int64_t n, void* stream) {
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "%s: n < 0", who);
  PC_REQUIRE(channels == 3, PC_ERR_UNSUPPORTED,
             "%s: %d channels (the pipeline's images are 3-channel uint8)", who, channels);
  PC_REQUIRE(canvas_w >= 1 && canvas_h >= 1 && canvas_w <= 16384 && canvas_h <= 16384,
             PC_ERR_INVALID_ARGUMENT, "%s: canvas %d x %d outside [1, 16384]", who, canvas_w,
             canvas_h);
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_src && d_src_offset && d_src_hw && d_dst_wh && d_dst, PC_ERR_INVALID_ARGUMENT,
             "%s: NULL tensor pointer", who);
  cudaStream_t st = (cudaStream_t)stream;
  // The uint8 column kernel writes a warp's 32 pixels as 24 aligned words: it needs every row
  // of the canvas to start on a word boundary.  Any other uint8 canvas takes the plain kernel.
  if (!norm && (canvas_w % 4 != 0 || (reinterpret_cast<uintptr_t>(d_dst) & 3u) != 0)) {
    const int tiles = (canvas_h + kGenericRows - 1) / kGenericRows;
    PC_REQUIRE(n * tiles < 0x7fffffffLL, PC_ERR_UNSUPPORTED, "%s: batch too large", who);
    rescale_pad_generic_kernel<<<(unsigned)(n * tiles), kGenericThreads, 0, st>>>(
        d_src, d_src_offset, d_src_hw, d_dst_wh, static_cast<uint8_t*>(d_dst), d_mask, canvas_w,
        canvas_h, tiles);
    PC_CUDA(cudaGetLastError());
    return PC_OK;
  }
  const int row_tiles = (canvas_h + kRescaleRows - 1) / kRescaleRows;
  const int col_blocks = (canvas_w + kRescaleMaxThreads - 1) / kRescaleMaxThreads;
  // columns per CTA: the canvas width split evenly, rounded up to whole warps
  const int threads = (((canvas_w + col_blocks - 1) / col_blocks) + 31) & ~31;
  PC_REQUIRE(n * row_tiles * col_blocks < 0x7fffffffLL, PC_ERR_UNSUPPORTED,
             "%s: batch too large", who);
  const unsigned grid = (unsigned)(n * row_tiles * col_blocks);
  if (norm) {
    rescale_pad_u8x3_kernel<true><<<grid, threads, 0, st>>>(
        d_src, d_src_offset, d_src_hw, d_dst_wh, d_dst, d_mask, canvas_w, canvas_h, col_blocks,
        row_tiles, *norm);
  } else {
    RescaleNorm none;
    memset(&none, 0, sizeof(none));
    rescale_pad_u8x3_kernel<false><<<grid, threads, 0, st>>>(
        d_src, d_src_offset, d_src_hw, d_dst_wh, d_dst, d_mask, canvas_w, canvas_h, col_blocks,
        row_tiles, none);
  }
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_rescale_pad_u8(const uint8_t* d_src, const int64_t* d_src_offset,
                                 const int32_t* d_src_hw, const int32_t* d_dst_wh,
                                 uint8_t* d_dst, uint8_t* d_mask, int32_t canvas_w,
                                 int32_t canvas_h, int32_t channels, int64_t n, void* stream) {
  return rescale_launch("pc_rescale_pad_u8", d_src, d_src_offset, d_src_hw, d_dst_wh, d_dst,
                        d_mask, canvas_w, canvas_h, channels, nullptr, n, stream);
}

extern "C" int pc_rescale_pad_u8_norm_chw(const uint8_t* d_src, const int64_t* d_src_offset,
                                          const int32_t* d_src_hw, const int32_t* d_dst_wh,
                                          float* d_dst, uint8_t* d_mask,
                                          const pc_warp_norm_params* params, int64_t n,
                                          void* stream) {
  PC_REQUIRE(params != nullptr, PC_ERR_INVALID_ARGUMENT,
             "pc_rescale_pad_u8_norm_chw: params is NULL");
  RescaleNorm na;
  for (int c = 0; c < 3; ++c) {
    na.mean[c] = params->mean[c];
    na.std[c] = params->std[c];
    PC_REQUIRE(params->std[c] != 0.f, PC_ERR_INVALID_ARGUMENT,
               "pc_rescale_pad_u8_norm_chw: std[%d] is 0", c);
  }
  na.fast = rescale_norm_fast_ok(&na) ? 1 : 0;
  return rescale_launch("pc_rescale_pad_u8_norm_chw", d_src, d_src_offset, d_src_hw, d_dst_wh,
                        d_dst, d_mask, params->dst_w, params->dst_h, params->channels, &na, n,
                        stream);
}
