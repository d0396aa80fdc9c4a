// bottomup_decode.cu -- fused HigherHRNet bottom-up decode for sm_100a.
//
// Replaces BottomUpHeatMapAEDecoder.construct
// (mindpose/models/decoders/bottom_up_decoder.py:67-203): decouple_output,
// multi-resolution aggregation (legacy asymmetric bilinear upsampling of the
// low-resolution heat plane + the high-resolution plane, / num_stages), mask,
// max-pool NMS, top-M selection (value desc, flat index asc), tag gather through
// the bilinearly resized tag plane, x = ind % W, y = ind // W.
//
// One CTA per (image, joint).  The aggregated plane is produced row by row into a
// 16-row ring in shared memory (each input element is read from HBM once; the
// low-resolution taps and the mask are re-read through L1/L2), the NMS window is
// evaluated out of the ring, and survivors stream into a top-M list that lives in
// the registers of warp 0 (lane i = rank i).  A running threshold (the current
// M-th value) keeps the candidate traffic tiny after the first tile; a tile that
// still produces many candidates is first cut down with the M-th largest
// per-thread maximum, which is a valid lower bound for the M-th largest element.
#include <math.h>

#include "common.cuh"

namespace pc {

constexpr int kBuThreads = 256;
constexpr int kBuTileRows = 8;
constexpr int kBuRing = 16;
constexpr int kBuMaxW = 512;
constexpr int kBuMaxNms = 7;
constexpr int kBuPxPerThread = (kBuTileRows * kBuMaxW) / kBuThreads;  // 16
constexpr int kBuRefineAbove = 96;

struct BuArgs {
  const float* out0;
  const float* out1;
  const uint8_t* mask;
  float* val_k;
  float* tag_k;
  float* ind_k;
  float* heatmap_raw;
  float* tagging;
  int32_t K, stages, h0, w0, h1, w1, mh, mw;
  int32_t use_nms, nms_k, M;
  FastDiv div_w1;
  float sy, sx;    // h0 / h1, w0 / w1 (float32, as the resize computes them)
  float msy, msx;  // mh / h1, mw / w1
};

// legacy asymmetric bilinear sample of a [h, w] plane at destination (y, x)
__device__ __forceinline__ float bilinear_legacy(const float* __restrict__ p, int h, int w,
                                                 float sy, float sx, int y, int x) {
  const float ys = __fmul_rn((float)y, sy), xs = __fmul_rn((float)x, sx);
  const float y0f = floorf(ys), x0f = floorf(xs);
  const int y0 = (int)y0f, x0 = (int)x0f;
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float fy = __fsub_rn(ys, y0f), fx = __fsub_rn(xs, x0f);
  const float tl = __ldg(p + y0 * w + x0), tr = __ldg(p + y0 * w + x1);
  const float bl = __ldg(p + y1 * w + x0), br = __ldg(p + y1 * w + x1);
  const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), fx));
  const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), fx));
  return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), fy));
}

__device__ __forceinline__ bool beats(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

__global__ void __launch_bounds__(kBuThreads) bottomup_decode_kernel(const BuArgs a) {
  extern __shared__ __align__(16) unsigned char bu_smem[];
  float* s_ring = reinterpret_cast<float*>(bu_smem);             // [kBuRing][W]
  float* s_cval = s_ring + kBuRing * a.w1;                         // [kBuTileRows * W]
  int* s_cidx = reinterpret_cast<int*>(s_cval + kBuTileRows * a.w1);
  float* s_tmax = reinterpret_cast<float*>(s_cidx + kBuTileRows * a.w1);  // [kBuThreads]
  __shared__ int s_ncand;
  __shared__ float s_thr, s_bound;
  __shared__ int s_full;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x / a.K, k = blockIdx.x - n * a.K;
  const int H = a.h1, W = a.w1, M = a.M;

  const float* heat_hi;  // [H, W] plane at output resolution
  const float* heat_lo;  // [h0, w0] plane to upsample (stages == 2)
  const float* tag_src;  // tag plane (resolution of out0)
  int th, tw;
  float tsy, tsx;
  if (a.stages == 2) {
    heat_lo = a.out0 + ((size_t)n * 2 * a.K + k) * a.h0 * a.w0;
    tag_src = a.out0 + ((size_t)n * 2 * a.K + a.K + k) * a.h0 * a.w0;
    heat_hi = a.out1 + ((size_t)n * a.K + k) * H * W;
    th = a.h0;
    tw = a.w0;
    tsy = a.sy;
    tsx = a.sx;
  } else {
    heat_lo = nullptr;
    heat_hi = a.out0 + ((size_t)n * 2 * a.K + k) * H * W;
    tag_src = a.out0 + ((size_t)n * 2 * a.K + a.K + k) * H * W;
    th = H;
    tw = W;
    tsy = 1.f;
    tsx = 1.f;
  }
  const uint8_t* mask = a.mask + (size_t)n * a.mh * a.mw;
  float* raw_out = a.heatmap_raw ? a.heatmap_raw + ((size_t)n * a.K + k) * H * W : nullptr;
  float* tag_out = a.tagging ? a.tagging + ((size_t)n * a.K + k) * H * W : nullptr;

  if (tid == 0) {
    s_thr = -INFINITY;
    s_full = 0;
    s_ncand = 0;
  }
  // top-M list: lane i of warp 0 holds rank i
  float top_v = -INFINITY;
  int top_i = 0x7fffffff;
  int top_count = 0;

  const int lo = a.use_nms ? (a.nms_k - 1) / 2 : 0;
  const int hi = a.use_nms ? a.nms_k - 1 - lo : 0;
  int filled = 0;  // rows [0, filled) have been aggregated into the ring
  __syncthreads();

  for (int r0 = 0; r0 < H; r0 += kBuTileRows) {
    const int rows = min(kBuTileRows, H - r0);
    // ---- 1. aggregate the rows this tile needs: [filled, min(H, r0 + rows + hi))
    const int need = min(H, r0 + rows + hi);
    const int new_px = (need - filled) * W;
    for (int e = tid; e < new_px; e += kBuThreads) {
      const int dy = (int)fdiv((uint32_t)e, a.div_w1);
      const int x = e - dy * W, y = filled + dy;
      float v = __ldg(heat_hi + y * W + x);
      if (a.stages == 2) {
        v = __fadd_rn(v, bilinear_legacy(heat_lo, a.h0, a.w0, a.sy, a.sx, y, x));
        v = __fmul_rn(v, 0.5f);  // / num_stages (2): exact
      }
      const int my = min((int)floorf(__fmul_rn((float)y, a.msy)), a.mh - 1);
      const int mx = min((int)floorf(__fmul_rn((float)x, a.msx)), a.mw - 1);
      if (mask[my * a.mw + mx] == 0) v = 0.f;
      s_ring[(y & (kBuRing - 1)) * W + x] = v;
      if (raw_out) raw_out[y * W + x] = v;
      if (tag_out) tag_out[y * W + x] = bilinear_legacy(tag_src, th, tw, tsy, tsx, y, x);
    }
    filled = need;
    __syncthreads();

    // ---- 2. NMS + candidate test for this tile's pixels
    const float thr = s_thr;
    const int full = s_full;
    const int npx = rows * W;
    float mv[kBuPxPerThread];
    int ncand_local = 0;
    float tmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < kBuPxPerThread; ++i) {
      const int p = tid + i * kBuThreads;
      mv[i] = -INFINITY;
      if (p < npx) {
        const int dy = (int)fdiv((uint32_t)p, a.div_w1);
        const int x = p - dy * W, y = r0 + dy;
        const float v = s_ring[(y & (kBuRing - 1)) * W + x];
        float m = v;
        if (a.use_nms) {
          float pooled = -INFINITY;
          for (int yy = max(0, y - lo); yy <= min(H - 1, y + hi); ++yy) {
            const float* row = s_ring + (yy & (kBuRing - 1)) * W;
            for (int xx = max(0, x - lo); xx <= min(W - 1, x + hi); ++xx)
              pooled = fmaxf(pooled, row[xx]);
          }
          m = __fmul_rn(v, pooled == v ? 1.f : 0.f);
        }
        if (!full || m > thr) {
          mv[i] = m;
          ++ncand_local;
          tmax = fmaxf(tmax, m);
        } else {
          mv[i] = __int_as_float(0xffc00000);  // NaN marks "not a candidate"
        }
      } else {
        mv[i] = __int_as_float(0xffc00000);
      }
    }
    int pos = 0;
    if (ncand_local) pos = atomicAdd(&s_ncand, ncand_local);
    __syncthreads();
    float bound = -INFINITY;
    if (s_ncand > kBuRefineAbove) {
      // ---- 3. many candidates: the M-th largest per-thread maximum bounds the M-th element
      s_tmax[tid] = tmax;
      __syncthreads();
      int rank = 0;
      for (int t = 0; t < kBuThreads; ++t) {
        const float o = s_tmax[t];
        rank += (o > tmax || (o == tmax && t < tid)) ? 1 : 0;
      }
      if (tid == 0) s_bound = -INFINITY;
      __syncthreads();
      if (rank == M - 1 && ncand_local) s_bound = tmax;
      if (tid == 0) s_ncand = 0;
      __syncthreads();
      bound = s_bound;
      ncand_local = 0;
#pragma unroll
      for (int i = 0; i < kBuPxPerThread; ++i)
        if (mv[i] == mv[i] && mv[i] >= bound) ++ncand_local;
      pos = ncand_local ? atomicAdd(&s_ncand, ncand_local) : 0;
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < kBuPxPerThread; ++i) {
      if (mv[i] == mv[i] && mv[i] >= bound) {
        s_cval[pos] = mv[i];
        s_cidx[pos] = r0 * W + tid + i * kBuThreads;
        ++pos;
      }
    }
    __syncthreads();

    // ---- 4. warp 0 merges the candidates into the sorted top-M list
    if (warp == 0) {
      const int nc = s_ncand;
      for (int c = 0; c < nc; ++c) {
        const float v = s_cval[c];
        const int idx = s_cidx[c];
        const bool mine_beats = lane < top_count && beats(top_v, top_i, v, idx);
        const int p = __popc(__ballot_sync(0xffffffffu, mine_beats));
        if (p < M) {
          const float uv = __shfl_up_sync(0xffffffffu, top_v, 1);
          const int ui = __shfl_up_sync(0xffffffffu, top_i, 1);
          if (lane > p) {
            top_v = uv;
            top_i = ui;
          } else if (lane == p) {
            top_v = v;
            top_i = idx;
          }
          top_count = min(top_count + 1, M);
        }
      }
      const float last = __shfl_sync(0xffffffffu, top_v, M - 1);
      if (lane == 0) {
        s_full = top_count == M;
        s_thr = top_count == M ? last : -INFINITY;
        s_ncand = 0;
      }
    }
    __syncthreads();
  }

  // ---- 5. results: value, (x, y), tag through the resized tag plane
  if (warp == 0 && lane < M) {
    const size_t o = ((size_t)n * a.K + k) * M + lane;
    const int y = (int)fdiv((uint32_t)top_i, a.div_w1);
    const int x = top_i - y * W;
    a.val_k[o] = top_v;
    a.ind_k[2 * o] = (float)x;
    a.ind_k[2 * o + 1] = (float)y;
    a.tag_k[o] = bilinear_legacy(tag_src, th, tw, tsy, tsx, y, x);
  }
}

}  // namespace pc

using namespace pc;

extern "C" int pc_bottomup_decode(const float* d_out0, const float* d_out1,
                                  const uint8_t* d_mask, float* d_val_k, float* d_tag_k,
                                  float* d_ind_k, float* d_heatmap_raw, float* d_tagging,
                                  const pc_bottomup_decode_params* p, int64_t n, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_bottomup_decode: params is NULL");
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_bottomup_decode: n < 0");
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_bottomup_decode: num_joints %d outside [1, %d]", p->num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(p->num_stages == 1 || p->num_stages == 2, PC_ERR_UNSUPPORTED,
             "pc_bottomup_decode: num_stages %d (only 1 or 2)", p->num_stages);
  PC_REQUIRE(p->h1 >= 1 && p->w1 >= 1 && p->w1 <= kBuMaxW, PC_ERR_UNSUPPORTED,
             "pc_bottomup_decode: output map %dx%d (width must be <= %d)", p->h1, p->w1, kBuMaxW);
  PC_REQUIRE(p->num_stages == 1 || (p->h0 >= 1 && p->w0 >= 1), PC_ERR_INVALID_ARGUMENT,
             "pc_bottomup_decode: bad stage-0 size");
  PC_REQUIRE(p->mask_h >= 1 && p->mask_w >= 1, PC_ERR_INVALID_ARGUMENT,
             "pc_bottomup_decode: bad mask size");
  PC_REQUIRE(p->max_num >= 1 && p->max_num <= 32, PC_ERR_UNSUPPORTED,
             "pc_bottomup_decode: max_num %d outside [1, 32]", p->max_num);
  PC_REQUIRE((int64_t)p->h1 * p->w1 >= p->max_num, PC_ERR_INVALID_ARGUMENT,
             "pc_bottomup_decode: map smaller than max_num");
  PC_REQUIRE(!p->use_nms || (p->nms_kernel >= 1 && p->nms_kernel <= kBuMaxNms),
             PC_ERR_UNSUPPORTED, "pc_bottomup_decode: nms_kernel %d outside [1, %d]",
             p->nms_kernel, kBuMaxNms);
  PC_REQUIRE(!p->shift_coordinate, PC_ERR_UNSUPPORTED,
             "pc_bottomup_decode: shift_coordinate=True is not supported (the reference pairs "
             "its offsets with the wrong candidates, bottom_up_decoder.py:195-201)");
  PC_REQUIRE((uint64_t)p->h1 * p->w1 * p->w1 < 0xffffffffull, PC_ERR_UNSUPPORTED,
             "pc_bottomup_decode: map too large");
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_out0 && (p->num_stages == 1 || d_out1) && d_mask && d_val_k && d_tag_k && d_ind_k,
             PC_ERR_INVALID_ARGUMENT, "pc_bottomup_decode: NULL tensor pointer");
  PC_REQUIRE(n * p->num_joints < 0x7fffffffLL, PC_ERR_UNSUPPORTED,
             "pc_bottomup_decode: batch too large");

  BuArgs a;
  a.out0 = d_out0;
  a.out1 = d_out1;
  a.mask = d_mask;
  a.val_k = d_val_k;
  a.tag_k = d_tag_k;
  a.ind_k = d_ind_k;
  a.heatmap_raw = d_heatmap_raw;
  a.tagging = d_tagging;
  a.K = p->num_joints;
  a.stages = p->num_stages;
  a.h0 = p->h0;
  a.w0 = p->w0;
  a.h1 = p->h1;
  a.w1 = p->w1;
  a.mh = p->mask_h;
  a.mw = p->mask_w;
  a.use_nms = p->use_nms;
  a.nms_k = p->nms_kernel;
  a.M = p->max_num;
  a.div_w1 = make_fastdiv((uint32_t)p->w1);
  a.sy = p->num_stages == 2 ? (float)p->h0 / (float)p->h1 : 1.f;
  a.sx = p->num_stages == 2 ? (float)p->w0 / (float)p->w1 : 1.f;
  a.msy = (float)p->mask_h / (float)p->h1;
  a.msx = (float)p->mask_w / (float)p->w1;
  const size_t smem = sizeof(float) * ((size_t)(kBuRing + 2 * kBuTileRows) * p->w1 + kBuThreads);
  if (smem > 48 * 1024)
    PC_CUDA(cudaFuncSetAttribute(bottomup_decode_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bottomup_decode_kernel<<<(unsigned)(n * p->num_joints), kBuThreads, smem,
                           (cudaStream_t)stream>>>(a);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
