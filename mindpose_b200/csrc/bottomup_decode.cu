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
  int32_t c0;        // channels of out0: 2K (a tag plane per joint) or K + 1 (one shared plane)
  int32_t tag_step;  // 1, or 0 when every joint reads the same tag plane (tag_per_joint False)
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
    heat_lo = a.out0 + ((size_t)n * a.c0 + k) * a.h0 * a.w0;
    tag_src = a.out0 + ((size_t)n * a.c0 + a.K + k * a.tag_step) * a.h0 * a.w0;
    heat_hi = a.out1 + ((size_t)n * a.K + k) * H * W;
    th = a.h0;
    tw = a.w0;
    tsy = a.sy;
    tsx = a.sx;
  } else {
    heat_lo = nullptr;
    heat_hi = a.out0 + ((size_t)n * a.c0 + k) * H * W;
    tag_src = a.out0 + ((size_t)n * a.c0 + a.K + k * a.tag_step) * H * W;
    th = H;
    tw = W;
    tsy = 1.f;
    tsx = 1.f;
  }
  const uint8_t* mask = a.mask + (size_t)n * a.mh * a.mw;
  float* raw_out = a.heatmap_raw ? a.heatmap_raw + ((size_t)n * a.K + k) * H * W : nullptr;
  // tagging_heatmap output: K planes per image, or the one shared plane (written by joint 0)
  float* tag_out = (a.tagging && (a.tag_step || k == 0))
                       ? a.tagging + ((size_t)n * (a.tag_step ? a.K : 1) + k * a.tag_step) * H * W
                       : nullptr;

  if (tid == 0) {
    s_thr = -INFINITY;
    s_full = 0;
    s_ncand = 0;
  }
  // top-M list (M <= 64): lane i of warp 0 holds ranks i and i + 32
  float top_v[2] = {-INFINITY, -INFINITY};
  int top_i[2] = {0x7fffffff, 0x7fffffff};
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
        const bool b0 = lane < top_count && beats(top_v[0], top_i[0], v, idx);
        const bool b1 = lane + 32 < top_count && beats(top_v[1], top_i[1], v, idx);
        const int p = __popc(__ballot_sync(0xffffffffu, b0)) + __popc(__ballot_sync(0xffffffffu, b1));
        if (p < M) {
          // shift ranks p.. up by one (rank 31 carries into rank 32) and insert at p
          const float cv = __shfl_sync(0xffffffffu, top_v[0], 31);
          const int ci = __shfl_sync(0xffffffffu, top_i[0], 31);
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            float uv = __shfl_up_sync(0xffffffffu, top_v[s], 1);
            int ui = __shfl_up_sync(0xffffffffu, top_i[s], 1);
            if (s == 1 && lane == 0) uv = cv, ui = ci;
            const int r = lane + 32 * s;
            if (r > p) {
              top_v[s] = uv;
              top_i[s] = ui;
            } else if (r == p) {
              top_v[s] = v;
              top_i[s] = idx;
            }
          }
          top_count = min(top_count + 1, M);
        }
      }
      const float last = M <= 32 ? __shfl_sync(0xffffffffu, top_v[0], (M - 1) & 31)
                                 : __shfl_sync(0xffffffffu, top_v[1], (M - 1) & 31);
      if (lane == 0) {
        s_full = top_count == M;
        s_thr = top_count == M ? last : -INFINITY;
        s_ncand = 0;
      }
    }
    __syncthreads();
  }

  // ---- 5. results: value, (x, y), tag through the resized tag plane
  if (warp == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int r = lane + 32 * s;
      if (r >= M) continue;
      const size_t o = ((size_t)n * a.K + k) * M + r;
      const int y = (int)fdiv((uint32_t)top_i[s], a.div_w1);
      const int x = top_i[s] - y * W;
      a.val_k[o] = top_v[s];
      a.ind_k[2 * o] = (float)x;
      a.ind_k[2 * o + 1] = (float)y;
      a.tag_k[o] = bilinear_legacy(tag_src, th, tw, tsy, tsx, y, x);
    }
  }
}


// ---------------------------------------------------------------------------
// Fast path: register-resident NMS, per-lane top-3, one merge per plane.
//
// The generic kernel above walks a plane tile by tile with several block-wide
// barriers per tile and funnels every candidate through warp 0; it reaches only a
// few percent of the HBM roofline.  Here one CTA of 8 warps owns one (image, joint)
// plane and every warp owns a band of rows on its own -- no block barrier until the
// merge:
//   * lane l holds C consecutive columns (32 * C >= W); a row of the aggregated map
//     is 128-bit loads of the high-resolution plane, the two low-resolution rows it
//     is interpolated from (the right neighbour comes from the next lane by
//     shuffle) and the mask bytes, combined in registers;
//   * the 3x3 max-pool is separable: the horizontal 3-max of each row (neighbour
//     columns by shuffle) is kept for the previous, current and next row, the
//     vertical max of the three decides the current row -- no shared memory;
//   * pass 1: every lane keeps the best THREE survivors of its own C x R pixel
//     region in registers (~4 instructions per pixel).  The union of these (768
//     entries) holds the plane's top M unless one lane owns four or more of them;
//   * merge: the lane bests are sorted per warp, ranked across warps by binary
//     search, the M-th gives a threshold T; the few entries >= T are compacted and
//     ranked exactly (value desc, flat index asc);
//   * check: if some lane's third entry beats the M-th result a fourth may have
//     been dropped -> pass 2 re-scans the plane (it is in L2) with exact per-warp
//     sorted lists (lane i = rank i), pre-filtered by the M-th result, and merges
//     those.  Rare on real maps (< 1 % of planes), bounded on any input.
// Taken when nms_kernel is 1 or 3 (or NMS is off), W % C == 0, and (two stages) the
// low-resolution map is exactly half the size; anything else uses the generic kernel.
constexpr int kFastWarps = 8;
constexpr int kFastThreads = kFastWarps * 32;

// planes that needed the exact second pass since the last reset (pc_bottomup_decode_stats)
__device__ unsigned long long g_bu_exact_planes = 0ull;

template <int C>
struct RowRegs {
  float v[C];
};

__device__ __forceinline__ void topm_insert(float v, int idx, float& top_v, int& top_i,
                                            int& count, int M, int lane) {
  const bool mine_beats = lane < count && beats(top_v, top_i, v, idx);
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
    count = min(count + 1, M);
  }
}

// C columns (x0 .. x0+C-1) of row y of the aggregated, masked map.
template <int C, bool TWO_STAGE, bool MASK2X>
__device__ __forceinline__ void aggregate_row(const BuArgs& a, const float* __restrict__ heat_hi,
                                              const float* __restrict__ heat_lo,
                                              const uint8_t* __restrict__ mask, int y, int x0,
                                              bool active, bool last_lane, float (&v)[C]) {
  const int W = a.w1;
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = -INFINITY;
  if (active) {
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
      const float4 t = ld_stream_f4(heat_hi + (size_t)y * W + x0 + 4 * q);
      v[4 * q] = t.x;
      v[4 * q + 1] = t.y;
      v[4 * q + 2] = t.z;
      v[4 * q + 3] = t.w;
    }
  }
  if (TWO_STAGE) {
    // legacy asymmetric bilinear, scale exactly 1/2: src = dst * 0.5
    const int w0 = a.w0, h0 = a.h0;
    const float ys = __fmul_rn((float)y, 0.5f);
    const float y0f = floorf(ys);
    const int y0 = (int)y0f, y1 = min(y0 + 1, h0 - 1);
    const float fy = __fsub_rn(ys, y0f);
    constexpr int L = C / 2;  // low-resolution columns owned by this lane
    float la[L + 1], lb[L + 1];
#pragma unroll
    for (int j = 0; j <= L; ++j) la[j] = lb[j] = 0.f;
    if (active) {
      const float* ra = heat_lo + (size_t)y0 * w0 + (x0 >> 1);
      const float* rb = heat_lo + (size_t)y1 * w0 + (x0 >> 1);
      if (L % 4 == 0) {
#pragma unroll
        for (int q = 0; q < L / 4; ++q) {
          const float4 ta = __ldg(reinterpret_cast<const float4*>(ra) + q);
          const float4 tb = __ldg(reinterpret_cast<const float4*>(rb) + q);
          la[4 * q] = ta.x, la[4 * q + 1] = ta.y, la[4 * q + 2] = ta.z, la[4 * q + 3] = ta.w;
          lb[4 * q] = tb.x, lb[4 * q + 1] = tb.y, lb[4 * q + 2] = tb.z, lb[4 * q + 3] = tb.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < L / 2; ++q) {
          const float2 ta = __ldg(reinterpret_cast<const float2*>(ra) + q);
          const float2 tb = __ldg(reinterpret_cast<const float2*>(rb) + q);
          la[2 * q] = ta.x, la[2 * q + 1] = ta.y;
          lb[2 * q] = tb.x, lb[2 * q + 1] = tb.y;
        }
      }
    }
    // right neighbour: first low-res column of the next lane, clamped at the edge
    const float na = __shfl_down_sync(0xffffffffu, la[0], 1);
    const float nb = __shfl_down_sync(0xffffffffu, lb[0], 1);
    la[L] = last_lane ? la[L - 1] : na;
    lb[L] = last_lane ? lb[L - 1] : nb;
    if (active) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int j = c >> 1;
        const float fx = (c & 1) ? 0.5f : 0.f;
        const float top = __fadd_rn(la[j], __fmul_rn(__fsub_rn(la[j + 1], la[j]), fx));
        const float bot = __fadd_rn(lb[j], __fmul_rn(__fsub_rn(lb[j + 1], lb[j]), fx));
        const float up = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), fy));
        v[c] = __fmul_rn(__fadd_rn(v[c], up), 0.5f);  // / num_stages (2): exact
      }
    }
  }
  if (active) {
    const int my = min((int)floorf(__fmul_rn((float)y, a.msy)), a.mh - 1);
    const uint8_t* mrow = mask + (size_t)my * a.mw;
    if (MASK2X) {  // mask is exactly twice as wide: nearest source column = 2 * x
#pragma unroll
      for (int q = 0; q < C / 8; ++q) {
        const uint4 mb = __ldg(reinterpret_cast<const uint4*>(mrow + 2 * x0) + q);
        const uint32_t w[4] = {mb.x, mb.y, mb.z, mb.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if ((w[t] & 0xffu) == 0) v[8 * q + 2 * t] = 0.f;
          if ((w[t] & 0xff0000u) == 0) v[8 * q + 2 * t + 1] = 0.f;
        }
      }
      if (C % 8 == 4) {
        const uint2 mb = __ldg(reinterpret_cast<const uint2*>(mrow + 2 * x0 + (C / 8) * 16));
        const uint32_t w[2] = {mb.x, mb.y};
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if ((w[t] & 0xffu) == 0) v[(C / 8) * 8 + 2 * t] = 0.f;
          if ((w[t] & 0xff0000u) == 0) v[(C / 8) * 8 + 2 * t + 1] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int mx = min((int)floorf(__fmul_rn((float)(x0 + c), a.msx)), a.mw - 1);
        if (__ldg(mrow + mx) == 0) v[c] = 0.f;
      }
    }
  }
}

// ---- lane-private cp.async ring (C = 8, two stages, 2x mask) ---------------------
// Every lane needs exactly 80 bytes of a row: its 8 high-resolution columns (2 x 16 B),
// its 4 low-resolution columns of the two source rows (2 x 16 B) and its 16 mask bytes.
// They are copied global -> shared with cp.async one row ahead, so the DRAM round trip
// of row r+1 overlaps the arithmetic of row r at no register cost.  A lane only ever
// reads back what it copied itself: no barrier, just cp.async.wait_group.
constexpr int kStageSeg = 512;                 // 32 lanes x 16 B
constexpr int kStageSlot = 5 * kStageSeg;      // hi0 | hi1 | lowA | lowB | mask
constexpr int kStageWarp = 2 * kStageSlot;     // two rows in flight per warp

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_pending() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_1() {
  asm volatile("cp.async.wait_group 1;" ::: "memory");
}

struct StagedRows {
  unsigned char* ring;  // this warp's kStageWarp bytes
  int first, last;      // rows are consumed in order first .. last
};

__device__ __forceinline__ void staged_issue(const BuArgs& a, const float* __restrict__ heat_hi,
                                             const float* __restrict__ heat_lo,
                                             const uint8_t* __restrict__ mask,
                                             const StagedRows& st, int r, int lane, bool active) {
  if (r <= st.last && active) {
    unsigned char* slot = st.ring + ((r - st.first) & 1) * kStageSlot + lane * 16;
    const int x0 = lane * 8;
    const float* hi = heat_hi + (size_t)r * a.w1 + x0;
    cp_async16(slot, hi);
    cp_async16(slot + kStageSeg, hi + 4);
    const int y0 = r >> 1, y1 = min(y0 + 1, a.h0 - 1);
    cp_async16(slot + 2 * kStageSeg, heat_lo + (size_t)y0 * a.w0 + (x0 >> 1));
    cp_async16(slot + 3 * kStageSeg, heat_lo + (size_t)y1 * a.w0 + (x0 >> 1));
    const int my = min((int)floorf(__fmul_rn((float)r, a.msy)), a.mh - 1);
    cp_async16(slot + 4 * kStageSeg, mask + (size_t)my * a.mw + 2 * x0);
  }
  cp_async_commit();  // one group per row, empty past the end: wait_group 1 stays exact
}

// Row r (the oldest row in flight) of the aggregated, masked map; then re-arms its slot
// with row r + 2.  Same arithmetic as aggregate_row<8, true, true>; for the exact 1/2
// scale the interpolation weights are 0 or 0.5 and "a + (b - a) * 0" is written as "a"
// (identical for finite maps; only the sign of an exact zero can differ).
__device__ __forceinline__ void staged_row(const BuArgs& a, const float* __restrict__ heat_hi,
                                           const float* __restrict__ heat_lo,
                                           const uint8_t* __restrict__ mask,
                                           const StagedRows& st, int r, int lane, bool active,
                                           bool last_lane, float (&v)[8]) {
  cp_async_wait_1();
  const unsigned char* slot = st.ring + ((r - st.first) & 1) * kStageSlot + lane * 16;
  float4 h0 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), h1 = h0;
  float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
  uint4 mb = make_uint4(0, 0, 0, 0);
  if (active) {
    h0 = *reinterpret_cast<const float4*>(slot);
    h1 = *reinterpret_cast<const float4*>(slot + kStageSeg);
    qa = *reinterpret_cast<const float4*>(slot + 2 * kStageSeg);
    qb = *reinterpret_cast<const float4*>(slot + 3 * kStageSeg);
    mb = *reinterpret_cast<const uint4*>(slot + 4 * kStageSeg);
  }
  staged_issue(a, heat_hi, heat_lo, mask, st, r + 2, lane, active);
  float la[5] = {qa.x, qa.y, qa.z, qa.w, 0.f}, lb[5] = {qb.x, qb.y, qb.z, qb.w, 0.f};
  const float na = __shfl_down_sync(0xffffffffu, la[0], 1);
  const float nb = __shfl_down_sync(0xffffffffu, lb[0], 1);
  la[4] = last_lane ? la[3] : na;
  lb[4] = last_lane ? lb[3] : nb;
  v[0] = h0.x, v[1] = h0.y, v[2] = h0.z, v[3] = h0.w;
  v[4] = h1.x, v[5] = h1.y, v[6] = h1.z, v[7] = h1.w;
  if (active) {
    const bool odd_row = r & 1;  // fy = 0.5 on odd rows, 0 on even rows
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = c >> 1;
      float top = la[j], bot = lb[j];
      if (c & 1) {
        top = __fadd_rn(la[j], __fmul_rn(__fsub_rn(la[j + 1], la[j]), 0.5f));
        bot = __fadd_rn(lb[j], __fmul_rn(__fsub_rn(lb[j + 1], lb[j]), 0.5f));
      }
      const float up = odd_row ? __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), 0.5f)) : top;
      v[c] = __fmul_rn(__fadd_rn(v[c], up), 0.5f);
    }
    const uint32_t w[4] = {mb.x, mb.y, mb.z, mb.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if ((w[t] & 0xffu) == 0) v[2 * t] = 0.f;
      if ((w[t] & 0xff0000u) == 0) v[2 * t + 1] = 0.f;
    }
  }
}

// horizontal 3-max of a row held C columns per lane
template <int C>
__device__ __forceinline__ void hmax3(const float (&v)[C], bool first_lane, bool last_lane,
                                      float (&h)[C]) {
  float left = __shfl_up_sync(0xffffffffu, v[C - 1], 1);
  float right = __shfl_down_sync(0xffffffffu, v[0], 1);
  if (first_lane) left = -INFINITY;
  if (last_lane) right = -INFINITY;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float l = c == 0 ? left : v[c - 1];
    const float r = c == C - 1 ? right : v[c + 1];
    h[c] = fmaxf(fmaxf(l, v[c]), r);
  }
}


struct Top3 {
  float v1, v2, v3;
  int i1, i2, i3;
};

struct ExactList {
  float top_v;  // lane i = rank i
  int top_i;
  int count;
  float last_v;
  int last_i;
  float pre_v;  // only candidates that are not beaten by (pre_v, pre_i) matter
  int pre_i;
};

// Scan rows [rb, re) of the plane: aggregate, NMS, feed survivors to the sink.
//
// NMS state per lane: A = max(hm[y-1], hm[y]) and hm[y] (hm = horizontal 3-max of a row),
// v[y]; when row y+1 arrives, pooled(y) = max(A, hm[y+1]) and A becomes max(hm[y], hm[y+1]).
// The row loop is unrolled by two so the (v, hm) buffers swap roles without register moves.
template <int C, bool TWO_STAGE, bool MASK2X, bool EXACT, bool STAGED>
__device__ __forceinline__ void scan_band(const BuArgs& a, const float* __restrict__ heat_hi,
                                          const float* __restrict__ heat_lo,
                                          const uint8_t* __restrict__ mask, float* raw_out,
                                          int rb, int re, int lane, bool nms, int M, Top3& t3,
                                          ExactList& ex, volatile float* s_kth, int warp_id,
                                          unsigned char* ring) {
  const int H = a.h1, W = a.w1;
  const int kth_rank = (M + kFastWarps - 1) / kFastWarps;
  float t_lb = -INFINITY;
  const int x0 = lane * C;
  const bool active = x0 < W;
  const bool first_lane = lane == 0;
  const bool last_lane = x0 + C >= W;  // also true for inactive lanes
  float A[C], hA[C], hB[C], vA[C], vB[C];
#pragma unroll
  for (int c = 0; c < C; ++c) A[c] = hA[c] = hB[c] = vA[c] = vB[c] = -INFINITY;
  StagedRows st;
  st.ring = ring;
  st.first = (nms && rb > 0) ? rb - 1 : rb;
  st.last = (nms && re < H) ? re : re - 1;
  if (STAGED) {
    staged_issue(a, heat_hi, heat_lo, mask, st, st.first, lane, active);
    staged_issue(a, heat_hi, heat_lo, mask, st, st.first + 1, lane, active);
  }
  // next row of the aggregated map, in order
#define PC_BU_ROW(ROW, OUT)                                                                 \
  do {                                                                                      \
    if constexpr (STAGED)                                                                   \
      staged_row(a, heat_hi, heat_lo, mask, st, (ROW), lane, active, last_lane, (OUT));     \
    else                                                                                    \
      aggregate_row<C, TWO_STAGE, MASK2X>(a, heat_hi, heat_lo, mask, (ROW), x0, active,     \
                                          last_lane, (OUT));                                \
  } while (0)

  // Publish this warp's kth-largest lane best (k = ceil(M / warps)) and refresh t_lb, a
  // lower bound of the plane's M-th best VALUE: at least k elements of every band are >=
  // its kth, so the minimum over the bands has >= M elements above it.  Stale reads of
  // other warps' slots are only smaller, i.e. still valid.
  auto refresh_bound = [&]() {
    float x = t3.v1, kth = -INFINITY;
    for (int r = 0; r < kth_rank; ++r) {
      float mx = x;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const unsigned b = __ballot_sync(0xffffffffu, x == mx);
      if (lane == __ffs(b) - 1) x = -INFINITY;
      kth = mx;
    }
    if (lane == 0) s_kth[warp_id] = kth;
    float o8 = lane < kFastWarps ? s_kth[lane] : INFINITY;
#pragma unroll
    for (int o = kFastWarps / 2; o > 0; o >>= 1)
      o8 = fminf(o8, __shfl_xor_sync(0xffffffffu, o8, o));
    o8 = __shfl_sync(0xffffffffu, o8, 0);
    t_lb = fmaxf(t_lb, o8);
  };

  // one row: vc / hc hold row y, vn / hn receive row y + 1
  auto step = [&](int y, float (&vc)[C], float (&hc)[C], float (&vn)[C], float (&hn)[C]) {
    if (!EXACT && raw_out && active) {
#pragma unroll
      for (int q = 0; q < C / 4; ++q)
        st_stream_f4(raw_out + (size_t)y * W + x0 + 4 * q,
                     make_float4(vc[4 * q], vc[4 * q + 1], vc[4 * q + 2], vc[4 * q + 3]));
    }
    const bool more = y + 1 < re || (nms && y + 1 < H);
    if (more) {
      PC_BU_ROW(y + 1, vn);
      if (nms) hmax3<C>(vn, first_lane, last_lane, hn);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) hn[c] = -INFINITY;
    }
    const int row_base = y * W + x0;
    float m[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      m[c] = vc[c];
      if (nms) {
        const float pooled = fmaxf(A[c], hn[c]);
        m[c] = __fmul_rn(m[c], pooled == m[c] ? 1.f : 0.f);
        A[c] = fmaxf(hc[c], hn[c]);
      }
    }
    if (!EXACT) {
      // Per-lane top 3.  A lane sees its pixels in increasing flat index and takes the
      // qualifying values of a row largest first (equal values left to right), so an
      // equal value never displaces an entry: strict compares.  Values under t_lb cannot
      // matter; values equal to it can (ties are resolved by index), hence >=.
      float rowmax = m[0];
#pragma unroll
      for (int c = 1; c < C; ++c) rowmax = fmaxf(rowmax, m[c]);
      bool want = active && rowmax > t3.v3 && rowmax >= t_lb;
      if (__any_sync(0xffffffffu, want)) {
        while (want) {
          int cb = 0;
#pragma unroll
          for (int c = C - 1; c >= 0; --c)
            if (m[c] == rowmax) cb = c;
          const int idx = row_base + cb;
          if (rowmax > t3.v1) {
            t3.v3 = t3.v2, t3.i3 = t3.i2;
            t3.v2 = t3.v1, t3.i2 = t3.i1;
            t3.v1 = rowmax, t3.i1 = idx;
          } else if (rowmax > t3.v2) {
            t3.v3 = t3.v2, t3.i3 = t3.i2;
            t3.v2 = rowmax, t3.i2 = idx;
          } else {
            t3.v3 = rowmax, t3.i3 = idx;
          }
          rowmax = -INFINITY;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if (c == cb) m[c] = -INFINITY;
            rowmax = fmaxf(rowmax, m[c]);
          }
          want = rowmax > t3.v3 && rowmax >= t_lb;
        }
      }
      if (((y - rb) & 7) == 7) refresh_bound();
    } else {
      bool cand[C];
      bool any = false;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int idx = row_base + c;
        cand[c] = active && !beats(ex.pre_v, ex.pre_i, m[c], idx) &&
                  (ex.count < M || beats(m[c], idx, ex.last_v, ex.last_i));
        any |= cand[c];
      }
      if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int idx = row_base + c;
          // the list may have moved since cand[c] was taken: test again
          const bool cnow = cand[c] &&
                            (ex.count < M || beats(m[c], idx, ex.last_v, ex.last_i));
          unsigned bits = __ballot_sync(0xffffffffu, cnow);
          if (bits) {
            while (bits) {
              const int src = __ffs(bits) - 1;
              bits &= bits - 1;
              const float cv = __shfl_sync(0xffffffffu, m[c], src);
              const int ci = __shfl_sync(0xffffffffu, idx, src);
              topm_insert(cv, ci, ex.top_v, ex.top_i, ex.count, M, lane);
            }
            ex.last_v = __shfl_sync(0xffffffffu, ex.top_v, M - 1);
            ex.last_i = __shfl_sync(0xffffffffu, ex.top_i, M - 1);
          }
        }
      }
    }
  };

  if (nms && rb > 0) {
    PC_BU_ROW(rb - 1, vA);
    hmax3<C>(vA, first_lane, last_lane, hB);  // hB = hm[rb - 1]
  }
  PC_BU_ROW(rb, vA);
  if (nms) {
    hmax3<C>(vA, first_lane, last_lane, hA);
#pragma unroll
    for (int c = 0; c < C; ++c) A[c] = fmaxf(hB[c], hA[c]);
  }
  for (int y = rb; y < re; y += 2) {
    step(y, vA, hA, vB, hB);
    if (y + 1 < re) step(y + 1, vB, hB, vA, hA);
  }
#undef PC_BU_ROW
  if (!EXACT) refresh_bound();  // final value of this band for the merge
  if (STAGED) asm volatile("cp.async.wait_all;" ::: "memory");
}

// entries of the sorted list (lv, li)[0..n) that beat (v, i)
__device__ __forceinline__ int count_beating(const float* lv, const int* li, int n, float v,
                                             int i) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (beats(lv[mid], li[mid], v, i))
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

constexpr int kFastBuf = 3 * kFastThreads;

template <int C, bool TWO_STAGE, bool MASK2X>
__global__ void __launch_bounds__(kFastThreads, (C <= 8) ? 3 : 1)
    bottomup_decode_fast_kernel(const BuArgs a) {
  __shared__ float s_lv[kFastWarps][32];
  __shared__ int s_li[kFastWarps][32];
  __shared__ int s_cnt[kFastWarps];
  __shared__ float s_bv[kFastBuf];
  __shared__ int s_bi[kFastBuf];
  __shared__ float s_ov[32];
  __shared__ int s_oi[32];
  __shared__ int s_nbuf, s_nout, s_ti;
  __shared__ float s_tv;
  __shared__ float s_kth[kFastWarps];

  constexpr bool STAGED = (C == 8) && TWO_STAGE && MASK2X;
  extern __shared__ __align__(16) unsigned char s_ring[];  // kFastWarps * kStageWarp if STAGED

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x / a.K, k = blockIdx.x - n * a.K;
  const int H = a.h1, W = a.w1, M = a.M;
  const bool nms = a.use_nms && a.nms_k == 3;
  unsigned char* ring = STAGED ? s_ring + warp * kStageWarp : nullptr;

  const float* heat_hi;
  const float* heat_lo = nullptr;
  const float* tag_src;
  int th, tw;
  float tsy, tsx;
  if (TWO_STAGE) {
    heat_lo = a.out0 + ((size_t)n * a.c0 + k) * a.h0 * a.w0;
    tag_src = a.out0 + ((size_t)n * a.c0 + a.K + k * a.tag_step) * a.h0 * a.w0;
    heat_hi = a.out1 + ((size_t)n * a.K + k) * H * W;
    th = a.h0, tw = a.w0, tsy = a.sy, tsx = a.sx;
  } else {
    heat_hi = a.out0 + ((size_t)n * a.c0 + k) * H * W;
    tag_src = a.out0 + ((size_t)n * a.c0 + a.K + k * a.tag_step) * H * W;
    th = H, tw = W, tsy = 1.f, tsx = 1.f;
  }
  const uint8_t* mask = a.mask + (size_t)n * a.mh * a.mw;
  float* raw_out = a.heatmap_raw ? a.heatmap_raw + ((size_t)n * a.K + k) * H * W : nullptr;

  const int R = (H + kFastWarps - 1) / kFastWarps;
  const int rb = min(H, warp * R), re = min(H, rb + R);

  if (tid == 0) {
    s_nbuf = 0;
    s_nout = 0;
    s_tv = -INFINITY;  // threshold "nothing is excluded"
    s_ti = 0x7fffffff;
  }
  if (tid < kFastWarps) s_kth[tid] = -INFINITY;
  __syncthreads();

  // ---- pass 1: per-lane top 3 ---------------------------------------------------
  Top3 t3;
  t3.v1 = t3.v2 = t3.v3 = -INFINITY;
  t3.i1 = t3.i2 = t3.i3 = 0x7fffffff;
  ExactList ex;
  ex.top_v = -INFINITY, ex.top_i = 0x7fffffff, ex.count = 0;
  ex.last_v = -INFINITY, ex.last_i = 0x7fffffff;
  ex.pre_v = -INFINITY, ex.pre_i = 0x7fffffff;
  if (rb < re)
    scan_band<C, TWO_STAGE, MASK2X, false, STAGED>(a, heat_hi, heat_lo, mask, raw_out, rb, re,
                                                   lane, nms, M, t3, ex, s_kth, warp, ring);

  // ---- merge: entries that reach the plane-wide bound are compacted ... -------------
  __syncthreads();  // every band has published its final kth
  {
    float tl = s_kth[0];
#pragma unroll
    for (int w = 1; w < kFastWarps; ++w) tl = fminf(tl, s_kth[w]);
    const float ev[3] = {t3.v1, t3.v2, t3.v3};
    const int ei[3] = {t3.i1, t3.i2, t3.i3};
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      if (ei[e] != 0x7fffffff && ev[e] >= tl) {
        const int pos = atomicAdd(&s_nbuf, 1);
        s_bv[pos] = ev[e];
        s_bi[pos] = ei[e];
      }
    }
  }
  __syncthreads();
  // ---- ... and ranked exactly (value desc, flat index asc) ----------------------------
  const int nb = s_nbuf;
  for (int t = tid; t < nb; t += kFastThreads) {
    const float v = s_bv[t];
    const int i = s_bi[t];
    int rank = 0;
    for (int j = 0; j < nb; ++j) rank += beats(s_bv[j], s_bi[j], v, i) ? 1 : 0;
    if (rank < M) {
      s_ov[rank] = v;
      s_oi[rank] = i;
    }
  }
  if (tid == 0) s_nout = min(nb, M);
  __syncthreads();
  // ---- check: could a lane have dropped a fourth entry that belongs to the top M? ----
  const int nout = s_nout;
  const bool has3 = t3.i3 != 0x7fffffff;
  // (a short union -- e.g. -inf pixels, which the strict compares never keep -- also
  // goes to the exact pass)
  const bool risk = nout < M || (has3 && beats(t3.v3, t3.i3, s_ov[M - 1], s_oi[M - 1]));
  const bool fallback = __syncthreads_or(risk ? 1 : 0) != 0;

  if (!fallback) {
    if (tid < nout) {
      const size_t o = ((size_t)n * a.K + k) * M + tid;
      const int ti = s_oi[tid];
      const int y = (int)fdiv((uint32_t)ti, a.div_w1);
      const int x = ti - y * W;
      a.val_k[o] = s_ov[tid];
      a.ind_k[2 * o] = (float)x;
      a.ind_k[2 * o + 1] = (float)y;
      a.tag_k[o] = bilinear_legacy(tag_src, th, tw, tsy, tsx, y, x);
    }
    return;
  }

  if (tid == 0) atomicAdd(&g_bu_exact_planes, 1ull);
  // ---- pass 2 (rare): exact per-warp lists, pre-filtered by the M-th result ---------
  if (nout == M) {
    ex.pre_v = s_ov[M - 1];
    ex.pre_i = s_oi[M - 1];
  }
  if (rb < re)
    scan_band<C, TWO_STAGE, MASK2X, true, STAGED>(a, heat_hi, heat_lo, mask, nullptr, rb, re,
                                                  lane, nms, M, t3, ex, s_kth, warp, ring);
  __syncthreads();  // everyone has read s_ov / s_oi
  s_lv[warp][lane] = ex.top_v;
  s_li[warp][lane] = ex.top_i;
  if (lane == 0) s_cnt[warp] = ex.count;
  __syncthreads();
  if (lane < ex.count) {
    int rank = lane;
    for (int w = 0; w < kFastWarps; ++w)
      if (w != warp) rank += count_beating(s_lv[w], s_li[w], s_cnt[w], ex.top_v, ex.top_i);
    if (rank < M) {
      const size_t o = ((size_t)n * a.K + k) * M + rank;
      const int y = (int)fdiv((uint32_t)ex.top_i, a.div_w1);
      const int x = ex.top_i - y * W;
      a.val_k[o] = ex.top_v;
      a.ind_k[2 * o] = (float)x;
      a.ind_k[2 * o + 1] = (float)y;
      a.tag_k[o] = bilinear_legacy(tag_src, th, tw, tsy, tsx, y, x);
    }
  }
}

// ---------------------------------------------------------------------------
// Pair scan (C = 8, two stages at exactly half size, 2x mask, 3x3 pool or none).
//
// Same decomposition as the fast kernel (one CTA per plane, one band of rows per warp,
// lane = 8 columns, per-lane top 3 + CTA merge), but pass 1 walks the band two output
// rows at a time, which is the natural unit of the 1/2-scale upsampling:
//   * the horizontally interpolated low-resolution row L[p] (8 values per lane) is
//     computed ONCE and used by output rows 2p-1, 2p and 2p+1 (the row-at-a-time scan
//     interpolated both source rows again for every output row);
//   * per pair a lane stages 5 x 16 B (rows 2p and 2p+1 of the high-resolution plane,
//     low-resolution row p+1) instead of 2 x 5 x 16 B;
//   * the mask is not read in the loop at all: mask_zero_rows_kernel reduces it to one
//     word per (image, output row) -- bit l set when lane l's 8 pixels contain a masked
//     one -- and only flagged rows fetch their mask bytes;
//   * the vertical pool is one 3-input max per pixel.
// A fourth register per lane (the largest value that was NOT kept in the lane's top 3)
// makes the "could a lane have dropped a member of the top M" check exact up to ties.
constexpr int kPairSeg = 512;                // 32 lanes x 16 B
constexpr int kPairSlot = 5 * kPairSeg;      // hi(2p) x2 | hi(2p+1) x2 | low(p+1)
static_assert(2 * kPairSlot >= kStageWarp, "the exact pass reuses the ring for its row staging");

// bit l of zrow[n * H + y]: some pixel x in [8l, 8l+8) of output row y is masked out.
// One warp per 4 rows (4 independent 16-byte loads in flight per lane).
__global__ void __launch_bounds__(256)
    mask_zero_rows_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ zrow, int H,
                          int W, int mh, int mw, float msy, int64_t rows) {
  // programmatic dependent launch: the decode kernel may start its prologue now; it waits
  // (griddepcontrol.wait) for this whole grid before it reads zrow
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4;
  uint4 mb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mb[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    const int64_t r = r0 + i;
    if (r < rows && lane * 8 < W) {
      const int64_t n = r / H;
      const int y = (int)(r - n * H);
      const int my = min((int)floorf(__fmul_rn((float)y, msy)), mh - 1);
      mb[i] = __ldg(reinterpret_cast<const uint4*>(mask + ((size_t)n * mh + my) * mw) + lane);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t w[4] = {mb[i].x, mb[i].y, mb[i].z, mb[i].w};
    bool z = false;
#pragma unroll
    for (int t = 0; t < 4; ++t) z |= (w[t] & 0xffu) == 0 || (w[t] & 0xff0000u) == 0;
    const unsigned b = __ballot_sync(0xffffffffu, z);
    if (lane == 0 && r0 + i < rows) zrow[r0 + i] = b;
  }
}

__device__ __forceinline__ void cp_async16_s(uint32_t saddr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem) : "memory");
}

// this lane's source pointers for the next pair to stage
struct PairSrc {
  const float* hi;  // row 2p, columns x0 .. x0+7
  const float* lo;  // low-resolution row min(p + 1, h0 - 1), columns x0/2 .. x0/2+3
  int p;
};

// (Round 2, measured and removed: staging a row as two warp-wide 512-byte halves instead of
// lane-private 2 x 16 bytes -- half the L2 sector requests, two extra __syncwarp per pair --
// ran 0.111 ms against 0.109 ms for this form on B200; profiles/README.md, r02a.)
template <bool ALL>
__device__ __forceinline__ void pair_issue(PairSrc& s, int W, int w0, int h0, int pe, int p_end,
                                           uint32_t slot, bool active) {
  if ((ALL || active) && s.p <= p_end) {
    cp_async16_s(slot, s.hi);
    cp_async16_s(slot + kPairSeg, s.hi + 4);
    if (s.p < pe) {
      cp_async16_s(slot + 2 * kPairSeg, s.hi + W);
      cp_async16_s(slot + 3 * kPairSeg, s.hi + W + 4);
      cp_async16_s(slot + 4 * kPairSeg, s.lo);
    }
  }
  cp_async_commit();  // one group per pair, empty past the end: wait_group 1 stays exact
  s.hi += 2 * W;
  if (s.p + 2 < h0) s.lo += w0;
  ++s.p;
}

// HALF of the horizontally interpolated low-resolution row.  The reference computes
// L = a + (b - a) * 0.5 on odd columns and the aggregate (hi + up) * 0.5; scaling by 0.5
// commutes with every rounding step (binary floating point, no underflow), so the loop
// works on hL = 0.5 * L and one fused multiply-add per term gives the same bits:
//   0.5 * (a + (b - a) * 0.5) == fma(0.5 b - 0.5 a, 0.5, 0.5 a),  (hi + L) * 0.5 == fma(hi, 0.5, hL).
// (Values within a factor 4 of the float32 underflow threshold are the only exception.)
__device__ __forceinline__ void lo_interp_half(const float4 q, bool last_lane, float (&L)[8]) {
  const float hx = __fmul_rn(q.x, 0.5f), hy = __fmul_rn(q.y, 0.5f);
  const float hz = __fmul_rn(q.z, 0.5f), hw = __fmul_rn(q.w, 0.5f);
  const float nx = __shfl_down_sync(0xffffffffu, hx, 1);
  const float h4 = last_lane ? hw : nx;
  L[0] = hx;
  L[1] = __fmaf_rn(__fsub_rn(hy, hx), 0.5f, hx);
  L[2] = hy;
  L[3] = __fmaf_rn(__fsub_rn(hz, hy), 0.5f, hy);
  L[4] = hz;
  L[5] = __fmaf_rn(__fsub_rn(hw, hz), 0.5f, hz);
  L[6] = hw;
  L[7] = __fmaf_rn(__fsub_rn(h4, hw), 0.5f, hw);
}

__device__ __forceinline__ void mask_row_slow(const BuArgs& a, const uint8_t* __restrict__ mask,
                                              int y, int x0, bool flagged, float (&v)[8]) {
  if (flagged) {
    const int my = min((int)floorf(__fmul_rn((float)y, a.msy)), a.mh - 1);
    const uint4 mb = __ldg(reinterpret_cast<const uint4*>(mask + (size_t)my * a.mw + 2 * x0));
    const uint32_t w[4] = {mb.x, mb.y, mb.z, mb.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if ((w[t] & 0xffu) == 0) v[2 * t] = 0.f;
      if ((w[t] & 0xff0000u) == 0) v[2 * t + 1] = 0.f;
    }
  }
}

// Warps (= row bands) per CTA of the pair kernel.  6 x 32 threads at 80 registers fit four
// CTAs per SM: 1088 planes are then 1.84 waves of CTAs instead of 2.45, which leaves the SMs
// less idle at the end of the grid.
constexpr int kPairWarps = 6;
constexpr int kPairThreads = kPairWarps * 32;

// Per-band candidate buffer in shared memory (one per warp, count in a register).
constexpr int kCandCap = 384;
struct CandBuf {
  float* v;
  int* i;
  int cnt;        // warp-uniform
  bool overflow;  // a row did not fit: the buffer is no longer complete
};

// Pass 1 of the pair kernel over rows [rb, re) (both even).
//
// Only pixels that reach t_lb -- a lower bound of the plane's M-th best value, shared by
// the CTA's bands -- can be in the result, and a pixel that reaches it is almost always a
// local maximum.  So the pool is evaluated lazily: per output row a lane only takes the
// maximum of its 8 aggregated values; the 3x3 max (vertical 3-max of the three rows, then
// the horizontal 3-max with the neighbour columns by shuffle) and the survivor test run
// only for rows in which some lane reaches t_lb (warp-uniform branch).  Survivors that
// reach t_lb are appended to the band's buffer by ballot compaction (no per-lane
// divergence); survivors below t_lb are never looked at, which is exact: they are strictly
// below the M-th value.  Suppressed pixels (value * 0) are not collected either; they can
// only belong to the top M when the M-th value is <= 0, and the merge sends exactly those
// planes to the exact pass -- as it does planes in which a band's buffer overflowed
// (plateaus: hundreds of survivors at or above the bound in one band).
//
// ALL: W == 256, every lane owns 8 columns (no per-lane activity predicates)
template <bool NMS, bool ALL, int kPairDepth>
__device__ __forceinline__ void scan_pairs(const BuArgs& a, const float* __restrict__ heat_hi,
                                           const float* __restrict__ heat_lo,
                                           const uint8_t* __restrict__ mask,
                                           const uint32_t* __restrict__ zrow, float* raw_out,
                                           int rb, int re, int lane, int M, CandBuf& cb,
                                           volatile float* s_kth, int warp_id,
                                           unsigned char* ring) {
  const int H = a.h1, W = ALL ? 256 : a.w1, h0 = a.h0, w0 = ALL ? 128 : a.w0;
  const int x0 = lane * 8;
  const bool active = ALL || x0 < W;
  const bool first_lane = lane == 0;
  const bool last_lane = ALL ? lane == 31 : x0 + 8 >= W;
  const int pb = rb >> 1, pe = re >> 1;
  const int p_end = re < H ? pe : pe - 1;
  const int kth_rank = (M + kPairWarps - 1) / kPairWarps;
  const unsigned lt_mask = (1u << lane) - 1u;
  constexpr int kPairWarp = kPairDepth * kPairSlot;
  unsigned char* slot0 = ring + lane * 16;
  const uint32_t sa0 = smem_u32(slot0);

  // Prologue loads first, all independent, so that their round trips overlap each other and
  // the first two staged pairs: mask flags of rows rb-1 .. re (bit j <-> row rb - 1 + j),
  // low-resolution rows pb-1 and pb, row rb-1 of the high-resolution plane (the band's
  // upper halo).
  const int zr0 = rb - 1, zlast = min(re, H - 1);
  const bool halo = NMS && rb > 0;
  uint32_t za = 0, zb = 0;
  float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
  float4 g0 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), g1 = g0;
  {
    if (active) {
      const float* lrow = heat_lo + (size_t)pb * w0 + (x0 >> 1);
      qb = __ldg(reinterpret_cast<const float4*>(lrow));
      if (halo) {
        qa = __ldg(reinterpret_cast<const float4*>(lrow - w0));
        const float* hrow = heat_hi + (size_t)(rb - 1) * W + x0;
        g0 = ld_stream_f4(hrow);
        g1 = ld_stream_f4(hrow + 4);
      }
    }
  }

  PairSrc src;
  src.hi = heat_hi + (size_t)rb * W + x0;
  src.lo = heat_lo + (size_t)min(pb + 1, h0 - 1) * w0 + (x0 >> 1);
  src.p = pb;
#pragma unroll
  for (int d = 0; d < kPairDepth; ++d)
    pair_issue<ALL>(src, W, w0, h0, pe, p_end, sa0 + d * kPairSlot, active);

  // zrow is written by mask_zero_rows_kernel, which this kernel may overlap (programmatic
  // dependent launch): wait for that grid, then read with plain (coherent) loads
  asm volatile("griddepcontrol.wait;" ::: "memory");
  {
    const int r = zr0 + lane;
    if (r >= 0 && r <= zlast) za = *(const volatile uint32_t*)(zrow + r);
    if (r + 32 <= zlast) zb = *(const volatile uint32_t*)(zrow + r + 32);
  }
  const uint64_t nz = (uint64_t)__ballot_sync(0xffffffffu, za != 0) |
                      ((uint64_t)__ballot_sync(0xffffffffu, zb != 0) << 32);
  auto masked = [&](int y, float (&v)[8]) {  // only called when row y is flagged (uniform)
    const int j = y - zr0;
    const uint32_t bits = __shfl_sync(0xffffffffu, j < 32 ? za : zb, j & 31);
    mask_row_slow(a, mask, y, x0, active && ((bits >> lane) & 1u), v);
  };

  // state carried from pair to pair: hL = half of the interpolated low-resolution row p,
  // vPP / vP = aggregated rows 2p-2 / 2p-1
  float hL[8], vPP[8], vP[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) vPP[c] = vP[c] = -INFINITY;
  lo_interp_half(qb, last_lane, hL);
  if (halo) {  // row rb-1 = 2(pb-1)+1: between low-resolution rows pb-1 and pb
    float hA[8];
    lo_interp_half(qa, last_lane, hA);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int c = 0; c < 8; ++c)
      vP[c] = __fmaf_rn(g[c], 0.5f, __fmaf_rn(__fsub_rn(hL[c], hA[c]), 0.5f, hA[c]));
    if (nz & 1ull) masked(rb - 1, vP);
  }

  // t_lb: each band publishes the kth largest of its lanes' best candidates (k = ceil(M/8),
  // distinct pixels); the minimum over the bands has >= M elements at or above it.  Stale
  // reads of other bands' slots are only smaller, i.e. still valid.
  float t_lb = -INFINITY;
  float best = -INFINITY;  // this lane's best candidate so far
  auto refresh_bound = [&]() {
    // positive values only: integer order == float order
    int key = best > 0.f ? __float_as_int(best) : 0;
    int mx = 0;
    for (int r = 0; r < kth_rank; ++r) {
      mx = __reduce_max_sync(0xffffffffu, key);
      const unsigned b = __ballot_sync(0xffffffffu, key == mx);
      if (lane == __ffs(b) - 1) key = 0;
    }
    if (lane == 0) s_kth[warp_id] = mx > 0 ? __int_as_float(mx) : -INFINITY;
    __syncwarp();
    // slots hold -inf or a positive float: signed integer order == float order
    int o = lane < kPairWarps ? __float_as_int(s_kth[lane]) : 0x7f800000;
    o = __reduce_min_sync(0xffffffffu, o);
    t_lb = fmaxf(t_lb, __int_as_float(o));
  };

  // row y = vb, with the rows above (va) and below (vc) it
  auto evaluate = [&](int y, const float (&va)[8], const float (&vb)[8], const float (&vc)[8]) {
    const float vm = fmaxf(fmaxf(fmaxf(vb[0], vb[1]), fmaxf(vb[2], vb[3])),
                           fmaxf(fmaxf(vb[4], vb[5]), fmaxf(vb[6], vb[7])));
    if (!__any_sync(0xffffffffu, active && vm >= t_lb)) return;
    float pl[8];
    if (NMS) {
      float cmx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) cmx[c] = fmaxf(fmaxf(va[c], vb[c]), vc[c]);
      hmax3<8>(cmx, first_lane, last_lane, pl);
    }
    // all eight ballots first (independent), then one branch for the common "no survivor" case
    bool pr[8];
    unsigned bal[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float bar = NMS ? fmaxf(pl[c], t_lb) : t_lb;
      pr[c] = active && vb[c] >= bar;
      bal[c] = __ballot_sync(0xffffffffu, pr[c]);
    }
    const unsigned any = (bal[0] | bal[1] | bal[2] | bal[3]) | (bal[4] | bal[5] | bal[6] | bal[7]);
    if (any == 0) return;
    if (cb.cnt > kCandCap - 256) {  // a row adds at most 256; heat maps collect ~150 per band
      int total = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) total += __popc(bal[c]);
      if (cb.cnt + total > kCandCap) {
        cb.overflow = true;  // -> the plane goes to the exact pass
        return;
      }
    }
    const int row_base = y * W + x0;
    int base = cb.cnt;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (bal[c]) {  // warp-uniform
        if (pr[c]) {
          const int pos = base + __popc(bal[c] & lt_mask);
          cb.v[pos] = vb[c];
          cb.i[pos] = row_base + c;
          best = fmaxf(best, vb[c]);
        }
        base += __popc(bal[c]);
      }
    }
    cb.cnt = base;
  };

  int tog = 0;  // byte offset of pair p's slot in the ring
  for (int p = pb; p < pe; ++p) {
    cp_async_wait_pending<kPairDepth - 1>();
    const unsigned char* slot = slot0 + tog;
    const float4 a0 = *reinterpret_cast<const float4*>(slot);
    const float4 a1 = *reinterpret_cast<const float4*>(slot + kPairSeg);
    const float4 b0 = *reinterpret_cast<const float4*>(slot + 2 * kPairSeg);
    const float4 b1 = *reinterpret_cast<const float4*>(slot + 3 * kPairSeg);
    const float4 q = *reinterpret_cast<const float4*>(slot + 4 * kPairSeg);
    // (inactive lanes read stale bytes: nothing they compute reaches an active lane)
    pair_issue<ALL>(src, W, w0, h0, pe, p_end, sa0 + tog, active);
    tog = tog + kPairSlot == kPairWarp ? 0 : tog + kPairSlot;
    float hN[8];
    lo_interp_half(q, last_lane, hN);
    float v0[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float v1[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v0[c] = __fmaf_rn(v0[c], 0.5f, hL[c]);
      const float hmid = __fmaf_rn(__fsub_rn(hN[c], hL[c]), 0.5f, hL[c]);
      v1[c] = __fmaf_rn(v1[c], 0.5f, hmid);
    }
    const int y = 2 * p;
    if ((nz >> (y - zr0)) & 3ull) {
      if ((nz >> (y - zr0)) & 1ull) masked(y, v0);
      if ((nz >> (y - zr0)) & 2ull) masked(y + 1, v1);
    }
    if (raw_out && active) {
      float* o = raw_out + (size_t)y * W + x0;
      st_stream_f4(o, make_float4(v0[0], v0[1], v0[2], v0[3]));
      st_stream_f4(o + 4, make_float4(v0[4], v0[5], v0[6], v0[7]));
      st_stream_f4(o + W, make_float4(v1[0], v1[1], v1[2], v1[3]));
      st_stream_f4(o + W + 4, make_float4(v1[4], v1[5], v1[6], v1[7]));
    }
    if (NMS) {
      if (p > pb) evaluate(y - 1, vPP, vP, v0);
      evaluate(y, vP, v0, v1);
    } else {
      evaluate(y, v0, v0, v0);
      evaluate(y + 1, v1, v1, v1);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) vPP[c] = v0[c], vP[c] = v1[c], hL[c] = hN[c];
    const int cnt = p - pb + 1;
    if (cnt <= 4 || (cnt <= 8 && (cnt & 1) == 0) || (cnt & 3) == 0) refresh_bound();
  }
  if (NMS) {  // the band's last row is still waiting for the row below it
    float vN[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) vN[c] = -INFINITY;
    if (re < H) {
      cp_async_wait_pending<kPairDepth - 1>();
      const unsigned char* slot = slot0 + tog;
      const float4 a0 = *reinterpret_cast<const float4*>(slot);
      const float4 a1 = *reinterpret_cast<const float4*>(slot + kPairSeg);
      float v0[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) vN[c] = __fmaf_rn(v0[c], 0.5f, hL[c]);
      if ((nz >> (re - zr0)) & 1ull) masked(re, vN);
    }
    evaluate(re - 1, vPP, vP, vN);
  }
  refresh_bound();  // final value of this band for the merge
  asm volatile("cp.async.wait_all;" ::: "memory");
}

constexpr int kPairBuf = kPairWarps * kCandCap;

template <bool NMS, bool ALL, int MINB>
__global__ void __launch_bounds__(kPairThreads, MINB)
    bottomup_decode_pairs_kernel(const BuArgs a, const uint32_t* __restrict__ zrow_all) {
  __shared__ float s_lv[kPairWarps][32];
  __shared__ int s_li[kPairWarps][32];
  __shared__ int s_cnt[kPairWarps];
  __shared__ float s_cv[kPairWarps][kCandCap];
  __shared__ int s_ci[kPairWarps][kCandCap];
  __shared__ float s_ov[32];
  __shared__ int s_oi[32];
  __shared__ int s_nbuf, s_nout;
  __shared__ float s_kth[kPairWarps];
  // pairs in flight per warp: 3 CTAs per SM only fit with a two-deep ring
  constexpr int kPairDepth = MINB >= 3 ? 2 : 3;
  constexpr int kPairWarp = kPairDepth * kPairSlot;
  extern __shared__ __align__(16) unsigned char s_ring[];  // kPairWarps * kPairWarp
  static_assert(kPairBuf * 8 <= kPairWarps * kPairWarp, "merge buffer must fit the ring");
  // merge buffer: reuses the staging ring once every band is done with it
  float* s_bv = reinterpret_cast<float*>(s_ring);
  int* s_bi = reinterpret_cast<int*>(s_ring) + kPairBuf;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x / a.K, k = blockIdx.x - n * a.K;
  const int H = a.h1, W = a.w1, M = a.M;
  unsigned char* ring = s_ring + warp * kPairWarp;

  const float* heat_lo = a.out0 + ((size_t)n * a.c0 + k) * a.h0 * a.w0;
  const float* tag_src = a.out0 + ((size_t)n * a.c0 + a.K + k * a.tag_step) * a.h0 * a.w0;
  const float* heat_hi = a.out1 + ((size_t)n * a.K + k) * H * W;
  const uint8_t* mask = a.mask + (size_t)n * a.mh * a.mw;
  const uint32_t* zrow = zrow_all + (size_t)n * H;
  float* raw_out = a.heatmap_raw ? a.heatmap_raw + ((size_t)n * a.K + k) * H * W : nullptr;

  // even number of rows per band
  const int R = (((H + kPairWarps - 1) / kPairWarps) + 1) & ~1;
  const int rb = min(H, warp * R), re = min(H, rb + R);

  if (tid == 0) s_nbuf = 0, s_nout = 0;
  if (tid < kPairWarps) s_kth[tid] = -INFINITY;
  __syncthreads();

  CandBuf cb;
  cb.v = s_cv[warp];
  cb.i = s_ci[warp];
  cb.cnt = 0;
  cb.overflow = false;
  if (rb < re)
    scan_pairs<NMS, ALL, kPairDepth>(a, heat_hi, heat_lo, mask, zrow, raw_out, rb, re, lane, M, cb,
                                     s_kth, warp, ring);

  __syncthreads();  // every band has published its final kth
  {
    float tl = s_kth[0];
#pragma unroll
    for (int w = 1; w < kPairWarps; ++w) tl = fminf(tl, s_kth[w]);
    for (int e = lane; e < cb.cnt; e += 32) {
      const float v = cb.v[e];
      if (v >= tl) {
        const int pos = atomicAdd(&s_nbuf, 1);
        s_bv[pos] = v;
        s_bi[pos] = cb.i[e];
      }
    }
  }
  __syncthreads();
  const int nb = s_nbuf;
  for (int t = tid; t < nb; t += kPairThreads) {
    const float v = s_bv[t];
    const int i = s_bi[t];
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < nb; ++j) rank += beats(s_bv[j], s_bi[j], v, i) ? 1 : 0;
    if (rank < M) {
      s_ov[rank] = v;
      s_oi[rank] = i;
    }
  }
  if (tid == 0) s_nout = min(nb, M);
  __syncthreads();
  // The buffers hold every survivor that reached the bound, so the result is exact unless
  // fewer than M were collected or the M-th value is <= 0: then suppressed pixels
  // (value * 0), which pass 1 never collects, could belong to the top M.
  const int nout = s_nout;
  const bool fallback =
      __syncthreads_or(cb.overflow ? 1 : 0) != 0 || nout < M || !(s_ov[M - 1] > 0.f);

  if (!fallback) {
    if (tid < nout) {
      const size_t o = ((size_t)n * a.K + k) * M + tid;
      const int ti = s_oi[tid];
      const int y = (int)fdiv((uint32_t)ti, a.div_w1);
      const int x = ti - y * W;
      a.val_k[o] = s_ov[tid];
      a.ind_k[2 * o] = (float)x;
      a.ind_k[2 * o + 1] = (float)y;
      a.tag_k[o] = bilinear_legacy(tag_src, a.h0, a.w0, a.sy, a.sx, y, x);
    }
    return;
  }

  if (tid == 0) atomicAdd(&g_bu_exact_planes, 1ull);
  // ---- pass 2 (rare): exact per-warp lists with the row-at-a-time scan ---------------
  ExactList ex;
  ex.top_v = -INFINITY, ex.top_i = 0x7fffffff, ex.count = 0;
  ex.last_v = -INFINITY, ex.last_i = 0x7fffffff;
  ex.pre_v = -INFINITY, ex.pre_i = 0x7fffffff;
  if (nout == M && s_ov[M - 1] > 0.f) {
    // every pass-1 candidate is a real survivor, so the M-th of them bounds the result
    // (also when a band overflowed: what it did collect is still real)
    ex.pre_v = s_ov[M - 1];
    ex.pre_i = s_oi[M - 1];
  }
  Top3 dummy;
  dummy.v1 = dummy.v2 = dummy.v3 = -INFINITY;
  dummy.i1 = dummy.i2 = dummy.i3 = 0x7fffffff;
  if (rb < re)
    scan_band<8, true, true, true, true>(a, heat_hi, heat_lo, mask, nullptr, rb, re, lane, NMS, M,
                                         dummy, ex, s_kth, warp, ring);
  __syncthreads();  // everyone has read s_ov / s_oi
  s_lv[warp][lane] = ex.top_v;
  s_li[warp][lane] = ex.top_i;
  if (lane == 0) s_cnt[warp] = ex.count;
  __syncthreads();
  if (lane < ex.count) {
    int rank = lane;
    for (int w = 0; w < kPairWarps; ++w)
      if (w != warp) rank += count_beating(s_lv[w], s_li[w], s_cnt[w], ex.top_v, ex.top_i);
    if (rank < M) {
      const size_t o = ((size_t)n * a.K + k) * M + rank;
      const int y = (int)fdiv((uint32_t)ex.top_i, a.div_w1);
      const int x = ex.top_i - y * W;
      a.val_k[o] = ex.top_v;
      a.ind_k[2 * o] = (float)x;
      a.ind_k[2 * o + 1] = (float)y;
      a.tag_k[o] = bilinear_legacy(tag_src, a.h0, a.w0, a.sy, a.sx, y, x);
    }
  }
}


// tagging_heatmap output (bottom_up_decoder.py:118-120): the tag planes resized to the
// output resolution; only _refine_missing and the visualiser read it.
__global__ void __launch_bounds__(256)
    resize_tags_kernel(const float* __restrict__ out0, float* __restrict__ tagging, int K,
                       int planes_per_image, int tag_first, int th, int tw, int H, int W,
                       float sy, float sx, FastDiv div_w, int64_t total_quads) {
  const int wq = W >> 2;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_quads;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = t / ((int64_t)H * wq);
    const int r = (int)(t - plane * (int64_t)H * wq);
    const int y = r / wq, x = (r - y * wq) << 2;
    const int64_t n = plane / K;
    const int k = (int)(plane - n * K);
    const float* src = out0 + ((size_t)n * planes_per_image + tag_first + k) * th * tw;
    float4 o;
    o.x = bilinear_legacy(src, th, tw, sy, sx, y, x);
    o.y = bilinear_legacy(src, th, tw, sy, sx, y, x + 1);
    o.z = bilinear_legacy(src, th, tw, sy, sx, y, x + 2);
    o.w = bilinear_legacy(src, th, tw, sy, sx, y, x + 3);
    st_stream_f4(tagging + ((size_t)plane * H + y) * W + x, o);
  }
}

// ---- A17: _shift_coordinate (bottom_up_decoder.py:180-203), quirk included -----------
// The reference takes sign(raw[y, x+1] - raw[y, x-1]) * 0.25 (and the same in y; zero on
// the border rows / columns) at the top-M positions with masked_select, i.e. in ROW-MAJOR
// order of the positions, and adds the resulting vector to the coordinates, which are in
// RANK order: entry t of the top M receives the offset of the t-th position in spatial
// order.  That pairing is what the reference computes, so it is what is reproduced.
// raw is the aggregated, masked, pre-NMS map; its values are recomputed here for the four
// neighbours of each position (same arithmetic as every decode kernel).
__device__ __forceinline__ float bu_raw_at(const BuArgs& a, const float* __restrict__ heat_hi,
                                           const float* __restrict__ heat_lo,
                                           const uint8_t* __restrict__ mask, int y, int x) {
  float v = __ldg(heat_hi + (size_t)y * a.w1 + x);
  if (a.stages == 2) {
    v = __fadd_rn(v, bilinear_legacy(heat_lo, a.h0, a.w0, a.sy, a.sx, y, x));
    v = __fmul_rn(v, 0.5f);
  }
  const int my = min((int)floorf(__fmul_rn((float)y, a.msy)), a.mh - 1);
  const int mx = min((int)floorf(__fmul_rn((float)x, a.msx)), a.mw - 1);
  if (mask[(size_t)my * a.mw + mx] == 0) v = 0.f;
  return v;
}

__device__ __forceinline__ float sign_of_diff(float hi, float lo) {
  const float d = __fsub_rn(hi, lo);
  return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
}

__global__ void __launch_bounds__(128) bottomup_shift_kernel(const BuArgs a, int planes) {
  __shared__ float s_ox[4][64], s_oy[4][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int plane = blockIdx.x * 4 + warp;
  if (plane >= planes) return;
  const int n = plane / a.K, k = plane - n * a.K;
  const int H = a.h1, W = a.w1, M = a.M;
  const float* heat_hi;
  const float* heat_lo = nullptr;
  if (a.stages == 2) {
    heat_lo = a.out0 + ((size_t)n * a.c0 + k) * a.h0 * a.w0;
    heat_hi = a.out1 + ((size_t)n * a.K + k) * H * W;
  } else {
    heat_hi = a.out0 + ((size_t)n * a.c0 + k) * H * W;
  }
  const uint8_t* mask = a.mask + (size_t)n * a.mh * a.mw;
  float* ind = a.ind_k + (size_t)plane * M * 2;
  // entries t = lane and lane + 32 (M <= 64)
  int flat[2] = {0x7fffffff, 0x7fffffff};
  float ox[2] = {0.f, 0.f}, oy[2] = {0.f, 0.f};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int t = lane + 32 * s;
    if (t < M) {
      const int x = (int)ind[2 * t], y = (int)ind[2 * t + 1];
      flat[s] = y * W + x;
      if (x >= 1 && x <= W - 2)
        ox[s] = sign_of_diff(bu_raw_at(a, heat_hi, heat_lo, mask, y, x + 1),
                             bu_raw_at(a, heat_hi, heat_lo, mask, y, x - 1));
      if (y >= 1 && y <= H - 2)
        oy[s] = sign_of_diff(bu_raw_at(a, heat_hi, heat_lo, mask, y + 1, x),
                             bu_raw_at(a, heat_hi, heat_lo, mask, y - 1, x));
    }
  }
  // spatial rank of each entry among the M positions (they are distinct)
  int rank[2] = {0, 0};
  for (int j = 0; j < M; ++j) {
    const int fj = j < 32 ? __shfl_sync(0xffffffffu, flat[0], j)
                          : __shfl_sync(0xffffffffu, flat[1], j - 32);
    rank[0] += fj < flat[0] ? 1 : 0;
    rank[1] += fj < flat[1] ? 1 : 0;
  }
#pragma unroll
  for (int s = 0; s < 2; ++s)
    if (lane + 32 * s < M) {
      s_ox[warp][rank[s]] = ox[s];
      s_oy[warp][rank[s]] = oy[s];
    }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int t = lane + 32 * s;
    if (t < M) {
      ind[2 * t] = __fadd_rn(ind[2 * t], __fmul_rn(s_ox[warp][t], 0.25f));
      ind[2 * t + 1] = __fadd_rn(ind[2 * t + 1], __fmul_rn(s_oy[warp][t], 0.25f));
    }
  }
}

}  // namespace pc

using namespace pc;

// Launch of the pair kernel as a programmatic dependent of mask_zero_rows_kernel (the
// previous launch on the stream): its CTAs may become resident and run their prologue
// while that grid drains; griddepcontrol.wait in the kernel orders the zrow reads.
template <typename Kernel>
static cudaError_t launch_pairs(Kernel kernel, unsigned grid, size_t dyn, cudaStream_t st,
                                const BuArgs& args, const uint32_t* zrow) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)dyn);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args, zrow);
}

extern "C" int pc_bottomup_decode_stats(int64_t* exact_pass_planes, int reset) {
  unsigned long long v = 0ull;
  PC_CUDA(cudaMemcpyFromSymbol(&v, g_bu_exact_planes, sizeof(v)));
  if (exact_pass_planes) *exact_pass_planes = (int64_t)v;
  if (reset) {
    v = 0ull;
    PC_CUDA(cudaMemcpyToSymbol(g_bu_exact_planes, &v, sizeof(v)));
  }
  return PC_OK;
}

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
  PC_REQUIRE(p->max_num >= 1 && p->max_num <= PC_MAX_DETECTIONS, PC_ERR_UNSUPPORTED,
             "pc_bottomup_decode: max_num %d outside [1, %d]", p->max_num, PC_MAX_DETECTIONS);
  PC_REQUIRE((int64_t)p->h1 * p->w1 >= p->max_num, PC_ERR_INVALID_ARGUMENT,
             "pc_bottomup_decode: map smaller than max_num");
  PC_REQUIRE(!p->use_nms || (p->nms_kernel >= 1 && p->nms_kernel <= kBuMaxNms),
             PC_ERR_UNSUPPORTED, "pc_bottomup_decode: nms_kernel %d outside [1, %d]",
             p->nms_kernel, kBuMaxNms);
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
  a.tag_step = p->tag_per_joint ? 1 : 0;
  a.c0 = p->tag_per_joint ? 2 * p->num_joints : p->num_joints + 1;
  a.div_w1 = make_fastdiv((uint32_t)p->w1);
  a.sy = p->num_stages == 2 ? (float)p->h0 / (float)p->h1 : 1.f;
  a.sx = p->num_stages == 2 ? (float)p->w0 / (float)p->w1 : 1.f;
  a.msy = (float)p->mask_h / (float)p->h1;
  a.msx = (float)p->mask_w / (float)p->w1;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)(n * p->num_joints);

  // ---- fast path selection ----------------------------------------------------
  const bool nms_ok = !p->use_nms || p->nms_kernel == 1 || p->nms_kernel == 3;
  const bool two = p->num_stages == 2;
  const bool half = !two || (p->h1 == 2 * p->h0 && p->w1 == 2 * p->w0);
  int C = 0;
  if (p->w1 % 4 == 0) {
    if (p->w1 <= 128 && p->w1 % 4 == 0) C = 4;
    if (p->w1 > 128 && p->w1 <= 256 && p->w1 % 8 == 0) C = 8;
    if (p->w1 > 256 && p->w1 <= 512 && p->w1 % 16 == 0) C = 16;
  }
  const bool aligned = ((uintptr_t)d_out0 % 16 == 0) && (!two || (uintptr_t)d_out1 % 16 == 0) &&
                       (!d_heatmap_raw || (uintptr_t)d_heatmap_raw % 16 == 0) &&
                       (!two || (p->w0 % 4 == 0));
  // (the fast kernels rank with one 32-lane ballot: max_num 33..64 takes the generic kernel)
  if (nms_ok && half && C != 0 && aligned && p->max_num <= 32) {
    const bool mask2x = p->mask_w == 2 * p->w1 && (uintptr_t)d_mask % 16 == 0;
    BuArgs b = a;
    if (p->use_nms && p->nms_kernel == 1) b.use_nms = 0;  // a 1x1 pool keeps every value
#define PC_BU_LAUNCH(CC, TWO, M2X)                                                        \
  do {                                                                                    \
    const size_t dyn = ((CC) == 8 && (TWO) && (M2X)) ? (size_t)kFastWarps * kStageWarp : 0; \
    if (dyn)                                                                              \
      PC_CUDA(cudaFuncSetAttribute(bottomup_decode_fast_kernel<CC, TWO, M2X>,             \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
    bottomup_decode_fast_kernel<CC, TWO, M2X><<<grid, kFastThreads, dyn, st>>>(b);        \
  } while (0)
#define PC_BU_PICK(CC)                                 \
  do {                                                 \
    if (two) {                                         \
      if (mask2x) PC_BU_LAUNCH(CC, true, true);        \
      else PC_BU_LAUNCH(CC, true, false);              \
    } else {                                           \
      if (mask2x) PC_BU_LAUNCH(CC, false, true);       \
      else PC_BU_LAUNCH(CC, false, false);             \
    }                                                  \
  } while (0)
    // (rows per band + 2 halo rows must fit the 64-bit masked-row flags)
    const bool pairs = C == 8 && two && mask2x && p->h1 % 2 == 0 && p->h1 <= 62 * kPairWarps;
    if (pairs) {
      // one word per (image, output row): which lanes see a masked pixel (stream-ordered scratch)
      uint32_t* zrow = nullptr;
      const int64_t rows = n * p->h1;
      cudaMemPool_t pool;
      PC_CUDA(scratch_pool(&pool));
      PC_CUDA(cudaMallocFromPoolAsync((void**)&zrow, sizeof(uint32_t) * (size_t)rows, pool, st));
      mask_zero_rows_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, st>>>(
          d_mask, zrow, p->h1, p->w1, p->mask_h, p->mask_w, a.msy, rows);
      const bool all = p->w1 == 256 && p->w0 == 128;
      // (errors are collected so that the scratch is returned to the pool on every path)
      cudaError_t le = cudaGetLastError();  // of the row-flag kernel
#define PC_BU_PAIRS(NMS_, ALL_)                                                              \
  le = launch_pairs(bottomup_decode_pairs_kernel<NMS_, ALL_, 4>, grid,                       \
                    (size_t)kPairWarps * 2 * kPairSlot, st, b, zrow)
      if (le == cudaSuccess) {
        if (b.use_nms) {
          if (all) PC_BU_PAIRS(true, true);
          else PC_BU_PAIRS(true, false);
        } else {
          if (all) PC_BU_PAIRS(false, true);
          else PC_BU_PAIRS(false, false);
        }
      }
#undef PC_BU_PAIRS
      if (le == cudaSuccess) le = cudaGetLastError();
      const cudaError_t fe = cudaFreeAsync(zrow, st);
      PC_CUDA(le);
      PC_CUDA(fe);
    } else if (C == 4) PC_BU_PICK(4);
    else if (C == 8) PC_BU_PICK(8);
    else PC_BU_PICK(16);
#undef PC_BU_PICK
#undef PC_BU_LAUNCH
    PC_CUDA(cudaGetLastError());
    if (d_tagging) {
      const int th = two ? p->h0 : p->h1, tw = two ? p->w0 : p->w1;
      const int tag_planes = a.tag_step ? p->num_joints : 1;
      const int64_t quads = n * tag_planes * (int64_t)p->h1 * (p->w1 / 4);
      int64_t blocks = (quads + 255) / 256;
      const int64_t cap = (int64_t)sm_count_cached() * 16;
      if (blocks > cap) blocks = cap;
      resize_tags_kernel<<<(unsigned)blocks, 256, 0, st>>>(
          d_out0, d_tagging, tag_planes, a.c0, p->num_joints, th, tw, p->h1,
          p->w1, two ? a.sy : 1.f, two ? a.sx : 1.f, a.div_w1, quads);
      PC_CUDA(cudaGetLastError());
    }
    if (p->shift_coordinate) {
      bottomup_shift_kernel<<<(grid + 3) / 4, 128, 0, st>>>(a, (int)grid);
      PC_CUDA(cudaGetLastError());
    }
    return PC_OK;
  }

  const size_t smem = sizeof(float) * ((size_t)(kBuRing + 2 * kBuTileRows) * p->w1 + kBuThreads);
  if (smem > 48 * 1024)
    PC_CUDA(cudaFuncSetAttribute(bottomup_decode_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bottomup_decode_kernel<<<grid, kBuThreads, smem, st>>>(a);
  PC_CUDA(cudaGetLastError());
  if (p->shift_coordinate) {
    bottomup_shift_kernel<<<(grid + 3) / 4, 128, 0, st>>>(a, (int)grid);
    PC_CUDA(cudaGetLastError());
  }
  return PC_OK;
}
