// topdown_decode.cu -- fused top-down heatmap decode for sm_100a.
//
// Replaces TopDownHeatMapDecoder.construct
// (mindpose/models/decoders/top_down_decoder.py:72-215) and the post-network
// half of the flip test (_MultiRunNet.construct,
// mindpose/engine/inferencer/topdown_inferencer.py:165-187).
//
// One persistent CTA per SM.  A work item is one joint map (n, k): its H*W
// float32 plane (and, for the flip test, the plane flipped[n, flip_index[k]])
// is contiguous in HBM, so a single elected producer thread moves it into a
// shared-memory stage with one or two 1-D TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx).  The eight consumer warps form groups of G warps; a
// group owns one item at a time and its members scan 1/G of the plane each
// (flip-average on the fly, float4-group argmax, value desc / flat index asc).
// G = 1 for small planes (many stages fit, every warp works on its own item);
// when a plane pair is 55 KB and only four stages fit, G = 4 keeps the time a
// stage is held short, so the other stages can be in flight.  The partial
// results meet in shared memory and ONE warp of the group (round robin)
// finishes the item: sub-pixel refinement (quarter offset or DARK/UDP Taylor
// step) out of the staged planes, back-projection, 12-byte result.  Every
// heatmap byte crosses HBM once; only the results are written.
#include <math.h>

#include "common.cuh"

namespace pc {

constexpr int kConsumerWarps = 16;
constexpr int kDecodeThreads = (kConsumerWarps + 1) * 32;  // + the producer warp
constexpr int kMaxStages = 12;
constexpr int kDarkSamples = 7;

constexpr int kGatherMaxPeers = 16;

struct DecodeArgs {
  const float* heatmap;
  const float* flipped;
  const float* center;
  const float* scale;
  const float* score;
  float* all_preds;
  float* all_boxes;
  int64_t num_items;  // n * K
  int32_t K, H, W, HW;
  FastDiv divW, divK;
  FastDiv divWq;           // W / 4 (float4 groups per row)
  FastDiv divKs, divWin;   // DARK kernel size, window size ks + 2
  int32_t step_xq, step_y;  // 32 float4 groups ahead = step_y rows + step_xq groups
  float pixel_std;
  int32_t to_original, use_udp;
  int32_t mode;  // 0 none, 1 quarter offset, 2 DARK/UDP
  int32_t ks;    // DARK kernel size
  int32_t shift_heatmap;
  int32_t stages;
  int32_t group;  // consumer warps per item (1, 2, 4, 8 or 16)
  int32_t win_floats, row_floats;  // per-warp DARK scratch (0 unless mode == 2)
  uint32_t stage_floats;  // floats per stage (HW or 2*HW)
  int32_t vec_ok;         // W % 4 == 0
  int32_t bulk_ok;        // planes 16-byte aligned in HBM (H*W % 4 == 0, bases % 16 == 0)
  // ---- fused all-gather of the results (pc_topdown_decode_gather; g_on = 0 otherwise) ----
  // Every result is ALSO stored into row g_row0 + crop of the gathered table [total, 3K + 6]
  // of every rank: through the NVSwitch multicast mapping (g_mc) or the peer-mapped tables.
  int32_t g_on, g_peers, g_nflags, g_rank;
  int64_t g_row0;
  float* g_mc;
  float* g_peer[kGatherMaxPeers];
  uint32_t* g_flags[kGatherMaxPeers];  // flag array (one word per source rank) on each rank
  uint32_t* g_step;                    // this rank's step number (device memory)
  int32_t* g_counter;                  // CTAs that have finished (zero between launches)
};

struct DecodeTables {
  int32_t flip_index[PC_MAX_JOINTS];
  float dark_kernel[PC_MAX_DARK_KERNEL * PC_MAX_DARK_KERNEL];
};

// Value of the (flip-averaged) map at (y, x); planes are in shared memory.
template <bool FLIP>
__device__ __forceinline__ float map_at(const float* hm, const float* fm, int W, int y, int x,
                                        int shift) {
  float v = hm[y * W + x];
  if (FLIP) {
    int xs = shift ? (x == 0 ? W - 1 : W - x) : (W - 1 - x);
    v = __fmul_rn(__fadd_rn(v, fm[y * W + xs]), 0.5f);
  }
  return v;
}

__device__ __forceinline__ void take_max(float v, int i, float& bv, int& bi) {
  if (v > bv) {
    bv = v;
    bi = i;
  }
}

// Four consecutive elements (row y, columns x0 .. x0+3, x0 % 4 == 0) of
//   FLIP ? heatmap + flip_back(flipped)   (the SUM: the * 0.5 is applied by the caller)
//        : heatmap
// read as 128-bit shared-memory loads.  The flipped plane is stored as it came from
// HBM; flip_back reverses x (fh[x] = fm[W-1-x]) and the optional 1-px shift makes
// fh[x] = fm[W-x] for x >= 1, fh[0] = fm[W-1] (topdown_inferencer.py:180-187).
template <bool FLIP>
__device__ __forceinline__ float4 quad_at(const float* hm, const float* fm, int W, int q, int y,
                                          int xq, int wq, int shift) {
  float4 v = reinterpret_cast<const float4*>(hm)[q];
  if (FLIP) {
    const int fq = q + wq - 1 - 2 * xq;  // same row, mirrored float4 column
    const float4 f = reinterpret_cast<const float4*>(fm)[fq];
    float f0 = f.w, f1 = f.z, f2 = f.y, f3 = f.x;
    if (shift) {
      f3 = f.y;
      f2 = f.z;
      f1 = f.w;
      f0 = xq == 0 ? f.w : fm[(fq << 2) + 4];
    }
    v.x = __fadd_rn(v.x, f0);
    v.y = __fadd_rn(v.y, f1);
    v.z = __fadd_rn(v.z, f2);
    v.w = __fadd_rn(v.w, f3);
  }
  return v;
}

// One row of the blur: sum over kx of kernel[ky][kx] * win[row][col0 + kx], the
// products added left to right in float32 (the oracle's order).
template <int KS>
__device__ __forceinline__ float blur_row(const float* __restrict__ wrow,
                                          const float* __restrict__ vrow, int ks) {
  float acc = __fmul_rn(wrow[0], vrow[0]);
  if (KS > 0) {
#pragma unroll
    for (int kx = 1; kx < KS; ++kx) acc = __fadd_rn(acc, __fmul_rn(wrow[kx], vrow[kx]));
  } else {
    for (int kx = 1; kx < ks; ++kx) acc = __fadd_rn(acc, __fmul_rn(wrow[kx], vrow[kx]));
  }
  return acc;
}

// DARK / UDP Taylor step (top_down_decoder.py:171-205) around the peak (px, py): returns
// inv(H + 1e-7 I) * d, valid on lane 0.
// (1) the (ks+2)^2 window of the averaged map around the peak, zero outside the map (the
//     blur's "same" padding), goes to a per-warp scratch once;
// (2) 7 samples x ks kernel rows = 7*ks independent row sums out of the scratch (each a
//     left-to-right float32 chain -- up to three chains per lane run interleaved);
// (3) lanes 0..6 add their rows top to bottom, clip, log; out-of-map samples are the zero
//     padding of the LOG map.
// samples: 0:i  1:ix1  2:iy1  3:ix1y1  4:ix1_y1_  5:ix1_  6:iy1_
// KS = 11 is the sigma = 2 recipe with everything unrolled; KS = 0 is any odd size.
template <bool FLIP, int KS>
__device__ __forceinline__ float2 dark_offset(const float* hm, const float* fm, int W, int H,
                                              int px, int py, int shift, const DecodeArgs& a,
                                              const float* s_kernel, float* my_rows,
                                              float* my_win, int lane) {
  const int ks = KS > 0 ? KS : a.ks;
  const int r = (ks - 1) >> 1, wsz = ks + 2;
  const int oy = py - r - 1, ox = px - r - 1;
  const int nwin = wsz * wsz;
  constexpr int kFillRounds = KS > 0 ? ((KS + 2) * (KS + 2) + 31) / 32 : 1;
  if constexpr (KS > 0) {
#pragma unroll
    for (int i = 0; i < kFillRounds; ++i) {
      const int e = lane + 32 * i;
      if (e < nwin) {
        const int wy = e / (KS + 2), wx = e - wy * (KS + 2);
        const int yy = oy + wy, xx = ox + wx;
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = map_at<FLIP>(hm, fm, W, yy, xx, shift);
        my_win[e] = v;
      }
    }
  } else {
    for (int e = lane; e < nwin; e += 32) {
      const int wy = (int)fdiv((uint32_t)e, a.divWin);
      const int wx = e - wy * wsz;
      const int yy = oy + wy, xx = ox + wx;
      float v = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = map_at<FLIP>(hm, fm, W, yy, xx, shift);
      my_win[e] = v;
    }
  }
  __syncwarp();
  const int ntask = kDarkSamples * ks;
  if constexpr (KS > 0) {
    constexpr int kChains = (kDarkSamples * KS + 31) / 32;  // 3 for KS = 11
    const float* vr[kChains];
    const float* wr[kChains];
    float acc[kChains];
    int slot[kChains];
#pragma unroll
    for (int u = 0; u < kChains; ++u) {
      const int t = lane + 32 * u;
      const int tt = t < ntask ? t : 0;
      const int smp = tt / KS, ky = tt - smp * KS;
      const int dxs = (smp == 1 || smp == 3) ? 1 : ((smp == 4 || smp == 5) ? -1 : 0);
      const int dys = (smp == 2 || smp == 3) ? 1 : ((smp == 4 || smp == 6) ? -1 : 0);
      vr[u] = my_win + (1 + dys + ky) * (KS + 2) + (1 + dxs);
      wr[u] = s_kernel + ky * KS;
      slot[u] = t < ntask ? smp * PC_MAX_DARK_KERNEL + ky : -1;
      acc[u] = __fmul_rn(wr[u][0], vr[u][0]);
    }
#pragma unroll
    for (int kx = 1; kx < KS; ++kx)
#pragma unroll
      for (int u = 0; u < kChains; ++u)
        acc[u] = __fadd_rn(acc[u], __fmul_rn(wr[u][kx], vr[u][kx]));
#pragma unroll
    for (int u = 0; u < kChains; ++u)
      if (slot[u] >= 0) my_rows[slot[u]] = acc[u];
  } else {
    for (int t = lane; t < ntask; t += 32) {
      const int smp = (int)fdiv((uint32_t)t, a.divKs), ky = t - smp * ks;
      const int dxs = (smp == 1 || smp == 3) ? 1 : ((smp == 4 || smp == 5) ? -1 : 0);
      const int dys = (smp == 2 || smp == 3) ? 1 : ((smp == 4 || smp == 6) ? -1 : 0);
      my_rows[smp * PC_MAX_DARK_KERNEL + ky] =
          blur_row<0>(s_kernel + ky * ks, my_win + (1 + dys + ky) * wsz + (1 + dxs), ks);
    }
  }
  __syncwarp();
  float lg = 0.f;
  if (lane < kDarkSamples) {
    const int smp = lane;
    const int dxs = (smp == 1 || smp == 3) ? 1 : ((smp == 4 || smp == 5) ? -1 : 0);
    const int dys = (smp == 2 || smp == 3) ? 1 : ((smp == 4 || smp == 6) ? -1 : 0);
    const int sy = py + dys, sx = px + dxs;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
      const float* rows = my_rows + smp * PC_MAX_DARK_KERNEL;
      float tot = rows[0];
      if (KS > 0) {
#pragma unroll
        for (int ky = 1; ky < KS; ++ky) tot = __fadd_rn(tot, rows[ky]);
      } else {
        for (int ky = 1; ky < ks; ++ky) tot = __fadd_rn(tot, rows[ky]);
      }
      tot = fminf(fmaxf(tot, 0.001f), 50.f);
      // correctly rounded float32 log (fp64 log rounded once): logf's last-ulp differences
      // become a full float32 ulp (> 1e-4 px) of image coordinates beyond 1024
      lg = (float)log((double)tot);
    }  // else: the reference zero-pads the LOG map -> 0.0
  }
  __syncwarp();  // my_rows / my_win are reused by this warp's next item
  const float i_ = __shfl_sync(0xffffffffu, lg, 0);
  const float ix1 = __shfl_sync(0xffffffffu, lg, 1);
  const float iy1 = __shfl_sync(0xffffffffu, lg, 2);
  const float ix1y1 = __shfl_sync(0xffffffffu, lg, 3);
  const float ix1_y1_ = __shfl_sync(0xffffffffu, lg, 4);
  const float ix1_ = __shfl_sync(0xffffffffu, lg, 5);
  const float iy1_ = __shfl_sync(0xffffffffu, lg, 6);
  const float dx = __fmul_rn(0.5f, __fsub_rn(ix1, ix1_));
  const float dy = __fmul_rn(0.5f, __fsub_rn(iy1, iy1_));
  const float two_i = __fmul_rn(2.f, i_);
  const float dxx = __fadd_rn(__fsub_rn(ix1, two_i), ix1_);
  const float dyy = __fadd_rn(__fsub_rn(iy1, two_i), iy1_);
  float t = __fsub_rn(ix1y1, ix1);
  t = __fsub_rn(t, iy1);
  t = __fadd_rn(t, i_);
  t = __fadd_rn(t, i_);
  t = __fsub_rn(t, ix1_);
  t = __fsub_rn(t, iy1_);
  t = __fadd_rn(t, ix1_y1_);
  const float dxy = __fmul_rn(0.5f, t);
  const float ha = __fadd_rn(dxx, 1e-7f), hd = __fadd_rn(dyy, 1e-7f), hb = dxy;
  const float det = __fsub_rn(__fmul_rn(ha, hd), __fmul_rn(hb, hb));
  const float i00 = __fdiv_rn(hd, det);
  const float i01 = __fdiv_rn(-hb, det);
  const float i11 = __fdiv_rn(ha, det);
  float2 o;
  o.x = __fadd_rn(__fmul_rn(i00, dx), __fmul_rn(i01, dy));
  o.y = __fadd_rn(__fmul_rn(i01, dx), __fmul_rn(i11, dy));
  return o;
}

// one float into element i of the gathered table of every rank
__device__ __forceinline__ void gather_store(const DecodeArgs& a, int64_t i, float v) {
  if (a.g_mc) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(a.g_mc + i), "f"(v)
                 : "memory");
  } else {
    for (int p = 0; p < a.g_peers; ++p) a.g_peer[p][i] = v;
  }
}

template <bool FLIP>
__global__ void __launch_bounds__(kDecodeThreads, 1)
    topdown_decode_kernel(const DecodeArgs a, const __grid_constant__ DecodeTables tab) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  // (rounded up: with H * W % 4 != 0 the stages do not end on a 16-byte boundary)
  const size_t stage_bytes_total =
      ((size_t)a.stages * a.stage_floats * sizeof(float) + 15) & ~(size_t)15;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + stage_bytes_total);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* part_bar = empty_bar + kMaxStages;
  volatile uint32_t* s_done = reinterpret_cast<volatile uint32_t*>(part_bar + kMaxStages);
  float* s_part_v = reinterpret_cast<float*>(const_cast<uint32_t*>(s_done) + kMaxStages);  // [stages][warps]
  int* s_part_i = reinterpret_cast<int*>(s_part_v + kMaxStages * kConsumerWarps);
  float* s_kernel = reinterpret_cast<float*>(s_part_i + kMaxStages * kConsumerWarps);
  float* s_rows = s_kernel + PC_MAX_DARK_KERNEL * PC_MAX_DARK_KERNEL;  // [groups][7][17]
  float* s_win = s_rows + (kConsumerWarps / a.group) * a.row_floats;  // [groups][(ks+2)^2]

  __shared__ int s_warps_done_v;
  int* const s_warps_done = &s_warps_done_v;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = a.stages;

  if (threadIdx.x == 0) {
    s_warps_done_v = 0;
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&part_bar[s], a.group);
      s_done[s] = 0;
    }
    fence_mbar_init();
  }
  if (a.mode == 2) {
    for (int i = threadIdx.x; i < a.ks * a.ks; i += blockDim.x) s_kernel[i] = tab.dark_kernel[i];
  }
  __syncthreads();

  // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int64_t first = blockIdx.x;
  const int64_t count =
      first < a.num_items ? (a.num_items - first + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t plane_bytes = (uint32_t)a.HW * sizeof(float);

  if (warp == 0) {
    // ---------------- producer: one thread issues all bulk copies -----------
    // In order, so a parity wait on empty_bar is never more than one phase ahead.
    if (!a.bulk_ok) {
      // Planes that are not 16-byte aligned in HBM (H * W % 4 != 0, or an offset base: the
      // reference takes any H, W -- top_down_decoder.py:96-116) cannot be bulk-copied.  The
      // producer WARP then copies them itself, element by element, into the same stages and
      // completes the same barrier by a plain arrival; the consumers do not change.  Slower
      // (one warp of loads per SM), exact.
      int s = 0;
      uint32_t round = 0;
      for (int64_t j = 0; j < count; ++j) {
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        const int64_t item = first + j * gridDim.x;
        const int64_t n = item / a.K;
        const int k = (int)(item - n * a.K);
        float* dst = stage_base + (size_t)s * a.stage_floats;
        const float* src0 = a.heatmap + item * a.HW;
        for (int i = lane; i < a.HW; i += 32) dst[i] = __ldg(src0 + i);
        if (FLIP) {
          const float* src1 = a.flipped + (n * a.K + tab.flip_index[k]) * a.HW;
          for (int i = lane; i < a.HW; i += 32) dst[a.HW + i] = __ldg(src1 + i);
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        if (++s == S) {
          s = 0;
          ++round;
        }
      }
      return;
    }
    if (lane == 0) {
      const uint64_t pol = l2_evict_first_policy();
      int s = 0;
      uint32_t round = 0;
      for (int64_t j = 0; j < count; ++j) {
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        const int64_t item = first + j * gridDim.x;
        const int64_t n = item / a.K;
        const int k = (int)(item - n * a.K);
        float* dst = stage_base + (size_t)s * a.stage_floats;
        mbar_arrive_expect_tx(&full_bar[s], FLIP ? 2 * plane_bytes : plane_bytes);
        bulk_g2s(dst, a.heatmap + item * a.HW, plane_bytes, &full_bar[s], pol);
        if (FLIP) {
          const int kf = tab.flip_index[k];
          bulk_g2s(dst + a.HW, a.flipped + (n * a.K + kf) * a.HW, plane_bytes, &full_bar[s],
                   pol);
        }
        if (++s == S) {
          s = 0;
          ++round;
        }
      }
    }
    return;
  }

  // ------------------------- consumers ---------------------------------------
  const int cw = warp - 1;
  const int W = a.W, H = a.H, HW = a.HW;
  const int shift = a.shift_heatmap;


  // group g of G warps owns items g, g + NG, ...; member r scans units [q_lo, q_hi)
  const int G = a.group, NG = kConsumerWarps / G;
  const int grp = cw / G, mem = cw - grp * G;
  // DARK scratch of the group (one finisher per group at a time)
  float* my_rows = s_rows + grp * a.row_floats;
  float* my_win = s_win + grp * a.win_floats;
  const int units = a.vec_ok ? (HW >> 2) : HW;
  const int per_warp = ((units + G * 32 - 1) / (G * 32)) * 32;
  const int q_lo = min(units, mem * per_warp), q_hi = min(units, q_lo + per_warp);

  for (int64_t j = grp; j < count; j += NG) {
    const int s = (int)(j % S);
    const uint32_t round = (uint32_t)(j / S);
    // Groups and stages are decoupled and bulk copies complete out of order, so a group
    // can get here while an EARLIER round of the same stage has not even landed or is
    // still being read by another group.  A parity wait is only meaningful one phase
    // ahead, so the rounds of a stage are first put in order with a plain counter
    // (s_done[s] = rounds of stage s finished so far); once round-1 is finished,
    // full_bar[s] is in phase `round` or has just completed it and the parity wait is
    // exact.
    while (s_done[s] != round) __nanosleep(32);
    mbar_wait(&full_bar[s], round & 1);
    const float* hm = stage_base + (size_t)s * a.stage_floats;
    const float* fm = hm + HW;

    // ---- pass 1: (flip-averaged) argmax, lowest flat index among equals ----
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (a.vec_ok) {
      // Each lane keeps the maximum of its float4 groups and the FIRST group that
      // reached it; the element inside the group is resolved after the loop.  With the
      // flip test the comparison runs on the sums h + fh: halving is exact (and so
      // order preserving) unless the result is subnormal, which is redone below.
      const int wq = W >> 2;
      int y = (int)fdiv((uint32_t)(q_lo + lane), a.divWq);
      int xq = q_lo + lane - y * wq;
      int bq = -1;
#pragma unroll 4
      for (int q = q_lo + lane; q < q_hi; q += 32) {
        const float4 v = quad_at<FLIP>(hm, fm, W, q, y, xq, wq, shift);
        const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (m > bv) {
          bv = m;
          bq = q;
        }
        xq += a.step_xq;
        y += a.step_y;
        if (xq >= wq) {
          xq -= wq;
          ++y;
        }
      }
      if (bq >= 0) {
        const int yb = (int)fdiv((uint32_t)bq, a.divWq);
        const float4 v = quad_at<FLIP>(hm, fm, W, bq, yb, bq - yb * wq, wq, shift);
        const int c = v.x == bv ? 0 : (v.y == bv ? 1 : (v.z == bv ? 2 : 3));
        bi = (bq << 2) + c;
      }
    } else {
      for (int idx = q_lo + lane; idx < q_hi; idx += 32) {
        const int y = (int)fdiv((uint32_t)idx, a.divW);
        const int x = idx - y * W;
        float v = hm[idx];
        if (FLIP) {
          const int xs = shift ? (x == 0 ? W - 1 : W - x) : (W - 1 - x);
          v = __fadd_rn(v, fm[y * W + xs]);
        }
        take_max(v, idx, bv, bi);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    // this warp is done with the planes (the reduction consumed every lane's loads):
    // publish the partial result; the item's finisher collects all of them
    if (lane == 0) {
      s_part_v[s * kConsumerWarps + mem] = bv;
      s_part_i[s * kConsumerWarps + mem] = bi;
      mbar_arrive(&part_bar[s]);
    }
    if (mem != (int)((j / NG) % G)) continue;

    // ---------------- finisher of item j ------------------------------------------
    const int64_t item = first + j * gridDim.x;
    const int64_t n = item / a.K;
    const int k = (int)(item - n * a.K);
    float cx = 0.f, cy = 0.f, sw = 0.f, sh = 0.f, sc = 0.f;
    if (lane == 0) {
      cx = __ldg(a.center + 2 * n);
      cy = __ldg(a.center + 2 * n + 1);
      sw = __ldg(a.scale + 2 * n);
      sh = __ldg(a.scale + 2 * n + 1);
      if (k == 0) sc = __ldg(a.score + n);
    }
    mbar_wait(&part_bar[s], round & 1);
    bv = lane < G ? s_part_v[s * kConsumerWarps + lane] : -INFINITY;
    bi = lane < G ? s_part_i[s * kConsumerWarps + lane] : 0x7fffffff;
#pragma unroll
    for (int o = kConsumerWarps / 2; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    bv = __shfl_sync(0xffffffffu, bv, 0);
    bi = __shfl_sync(0xffffffffu, bi, 0);
    if (FLIP) {
      if (bi != 0x7fffffff && fabsf(bv) < 4.7019774e-38f /* 2^-124 */) {
        // halves of sums this small can round: redo the scan on the exact averages
        bv = -INFINITY;
        bi = 0x7fffffff;
        for (int idx = lane; idx < HW; idx += 32) {
          const int y = (int)fdiv((uint32_t)idx, a.divW);
          take_max(map_at<true>(hm, fm, W, y, idx - y * W, shift), idx, bv, bi);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
          }
        }
      } else {
        bv = __fmul_rn(bv, 0.5f);
      }
    }
    if (bi == 0x7fffffff) {  // no element compared greater than -inf: index 0, as numpy.argmax
      bi = 0;
      bv = map_at<FLIP>(hm, fm, W, 0, 0, shift);
    }
    const int py = (int)fdiv((uint32_t)bi, a.divW);
    const int px = bi - py * W;

    float rx = (float)px, ry = (float)py;

    if (a.mode == 1) {
      // ---- quarter offset (top_down_decoder.py:118-141): interior only ------
      if (lane == 0) {
        if (px >= 1 && px <= W - 2) {
          const float d = __fsub_rn(map_at<FLIP>(hm, fm, W, py, px + 1, shift),
                                    map_at<FLIP>(hm, fm, W, py, px - 1, shift));
          rx = __fadd_rn(rx, d > 0.f ? 0.25f : (d < 0.f ? -0.25f : 0.f));
        }
        if (py >= 1 && py <= H - 2) {
          const float d = __fsub_rn(map_at<FLIP>(hm, fm, W, py + 1, px, shift),
                                    map_at<FLIP>(hm, fm, W, py - 1, px, shift));
          ry = __fadd_rn(ry, d > 0.f ? 0.25f : (d < 0.f ? -0.25f : 0.f));
        }
      }
    } else if (a.mode == 2) {
      // ---- DARK / UDP Taylor step (top_down_decoder.py:171-205) -------------
      float2 off;
      if (a.ks == 11)
        off = dark_offset<FLIP, 11>(hm, fm, W, H, px, py, shift, a, s_kernel, my_rows, my_win,
                                    lane);
      else
        off = dark_offset<FLIP, 0>(hm, fm, W, H, px, py, shift, a, s_kernel, my_rows, my_win,
                                   lane);
      rx = __fsub_rn(rx, off.x);
      ry = __fsub_rn(ry, off.y);
    }

    // the whole group has scanned the stage and the refinement is done: let the next
    // round's group in, hand the stage back to the producer
    __syncwarp();
    if (lane == 0) {
      s_done[s] = round + 1;
      mbar_arrive(&empty_bar[s]);
    }

    if (lane == 0) {
      const float s_w = __fmul_rn(sw, a.pixel_std);
      const float s_h = __fmul_rn(sh, a.pixel_std);
      if (a.to_original) {
        const float den_x = a.use_udp ? (float)(W - 1) : (float)W;
        const float den_y = a.use_udp ? (float)(H - 1) : (float)H;
        const float kx = __fdiv_rn(s_w, den_x);
        const float ky = __fdiv_rn(s_h, den_y);
        rx = __fsub_rn(__fadd_rn(__fmul_rn(rx, kx), cx), __fmul_rn(s_w, 0.5f));
        ry = __fsub_rn(__fadd_rn(__fmul_rn(ry, ky), cy), __fmul_rn(s_h, 0.5f));
      }
      float* o = a.all_preds + item * 3;
      o[0] = rx;
      o[1] = ry;
      o[2] = bv;
      const float area = __fmul_rn(s_w, s_h);
      if (k == 0) {
        float* b = a.all_boxes + n * 6;
        b[0] = cx;
        b[1] = cy;
        b[2] = sw;
        b[3] = sh;
        b[4] = area;
        b[5] = sc;
      }
      if (a.g_on) {  // the same values straight into every rank's gathered table
        const int64_t row = (a.g_row0 + n) * (int64_t)(a.K * 3 + 6);
        gather_store(a, row + k * 3, rx);
        gather_store(a, row + k * 3 + 1, ry);
        gather_store(a, row + k * 3 + 2, bv);
        if (k == 0) {
          const int64_t b0 = row + a.K * 3;
          gather_store(a, b0, cx);
          gather_store(a, b0 + 1, cy);
          gather_store(a, b0 + 2, sw);
          gather_store(a, b0 + 3, sh);
          gather_store(a, b0 + 4, area);
          gather_store(a, b0 + 5, sc);
        }
      }
    }
  }
  if (a.g_on) {
    // The remote stores of this warp are made visible system-wide, then the last warp of
    // the last CTA publishes the step number on every rank (release, system scope): the
    // consumers wait for it with pc_wait_peer_flags.  One fence per warp (by the lane that
    // stored), the counters only count.
    if (lane == 0) {
      __threadfence_system();
      if (atomicAdd(s_warps_done, 1) == kConsumerWarps - 1) {
        if (atomicAdd(a.g_counter, 1) == (int)gridDim.x - 1) {
          *a.g_counter = 0;  // for the next launch (stream order)
          const uint32_t step = *a.g_step + 1u;
          *a.g_step = step;
          __threadfence_system();
          for (int p = 0; p < a.g_nflags; ++p)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.g_flags[p] + a.g_rank),
                         "r"(step)
                         : "memory");
        }
      }
    }
  }
}

// Blur kernel exactly as _create_gaussian_kernel builds it
// (top_down_decoder.py:207-215): fp64 exp, fp64 normalisation, one rounding.
static void build_dark_kernel(int ks, float* out) {
  const double sigma = 0.3 * ((ks - 1) * 0.5 - 1) + 0.8;
  const int r = (ks - 1) / 2;
  double tmp[PC_MAX_DARK_KERNEL * PC_MAX_DARK_KERNEL];
  double sum = 0.0;
  for (int y = 0; y < ks; ++y)
    for (int x = 0; x < ks; ++x) {
      const double d2 = (double)((x - r) * (x - r) + (y - r) * (y - r));
      tmp[y * ks + x] = exp(-d2 / (2 * sigma * sigma));
    }
  // numpy's float64 sum over the flattened array is pairwise; 289 terms of
  // similar magnitude agree with a plain sum to well below float32 rounding.
  for (int i = 0; i < ks * ks; ++i) sum += tmp[i];
  for (int i = 0; i < ks * ks; ++i) out[i] = (float)(tmp[i] / sum);
}

}  // namespace pc

using namespace pc;

static int decode_launch(const float* d_heatmap, const float* d_flipped, const float* d_center,
                         const float* d_scale, const float* d_score, float* d_all_preds,
                         float* d_all_boxes, const pc_topdown_decode_params* p, int64_t n,
                         const pc_gather_target* g, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode: params is NULL");
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode: n = %lld < 0", (long long)n);
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_decode: num_joints %d outside [1, %d]", p->num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(p->height >= 1 && p->width >= 1, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_decode: bad map size %dx%d", p->height, p->width);
  PC_REQUIRE(!(p->dark_udp_refine && p->shift_coordinate), PC_ERR_INVALID_ARGUMENT,
             "`udp_refine` and `shift_coordinate` cannot be `true` in the same time.");
  if (p->dark_udp_refine)
    PC_REQUIRE(p->kernel_size >= 1 && p->kernel_size <= PC_MAX_DARK_KERNEL &&
                   (p->kernel_size & 1),
               PC_ERR_UNSUPPORTED, "pc_topdown_decode: kernel_size %d must be odd and <= %d",
               p->kernel_size, PC_MAX_DARK_KERNEL);
  if (g) {
    PC_REQUIRE(n >= 1, PC_ERR_INVALID_ARGUMENT,
               "pc_topdown_decode_gather: a rank without crops must still signal -- use "
               "pc_scatter_results_signal with n = 0");
    PC_REQUIRE(g->row_offset >= 0 && g->d_step && g->d_counter && g->h_peer_flags &&
                   g->num_flag_peers >= 1 && g->num_flag_peers <= kGatherMaxPeers &&
                   g->my_rank >= 0 && g->my_rank < g->num_flag_peers,
               PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode_gather: bad flags / step / counter");
    PC_REQUIRE(g->d_multicast_table || (g->h_peer_tables && g->num_peers >= 1 &&
                                        g->num_peers <= kGatherMaxPeers),
               PC_ERR_INVALID_ARGUMENT,
               "pc_topdown_decode_gather: need a multicast table or 1..%d peer tables",
               kGatherMaxPeers);
  }
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_heatmap && d_center && d_scale && d_score && d_all_preds && d_all_boxes,
             PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode: NULL tensor pointer");
  PC_REQUIRE(!p->flip_test || d_flipped, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_decode: flip_test needs the flipped heatmap");
  const int64_t hw = (int64_t)p->height * p->width;
  PC_REQUIRE(((uintptr_t)d_heatmap % 4 == 0) && (!p->flip_test || (uintptr_t)d_flipped % 4 == 0),
             PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode: heatmaps must be float32-aligned");
  // bulk copies need 16-byte aligned planes; any other H, W / base takes the manual copy
  const bool bulk_ok = hw % 4 == 0 && ((uintptr_t)d_heatmap % 16 == 0) &&
                       (!p->flip_test || (uintptr_t)d_flipped % 16 == 0);
  PC_REQUIRE((uint64_t)hw * p->width < 0xffffffffull, PC_ERR_UNSUPPORTED,
             "pc_topdown_decode: map %dx%d too large", p->height, p->width);

  DecodeTables tab;
  memset(&tab, 0, sizeof(tab));
  for (int k = 0; k < p->num_joints; ++k) {
    tab.flip_index[k] = p->flip_test ? p->flip_index[k] : k;
    PC_REQUIRE(tab.flip_index[k] >= 0 && tab.flip_index[k] < p->num_joints,
               PC_ERR_INVALID_ARGUMENT, "pc_topdown_decode: flip_index[%d] = %d out of range", k,
               tab.flip_index[k]);
  }
  if (p->dark_udp_refine) {
    if (p->dark_kernel_set)
      memcpy(tab.dark_kernel, p->dark_kernel, sizeof(float) * p->kernel_size * p->kernel_size);
    else
      build_dark_kernel(p->kernel_size, tab.dark_kernel);
  }

  DecodeArgs a;
  a.heatmap = d_heatmap;
  a.flipped = d_flipped;
  a.center = d_center;
  a.scale = d_scale;
  a.score = d_score;
  a.all_preds = d_all_preds;
  a.all_boxes = d_all_boxes;
  a.K = p->num_joints;
  a.H = p->height;
  a.W = p->width;
  a.HW = (int32_t)hw;
  a.num_items = n * a.K;
  a.divW = make_fastdiv((uint32_t)a.W);
  a.divK = make_fastdiv((uint32_t)a.K);
  a.pixel_std = p->pixel_std;
  a.to_original = p->to_original;
  a.use_udp = p->use_udp;
  a.mode = p->shift_coordinate ? 1 : (p->dark_udp_refine ? 2 : 0);
  a.ks = p->kernel_size;
  a.shift_heatmap = p->flip_test ? p->shift_heatmap : 0;
  a.vec_ok = (a.W % 4 == 0);
  a.bulk_ok = bulk_ok ? 1 : 0;
  a.g_on = g ? 1 : 0;
  a.g_peers = a.g_nflags = a.g_rank = 0;
  a.g_row0 = 0;
  a.g_mc = nullptr;
  a.g_step = nullptr;
  a.g_counter = nullptr;
  for (int i = 0; i < kGatherMaxPeers; ++i) {
    a.g_peer[i] = nullptr;
    a.g_flags[i] = nullptr;
  }
  if (g) {
    a.g_row0 = g->row_offset;
    a.g_mc = static_cast<float*>(g->d_multicast_table);
    if (!a.g_mc) {
      a.g_peers = g->num_peers;
      for (int i = 0; i < g->num_peers; ++i) {
        PC_REQUIRE(g->h_peer_tables[i] != nullptr, PC_ERR_INVALID_ARGUMENT,
                   "pc_topdown_decode_gather: peer table %d is NULL", i);
        a.g_peer[i] = static_cast<float*>(g->h_peer_tables[i]);
      }
    }
    a.g_nflags = g->num_flag_peers;
    a.g_rank = g->my_rank;
    for (int i = 0; i < g->num_flag_peers; ++i) {
      PC_REQUIRE(g->h_peer_flags[i] != nullptr, PC_ERR_INVALID_ARGUMENT,
                 "pc_topdown_decode_gather: flag array %d is NULL", i);
      a.g_flags[i] = static_cast<uint32_t*>(g->h_peer_flags[i]);
    }
    a.g_step = g->d_step;
    a.g_counter = g->d_counter;
  }
  a.divWq = make_fastdiv((uint32_t)(a.vec_ok ? a.W / 4 : 1));
  a.step_xq = a.vec_ok ? 32 % (a.W / 4) : 0;
  a.step_y = a.vec_ok ? 32 / (a.W / 4) : 0;
  a.divKs = make_fastdiv((uint32_t)(p->dark_udp_refine ? p->kernel_size : 1));
  a.divWin = make_fastdiv((uint32_t)(p->dark_udp_refine ? p->kernel_size + 2 : 1));
  a.stage_floats = (uint32_t)(p->flip_test ? 2 * hw : hw);

  // ---- stages and consumer groups ------------------------------------------------
  // G warps share an item.  A stage is busy from the moment its copy is issued until
  // the item's finisher releases it; with items arriving every T_arr = bytes / (HBM
  // share of one SM) cycles, the stages busy being CONSUMED are (Little's law)
  //     busy(G) = (T_scan / G + T_refine) / T_arr,
  // where one warp scans at about 1/4.2 of the arrival rate (T_scan = 4.2 T_arr, measured)
  // and the DARK step costs about 2000 cycles.  Pick the smallest G that leaves two
  // stages to be in flight; the DARK scratch is per group (one finisher per group at a
  // time), so fewer, larger groups also free shared memory for stages.
  a.row_floats = a.mode == 2 ? kDarkSamples * PC_MAX_DARK_KERNEL : 0;
  a.win_floats = a.mode == 2 ? (p->kernel_size + 2) * (p->kernel_size + 2) : 0;
  const size_t smem_cap = 227 * 1024;
  const size_t stage_bytes = (size_t)a.stage_floats * sizeof(float);
  const double t_arr = (double)stage_bytes / 23.0;  // cycles per item at 6.5 TB/s / 148 SMs
  const double refine = (a.mode == 2 ? 2000.0 : 250.0) / t_arr;
  int stages = 0, group = 1;
  size_t tail_bytes = 0;
  for (group = 1;; group <<= 1) {
    const int ng = kConsumerWarps / group;
    tail_bytes = 3 * kMaxStages * sizeof(uint64_t) + kMaxStages * sizeof(uint32_t) +
                 2 * kMaxStages * kConsumerWarps * sizeof(float) +
                 sizeof(float) * PC_MAX_DARK_KERNEL * PC_MAX_DARK_KERNEL +
                 sizeof(float) * ng * (a.row_floats + a.win_floats);
    stages = (int)((smem_cap - tail_bytes - 16) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    const double busy = 4.2 / group + refine;
    if (busy <= stages - 2 || group == kConsumerWarps) break;
  }
  PC_REQUIRE(stages >= 2, PC_ERR_UNSUPPORTED,
             "pc_topdown_decode: a %dx%d plane%s does not fit two shared-memory stages",
             p->height, p->width, p->flip_test ? " pair" : "");
  a.stages = stages;
  a.group = group;
  const size_t smem = ((stages * stage_bytes + 15) & ~(size_t)15) + tail_bytes;

  const int sms = sm_count_cached();
  PC_REQUIRE(sms > 0, PC_ERR_NO_DEVICE, "pc_topdown_decode: no CUDA device");
  int64_t grid = a.num_items < sms ? a.num_items : sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->flip_test) {
    PC_CUDA(cudaFuncSetAttribute(topdown_decode_kernel<true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topdown_decode_kernel<true><<<(unsigned)grid, kDecodeThreads, smem, st>>>(a, tab);
  } else {
    PC_CUDA(cudaFuncSetAttribute(topdown_decode_kernel<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topdown_decode_kernel<false><<<(unsigned)grid, kDecodeThreads, smem, st>>>(a, tab);
  }
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_topdown_decode(const float* d_heatmap, const float* d_flipped,
                                 const float* d_center, const float* d_scale,
                                 const float* d_score, float* d_all_preds, float* d_all_boxes,
                                 const pc_topdown_decode_params* p, int64_t n, void* stream) {
  return decode_launch(d_heatmap, d_flipped, d_center, d_scale, d_score, d_all_preds,
                       d_all_boxes, p, n, nullptr, stream);
}

extern "C" int pc_topdown_decode_gather(const float* d_heatmap, const float* d_flipped,
                                        const float* d_center, const float* d_scale,
                                        const float* d_score, float* d_all_preds,
                                        float* d_all_boxes, const pc_topdown_decode_params* p,
                                        int64_t n, const pc_gather_target* gather,
                                        void* stream) {
  PC_REQUIRE(gather != nullptr, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_decode_gather: gather target is NULL");
  return decode_launch(d_heatmap, d_flipped, d_center, d_scale, d_score, d_all_preds,
                       d_all_boxes, p, n, gather, stream);
}
