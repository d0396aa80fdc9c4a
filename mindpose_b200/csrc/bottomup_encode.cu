// bottomup_encode.cu -- multi-resolution bottom-up target encoding for sm_100a.
//
// Replaces BottomUpGenerateTarget._encoding / ._generate_heatmap_and_tag_ind
// (mindpose/data/transform/bottomup_transform.py:504-598) and pad_to_same
// (mindpose/data/transform/utils.py:213-232) -- SURVEY.md section 8(f), row N1.
//
// One CTA per (image, scale, joint) plane of the padded output [Hmax, Wmax].  The plane
// is streamed out once with 128-bit st.global.cs zero stores (the write pass that bounds
// the kernel: 262,144 B per 256x256 plane), then -- after a block barrier, which orders
// the two sets of stores -- the at most max_num 13x13 windows are written on top.  A
// window pixel takes the maximum over every person whose window covers it (the
// reference's np.maximum merge, order independent), so overlapping windows store the same
// value and need no atomics.
#include <math.h>

#include "common.cuh"

namespace pc {

constexpr int kBueThreads = 256;
constexpr int kBueMaxPeople = 64;

struct BueArgs {
  const float* keypoints;  // [N, S, M, K, 3]
  float* target;           // [N, S, K, Hmax, Wmax]
  int32_t* tag_ind;        // [N, S, max_num, K, 2] or [N, S, max_num, 2]
  int32_t S, M, K, max_num, hmax, wmax;
  int32_t w[PC_MAX_SCALES], h[PC_MAX_SCALES];
  int32_t tmp, size;  // 3 sigma, 2 * tmp + 1
  float c0;           // size // 2
  float two_sigma2;
  int32_t tag_per_joint;
  int32_t vec_ok;
  int32_t bands;      // CTAs per plane: each fills and pastes hmax / bands rows
};

struct Window {
  float cx, cy;          // sub-pixel centre inside the patch (float32, as numpy >= 2 computes it)
  int ul_x, ul_y;        // patch origin in the map
  int x_lo, x_hi, y_lo, y_hi;  // clipped window; x_lo >= x_hi marks "nothing to paste"
  int mu_x, mu_y;
};

__device__ __forceinline__ Window make_window(float px, float py, float vis, int W, int H,
                                              int tmp, float c0) {
  Window wd;
  wd.x_lo = wd.x_hi = wd.y_lo = wd.y_hi = 0;
  wd.ul_x = wd.ul_y = 0;
  wd.cx = wd.cy = 0.f;
  // Python round(): half to even; far-away centres are clamped so the int math is safe
  const float rx = fminf(fmaxf(rintf(px), -1.0e6f), 1.0e6f);
  const float ry = fminf(fmaxf(rintf(py), -1.0e6f), 1.0e6f);
  wd.mu_x = (int)rx;
  wd.mu_y = (int)ry;
  if (!(vis > 0.f)) {
    wd.mu_x = -1;  // not visible: no window, no tag
    return wd;
  }
  const int ul_x = wd.mu_x - tmp, ul_y = wd.mu_y - tmp;
  const int br_x = wd.mu_x + tmp + 1, br_y = wd.mu_y + tmp + 1;
  if (ul_x >= W || ul_y >= H || br_x < 0 || br_y < 0) {
    wd.mu_x = -1;  // the reference `continue`s before the tag assignment
    return wd;
  }
  wd.ul_x = ul_x;
  wd.ul_y = ul_y;
  wd.x_lo = max(0, ul_x);
  wd.x_hi = min(br_x, W);
  wd.y_lo = max(0, ul_y);
  wd.y_hi = min(br_y, H);
  wd.cx = __fsub_rn(__fadd_rn(c0, px), rx);  // x0 + pt[0] - mu_x, left to right in float32
  wd.cy = __fsub_rn(__fadd_rn(c0, py), ry);
  return wd;
}

__device__ __forceinline__ float window_value(const Window& wd, int x, int y, float two_sigma2) {
  const float dx = __fsub_rn((float)(x - wd.ul_x), wd.cx);
  const float dy = __fsub_rn((float)(y - wd.ul_y), wd.cy);
  const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  // expf is within 2 ulp (2.4e-7 relative) of numpy's float32 exp: far inside the 1e-5 bar
  return expf(__fdiv_rn(-s, two_sigma2));
}

__global__ void __launch_bounds__(kBueThreads) bottomup_encode_kernel(const BueArgs a) {
  __shared__ Window s_win[kBueMaxPeople];
  __shared__ int s_valid[kBueMaxPeople];  // people with something to paste, any order
  __shared__ int s_nvalid;

  // A plane is cut into `bands` row bands, one CTA each: many short CTAs whose load / fill /
  // paste phases interleave across the SM instead of a few long ones in lockstep.
  const int tid = threadIdx.x;
  const int band = blockIdx.x % a.bands;
  const int plane_id = blockIdx.x / a.bands;
  const int k = plane_id % a.K;
  const int s = (plane_id / a.K) % a.S;
  const int64_t n = plane_id / (a.K * a.S);
  const int band_rows = a.hmax / a.bands;
  const int y_band0 = band * band_rows, y_band1 = y_band0 + band_rows;
  const int W = a.w[s], H = a.h[s];

  if (tid == 0) s_nvalid = 0;
  __syncthreads();
  if (tid < a.M) {
    const float* kp = a.keypoints + ((((size_t)n * a.S + s) * a.M + tid) * a.K + k) * 3;
    const Window wd = make_window(__ldg(kp), __ldg(kp + 1), __ldg(kp + 2), W, H, a.tmp, a.c0);
    s_win[tid] = wd;
    if (wd.x_lo < wd.x_hi && wd.y_lo < wd.y_hi) s_valid[atomicAdd(&s_nvalid, 1)] = tid;
  }

  // ---- the write pass: zeros over the whole padded plane ------------------------------
  float* plane = a.target + (((size_t)n * a.S + s) * a.K + k) * a.hmax * a.wmax;
  float* band_base = plane + (size_t)y_band0 * a.wmax;
  const int total = band_rows * a.wmax;
  if (a.vec_ok) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < (total >> 2); i += kBueThreads) st_stream_f4(band_base + 4 * i, z);
  } else {
    for (int i = tid; i < total; i += kBueThreads) band_base[i] = 0.f;
  }
  __syncthreads();  // windows are visible; the zero stores are ordered before the paste

  // ---- paste: every window pixel = max over the windows that cover it -------------------
  const int area = a.size * a.size;
  const int nv = s_nvalid;
  for (int t = tid; t < nv * area; t += kBueThreads) {
    const int vi = t / area, e = t - vi * area;
    const Window wd = s_win[s_valid[vi]];
    const int y = wd.ul_y + e / a.size, x = wd.ul_x + e % a.size;
    if (x < wd.x_lo || x >= wd.x_hi || y < wd.y_lo || y >= wd.y_hi) continue;
    if (y < y_band0 || y >= y_band1) continue;  // another band's rows
    float v = window_value(wd, x, y, a.two_sigma2);
    for (int o = 0; o < nv; ++o) {
      if (o == vi) continue;
      const Window& w2 = s_win[s_valid[o]];
      if (x >= w2.x_lo && x < w2.x_hi && y >= w2.y_lo && y < w2.y_hi)
        v = fmaxf(v, window_value(w2, x, y, a.two_sigma2));
    }
    plane[y * a.wmax + x] = v;
  }

  // ---- tag_ind (first band only) ---------------------------------------------------------
  if (band != 0) return;
  if (a.tag_per_joint) {
    if (tid < a.max_num) {
      int2 r = make_int2(0, 0);
      if (tid < a.M) {
        const Window wd = s_win[tid];
        if (wd.mu_x >= 0 && wd.mu_x < W && wd.mu_y >= 0 && wd.mu_y < H)
          r = make_int2(wd.mu_y * W + wd.mu_x, 1);
      }
      int32_t* o = a.tag_ind + ((((size_t)n * a.S + s) * a.max_num + tid) * a.K + k) * 2;
      o[0] = r.x;
      o[1] = r.y;
    }
  } else if (k == 0 && tid < a.max_num) {
    // one tag per person: the last joint (in index order) with a usable centre wins
    int2 r = make_int2(0, 0);
    if (tid < a.M) {
      for (int kk = 0; kk < a.K; ++kk) {
        const float* kp = a.keypoints + ((((size_t)n * a.S + s) * a.M + tid) * a.K + kk) * 3;
        const Window wd = make_window(__ldg(kp), __ldg(kp + 1), __ldg(kp + 2), W, H, a.tmp, a.c0);
        if (wd.mu_x >= 0 && wd.mu_x < W && wd.mu_y >= 0 && wd.mu_y < H)
          r = make_int2(wd.mu_y * W + wd.mu_x, 1);
      }
    }
    int32_t* o = a.tag_ind + (((size_t)n * a.S + s) * a.max_num + tid) * 2;
    o[0] = r.x;
    o[1] = r.y;
  }
}

}  // namespace pc

using namespace pc;

extern "C" int pc_bottomup_encode(const float* d_keypoints, float* d_target, int32_t* d_tag_ind,
                                  const pc_bottomup_encode_params* p, int64_t n, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_bottomup_encode: params is NULL");
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_bottomup_encode: n < 0");
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_bottomup_encode: num_joints %d outside [1, %d]", p->num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(p->num_scales >= 1 && p->num_scales <= PC_MAX_SCALES, PC_ERR_UNSUPPORTED,
             "pc_bottomup_encode: num_scales %d outside [1, %d]", p->num_scales, PC_MAX_SCALES);
  PC_REQUIRE(p->max_num >= 1 && p->max_num <= kBueMaxPeople, PC_ERR_UNSUPPORTED,
             "pc_bottomup_encode: max_num %d outside [1, %d]", p->max_num, kBueMaxPeople);
  PC_REQUIRE(p->num_people >= 0, PC_ERR_INVALID_ARGUMENT, "pc_bottomup_encode: num_people < 0");
  // the reference raises ValueError here (bottomup_transform.py:536-540)
  PC_REQUIRE(p->num_people <= p->max_num, PC_ERR_INVALID_ARGUMENT,
             "Number of keypoints in one image `%d` exeeds the maximum num: `%d`", p->num_people,
             p->max_num);
  const double tmpd = (double)p->sigma * 3.0;
  PC_REQUIRE(p->sigma > 0.f && tmpd == floor(tmpd) && tmpd <= 15.0, PC_ERR_UNSUPPORTED,
             "pc_bottomup_encode: 3*sigma must be an integer in [1, 15] (sigma = %g)", p->sigma);
  int hmax = 0, wmax = 0;
  for (int s = 0; s < p->num_scales; ++s) {
    PC_REQUIRE(p->heatmap_w[s] >= 1 && p->heatmap_h[s] >= 1 &&
                   (int64_t)p->heatmap_w[s] * p->heatmap_h[s] < (1 << 30),
               PC_ERR_INVALID_ARGUMENT, "pc_bottomup_encode: bad heatmap size at scale %d", s);
    hmax = p->heatmap_h[s] > hmax ? p->heatmap_h[s] : hmax;
    wmax = p->heatmap_w[s] > wmax ? p->heatmap_w[s] : wmax;
  }
  PC_REQUIRE((int64_t)hmax * wmax < (1 << 30), PC_ERR_UNSUPPORTED,
             "pc_bottomup_encode: padded map too large");
  if (n == 0) return PC_OK;
  PC_REQUIRE((p->num_people == 0 || d_keypoints) && d_target && d_tag_ind,
             PC_ERR_INVALID_ARGUMENT, "pc_bottomup_encode: NULL tensor pointer");
  const int64_t grid = n * p->num_scales * p->num_joints;
  PC_REQUIRE(grid < 0x7fffffffLL, PC_ERR_UNSUPPORTED, "pc_bottomup_encode: batch too large");

  BueArgs a;
  a.keypoints = d_keypoints;
  a.target = d_target;
  a.tag_ind = d_tag_ind;
  a.S = p->num_scales;
  a.M = p->num_people;
  a.K = p->num_joints;
  a.max_num = p->max_num;
  a.hmax = hmax;
  a.wmax = wmax;
  for (int s = 0; s < PC_MAX_SCALES; ++s) {
    a.w[s] = s < p->num_scales ? p->heatmap_w[s] : 1;
    a.h[s] = s < p->num_scales ? p->heatmap_h[s] : 1;
  }
  a.tmp = (int32_t)tmpd;
  a.size = 2 * a.tmp + 1;
  a.c0 = (float)(a.size / 2);
  a.two_sigma2 = (float)(2.0 * (double)p->sigma * (double)p->sigma);
  a.tag_per_joint = p->tag_per_joint;
  a.bands = (hmax % 4 == 0 && hmax >= 64) ? 4 : 1;
  a.vec_ok = ((int64_t)(hmax / a.bands) * wmax) % 4 == 0 && ((uintptr_t)d_target % 16 == 0);
  PC_REQUIRE(grid * a.bands < 0x7fffffffLL, PC_ERR_UNSUPPORTED,
             "pc_bottomup_encode: batch too large");
  bottomup_encode_kernel<<<(unsigned)(grid * a.bands), kBueThreads, 0, (cudaStream_t)stream>>>(a);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
