// topdown_encode.cu -- fused heatmap target encoding for sm_100a.
//
// Replaces TopDownGenerateTarget._encoding / ._udp_encoding
// (mindpose/data/transform/topdown_transform.py:324-375 / :377-430).
//
// A work item is one joint map (n, k).  One warp owns an item: it derives the
// window from the keypoint (fp64 scalar math, the reference's rounding rules),
// then streams the whole H*W plane out with 128-bit st.global.cs stores --
// zeros outside the window, the Gaussian inside -- so zero-fill, splat and
// target_weight are a single write pass over HBM (208,896 B per 17x64x48 crop).
#include <math.h>

#include "common.cuh"

namespace pc {

constexpr int kEncodeThreads = 256;
constexpr int kEncodeWarps = kEncodeThreads / 32;
constexpr int kMaxPatch = 31;  // 2 * (3 * sigma) + 1 with 3 * sigma <= 15

struct EncodeArgs {
  const float* keypoints;
  float* target;
  float* target_weight;
  int64_t num_items;
  int32_t K, H, W, HW;
  FastDiv divW;
  double stride_x, stride_y;  // feat_stride (fp64, as numpy computes it)
  double two_sigma2;          // 2 * sigma^2
  int32_t tmp;                // 3 * sigma
  int32_t size;               // 2 * tmp + 1
  int32_t use_joint_weights;
  int32_t vec_ok;
};

struct EncodeTables {
  float joint_weights[PC_MAX_JOINTS];
};

template <bool UDP>
__global__ void __launch_bounds__(kEncodeThreads)
    topdown_encode_kernel(const EncodeArgs a, const __grid_constant__ EncodeTables tab) {
  __shared__ float s_patch[UDP ? 1 : kMaxPatch * kMaxPatch];
  __shared__ double s_g[UDP ? kEncodeWarps * 2 * 32 : 1];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int size = a.size, tmp = a.tmp;
  const int W = a.W, H = a.H;

  if (!UDP) {
    // 13x13 (sigma = 2) patch, centre value 1: exp(-((x-c)^2 + (y-c)^2) / (2 sigma^2)),
    // float32 argument as the reference, exponential rounded once.
    for (int i = threadIdx.x; i < size * size; i += blockDim.x) {
      const int py = i / size, px = i - py * size;
      const float d2 = (float)((px - tmp) * (px - tmp) + (py - tmp) * (py - tmp));
      const float arg = __fdiv_rn(-d2, (float)a.two_sigma2);
      s_patch[i] = (float)exp((double)arg);
    }
    __syncthreads();
  }

  const int64_t warp_global = (int64_t)blockIdx.x * kEncodeWarps + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * kEncodeWarps;
  for (int64_t item = warp_global; item < a.num_items; item += warp_stride) {
    const int k = (int)(item % a.K);
    const float* kp = a.keypoints + item * 3;
    const float kx = __ldg(kp), ky = __ldg(kp + 1), vis = __ldg(kp + 2);

    const double qx = __ddiv_rn((double)kx, a.stride_x);
    const double qy = __ddiv_rn((double)ky, a.stride_y);
    double mxd, myd;
    if (UDP) {  // int(q + 0.5): truncation toward zero
      mxd = trunc(__dadd_rn(qx, 0.5));
      myd = trunc(__dadd_rn(qy, 0.5));
    } else {  // Python round(): half to even
      mxd = rint(qx);
      myd = rint(qy);
    }
    // centres far outside the map cannot touch it; clamp so the int math is safe
    mxd = fmin(fmax(mxd, -1.0e6), 1.0e6);
    myd = fmin(fmax(myd, -1.0e6), 1.0e6);
    const int mu_x = (int)mxd, mu_y = (int)myd;
    const int ul_x = mu_x - tmp, ul_y = mu_y - tmp;
    const int br_x = mu_x + tmp + 1, br_y = mu_y + tmp + 1;
    const bool outside = ul_x >= W || ul_y >= H || br_x < 0 || br_y < 0;
    float weight = outside ? 0.f : vis;
    const bool paste = !outside && weight > 0.5f;
    int x_lo = 0, x_hi = 0, y_lo = 0, y_hi = 0;
    double* gx = s_g + warp * 64;
    double* gy = gx + 32;
    if (paste) {
      x_lo = max(0, ul_x);
      x_hi = min(br_x, W);
      y_lo = max(0, ul_y);
      y_hi = min(br_y, H);
      if (UDP) {
        // sub-pixel centre inside the patch: c0 + q - mu (fp64, left to right)
        const double cx = __dsub_rn(__dadd_rn((double)tmp, qx), mxd);
        const double cy = __dsub_rn(__dadd_rn((double)tmp, qy), myd);
        if (lane < size) {
          const double dx = __dsub_rn((double)lane, cx);
          const double dy = __dsub_rn((double)lane, cy);
          gx[lane] = exp(-__dmul_rn(dx, dx) / a.two_sigma2);
          gy[lane] = exp(-__dmul_rn(dy, dy) / a.two_sigma2);
        }
        __syncwarp();
      }
    }

    float* out = a.target + item * a.HW;
    if (a.vec_ok) {
      const int nvec = a.HW >> 2;
#pragma unroll 4
      for (int q = lane; q < nvec; q += 32) {
        const int idx = q << 2;
        const int y = (int)fdiv((uint32_t)idx, a.divW);
        const int x0 = idx - y * W;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (paste && y >= y_lo && y < y_hi && x0 + 3 >= x_lo && x0 < x_hi) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int x = x0 + c;
            if (x >= x_lo && x < x_hi) {
              if (UDP)
                v[c] = (float)__dmul_rn(gx[x - ul_x], gy[y - ul_y]);
              else
                v[c] = s_patch[(y - ul_y) * size + (x - ul_x)];
            }
          }
        }
        st_stream_f4(out + idx, make_float4(v[0], v[1], v[2], v[3]));
      }
    } else {
      for (int idx = lane; idx < a.HW; idx += 32) {
        const int y = (int)fdiv((uint32_t)idx, a.divW);
        const int x = idx - y * W;
        float v = 0.f;
        if (paste && y >= y_lo && y < y_hi && x >= x_lo && x < x_hi) {
          if (UDP)
            v = (float)__dmul_rn(gx[x - ul_x], gy[y - ul_y]);
          else
            v = s_patch[(y - ul_y) * size + (x - ul_x)];
        }
        out[idx] = v;
      }
    }
    if (UDP) __syncwarp();  // gx / gy are reused by the next item
    if (lane == 0) {
      if (a.use_joint_weights) weight = __fmul_rn(weight, tab.joint_weights[k]);
      a.target_weight[item] = weight;
    }
  }
}

}  // namespace pc

using namespace pc;

extern "C" int pc_topdown_encode(const float* d_keypoints, float* d_target,
                                 float* d_target_weight, const pc_encode_params* p, int64_t n,
                                 void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_topdown_encode: params is NULL");
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_topdown_encode: n < 0");
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_encode: num_joints %d outside [1, %d]", p->num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(p->heatmap_w >= 1 && p->heatmap_h >= 1 && p->image_w >= 1 && p->image_h >= 1,
             PC_ERR_INVALID_ARGUMENT, "pc_topdown_encode: bad image / heatmap size");
  PC_REQUIRE(!p->use_udp || (p->heatmap_w > 1 && p->heatmap_h > 1), PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_encode: UDP needs a heatmap larger than 1x1");
  const double tmpd = (double)p->sigma * 3.0;
  PC_REQUIRE(p->sigma > 0.f && tmpd == floor(tmpd) && tmpd <= 15.0, PC_ERR_UNSUPPORTED,
             "pc_topdown_encode: 3*sigma must be an integer in [1, 15] (sigma = %g)", p->sigma);
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_keypoints && d_target && d_target_weight, PC_ERR_INVALID_ARGUMENT,
             "pc_topdown_encode: NULL tensor pointer");
  const int64_t hw = (int64_t)p->heatmap_h * p->heatmap_w;
  PC_REQUIRE((uint64_t)hw * p->heatmap_w < 0xffffffffull, PC_ERR_UNSUPPORTED,
             "pc_topdown_encode: heatmap too large");

  EncodeArgs a;
  a.keypoints = d_keypoints;
  a.target = d_target;
  a.target_weight = d_target_weight;
  a.K = p->num_joints;
  a.H = p->heatmap_h;
  a.W = p->heatmap_w;
  a.HW = (int32_t)hw;
  a.num_items = n * a.K;
  a.divW = make_fastdiv((uint32_t)a.W);
  if (p->use_udp) {
    a.stride_x = ((double)p->image_w - 1.0) / ((double)p->heatmap_w - 1.0);
    a.stride_y = ((double)p->image_h - 1.0) / ((double)p->heatmap_h - 1.0);
  } else {
    a.stride_x = (double)p->image_w / (double)p->heatmap_w;
    a.stride_y = (double)p->image_h / (double)p->heatmap_h;
  }
  a.two_sigma2 = 2.0 * (double)p->sigma * (double)p->sigma;
  a.tmp = (int32_t)tmpd;
  a.size = 2 * a.tmp + 1;
  a.use_joint_weights = p->use_joint_weights;
  a.vec_ok = (a.W % 4 == 0) && ((uintptr_t)d_target % 16 == 0);
  EncodeTables tab;
  memcpy(tab.joint_weights, p->joint_weights, sizeof(tab.joint_weights));

  const int sms = sm_count_cached();
  PC_REQUIRE(sms > 0, PC_ERR_NO_DEVICE, "pc_topdown_encode: no CUDA device");
  int64_t blocks = (a.num_items + kEncodeWarps - 1) / kEncodeWarps;
  const int64_t cap = (int64_t)sms * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->use_udp)
    topdown_encode_kernel<true><<<(unsigned)blocks, kEncodeThreads, 0, st>>>(a, tab);
  else
    topdown_encode_kernel<false><<<(unsigned)blocks, kEncodeThreads, 0, st>>>(a, tab);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
