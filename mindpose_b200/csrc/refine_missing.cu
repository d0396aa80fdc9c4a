// refine_missing.cu -- missing-joint refinement of the bottom-up inferencer (sm_100a).
//
// Replaces BottomUpHeatMapAEInferencer._refine_missing
// (mindpose/engine/inferencer/bottomup_inferencer.py:189-249) -- SURVEY.md section 8(f),
// row N3 -- for every person of every image at once.
//
// The reference re-reads the K x H x W heat-map and tag stacks once PER PERSON (8.9 MB per
// person at 17 x 256 x 256).  Here a CTA owns one (image, joint) plane and evaluates all
// people of the image against it: each pixel of the plane is loaded once per chunk of 8
// people and scored for the 8 of them (heat - round(|tag - mean_tag_p|)), a running
// argmax per person lives in registers, and one block reduction per chunk picks
// (value desc, flat index asc) -- numpy.argmax's first occurrence.
#include <math.h>

#include "common.cuh"

namespace pc {

constexpr int kRefThreads = 256;
constexpr int kRefChunk = 8;

// numpy's float32 add.reduce over n <= 64 contiguous values (pairwise_sum)
__device__ float np_sum_f32_local(const float* a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  float r[8];
  for (int t = 0; t < 8; ++t) r[t] = a[t];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int t = 0; t < 8; ++t) r[t] = __fadd_rn(r[t], a[i + t]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}

// mean tag of every person: np.mean over the tags at the detected joints (:204-210)
__global__ void refine_mean_tag_kernel(const float* __restrict__ tagging,
                                       const float* __restrict__ ans,
                                       const int32_t* __restrict__ num_groups,
                                       float* __restrict__ mean_tag, int K, int H, int W, int G) {
  const int n = blockIdx.x, p = blockIdx.y * blockDim.x + threadIdx.x;
  const int np_ = num_groups[n];
  if (p >= np_ || p >= G) return;
  const float* person = ans + ((size_t)n * G + p) * K * 4;
  float tags[PC_MAX_JOINTS];
  int nv = 0;
  for (int k = 0; k < K; ++k) {
    if (person[4 * k + 2] > 0.f) {
      // keypoints[:, :2].astype(np.int32): truncation toward zero
      int x = (int)person[4 * k], y = (int)person[4 * k + 1];
      x = min(max(x, 0), W - 1);  // the reference would raise IndexError outside the map
      y = min(max(y, 0), H - 1);
      tags[nv++] = __ldg(tagging + (((size_t)n * K + k) * H + y) * W + x);
    }
  }
  mean_tag[(size_t)n * G + p] = __fdiv_rn(np_sum_f32_local(tags, nv), (float)nv);
}

__device__ __forceinline__ void score_px(float heat, float tag, int idx, const float (&mt)[kRefChunk],
                                         float (&bv)[kRefChunk], int (&bi)[kRefChunk]) {
#pragma unroll
  for (int c = 0; c < kRefChunk; ++c) {
    // np.linalg.norm over one tag channel is sqrt(d * d), which in binary floating point
    // equals |d| exactly unless d * d over- or underflows (|d| > 1.8e19: the reference gets
    // inf there; |d| < 1e-19 rounds to 0 either way)
    const float nd = fabsf(__fsub_rn(tag, mt[c]));
    // np.round: half to even.  (rintf is a quarter-rate conversion-pipe instruction; the
    // adder form (nd + 2^23) - 2^23 with a select for nd >= 2^23 was measured: 0.248 ms
    // against 0.203 ms -- the kernel is bound by instruction issue, and that form is three
    // instructions more.)
    const float s = __fsub_rn(heat, rintf(nd));
    if (s > bv[c]) {
      bv[c] = s;
      bi[c] = idx;
    }
  }
}

__global__ void __launch_bounds__(kRefThreads)
    refine_missing_kernel(const float* __restrict__ heatmap, const float* __restrict__ tagging,
                          float* __restrict__ ans, const int32_t* __restrict__ num_groups,
                          const float* __restrict__ mean_tag, int K, int H, int W, int vec_ok,
                          int G) {
  __shared__ float s_v[kRefThreads / 32][kRefChunk];
  __shared__ int s_i[kRefThreads / 32][kRefChunk];
  const int n = blockIdx.x / K, k = blockIdx.x - n * K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int np_ = num_groups[n];
  if (np_ <= 0) return;
  np_ = min(np_, G);
  const int HW = H * W;
  const float* heat = heatmap + ((size_t)n * K + k) * HW;
  const float* tagp = tagging + ((size_t)n * K + k) * HW;

  for (int p0 = 0; p0 < np_; p0 += kRefChunk) {
    float mt[kRefChunk], bv[kRefChunk];
    int bi[kRefChunk];
#pragma unroll
    for (int c = 0; c < kRefChunk; ++c) {
      // people past the end get a NaN mean: never better than -inf
      mt[c] = p0 + c < np_ ? __ldg(mean_tag + (size_t)n * G + p0 + c)
                           : __int_as_float(0x7fc00000);
      bv[c] = -INFINITY;
      bi[c] = 0x7fffffff;
    }
    if (vec_ok) {
      for (int q = tid; q < (HW >> 2); q += kRefThreads) {
        const float4 h4 = __ldg(reinterpret_cast<const float4*>(heat) + q);
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(tagp) + q);
        score_px(h4.x, t4.x, 4 * q, mt, bv, bi);
        score_px(h4.y, t4.y, 4 * q + 1, mt, bv, bi);
        score_px(h4.z, t4.z, 4 * q + 2, mt, bv, bi);
        score_px(h4.w, t4.w, 4 * q + 3, mt, bv, bi);
      }
    } else {
      for (int i = tid; i < HW; i += kRefThreads)
        score_px(__ldg(heat + i), __ldg(tagp + i), i, mt, bv, bi);
    }
#pragma unroll
    for (int c = 0; c < kRefChunk; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv[c], o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi[c], o);
        if (ov > bv[c] || (ov == bv[c] && oi < bi[c])) {
          bv[c] = ov;
          bi[c] = oi;
        }
      }
      if (lane == 0) {
        s_v[warp][c] = bv[c];
        s_i[warp][c] = bi[c];
      }
    }
    __syncthreads();
    if (tid < kRefChunk && p0 + tid < np_) {
      float v = s_v[0][tid];
      int i = s_i[0][tid];
      for (int w = 1; w < kRefThreads / 32; ++w)
        if (s_v[w][tid] > v || (s_v[w][tid] == v && s_i[w][tid] < i)) {
          v = s_v[w][tid];
          i = s_i[w][tid];
        }
      if (i == 0x7fffffff) i = 0;  // nothing compared greater than -inf: argmax 0
      const int y = i / W, x = i - y * W;
      // +0.5, then +-0.25 toward the higher neighbour (minus on ties), borders clamped
      float fx = __fadd_rn((float)x, 0.5f), fy = __fadd_rn((float)y, 0.5f);
      const bool px = heat[y * W + min(x + 1, W - 1)] > heat[y * W + max(x - 1, 0)];
      const bool py = heat[min(y + 1, H - 1) * W + x] > heat[max(y - 1, 0) * W + x];
      fx = __fadd_rn(fx, px ? 0.25f : -0.25f);
      fy = __fadd_rn(fy, py ? 0.25f : -0.25f);
      const float val = heat[y * W + x];
      float* row = ans + (((size_t)n * G + p0 + tid) * K + k) * 4;
      if (val > 0.f && row[2] == 0.f) {
        row[0] = fx;
        row[1] = fy;
        row[2] = val;
      }
    }
    __syncthreads();
  }
}

}  // namespace pc

using namespace pc;

extern "C" int pc_refine_missing(const float* d_heatmap, const float* d_tagging, float* d_ans,
                                 const int32_t* d_num_groups, float* d_mean_tag,
                                 const pc_refine_params* p, int64_t n, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_refine_missing: params is NULL");
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_refine_missing: n < 0");
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_refine_missing: num_joints %d outside [1, %d]", p->num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(p->height >= 1 && p->width >= 1 && (int64_t)p->height * p->width < (1 << 30),
             PC_ERR_INVALID_ARGUMENT, "pc_refine_missing: bad map size");
  PC_REQUIRE(p->max_groups >= 0, PC_ERR_INVALID_ARGUMENT, "pc_refine_missing: max_groups < 0");
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_heatmap && d_tagging && d_ans && d_num_groups && d_mean_tag,
             PC_ERR_INVALID_ARGUMENT, "pc_refine_missing: NULL tensor pointer");
  PC_REQUIRE(n * p->num_joints < 0x7fffffffLL, PC_ERR_UNSUPPORTED,
             "pc_refine_missing: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int G = p->max_groups > 0 ? p->max_groups : PC_MAX_GROUPS;
  refine_mean_tag_kernel<<<dim3((unsigned)n, (unsigned)((G + 127) / 128)), 128, 0, st>>>(
      d_tagging, d_ans, d_num_groups, d_mean_tag, p->num_joints, p->height, p->width, G);
  PC_CUDA(cudaGetLastError());
  const int vec_ok = ((int64_t)p->height * p->width) % 4 == 0 &&
                     ((uintptr_t)d_heatmap % 16 == 0) && ((uintptr_t)d_tagging % 16 == 0);
  refine_missing_kernel<<<(unsigned)(n * p->num_joints), kRefThreads, 0, st>>>(
      d_heatmap, d_tagging, d_ans, d_num_groups, d_mean_tag, p->num_joints, p->height, p->width,
      vec_ok, G);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
