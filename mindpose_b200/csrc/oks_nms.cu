// oks_nms.cu -- OKS rescoring + OKS NMS / soft OKS NMS on the device (sm_100a).
//
// the first consumer of the gathered keypoints.  Replaces
//   * the rescoring loop of TopDownEvaluator.eval
//     (mindpose/engine/evaluator/topdown_evaluator.py:93-110),
//   * oks_iou / oks_nms / _rescore / soft_oks_nms (mindpose/utils/nms.py:7-190),
// for all images of an evaluation at once: one CTA per image, the image's people stay in
// place in HBM (204 B each at K = 17, read through L1/L2), only scores / order / flags live
// in shared memory.  The work is O(P^2 K) per image and latency bound; it is reported in
// images per second, not against the HBM roofline.
//
// Arithmetic follows the reference operation by operation (see oracle/nms.py for the dtype
// chain): float32 differences and squares, float64 from the division by (2 sigma)^2 on,
// numpy's pairwise sum of the exp terms, float32 result, float32 comparisons.  Equal scores
// are ordered as a stable ascending sort reversed (score desc, current position desc).
#include <math.h>

#include "common.cuh"

namespace pc {

constexpr int kNmsThreads = 256;

struct NmsArgs {
  const float* kpts;           // [P, K, 3]
  const float* area;           // [P]
  float* score;                // [P]
  const int32_t* image_offset;  // [I + 1]
  int32_t* keep;               // [P]
  int32_t* num_keep;           // [I]
  int32_t K, rescore, use_nms, soft, max_dets, use_iou_vis, cap;
  int32_t use_matrix;  // small images: all pairwise OKS values up front, in shared memory
  float rescore_vis_thr, oks_thr, iou_vis_thr;
  double key_vars[PC_MAX_JOINTS];  // (2 sigma)^2
};

// numpy's float64 add.reduce over n <= 64 contiguous values (pairwise_sum)
__device__ double np_sum_f64_local(const double* a, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
    return r;
  }
  double r[8];
  for (int t = 0; t < 8; ++t) r[t] = a[t];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int t = 0; t < 8; ++t) r[t] = __dadd_rn(r[t], a[i + t]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}

// oks_iou(g, d) for one detection d (nms.py:52-68)
__device__ float oks_pair(const NmsArgs& a, const float* __restrict__ g,
                          const float* __restrict__ d, float area_g, float area_d) {
  double ex[PC_MAX_JOINTS];
  // (a_g + a_d) / 2 in float32, + np.spacing(1) in float64
  const double area =
      __dadd_rn((double)__fmul_rn(__fadd_rn(area_g, area_d), 0.5f), 2.220446049250313e-16);
  int m = 0;
  for (int k = 0; k < a.K; ++k) {
    const float dx = __fsub_rn(__ldg(d + 3 * k), __ldg(g + 3 * k));
    const float dy = __fsub_rn(__ldg(d + 3 * k + 1), __ldg(g + 3 * k + 1));
    const float sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (a.use_iou_vis && !(__ldg(d + 3 * k + 2) > a.iou_vis_thr)) continue;
    const double e = __ddiv_rn(__ddiv_rn((double)sq, a.key_vars[k]), area) * 0.5;
    ex[m++] = exp(-e);
  }
  if (m == 0) return 0.f;
  return (float)__ddiv_rn(np_sum_f64_local(ex, m), (double)m);
}

// dst order = stable ascending argsort of sc[0..n) reversed: rank by counting
__device__ __forceinline__ void rank_sort(const float* sc, const int* ord, float* sc_out,
                                          int* ord_out, int n, int tid) {
  for (int j = tid; j < n; j += blockDim.x) {
    const float s = sc[j];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const float o = sc[k];
      rank += (o > s || (o == s && k > j)) ? 1 : 0;
    }
    sc_out[rank] = s;
    ord_out[rank] = ord[j];
  }
}

__global__ void __launch_bounds__(kNmsThreads) oks_nms_kernel(const NmsArgs a) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  const int cap = a.cap;
  float* sc0 = reinterpret_cast<float*>(nms_smem);
  float* sc1 = sc0 + cap;
  int* ord0 = reinterpret_cast<int*>(sc1 + cap);
  int* ord1 = ord0 + cap;
  int* flag = ord1 + cap;
  float* mat = reinterpret_cast<float*>(flag + cap);  // [n][n] when use_matrix
  __shared__ int s_nkeep;

  const int tid = threadIdx.x;
  const int img = blockIdx.x;
  const int begin = a.image_offset[img];
  const int n = a.image_offset[img + 1] - begin;
  const float* kp = a.kpts + (size_t)begin * a.K * 3;
  const float* area = a.area + begin;
  float* score = a.score + begin;
  int32_t* keep = a.keep + begin;
  if (n <= 0) {
    if (tid == 0) a.num_keep[img] = 0;
    return;
  }
  if (n > cap) {  // the host promised max_people_per_image: flag and leave
    if (tid == 0) a.num_keep[img] = -1;
    return;
  }

  // ---- rescoring: mean of the joint scores above vis_thr, times the box score --------
  for (int p = tid; p < n; p += blockDim.x) {
    float s = score[p];
    if (a.rescore) {
      float acc = 0.f;
      int cnt = 0;
      for (int k = 0; k < a.K; ++k) {
        const float t = __ldg(kp + ((size_t)p * a.K + k) * 3 + 2);
        if (t > a.rescore_vis_thr) {
          acc = __fadd_rn(acc, t);
          ++cnt;
        }
      }
      if (cnt) acc = __fdiv_rn(acc, (float)cnt);
      s = __fmul_rn(acc, s);
      score[p] = s;
    }
    sc1[p] = s;
    ord1[p] = p;
    flag[p] = 0;
    keep[p] = -1;
  }
  if (tid == 0) s_nkeep = 0;
  __syncthreads();
  if (!a.use_nms) {
    for (int p = tid; p < n; p += blockDim.x) keep[p] = p;
    if (tid == 0) a.num_keep[img] = n;
    return;
  }
  rank_sort(sc1, ord1, sc0, ord0, n, tid);
  __syncthreads();

  const size_t stride = (size_t)a.K * 3;
  // The greedy loops below are chains of barriers with one OKS evaluation (K float64 exps)
  // per thread and step.  For small images every ordered pair (g = i, d = j) is evaluated
  // up front instead, all threads busy, and the loops only look values up.
  if (a.use_matrix) {
    for (int e = tid; e < n * n; e += blockDim.x) {
      const int i = e / n, j = e - i * n;
      if (i != j) mat[e] = oks_pair(a, kp + i * stride, kp + j * stride, area[i], area[j]);
    }
    __syncthreads();
  }
  auto oks = [&](int i, int j, float ai) {
    return a.use_matrix ? mat[i * n + j]
                        : oks_pair(a, kp + i * stride, kp + j * stride, ai, area[j]);
  };
  if (!a.soft) {
    // ---- oks_nms (nms.py:72-111): walk the order, drop what overlaps a kept person ----
    for (int t = 0; t < n; ++t) {
      if (flag[t]) continue;  // block-uniform (shared memory, read after a barrier)
      const int i = ord0[t];
      if (tid == 0) keep[s_nkeep++] = i;
      const float ai = area[i];
      for (int u = t + 1 + tid; u < n; u += blockDim.x) {
        if (flag[u]) continue;
        const int j = ord0[u];
        const float ov = oks(i, j, ai);
        if (!(ov <= a.oks_thr)) flag[u] = 1;
      }
      __syncthreads();
    }
  } else {
    // ---- soft_oks_nms (nms.py:141-190): rescore the rest, re-sort, repeat ---------------
    float* sc_cur = sc0;
    float* sc_nxt = sc1;
    int* ord_cur = ord0;
    int* ord_nxt = ord1;
    int* ord_tmp = flag;  // unsorted order of the rescored rest
    int cur = n, kept = 0;
    while (cur > 0 && kept < a.max_dets) {
      const int i = ord_cur[0];
      if (tid == 0) keep[kept] = i;
      ++kept;
      const float ai = area[i];
      for (int u = 1 + tid; u < cur; u += blockDim.x) {
        const int j = ord_cur[u];
        const float ov = oks(i, j, ai);
        // scores * np.exp(-(overlap**2) / thr), float32
        const float x = __fdiv_rn(-__fmul_rn(ov, ov), a.oks_thr);
        sc_nxt[u - 1] = __fmul_rn(sc_cur[u], (float)exp((double)x));
        ord_tmp[u - 1] = j;
      }
      __syncthreads();
      --cur;
      rank_sort(sc_nxt, ord_tmp, sc_cur, ord_nxt, cur, tid);
      __syncthreads();
      int* t2 = ord_cur;
      ord_cur = ord_nxt;
      ord_nxt = t2;
    }
    if (tid == 0) s_nkeep = kept;
  }
  __syncthreads();
  if (tid == 0) a.num_keep[img] = s_nkeep;
}

}  // namespace pc

using namespace pc;

extern "C" int pc_oks_nms(const float* d_kpts, const float* d_area, float* d_score,
                          const int32_t* d_image_offset, int32_t* d_keep, int32_t* d_num_keep,
                          const pc_oks_nms_params* p, int64_t num_images, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_oks_nms: params is NULL");
  PC_REQUIRE(num_images >= 0 && num_images < 0x7fffffffLL, PC_ERR_INVALID_ARGUMENT,
             "pc_oks_nms: bad image count");
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_oks_nms: num_joints %d outside [1, %d]", p->num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(p->max_people_per_image >= 0 && p->max_people_per_image <= PC_NMS_MAX_PEOPLE,
             PC_ERR_UNSUPPORTED, "pc_oks_nms: max_people_per_image %d outside [0, %d]",
             p->max_people_per_image, PC_NMS_MAX_PEOPLE);
  PC_REQUIRE(!p->soft || p->max_dets >= 0, PC_ERR_INVALID_ARGUMENT, "pc_oks_nms: max_dets < 0");
  PC_REQUIRE(!(p->use_nms && p->soft) || p->oks_thr != 0.f, PC_ERR_INVALID_ARGUMENT,
             "pc_oks_nms: soft NMS divides by oks_thr, which is 0");
  if (num_images == 0) return PC_OK;
  PC_REQUIRE(d_kpts && d_area && d_score && d_image_offset && d_keep && d_num_keep,
             PC_ERR_INVALID_ARGUMENT, "pc_oks_nms: NULL tensor pointer");
  NmsArgs a;
  a.kpts = d_kpts;
  a.area = d_area;
  a.score = d_score;
  a.image_offset = d_image_offset;
  a.keep = d_keep;
  a.num_keep = d_num_keep;
  a.K = p->num_joints;
  a.rescore = p->rescore;
  a.use_nms = p->use_nms;
  a.soft = p->soft;
  a.max_dets = p->max_dets;
  a.use_iou_vis = p->use_iou_vis_thr;
  a.cap = p->max_people_per_image > 0 ? p->max_people_per_image : 1;
  a.rescore_vis_thr = p->rescore_vis_thr;
  a.oks_thr = p->oks_thr;
  a.iou_vis_thr = p->iou_vis_thr;
  for (int k = 0; k < p->num_joints; ++k) {
    const double s2 = p->sigmas[k] * 2;  // key_vars = (sigmas * 2) ** 2
    a.key_vars[k] = s2 * s2;
  }
  // images of up to 96 people (36 KB of pairwise values): matrix variant with full CTAs;
  // else one thread per person up to 256, so that small CTAs share an SM
  a.use_matrix = p->use_nms && a.cap <= 96;
  const size_t smem = ((size_t)a.cap * 5 + (a.use_matrix ? (size_t)a.cap * a.cap : 0)) * sizeof(float);
  int threads = kNmsThreads;
  if (!a.use_matrix) {
    threads = ((a.cap + 31) / 32) * 32;
    if (threads > kNmsThreads) threads = kNmsThreads;
  }
  oks_nms_kernel<<<(unsigned)num_images, threads, smem, (cudaStream_t)stream>>>(a);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
