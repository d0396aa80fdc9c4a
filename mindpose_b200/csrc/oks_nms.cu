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
// chain), in the precision numpy gives the caller's data:
//   T = float  (records holding float32 ndarrays): float32 differences and squares, float64
//              from the division by (2 sigma)^2 on;
//   T = double (records holding Python floats -- what the reference's inferencer emits,
//              `pred.tolist()` / `box.tolist()`, topdown_inferencer.py:135-140): float64
//              rescoring, differences, squares, areas and sort keys.
// Both: numpy's pairwise sum of the exp terms, float32 OKS values, float32 comparisons.  Equal
// scores are ordered as a stable ascending sort reversed (score desc, current position desc).
#include <math.h>

#include "common.cuh"

namespace pc {

constexpr int kNmsThreads = 256;

template <typename T>
struct NmsArgs {
  const T* kpts;               // [P, K, 3]
  const T* area;               // [P]
  T* score;                    // [P]
  const int32_t* image_offset;  // [I + 1]
  int32_t* keep;               // [P]
  int32_t* num_keep;           // [I]
  int32_t K, rescore, use_nms, soft, max_dets, use_iou_vis, cap;
  int32_t use_matrix;  // small images: all pairwise OKS values up front, in shared memory
  float oks_thr;                   // compared with float32 OKS values in both precisions
  T rescore_vis_thr, iou_vis_thr;  // compared with values of the caller's precision
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

// the caller's precision, rounding by rounding (no contraction)
__device__ __forceinline__ float r_add(float x, float y) { return __fadd_rn(x, y); }
__device__ __forceinline__ double r_add(double x, double y) { return __dadd_rn(x, y); }
__device__ __forceinline__ float r_sub(float x, float y) { return __fsub_rn(x, y); }
__device__ __forceinline__ double r_sub(double x, double y) { return __dsub_rn(x, y); }
__device__ __forceinline__ float r_mul(float x, float y) { return __fmul_rn(x, y); }
__device__ __forceinline__ double r_mul(double x, double y) { return __dmul_rn(x, y); }
__device__ __forceinline__ float r_div(float x, float y) { return __fdiv_rn(x, y); }
__device__ __forceinline__ double r_div(double x, double y) { return __ddiv_rn(x, y); }

// oks_iou(g, d) for one detection d (nms.py:52-68)
template <typename T>
__device__ float oks_pair(const NmsArgs<T>& a, const T* __restrict__ g, const T* __restrict__ d,
                          T area_g, T area_d) {
  double ex[PC_MAX_JOINTS];
  // (a_g + a_d) / 2 in the caller's precision, + np.spacing(1) in float64
  const double area =
      __dadd_rn((double)r_mul(r_add(area_g, area_d), (T)0.5), 2.220446049250313e-16);
  int m = 0;
  for (int k = 0; k < a.K; ++k) {
    const T dx = r_sub(__ldg(d + 3 * k), __ldg(g + 3 * k));
    const T dy = r_sub(__ldg(d + 3 * k + 1), __ldg(g + 3 * k + 1));
    const T sq = r_add(r_mul(dx, dx), r_mul(dy, dy));
    // float32 data: compared with float32(vis_thr); Python floats: with the float64 value
    if (a.use_iou_vis && !(__ldg(d + 3 * k + 2) > a.iou_vis_thr)) continue;
    const double e = __ddiv_rn(__ddiv_rn((double)sq, a.key_vars[k]), area) * 0.5;
    ex[m++] = exp(-e);
  }
  if (m == 0) return 0.f;
  return (float)__ddiv_rn(np_sum_f64_local(ex, m), (double)m);
}

// dst order = stable ascending argsort of sc[0..n) reversed: rank by counting
template <typename T>
__device__ __forceinline__ void rank_sort(const T* sc, const int* ord, T* sc_out, int* ord_out,
                                          int n, int tid) {
  for (int j = tid; j < n; j += blockDim.x) {
    const T s = sc[j];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const T o = sc[k];
      rank += (o > s || (o == s && k > j)) ? 1 : 0;
    }
    sc_out[rank] = s;
    ord_out[rank] = ord[j];
  }
}

template <typename T>
__global__ void __launch_bounds__(kNmsThreads) oks_nms_kernel(const NmsArgs<T> a) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  const int cap = a.cap;
  T* sc0 = reinterpret_cast<T*>(nms_smem);
  T* sc1 = sc0 + cap;
  int* ord0 = reinterpret_cast<int*>(sc1 + cap);
  int* ord1 = ord0 + cap;
  int* flag = ord1 + cap;
  float* mat = reinterpret_cast<float*>(flag + cap);  // [n][n] when use_matrix
  __shared__ int s_nkeep;

  const int tid = threadIdx.x;
  const int img = blockIdx.x;
  const int begin = a.image_offset[img];
  const int n = a.image_offset[img + 1] - begin;
  const T* kp = a.kpts + (size_t)begin * a.K * 3;
  const T* area = a.area + begin;
  T* score = a.score + begin;
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
    T s = score[p];
    if (a.rescore) {
      T acc = 0;
      int cnt = 0;
      for (int k = 0; k < a.K; ++k) {
        const T t = __ldg(kp + ((size_t)p * a.K + k) * 3 + 2);
        if (t > a.rescore_vis_thr) {
          acc = r_add(acc, t);
          ++cnt;
        }
      }
      if (cnt) acc = r_div(acc, (T)cnt);
      s = r_mul(acc, s);
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
  auto oks = [&](int i, int j, T ai) {
    return a.use_matrix ? mat[i * n + j]
                        : oks_pair(a, kp + i * stride, kp + j * stride, ai, area[j]);
  };
  if (!a.soft) {
    // ---- oks_nms (nms.py:72-111): walk the order, drop what overlaps a kept person ----
    for (int t = 0; t < n; ++t) {
      if (flag[t]) continue;  // block-uniform (shared memory, read after a barrier)
      const int i = ord0[t];
      if (tid == 0) keep[s_nkeep++] = i;
      const T ai = area[i];
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
    T* sc_cur = sc0;
    T* sc_nxt = sc1;
    int* ord_cur = ord0;
    int* ord_nxt = ord1;
    int* ord_tmp = flag;  // unsorted order of the rescored rest
    int cur = n, kept = 0;
    while (cur > 0 && kept < a.max_dets) {
      const int i = ord_cur[0];
      if (tid == 0) keep[kept] = i;
      ++kept;
      const T ai = area[i];
      for (int u = 1 + tid; u < cur; u += blockDim.x) {
        const int j = ord_cur[u];
        const float ov = oks(i, j, ai);
        // scores * np.exp(-(overlap**2) / thr): the weight is float32 (overlap is), the
        // product is in the precision of the scores
        const float x = __fdiv_rn(-__fmul_rn(ov, ov), a.oks_thr);
        sc_nxt[u - 1] = r_mul(sc_cur[u], (T)(float)exp((double)x));
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

template <typename T>
static int oks_nms_launch(const char* who, const T* d_kpts, const T* d_area, T* d_score,
                          const int32_t* d_image_offset, int32_t* d_keep, int32_t* d_num_keep,
                          const pc_oks_nms_params* p, int64_t num_images, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "%s: params is NULL", who);
  PC_REQUIRE(num_images >= 0 && num_images < 0x7fffffffLL, PC_ERR_INVALID_ARGUMENT,
             "%s: bad image count", who);
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "%s: num_joints %d outside [1, %d]", who, p->num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(p->max_people_per_image >= 0 && p->max_people_per_image <= PC_NMS_MAX_PEOPLE,
             PC_ERR_UNSUPPORTED, "%s: max_people_per_image %d outside [0, %d]", who,
             p->max_people_per_image, PC_NMS_MAX_PEOPLE);
  PC_REQUIRE(!p->soft || p->max_dets >= 0, PC_ERR_INVALID_ARGUMENT, "%s: max_dets < 0", who);
  PC_REQUIRE(!(p->use_nms && p->soft) || p->oks_thr != 0.f, PC_ERR_INVALID_ARGUMENT,
             "%s: soft NMS divides by oks_thr, which is 0", who);
  if (num_images == 0) return PC_OK;
  PC_REQUIRE(d_kpts && d_area && d_score && d_image_offset && d_keep && d_num_keep,
             PC_ERR_INVALID_ARGUMENT, "%s: NULL tensor pointer", who);
  NmsArgs<T> a;
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
  a.oks_thr = p->oks_thr;
  if (sizeof(T) == sizeof(double)) {
    a.rescore_vis_thr = (T)p->rescore_vis_thr_f64;
    a.iou_vis_thr = (T)p->iou_vis_thr_f64;
  } else {
    a.rescore_vis_thr = (T)p->rescore_vis_thr;
    a.iou_vis_thr = (T)p->iou_vis_thr;
  }
  for (int k = 0; k < p->num_joints; ++k) {
    const double s2 = p->sigmas[k] * 2;  // key_vars = (sigmas * 2) ** 2
    a.key_vars[k] = s2 * s2;
  }
  // images of up to 96 people (36 KB of pairwise values): matrix variant with full CTAs;
  // else one thread per person up to 256, so that small CTAs share an SM
  a.use_matrix = p->use_nms && a.cap <= 96;
  // two score arrays of T (8-byte aligned at the front), then three int arrays, then the matrix
  const size_t smem = (size_t)a.cap * (2 * sizeof(T) + 3 * sizeof(int)) +
                      (a.use_matrix ? (size_t)a.cap * a.cap * sizeof(float) : 0);
  int threads = kNmsThreads;
  if (!a.use_matrix) {
    threads = ((a.cap + 31) / 32) * 32;
    if (threads > kNmsThreads) threads = kNmsThreads;
  }
  oks_nms_kernel<T><<<(unsigned)num_images, threads, smem, (cudaStream_t)stream>>>(a);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_oks_nms(const float* d_kpts, const float* d_area, float* d_score,
                          const int32_t* d_image_offset, int32_t* d_keep, int32_t* d_num_keep,
                          const pc_oks_nms_params* p, int64_t num_images, void* stream) {
  return oks_nms_launch<float>("pc_oks_nms", d_kpts, d_area, d_score, d_image_offset, d_keep,
                               d_num_keep, p, num_images, stream);
}

extern "C" int pc_oks_nms_f64(const double* d_kpts, const double* d_area, double* d_score,
                              const int32_t* d_image_offset, int32_t* d_keep,
                              int32_t* d_num_keep, const pc_oks_nms_params* p,
                              int64_t num_images, void* stream) {
  return oks_nms_launch<double>("pc_oks_nms_f64", d_kpts, d_area, d_score, d_image_offset,
                                d_keep, d_num_keep, p, num_images, stream);
}
