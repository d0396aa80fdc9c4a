// grouping.cu -- associative-embedding tag grouping on the device (sm_100a).
//
// Replaces match_by_tag (mindpose/utils/match.py:14-116), the instance score of
// BottomUpHeatMapAEInferencer._parse (engine/inferencer/bottomup_inferencer.py:
// 153-156) and transform_keypoints (data/transform/utils.py:235-274).
//
// The algorithm is a 17-step greedy over joints with a rectangular assignment
// problem per step; it is serial per image and latency bound (8 KB of input per
// image), so one warp owns one image: the image's groups, tag lists, cost matrix
// and the assignment state live in shared memory, the column scans of the
// shortest-augmenting-path solver and the per-group reductions run across the 32
// lanes, and the order-dependent parts (dict-key collisions, augmentation) are
// executed by lane 0.  The assignment solver is scipy's rectangular LSAP
// (Crouse 2016) including its scan order and tie rule, because with rounded
// norms the cost matrix is integer valued and the optimum is not unique.
#include <math.h>

#include "common.cuh"

namespace pc {

// (the group capacity G of an image is a run-time argument: pc_group_params.max_groups)
// detections per joint: shared memory is carved for 32, or for 64 when max_num > 32
constexpr int kMaxDetCap = PC_MAX_DETECTIONS;

struct GroupArgs {
  const float* val_k;
  const float* tag_k;
  const float* ind_k;
  float* ans;
  int32_t* num_groups;
  float* scores;
  int32_t K, M;
  int32_t G;  // people an image can hold (capacity of ans / scores per image)
  float vis_thr, tag_thr;
  int32_t ignore_too_much, use_rounded_norm;
};

struct GroupTables {
  int32_t joint_order[PC_MAX_JOINTS];
};

// numpy's float32 add.reduce over n <= 64 contiguous values (pairwise_sum)
__device__ float np_sum_f32(const float* a, int n, int stride) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i * stride]);
    return r;
  }
  float r[8];
  for (int t = 0; t < 8; ++t) r[t] = a[t * stride];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int t = 0; t < 8; ++t) r[t] = __fadd_rn(r[t], a[(i + t) * stride]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i * stride]);
  return res;
}

// Warp reductions of the solver's inner loop, as single REDUX instructions instead of five
// shuffle steps each (the loop is a chain of dependent reductions: their latency is the
// run time of the kernel).  The float64 minimum goes through the usual order-preserving map
// to a signed 64-bit key (no NaNs here), reduced as a signed high word and an unsigned low
// word among the lanes that hold the minimal high word.
__device__ __forceinline__ double warp_min_f64(double v) {
  long long b = __double_as_longlong(v);
  b ^= (b >> 63) & 0x7fffffffffffffffLL;  // negative values: reverse their order
  const int hi = (int)(b >> 32);
  const unsigned lo = (unsigned)b;
  const int mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  long long m = ((long long)mhi << 32) | (long long)mlo;
  m ^= (m >> 63) & 0x7fffffffffffffffLL;
  return __longlong_as_double(m);
}
__device__ __forceinline__ int warp_min_i32(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_max_i32(int v) { return __reduce_max_sync(0xffffffffu, v); }

struct LsapState {
  double* u;         // [kMaxDet]
  double* v;         // [kG]
  double* shortest;  // [kG]
  int* path;         // [kG]
  int* col4row;      // [kMaxDet]
  int* row4col;      // [kG]
  int* remaining;    // [kG]
  int* sr;           // [kMaxDet]
  int* sc;           // [kG]
};

// Warp-cooperative rectangular LSAP (nr <= nc); result in st.col4row[0..nr).
__device__ void lsap_warp(const float* cost, int ldc, int nr, int nc, LsapState st, int lane) {
  for (int j = lane; j < nc; j += 32) {
    st.v[j] = 0.0;
    st.row4col[j] = -1;
  }
  for (int r = lane; r < nr; r += 32) {
    st.u[r] = 0.0;
    st.col4row[r] = -1;
  }
  __syncwarp();
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  for (int cur = 0; cur < nr; ++cur) {
    for (int j = lane; j < nc; j += 32) {
      st.remaining[j] = nc - j - 1;
      st.sc[j] = 0;
      st.shortest[j] = kInf;
    }
    for (int r = lane; r < nr; r += 32) st.sr[r] = 0;
    __syncwarp();
    int num_rem = nc, sink = -1, i = cur;
    double min_val = 0.0;
    while (sink == -1) {
      if (lane == 0) st.sr[i] = 1;
      const double ui = st.u[i];
      double best = kInf;
      int first = 0x7fffffff, unassigned = -1;
      for (int it = lane; it < num_rem; it += 32) {
        const int j = st.remaining[it];
        const double r =
            __dsub_rn(__dsub_rn(__dadd_rn(min_val, (double)cost[i * ldc + j]), ui), st.v[j]);
        double sj = st.shortest[j];
        if (r < sj) {
          st.path[j] = i;
          st.shortest[j] = r;
          sj = r;
        }
        const bool open = st.row4col[j] == -1;
        if (sj < best) {
          best = sj;
          first = it;
          unassigned = open ? it : -1;
        } else if (sj == best && open) {
          unassigned = it;
        }
      }
      const double lowest = warp_min_f64(best);
      const int f = warp_min_i32(best == lowest ? first : 0x7fffffff);
      const int un = warp_max_i32(best == lowest ? unassigned : -1);
      const int index = un >= 0 ? un : f;
      min_val = lowest;
      const int j = st.remaining[index];
      if (st.row4col[j] == -1)
        sink = j;
      else
        i = st.row4col[j];
      --num_rem;
      __syncwarp();
      if (lane == 0) {
        st.sc[j] = 1;
        st.remaining[index] = st.remaining[num_rem];
      }
      __syncwarp();
    }
    // dual variables
    if (lane == 0) st.u[cur] = __dadd_rn(st.u[cur], min_val);
    for (int r = lane; r < nr; r += 32)
      if (st.sr[r] && r != cur)
        st.u[r] = __dadd_rn(st.u[r], __dsub_rn(min_val, st.shortest[st.col4row[r]]));
    for (int j = lane; j < nc; j += 32)
      if (st.sc[j]) st.v[j] = __dsub_rn(st.v[j], __dsub_rn(min_val, st.shortest[j]));
    __syncwarp();
    // augment along the path
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = st.path[j];
        st.row4col[j] = r;
        const int prev = st.col4row[r];
        st.col4row[r] = j;
        j = prev;
        if (r == cur) break;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32)
    group_by_tag_kernel(const GroupArgs a, const __grid_constant__ GroupTables tab) {
  extern __shared__ __align__(16) unsigned char g_smem[];
  const int lane = threadIdx.x;
  const int K = a.K, M = a.M, kG = a.G;
  const int kMaxDet = M <= 32 ? 32 : kMaxDetCap;
  // ---- shared-memory carve-up
  double* s_v = reinterpret_cast<double*>(g_smem);
  double* s_short = s_v + kG;
  double* s_u = s_short + kG;
  float* s_key = reinterpret_cast<float*>(s_u + kMaxDet);
  float* s_ref = s_key + kG;
  float* s_cost = s_ref + kG;            // [kMaxDet][kG]
  float* s_tags = s_cost + kMaxDet * kG;  // [kG][K]
  float* s_det = s_tags + kG * K;         // [4][kMaxDet]: x, y, val, tag
  int* s_ntag = reinterpret_cast<int*>(s_det + 4 * kMaxDet);
  int* s_path = s_ntag + kG;
  int* s_row4col = s_path + kG;
  int* s_remaining = s_row4col + kG;
  int* s_sc = s_remaining + kG;
  int* s_col4row = s_sc + kG;
  int* s_sr = s_col4row + kMaxDet;
  LsapState st = {s_u, s_v, s_short, s_path, s_col4row, s_row4col, s_remaining, s_sr, s_sc};

  const int64_t img = blockIdx.x;
  const float* val = a.val_k + img * K * M;
  const float* tag = a.tag_k + img * K * M;
  const float* ind = a.ind_k + img * K * M * 2;
  float* ans = a.ans + img * kG * K * 4;
  int ngroups = 0;
  bool overflow = false;

  // Group bookkeeping.  `open_or_overwrite` is the reference's
  //   key = tags[row, 0]; joint_dict[key][idx] = joints[row]; tag_dict[key] = [tags[row]]
  // with float keys: an equal key re-uses that group and resets its tag list.
  auto open_or_overwrite = [&](int d, int idx) {
    const float key = s_det[3 * kMaxDet + d];
    int found = -1;
    for (int g0 = 0; g0 < ngroups; g0 += 32) {
      const int g = g0 + lane;
      const unsigned hit = __ballot_sync(0xffffffffu, g < ngroups && s_key[g] == key);
      if (hit) {
        found = g0 + __ffs(hit) - 1;
        break;
      }
    }
    if (found < 0) {
      if (ngroups >= kG) {
        overflow = true;
        return;
      }
      found = ngroups++;
      for (int e = lane; e < K * 4; e += 32) ans[found * K * 4 + e] = 0.f;
      if (lane == 0) s_key[found] = key;
      __syncwarp();  // the zero fill must land before lane 0 writes the joint row
    }
    if (lane == 0) {
      float* row = ans + (found * K + idx) * 4;
      row[0] = s_det[d];
      row[1] = s_det[kMaxDet + d];
      row[2] = s_det[2 * kMaxDet + d];
      row[3] = key;
      s_ntag[found] = 1;
      s_tags[found * K] = key;
    }
    __syncwarp();
  };

  for (int step = 0; step < K && !overflow; ++step) {
    const int idx = tab.joint_order[step];
    // ---- detections of this joint with val > vis_thr, compacted in rank order
    int na = 0;
    __syncwarp();
    for (int m0 = 0; m0 < M; m0 += 32) {
      const int m = m0 + lane;
      float dv = 0.f, dt = 0.f, dx = 0.f, dy = 0.f;
      bool keep = false;
      if (m < M) {
        dv = __ldg(val + idx * M + m);
        dt = __ldg(tag + idx * M + m);
        dx = __ldg(ind + (idx * M + m) * 2);
        dy = __ldg(ind + (idx * M + m) * 2 + 1);
        keep = dv > a.vis_thr;
      }
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int d = na + __popc(km & ((1u << lane) - 1));
        s_det[d] = dx;
        s_det[kMaxDet + d] = dy;
        s_det[2 * kMaxDet + d] = dv;
        s_det[3 * kMaxDet + d] = dt;
      }
      na += __popc(km);
    }
    if (na == 0) continue;
    __syncwarp();

    if (step == 0 || ngroups == 0) {
      for (int d = 0; d < na && !overflow; ++d) open_or_overwrite(d, idx);
      continue;
    }
    const int ng = ngroups;
    if (a.ignore_too_much && ng == M) continue;

    // ---- reference tag of every group: float32 mean of its tag list
    for (int g = lane; g < ng; g += 32) {
      const int nt = s_ntag[g];
      s_ref[g] = __fdiv_rn(np_sum_f32(s_tags + g * K, nt, 1), (float)nt);
    }
    __syncwarp();
    // ---- cost matrix: |tag - ref| (sqrt of the float32 square), rounded; 1e10 padding
    const int nc = max(ng, na);
    for (int e = lane; e < na * nc; e += 32) {
      const int r = e / nc, c = e - r * nc;
      float cost = 1e10f;
      if (c < ng) {
        const float diff = __fsub_rn(s_det[3 * kMaxDet + r], s_ref[c]);
        cost = __fsqrt_rn(__fmul_rn(diff, diff));
        if (a.use_rounded_norm) cost = rintf(cost);
      }
      s_cost[r * kG + c] = cost;
    }
    __syncwarp();
    lsap_warp(s_cost, kG, na, nc, st, lane);

    // ---- apply the pairs in row order
    for (int r = 0; r < na && !overflow; ++r) {
      const int c = s_col4row[r];
      bool accept = false;
      if (c < ng) {
        const float diff = __fsub_rn(s_det[3 * kMaxDet + r], s_ref[c]);
        accept = __fsqrt_rn(__fmul_rn(diff, diff)) < a.tag_thr;
      }
      if (accept) {
        if (lane == 0) {
          float* row = ans + (c * K + idx) * 4;
          row[0] = s_det[r];
          row[1] = s_det[kMaxDet + r];
          row[2] = s_det[2 * kMaxDet + r];
          row[3] = s_det[3 * kMaxDet + r];
          s_tags[c * K + s_ntag[c]] = s_det[3 * kMaxDet + r];
          s_ntag[c] += 1;
        }
        __syncwarp();
      } else {
        open_or_overwrite(r, idx);
      }
    }
  }

  __syncwarp();
  __threadfence_block();
  if (overflow) {
    if (lane == 0) a.num_groups[img] = -1;
    return;
  }
  if (lane == 0) a.num_groups[img] = ngroups;
  // instance score: numpy float32 mean of the value column over all K joints
  for (int g = lane; g < ngroups; g += 32)
    a.scores[img * kG + g] = __fdiv_rn(np_sum_f32(ans + g * K * 4 + 2, K, 4), (float)K);
}

__global__ void transform_keypoints_kernel(float* __restrict__ ans,
                                           const int32_t* __restrict__ num_groups,
                                           const double* __restrict__ center,
                                           const double* __restrict__ scale,
                                           const double* __restrict__ hm_wh, double pixel_std,
                                           int K, int kG) {
  const int64_t img = blockIdx.x;
  const int p = num_groups[img];
  if (p <= 0) return;
  // scale = scale * pixel_std; sx = scale_w / heatmap_w  (float64, as numpy computes it)
  const double sw = __dmul_rn(scale[2 * img], pixel_std);
  const double sh = __dmul_rn(scale[2 * img + 1], pixel_std);
  const double kx = __ddiv_rn(sw, hm_wh[2 * img]);
  const double ky = __ddiv_rn(sh, hm_wh[2 * img + 1]);
  const double cx = center[2 * img], cy = center[2 * img + 1];
  float* base = ans + img * kG * K * 4;
  for (int e = threadIdx.x; e < p * K; e += blockDim.x) {
    float* row = base + e * 4;
    const double x = (double)row[0], y = (double)row[1];
    row[0] = (float)__dsub_rn(__dadd_rn(__dmul_rn(x, kx), cx), __dmul_rn(sw, 0.5));
    row[1] = (float)__dsub_rn(__dadd_rn(__dmul_rn(y, ky), cy), __dmul_rn(sh, 0.5));
  }
}

}  // namespace pc

using namespace pc;

extern "C" int pc_group_by_tag(const float* d_val_k, const float* d_tag_k, const float* d_ind_k,
                               float* d_ans, int32_t* d_num_groups, float* d_scores,
                               const pc_group_params* p, int64_t n, void* stream) {
  PC_REQUIRE(p != nullptr, PC_ERR_INVALID_ARGUMENT, "pc_group_by_tag: params is NULL");
  PC_REQUIRE(n >= 0, PC_ERR_INVALID_ARGUMENT, "pc_group_by_tag: n < 0");
  PC_REQUIRE(p->num_joints >= 1 && p->num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_group_by_tag: num_joints %d outside [1, %d]", p->num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(p->max_num >= 1 && p->max_num <= kMaxDetCap, PC_ERR_UNSUPPORTED,
             "pc_group_by_tag: max_num %d outside [1, %d]", p->max_num, kMaxDetCap);
  PC_REQUIRE(p->max_groups >= 0, PC_ERR_INVALID_ARGUMENT, "pc_group_by_tag: max_groups < 0");
  GroupTables tab;
  memset(&tab, 0, sizeof(tab));
  uint64_t seen = 0;
  for (int i = 0; i < p->num_joints; ++i) {
    const int j = p->joint_order[i];
    PC_REQUIRE(j >= 0 && j < p->num_joints && !((seen >> j) & 1), PC_ERR_INVALID_ARGUMENT,
               "pc_group_by_tag: joint_order must be a permutation of 0..%d", p->num_joints - 1);
    seen |= 1ull << j;
    tab.joint_order[i] = j;
  }
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_val_k && d_tag_k && d_ind_k && d_ans && d_num_groups && d_scores,
             PC_ERR_INVALID_ARGUMENT, "pc_group_by_tag: NULL tensor pointer");
  PC_REQUIRE(n < 0x7fffffffLL, PC_ERR_UNSUPPORTED, "pc_group_by_tag: batch too large");
  GroupArgs a;
  a.val_k = d_val_k;
  a.tag_k = d_tag_k;
  a.ind_k = d_ind_k;
  a.ans = d_ans;
  a.num_groups = d_num_groups;
  a.scores = d_scores;
  a.K = p->num_joints;
  a.M = p->max_num;
  a.vis_thr = p->vis_thr;
  a.tag_thr = p->tag_thr;
  a.ignore_too_much = p->ignore_too_much;
  a.use_rounded_norm = p->use_rounded_norm;
  const int kG = p->max_groups > 0 ? p->max_groups : PC_MAX_GROUPS;
  a.G = kG;
  const int kMaxDet = p->max_num <= 32 ? 32 : kMaxDetCap;
  const size_t smem = sizeof(double) * (2 * kG + kMaxDet) +
                      sizeof(float) * (2 * kG + kMaxDet * kG + (size_t)kG * a.K + 4 * kMaxDet) +
                      sizeof(int) * (5 * kG + 2 * kMaxDet);
  PC_REQUIRE(smem <= 227 * 1024, PC_ERR_UNSUPPORTED,
             "pc_group_by_tag: max_groups %d needs %zu bytes of shared memory (> 227 KB)", kG, smem);
  if (smem > 48 * 1024)
    PC_CUDA(cudaFuncSetAttribute(group_by_tag_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  group_by_tag_kernel<<<(unsigned)n, 32, smem, (cudaStream_t)stream>>>(a, tab);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_transform_keypoints(float* d_ans, const int32_t* d_num_groups,
                                      const double* d_center, const double* d_scale,
                                      const double* d_heatmap_wh, float pixel_std,
                                      int32_t num_joints, int32_t max_groups, int64_t n,
                                      void* stream) {
  PC_REQUIRE(n >= 0 && num_joints >= 1 && num_joints <= PC_MAX_JOINTS && max_groups >= 0,
             PC_ERR_INVALID_ARGUMENT, "pc_transform_keypoints: bad n / num_joints / max_groups");
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_ans && d_num_groups && d_center && d_scale && d_heatmap_wh,
             PC_ERR_INVALID_ARGUMENT, "pc_transform_keypoints: NULL tensor pointer");
  transform_keypoints_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(
      d_ans, d_num_groups, d_center, d_scale, d_heatmap_wh, (double)pixel_std, num_joints,
      max_groups > 0 ? max_groups : PC_MAX_GROUPS);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
