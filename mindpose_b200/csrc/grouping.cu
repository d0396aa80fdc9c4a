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

// The same solver for nc <= 32 * S (S = 1: every image with at most 32 people; S = 2: up to
// 64), entirely in registers: lane l owns columns l + 32 s (v, shortest, path, row4col and the
// column's position in the solver's `remaining` list) and rows l + 32 s (u, col4row); the
// SR / SC sets are warp-uniform bit masks.  No shared-memory round trips and no __syncwarp in
// the search loop: its chain of dependent steps is the run time of the kernel.  Same
// arithmetic, same scan order and tie-breaks (the `remaining` list is permuted exactly as the
// array version permutes it), so the assignment is the same.
template <int S>
__device__ void lsap_warp_regs(const float* cost, int ldc, int nr, int nc, int* col4row_out,
                               int lane) {
  const unsigned full = 0xffffffffu;
  const double kInf = __longlong_as_double(0x7ff0000000000000LL);
  double u[S], v[S];
  int col4row[S], row4col[S];
#pragma unroll
  for (int s = 0; s < S; ++s) u[s] = v[s] = 0.0, col4row[s] = row4col[s] = -1;
  // value of a per-slot register of row / column x (x warp-uniform)
  auto at = [&](auto (&reg)[S], int x) {
    auto val = reg[0];
#pragma unroll
    for (int s = 1; s < S; ++s) val = (x >> 5) == s ? reg[s] : val;
    return __shfl_sync(full, val, x & 31);
  };
  for (int cur = 0; cur < nr; ++cur) {
    int pos[S], path[S];
    bool rem[S];
    double shortest[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      pos[s] = nc - 1 - (lane + 32 * s);  // remaining[it] = nc - it - 1
      rem[s] = lane + 32 * s < nc;
      shortest[s] = kInf;
      path[s] = -1;
    }
    unsigned long long SR = 0ull, SC = 0ull;
    int num_rem = nc, sink = -1, i = cur;
    double min_val = 0.0;
    while (sink == -1) {
      SR |= 1ull << i;
      const double ui = at(u, i);
      double sj = kInf;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (rem[s]) {
          const double r = __dsub_rn(
              __dsub_rn(__dadd_rn(min_val, (double)cost[i * ldc + lane + 32 * s]), ui), v[s]);
          if (r < shortest[s]) {
            path[s] = i;
            shortest[s] = r;
          }
          sj = fmin(sj, shortest[s]);
        }
      }
      const double lowest = warp_min_f64(sj);
      // sequential scan semantics: the first position that reaches the minimum, unless some
      // position at the minimum is an unassigned column -- then the last of those.  As ONE
      // maximum: (unassigned, position or its reverse, column) packed into a key.
      int key = -1;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (rem[s] && shortest[s] == lowest) {
          const int open = row4col[s] == -1;
          const int k = (open << 16) | ((open ? pos[s] : 127 - pos[s]) << 8) | (lane + 32 * s);
          key = max(key, k);
        }
      }
      key = warp_max_i32(key);
      const int jsel = key & 0xff;
      const int pe = (key >> 8) & 0xff;
      const int index = (key >> 16) ? pe : 127 - pe;
      min_val = lowest;
      if (key >> 16) {
        sink = jsel;
      } else {
        i = at(row4col, jsel);
      }
      --num_rem;
      SC |= 1ull << jsel;
      // remaining[index] = remaining[num_rem]
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (lane + 32 * s == jsel)
          rem[s] = false;
        else if (rem[s] && pos[s] == num_rem)
          pos[s] = index;
      }
    }
    // dual variables (before the augmentation: col4row is still the old assignment)
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int c = col4row[s] >= 0 ? col4row[s] : 0;
      double sh_c = __shfl_sync(full, shortest[0], c & 31);
#pragma unroll
      for (int t = 1; t < S; ++t) {
        const double o = __shfl_sync(full, shortest[t], c & 31);
        sh_c = (c >> 5) == t ? o : sh_c;
      }
      const int r = lane + 32 * s;
      if (r == cur)
        u[s] = __dadd_rn(u[s], min_val);
      else if ((SR >> r) & 1ull)
        u[s] = __dadd_rn(u[s], __dsub_rn(min_val, sh_c));
    }
#pragma unroll
    for (int s = 0; s < S; ++s)
      if ((SC >> (lane + 32 * s)) & 1ull) v[s] = __dsub_rn(v[s], __dsub_rn(min_val, shortest[s]));
    // augment along the path
    int j = sink;
    while (true) {
      const int r = at(path, j);
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (lane + 32 * s == j) row4col[s] = r;
      const int prev = at(col4row, r);
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (lane + 32 * s == r) col4row[s] = j;
      j = prev;
      if (r == cur) break;
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s)
    if (lane + 32 * s < nr) col4row_out[lane + 32 * s] = col4row[s];
  __syncwarp();
}

// np.linalg.norm over one tag channel: sqrt(d * d) in float32.  With a correctly rounded
// square root that is |d| exactly whenever d * d stays in the normal range (radix 2), so the
// software square root only runs outside it.
__device__ __forceinline__ float tag_norm(float d) {
  const float ad = fabsf(d);
  return (ad > 1e-18f && ad < 1e18f) ? ad : __fsqrt_rn(__fmul_rn(d, d));
}

#ifdef PC_GROUP_PROFILE
// cycles per phase, summed over the images (experiment builds only)
__device__ unsigned long long g_group_prof[8];
#define PC_PROF_MARK(slot)                                             \
  do {                                                                 \
    const long long _t = clock64();                                    \
    if (lane == 0) atomicAdd(&g_group_prof[slot], (unsigned long long)(_t - prof_t)); \
    prof_t = _t;                                                       \
  } while (0)
#else
#define PC_PROF_MARK(slot) \
  do {                     \
  } while (0)
#endif

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
#ifdef PC_GROUP_PROFILE
  long long prof_t = clock64();
#endif

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

  // The detections of a joint (lane l: ranks l and l + 32) are loaded one step ahead, so that
  // the round trip to L2 overlaps the previous joint's assignment problem.
  float pv[2], pt[2], px[2], py[2];
  auto fetch = [&](int step) {
    const int j = tab.joint_order[step];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      pv[h] = pt[h] = px[h] = py[h] = 0.f;
      if (m < M) {
        pv[h] = __ldg(val + j * M + m);
        pt[h] = __ldg(tag + j * M + m);
        px[h] = __ldg(ind + (j * M + m) * 2);
        py[h] = __ldg(ind + (j * M + m) * 2 + 1);
      }
    }
  };
  fetch(0);

  for (int step = 0; step < K && !overflow; ++step) {
    const int idx = tab.joint_order[step];
    // ---- detections of this joint with val > vis_thr, compacted in rank order
    int na = 0;
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool keep = lane + 32 * h < M && pv[h] > a.vis_thr;
      const unsigned km = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int d = na + __popc(km & ((1u << lane) - 1));
        s_det[d] = px[h];
        s_det[kMaxDet + d] = py[h];
        s_det[2 * kMaxDet + d] = pv[h];
        s_det[3 * kMaxDet + d] = pt[h];
      }
      na += __popc(km);
    }
    if (step + 1 < K) fetch(step + 1);
    if (na == 0) continue;
    __syncwarp();
    PC_PROF_MARK(0);

    if (step == 0 || ngroups == 0) {
      for (int d = 0; d < na && !overflow; ++d) open_or_overwrite(d, idx);
      PC_PROF_MARK(3);
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
    const int nc = max(ng, na);
    // ---- Quick solve.  If every detection's cheapest group is cheapest STRICTLY and no two
    // detections share it, that assignment is the unique optimum (any other one pays strictly
    // more in some row and never less in the others), so it is what scipy returns -- no
    // solver run, no cost matrix.  Well separated tags, the normal case, end here; ties (tags
    // that round to the same distance) fall through to the solver.  Only without padded
    // columns (na <= ng): next to 1e10 entries the solver's float64 sums no longer resolve
    // small gaps, and its answer is then the one to reproduce, optimal or not.
    bool solved = false;
    if (nc <= 32 && na <= ng) {
      const float ref = lane < ng ? s_ref[lane] : 0.f;
      unsigned taken = 0u;
      int mine = -1;
      bool ok = true;
      for (int r = 0; r < na; ++r) {
        float cost = tag_norm(__fsub_rn(s_det[3 * kMaxDet + r], ref));
        if (a.use_rounded_norm) cost = rintf(cost);
        // cost >= +0: integer order of the bits == float order
        const int key = lane < ng ? __float_as_int(cost) : 0x7fffffff;
        const int mn = warp_min_i32(key);
        const unsigned at = __ballot_sync(0xffffffffu, key == mn);
        if ((at & (at - 1u)) != 0u || (at & taken) != 0u) {  // (uniform) a tie, or a shared group
          ok = false;
          break;
        }
        taken |= at;
        if (lane == r) mine = __ffs(at) - 1;
      }
      if (ok) {
        if (lane < na) s_col4row[lane] = mine;
        __syncwarp();
        solved = true;
      }
    }
    if (!solved) {
      // ---- cost matrix: |tag - ref| (sqrt of the float32 square), rounded; 1e10 padding
      if (nc <= 64) {  // lane = column: no index division, conflict-free stores
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int c = lane + 32 * s;
          if (c < nc) {
            const float ref = c < ng ? s_ref[c] : 0.f;
            for (int r = 0; r < na; ++r) {
              float cost = 1e10f;
              if (c < ng) {
                cost = tag_norm(__fsub_rn(s_det[3 * kMaxDet + r], ref));
                if (a.use_rounded_norm) cost = rintf(cost);
              }
              s_cost[r * kG + c] = cost;
            }
          }
        }
      } else {
        for (int e = lane; e < na * nc; e += 32) {
          const int r = e / nc, c = e - r * nc;
          float cost = 1e10f;
          if (c < ng) {
            cost = tag_norm(__fsub_rn(s_det[3 * kMaxDet + r], s_ref[c]));
            if (a.use_rounded_norm) cost = rintf(cost);
          }
          s_cost[r * kG + c] = cost;
        }
      }
      __syncwarp();
      PC_PROF_MARK(1);
      if (nc <= 32)
        lsap_warp_regs<1>(s_cost, kG, na, nc, s_col4row, lane);
      else if (nc <= 64)
        lsap_warp_regs<2>(s_cost, kG, na, nc, s_col4row, lane);
      else
        lsap_warp(s_cost, kG, na, nc, st, lane);
    }
    PC_PROF_MARK(2);

    // ---- apply the pairs.  The reference walks them in row order; an accepted pair writes
    // the joint into its (old) group and appends the tag, a rejected one opens a group or
    // overwrites the group whose key equals its tag.  Accepted pairs touch distinct old
    // groups, rejected ones new groups -- unless a rejected tag EQUALS the key of an old
    // group, the one case in which the order between the two kinds matters.  So: accepted
    // pairs all at once (lane = row), then the rejected ones in row order; the rare key
    // collision takes the plain sequential walk.
    bool fast_apply = false;
    if (na <= 32) {
      const int r = lane;
      int c = -1;
      bool accept = false;
      float rt = 0.f;
      if (r < na) {
        c = s_col4row[r];
        rt = s_det[3 * kMaxDet + r];
        if (c < ng) {
          accept = tag_norm(__fsub_rn(rt, s_ref[c])) < a.tag_thr;
        }
      }
      const unsigned acc = __ballot_sync(0xffffffffu, accept);
      const unsigned rej = ~acc & (na == 32 ? 0xffffffffu : (1u << na) - 1u);
      bool collide = false;
      if (rej && r < na && !accept)
        for (int g = 0; g < ng; ++g) collide |= s_key[g] == rt;
      if (!__any_sync(0xffffffffu, collide)) {
        fast_apply = true;
        if (accept) {
          float* row = ans + (c * K + idx) * 4;
          row[0] = s_det[r];
          row[1] = s_det[kMaxDet + r];
          row[2] = s_det[2 * kMaxDet + r];
          row[3] = rt;
          s_tags[c * K + s_ntag[c]] = rt;
          s_ntag[c] += 1;
        }
        __syncwarp();
        for (unsigned m = rej; m && !overflow; m &= m - 1) open_or_overwrite(__ffs(m) - 1, idx);
      }
    }
    if (!fast_apply) {
      for (int r = 0; r < na && !overflow; ++r) {
        const int c = s_col4row[r];
        bool accept = false;
        if (c < ng) {
          accept = tag_norm(__fsub_rn(s_det[3 * kMaxDet + r], s_ref[c])) < a.tag_thr;
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
    PC_PROF_MARK(3);
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

#ifdef PC_GROUP_PROFILE
// experiment builds only (python -m mindpose_b200.csrc.build --variant ... with
// PC_NVCC_DEFINES=-DPC_GROUP_PROFILE=1): cycles per phase of group_by_tag_kernel
extern "C" int pc_group_profile(unsigned long long* out8, int reset) {
  PC_CUDA(cudaDeviceSynchronize());
  PC_CUDA(cudaMemcpyFromSymbol(out8, g_group_prof, sizeof(unsigned long long) * 8));
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    PC_CUDA(cudaMemcpyToSymbol(g_group_prof, z, sizeof(z)));
  }
  return PC_OK;
}
#endif
