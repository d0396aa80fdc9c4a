// peer_gather.cu -- the keypoint all-gather as direct stores into peer memory (sm_100a).
//
// the only exchange of the codec is the all-gather of the decoded
// keypoints for COCO-style evaluation (228 B per crop at K = 17; the reference evaluates on
// rank 0 only, mindpose/callbacks/eval_callback.py:142-145).  An NCCL all-gather of such a
// small table is launch- and protocol-latency bound (0.2 ms per step at 8 GPUs, measured).
// Here the packing of (preds, boxes) into table rows and the exchange are ONE kernel: every
// rank stores its rows straight into the gathered table of every rank, through
//   * the NVSwitch multicast mapping of the table (NVLS; one multimem.st reaches all
//     replicas), when the allocation has one, or
//   * the peer-mapped address of the table on each rank (P2P stores over NVLink).
// The tables are symmetric-memory allocations made by the caller
// (mindpose_b200/dist.py::PeerGather, torch.distributed._symmetric_memory).
//
// Ordering between ranks, two forms:
//   * pc_scatter_results: the caller runs a barrier over the allocation's signal pads after
//     the kernel (one more launch, and every rank waits for the slowest one at once);
//   * pc_scatter_results_signal + pc_wait_peer_flags: the last CTA of the scatter kernel
//     publishes the step number in a flag word on every rank (release, system scope) and the
//     consumer waits for the flags of a step only when it reads that step's table -- one
//     step later in bench.py, so the wait is off the critical path of the step.
#include "common.cuh"

namespace pc {

constexpr int kMaxPeers = 16;

struct PeerTables {
  float* table[kMaxPeers];
};

struct PeerFlags {
  uint32_t* flags[kMaxPeers];  // flags[p]: the flag array (one word per source rank) on rank p
};

template <bool SIGNAL>
__global__ void __launch_bounds__(256)
    scatter_results_kernel(const float* __restrict__ preds, const float* __restrict__ boxes,
                           const PeerTables peers, int num_peers, float* multicast,
                           int64_t row_offset, int kw, int width, int64_t total,
                           const PeerFlags pf, int num_flag_peers, int my_rank,
                           uint32_t* __restrict__ step_word, int* __restrict__ counter) {
  auto value_at = [&](int64_t i) {
    const int64_t row = i / width;
    const int j = (int)(i - row * width);
    return j < kw ? preds[row * kw + j] : boxes[row * 6 + (j - kw)];
  };
  const int64_t base = row_offset * width;  // first element of this rank's rows in the table
  if (((base | total) & 3) == 0) {
    // four table elements (16 bytes, aligned) per thread and store
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < (total >> 2);
         q += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i = q << 2;
      const float4 v = make_float4(value_at(i), value_at(i + 1), value_at(i + 2), value_at(i + 3));
      if (multicast) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                         multicast + base + i),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
      } else {
        for (int p = 0; p < num_peers; ++p)
          *reinterpret_cast<float4*>(peers.table[p] + base + i) = v;
      }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
      const float v = value_at(i);
      if (multicast) {
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(multicast + base + i),
                     "f"(v)
                     : "memory");
      } else {
        for (int p = 0; p < num_peers; ++p) peers.table[p][base + i] = v;
      }
    }
  }
  // Make the remote stores visible system-wide before the kernel retires (the caller's
  // barrier kernel then signals the peers) or before the step is published: ONE system-scope
  // fence per CTA, by the thread that has synchronised with all the others (fences are
  // cumulative) -- a fence per thread serialises on the SM and cost 40 us per step.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (SIGNAL) {
      const int done = atomicAdd(counter, 1);
      if (done == (int)gridDim.x - 1) {  // the last CTA publishes the step on every rank
        *counter = 0;  // for the next launch (stream order)
        // the step number lives in device memory, so that a captured CUDA graph of the step
        // publishes a new number at every replay
        const uint32_t step = *step_word + 1u;
        *step_word = step;
        __threadfence_system();
        for (int p = 0; p < num_flag_peers; ++p)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[p] + my_rank),
                       "r"(step)
                       : "memory");
      }
    }
  }
}

// One thread per source rank: wait until that rank has published step `*step_word - lag` (or
// a later one) in this rank's flag array: lag 0 is the scatter this rank issued last before
// this kernel in stream order, lag 1 the one before.  The awaited kernels run on OTHER GPUs
// (one rank per GPU).
__global__ void wait_peer_flags_kernel(const uint32_t* __restrict__ flags, int num_peers,
                                       const uint32_t* __restrict__ step_word, uint32_t lag) {
  const int p = threadIdx.x;
  if (p >= num_peers) return;
  const uint32_t step = *step_word - lag;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory");
    if ((int32_t)(v - step) >= 0) break;
    __nanosleep(200);
    if (spins > (1u << 26)) __trap();  // > 13 s: a peer died; fail instead of hanging the GPU
  }
}

}  // namespace pc

using namespace pc;

static int scatter_launch(const float* d_preds, const float* d_boxes, void* const* h_peer_tables,
                          int32_t num_peers, void* d_multicast_table, int64_t row_offset,
                          int32_t num_joints, int64_t n, void* const* h_peer_flags,
                          int32_t num_flag_peers, int32_t my_rank, uint32_t* d_step,
                          int32_t* d_counter, void* stream) {
  const bool signal = h_peer_flags != nullptr;
  PC_REQUIRE(n >= 0 && row_offset >= 0, PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results: negative n / row_offset");
  PC_REQUIRE(!signal || (num_flag_peers >= 1 && num_flag_peers <= kMaxPeers && my_rank >= 0 &&
                         my_rank < num_flag_peers && d_counter && d_step),
             PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results_signal: need 1..%d flag arrays, my_rank inside, the step word "
             "and a counter",
             kMaxPeers);
  PC_REQUIRE(num_joints >= 1 && num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results: num_joints %d outside [1, %d]", num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(d_multicast_table || (num_peers >= 1 && num_peers <= kMaxPeers && h_peer_tables),
             PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results: need a multicast table or 1..%d peer tables", kMaxPeers);
  PC_REQUIRE(n == 0 || (d_preds && d_boxes), PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results: NULL tensor pointer");
  if (n == 0 && !signal) return PC_OK;
  PeerTables pt;
  for (int p = 0; p < kMaxPeers; ++p) pt.table[p] = nullptr;
  if (!d_multicast_table)
    for (int p = 0; p < num_peers; ++p) {
      PC_REQUIRE(h_peer_tables[p] != nullptr, PC_ERR_INVALID_ARGUMENT,
                 "pc_scatter_results: peer table %d is NULL", p);
      pt.table[p] = static_cast<float*>(h_peer_tables[p]);
    }
  PeerFlags pf;
  for (int p = 0; p < kMaxPeers; ++p) pf.flags[p] = nullptr;
  if (signal)
    for (int p = 0; p < num_flag_peers; ++p) {
      PC_REQUIRE(h_peer_flags[p] != nullptr, PC_ERR_INVALID_ARGUMENT,
                 "pc_scatter_results_signal: flag array %d is NULL", p);
      pf.flags[p] = static_cast<uint32_t*>(h_peer_flags[p]);
    }
  const int kw = num_joints * 3, width = kw + 6;
  const int64_t total = n * width;
  int64_t blocks = (total / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;  // an empty shard still signals
  const int64_t cap = (int64_t)sm_count_cached() * 2;
  if (cap > 0 && blocks > cap) blocks = cap;
  if (signal)
    scatter_results_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_preds, d_boxes, pt, d_multicast_table ? 0 : num_peers,
        static_cast<float*>(d_multicast_table), row_offset, kw, width, total, pf,
        num_flag_peers, my_rank, d_step, d_counter);
  else
    scatter_results_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_preds, d_boxes, pt, d_multicast_table ? 0 : num_peers,
        static_cast<float*>(d_multicast_table), row_offset, kw, width, total, pf, 0, 0, nullptr,
        nullptr);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_scatter_results(const float* d_preds, const float* d_boxes,
                                  void* const* h_peer_tables, int32_t num_peers,
                                  void* d_multicast_table, int64_t row_offset,
                                  int32_t num_joints, int64_t n, void* stream) {
  return scatter_launch(d_preds, d_boxes, h_peer_tables, num_peers, d_multicast_table,
                        row_offset, num_joints, n, nullptr, 0, 0, nullptr, nullptr, stream);
}

extern "C" int pc_scatter_results_signal(const float* d_preds, const float* d_boxes,
                                         void* const* h_peer_tables, int32_t num_peers,
                                         void* d_multicast_table, int64_t row_offset,
                                         int32_t num_joints, int64_t n,
                                         void* const* h_peer_flags, int32_t num_flag_peers,
                                         int32_t my_rank, uint32_t* d_step, int32_t* d_counter,
                                         void* stream) {
  PC_REQUIRE(h_peer_flags != nullptr, PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results_signal: flag arrays are NULL");
  return scatter_launch(d_preds, d_boxes, h_peer_tables, num_peers, d_multicast_table,
                        row_offset, num_joints, n, h_peer_flags, num_flag_peers, my_rank, d_step,
                        d_counter, stream);
}

extern "C" int pc_wait_peer_flags(const uint32_t* d_flags, int32_t num_peers,
                                  const uint32_t* d_step, uint32_t lag, void* stream) {
  PC_REQUIRE(d_flags != nullptr && d_step != nullptr && num_peers >= 1 && num_peers <= kMaxPeers,
             PC_ERR_INVALID_ARGUMENT, "pc_wait_peer_flags: need flags of 1..%d ranks", kMaxPeers);
  wait_peer_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_flags, num_peers, d_step, lag);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
