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
// (mindpose_b200/dist.py::PeerGather, torch.distributed._symmetric_memory); a barrier over
// the same allocation's signal pads orders the ranks afterwards.
#include "common.cuh"

namespace pc {

constexpr int kMaxPeers = 16;

struct PeerTables {
  float* table[kMaxPeers];
};

__global__ void __launch_bounds__(256)
    scatter_results_kernel(const float* __restrict__ preds, const float* __restrict__ boxes,
                           const PeerTables peers, int num_peers, float* multicast,
                           int64_t row_offset, int kw, int width, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / width;
    const int j = (int)(i - row * width);
    const float v = j < kw ? preds[row * kw + j] : boxes[row * 6 + (j - kw)];
    const int64_t o = (row_offset + row) * width + j;
    if (multicast) {
      asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(multicast + o), "f"(v)
                   : "memory");
    } else {
      for (int p = 0; p < num_peers; ++p) peers.table[p][o] = v;
    }
  }
  // make the remote stores visible system-wide before the kernel retires (the caller's
  // barrier kernel then signals the peers)
  __threadfence_system();
}

}  // namespace pc

using namespace pc;

extern "C" int pc_scatter_results(const float* d_preds, const float* d_boxes,
                                  void* const* h_peer_tables, int32_t num_peers,
                                  void* d_multicast_table, int64_t row_offset,
                                  int32_t num_joints, int64_t n, void* stream) {
  PC_REQUIRE(n >= 0 && row_offset >= 0, PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results: negative n / row_offset");
  PC_REQUIRE(num_joints >= 1 && num_joints <= PC_MAX_JOINTS, PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results: num_joints %d outside [1, %d]", num_joints, PC_MAX_JOINTS);
  PC_REQUIRE(d_multicast_table || (num_peers >= 1 && num_peers <= kMaxPeers && h_peer_tables),
             PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results: need a multicast table or 1..%d peer tables", kMaxPeers);
  if (n == 0) return PC_OK;
  PC_REQUIRE(d_preds && d_boxes, PC_ERR_INVALID_ARGUMENT,
             "pc_scatter_results: NULL tensor pointer");
  PeerTables pt;
  for (int p = 0; p < kMaxPeers; ++p) pt.table[p] = nullptr;
  if (!d_multicast_table)
    for (int p = 0; p < num_peers; ++p) {
      PC_REQUIRE(h_peer_tables[p] != nullptr, PC_ERR_INVALID_ARGUMENT,
                 "pc_scatter_results: peer table %d is NULL", p);
      pt.table[p] = static_cast<float*>(h_peer_tables[p]);
    }
  const int kw = num_joints * 3, width = kw + 6;
  const int64_t total = n * width;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count_cached() * 8;
  if (cap > 0 && blocks > cap) blocks = cap;
  scatter_results_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      d_preds, d_boxes, pt, d_multicast_table ? 0 : num_peers,
      static_cast<float*>(d_multicast_table), row_offset, kw, width, total);
  PC_CUDA(cudaGetLastError());
  return PC_OK;
}
