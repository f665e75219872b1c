// peer_exchange.cu — one-shot all-reduce(sum) of a short float64 statistics vector across the
// GPUs of one NVSwitch box, over peer memory (NVLink P2P stores), in ONE small kernel.
//
// The clustering passes exchange only O(K d^2) doubles (<= 72 KB), so the collective is
// latency-bound: an NCCL all-reduce costs ~15 us of launch + protocol per call, two per DEC step,
// against an ~80 us step.  Here every rank pushes its vector straight into a slot of every
// peer's exchange window (symmetric allocation), publishes a sequence flag with system-scope
// release semantics, waits for the world's flags and sums the slots in RANK ORDER — every rank
// gets the bit-identical result (replicas stay in lockstep), no ring, no staging copies.
//
// Window layout (identical on every rank; slots double-buffered by sequence parity so a fast rank
// can run at most one exchange ahead of the slowest without overwriting unread data):
//   uint32 seq; uint32 flags[2][16]; double slots[2][16][max_len]; uint64 ll_cells[2][16][2*min(max_len,1024)]
// Vectors of up to 1024 doubles (all DEC statistics) use the flag-in-data cells: no system fence and no
// flag round trip, see scc_common.cuh.
// One process per GPU: kernels of different ranks run on different devices, so the spin-wait is safe.
#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const double* __restrict__ local, int len, double* __restrict__ out, PeerCtx ex) {
    // Launched with the PDL attribute: this one-CTA kernel is already resident when the statistics kernel
    // before it drains, and the kernel after it may be scheduled (up to its own dependency wait) while
    // the exchange is in flight — the two launch latencies around the exchange leave the critical path.
    pdl_wait();
    pdl_trigger();
    // push my vector into slot [parity][rank] of every window (own window included), wait for the world's
    // vectors and sum them in rank order; short vectors travel flag-in-data (scc_common.cuh)
    const unsigned int seq = peer_push(ex, local, len);
    peer_pull(ex, out, len, seq);
}

// Second half of a fused exchange: the producing kernel's last CTA already pushed (grid_publish with a
// PeerCtx, mode 1); wait for the world and write the rank-ordered sum.
__global__ void __launch_bounds__(256)
peer_finish_kernel(double* __restrict__ out, int len, PeerCtx ex) {
    pdl_wait();
    pdl_trigger();
    peer_pull(ex, out, len);
}

// Push-only half for a rank whose shard is empty (its statistics are all zero but it must still take
// part in the exchange).
__global__ void __launch_bounds__(256)
peer_push_kernel(const double* __restrict__ local, int len, PeerCtx ex) {
    peer_push(ex, local, len);
}

// One-CTA launch with programmatic stream serialization allowed (see scc_common.cuh, PDL).
template <typename Kern, typename... Args>
static cudaError_t launch_one_cta_pdl(Kern kern, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

size_t peer_window_bytes(int max_len) {
    if (max_len < 1) return 0;
    return peer_ll_offset(max_len) + sizeof(unsigned long long) * 2 * kPeerMaxWorld * 2 * (size_t)peer_ll_len(max_len);
}

int peer_allreduce(const double* local, int len, double* out, void* const* windows_dev, int rank, int world,
                   int max_len, cudaStream_t st) {
    if (!local || !out || !windows_dev || len < 1 || len > max_len) return SCC_ERR_INVALID;
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return SCC_ERR_INVALID;
    PeerCtx ex{reinterpret_cast<unsigned char* const*>(windows_dev), rank, world, max_len};
    SCC_CUDA(launch_one_cta_pdl(peer_allreduce_kernel, st, local, len, out, ex));
    return SCC_OK;
}

int peer_push_only(const double* local, int len, void* const* windows_dev, int rank, int world, int max_len,
                   cudaStream_t st) {
    if (!local || !windows_dev || len < 1 || len > max_len) return SCC_ERR_INVALID;
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return SCC_ERR_INVALID;
    PeerCtx ex{reinterpret_cast<unsigned char* const*>(windows_dev), rank, world, max_len};
    peer_push_kernel<<<1, 256, 0, st>>>(local, len, ex);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

int peer_finish(double* out, int len, void* const* windows_dev, int rank, int world, int max_len, cudaStream_t st) {
    if (!out || !windows_dev || len < 1 || len > max_len) return SCC_ERR_INVALID;
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return SCC_ERR_INVALID;
    PeerCtx ex{reinterpret_cast<unsigned char* const*>(windows_dev), rank, world, max_len};
    SCC_CUDA(launch_one_cta_pdl(peer_finish_kernel, st, out, len, ex));
    return SCC_OK;
}

}  // namespace scc
