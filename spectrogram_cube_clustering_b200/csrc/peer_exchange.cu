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
//   uint32 seq; uint32 flags[2][16]; double slots[2][16][max_len]
// One process per GPU: kernels of different ranks run on different devices, so the spin-wait is safe.
#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const double* __restrict__ local, int len, double* __restrict__ out,
                      unsigned char* const* __restrict__ windows, int rank, int world, int max_len) {
    // Launched with the PDL attribute: this one-CTA kernel is already resident when the statistics kernel
    // before it drains, and the kernel after it may be scheduled (up to its own dependency wait) while
    // the exchange is in flight — the two launch latencies around the exchange leave the critical path.
    pdl_wait();
    pdl_trigger();
    PeerHeader* me = reinterpret_cast<PeerHeader*>(windows[rank]);
    const unsigned int seq = me->seq + 1u;
    const int parity = seq & 1u;
    const size_t slot_stride = (size_t)max_len;
    // push my vector into slot [parity][rank] of every window (own window included)
    for (int p = 0; p < world; ++p) {
        double* dst = reinterpret_cast<double*>(windows[p] + kPeerHeaderBytes) +
                      ((size_t)parity * kPeerMaxWorld + rank) * slot_stride;
        for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = local[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        PeerHeader* peer = reinterpret_cast<PeerHeader*>(windows[threadIdx.x]);
        st_release_sys(&peer->flags[parity][rank], seq);
        while (ld_acquire_sys(&me->flags[parity][threadIdx.x]) != seq) {}
    }
    __syncthreads();
    const double* slots = reinterpret_cast<const double*>(windows[rank] + kPeerHeaderBytes) +
                          (size_t)parity * kPeerMaxWorld * slot_stride;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < world; ++r) acc += __ldcv(slots + (size_t)r * slot_stride + i);
        out[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) me->seq = seq;
}

// Second half of a fused exchange: the producing kernel's last CTA already pushed (grid_publish with a
// PeerCtx); wait for the world and write the rank-ordered sum.
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
    peer_push(ex, ((int)threadIdx.x < len) ? local[threadIdx.x] : 0.0, len);
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
    return kPeerHeaderBytes + sizeof(double) * 2 * kPeerMaxWorld * (size_t)max_len;
}

int peer_allreduce(const double* local, int len, double* out, void* const* windows_dev, int rank, int world,
                   int max_len, cudaStream_t st) {
    if (!local || !out || !windows_dev || len < 1 || len > max_len) return SCC_ERR_INVALID;
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return SCC_ERR_INVALID;
    SCC_CUDA(launch_one_cta_pdl(peer_allreduce_kernel, st, local, len, out,
                                reinterpret_cast<unsigned char* const*>(windows_dev), rank, world, max_len));
    return SCC_OK;
}

int peer_push_only(const double* local, int len, void* const* windows_dev, int rank, int world, int max_len,
                   cudaStream_t st) {
    if (!local || !windows_dev || len < 1 || len > 256 || len > max_len) return SCC_ERR_INVALID;
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
