// peer_exchange.cu — one-shot all-reduce(sum) of a short float64 statistics vector across the
// GPUs of one NVSwitch box, over peer memory (NVLink P2P stores), in ONE small kernel.
//
// The clustering passes exchange only O(K d^2) doubles (<= 72 KB), so the collective is
// latency-bound: an NCCL all-reduce costs ~17-32 us of launch + protocol per call, two per DEC step,
// against a ~50 us step.  Here every rank pushes its vector straight into the cells of every
// peer's exchange window (symmetric allocation) as {data, sequence number} words, polls its own
// window for the world's words and sums them in RANK ORDER — every rank gets the bit-identical
// result (replicas stay in lockstep), no ring, no staging copies, no system-scope fence.
//
// Window layout (identical on every rank; slots double-buffered by sequence parity so a fast rank
// can run at most one exchange ahead of the slowest without overwriting unread data):
//   uint32 seq, ticket; uint64 ll_cells[2][16][2*max_len]    (flag-in-data cells, see scc_common.cuh: no system
//   fence, no flag round trip; long vectors are exchanged slice-wise by several CTAs)
// One process per GPU: kernels of different ranks run on different devices, so the spin-wait is safe.
#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

constexpr int kExchangeSlice = 256;        // elements per CTA of the stand-alone exchange kernel

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const double* __restrict__ local, int len, double* __restrict__ out, PeerCtx ex) {
    // Launched with the PDL attribute: the kernel is already resident when the statistics kernel before it
    // drains, and the kernel after it may be scheduled (up to its own dependency wait) while the exchange is in
    // flight — the two launch latencies around the exchange leave the critical path.
    pdl_wait();
    pdl_trigger();
    // every CTA owns a slice: push it into cell [parity][rank] of every window (own window included), poll the
    // same slice of every rank and sum in rank order.  Flag-in-data needs no ordering between elements, so the
    // slices are independent; the sequence number is advanced by the CTA that finishes last.
    __shared__ double slice[kExchangeSlice];
    PeerHeader* me = reinterpret_cast<PeerHeader*>(ex.windows[ex.rank]);
    const unsigned int seq = ld_relaxed_gpu_u32(&me->seq) + 1u;
    const int lo = blockIdx.x * kExchangeSlice, hi = min(len, lo + kExchangeSlice);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) slice[i - lo] = local[i];
    __syncthreads();
    peer_push_slice(ex, slice, lo, hi, seq);
    peer_pull_slice(ex, out + lo, lo, hi, seq);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0u;
        if (gridDim.x > 1)
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(&me->ticket) : "memory");
        if (t == gridDim.x - 1) {                   // every CTA has read me->seq (it takes its ticket at the end)
            if (gridDim.x > 1) me->ticket = 0u;
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&me->seq), "r"(seq) : "memory");
        }
    }
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

// Small launch with programmatic stream serialization allowed (see scc_common.cuh, PDL).
template <typename Kern, typename... Args>
static cudaError_t launch_ctas_pdl(int ctas, Kern kern, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ctas);
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
    return kPeerHeaderBytes + sizeof(unsigned long long) * 2 * kPeerMaxWorld * 2 * (size_t)max_len;
}

int peer_allreduce(const double* local, int len, double* out, void* const* windows_dev, int rank, int world,
                   int max_len, cudaStream_t st) {
    if (!local || !out || !windows_dev || len < 1 || len > max_len) return SCC_ERR_INVALID;
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return SCC_ERR_INVALID;
    PeerCtx ex{reinterpret_cast<unsigned char* const*>(windows_dev), rank, world, max_len};
    SCC_CUDA(launch_ctas_pdl((len + kExchangeSlice - 1) / kExchangeSlice, peer_allreduce_kernel, st, local, len, out, ex));
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
    SCC_CUDA(launch_ctas_pdl(1, peer_finish_kernel, st, out, len, ex));
    return SCC_OK;
}

}  // namespace scc
