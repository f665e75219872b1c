// scc_common.cuh — shared device helpers for the sm_100a clustering kernels.
//
// Data movement model: every N-sized pass is a persistent grid
// of CTAs that stream TILE-point tiles of the row-major latent buffer z[n,d]
// through a shared-memory ring.  Dense row layouts (odd d, or d = 4 mod 8) are
// filled by one elected thread with a 1-D TMA bulk copy (cp.async.bulk ->
// UBLKCP) completing on an mbarrier; layouts that need padding to stay free of
// bank conflicts (d = 0 mod 8, even d) are filled cooperatively with coalesced
// 128-bit loads.  Each thread then owns one latent point (its row lives in
// registers), centroids / Cholesky factors are broadcast from shared memory,
// and per-cluster statistics are reduced warp -> CTA -> grid in a fixed order.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "scc_b200.h"

namespace scc {

constexpr int kWarp = 32;

// ----------------------------------------------------------------------------
// Shared-memory row layout of a tile of latent points.
// ----------------------------------------------------------------------------
template <int D>
struct RowLayout {
    static constexpr bool kVec4 = (D % 4 == 0);
    // Row stride (floats) chosen so that "lane i reads row i" is conflict free:
    //  * float4 reads: (LD/4) must be odd  -> d = 0 mod 8 gets 4 floats of padding
    //  * scalar reads: LD must be odd      -> even d (not multiple of 4) gets 1
    static constexpr int LD = kVec4 ? (((D / 4) % 2 == 1) ? D : D + 4) : ((D % 2 == 1) ? D : D + 1);
    static constexpr bool kDense = (LD == D);
};

template <int D>
__device__ __forceinline__ void load_row(const float* __restrict__ tile, int t, float (&r)[D]) {
    const float* p = tile + t * RowLayout<D>::LD;
    if constexpr (RowLayout<D>::kVec4) {
#pragma unroll
        for (int c = 0; c < D / 4; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(p + 4 * c);
            r[4 * c + 0] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) r[c] = p[c];
    }
}

template <int D>
__device__ __forceinline__ void store_row(float* __restrict__ tile, int t, const float (&r)[D]) {
    float* p = tile + t * RowLayout<D>::LD;
    if constexpr (RowLayout<D>::kVec4) {
#pragma unroll
        for (int c = 0; c < D / 4; ++c)
            *reinterpret_cast<float4*>(p + 4 * c) = make_float4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = r[c];
    }
}

// ----------------------------------------------------------------------------
// mbarrier + TMA bulk copy (1-D) wrappers.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared::cta bulk copy; dst/src 16-byte aligned, bytes multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared::cta -> global bulk store (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Ampere-style asynchronous 16-byte global -> shared copy (LDGSTS), used for the padded row layouts
// that a dense TMA bulk copy cannot produce.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ----------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The DEC kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: a kernel may START while its predecessor in
// the stream is still in its epilogue.  Rules followed by every kernel here:
//   * before pdl_wait(): NO global-memory access at all — only the shared-memory carve-up and the
//     mbarrier initialisation.  (The predecessor may be a foreign kernel that wrote z or mu and never
//     triggers; its writes are only guaranteed visible after the wait.)
//   * pdl_wait() (griddepcontrol.wait) before the first global read or write;
//   * pdl_trigger() (griddepcontrol.launch_dependents) after the main loop, so the successor's
//     launch latency + prologue overlap this kernel's CTA/grid reduction tail.
// Without the launch attribute both instructions are no-ops.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------
// In-kernel timeline (profiling builds only: `make timeline` -> libscc_b200_timeline.so, used by
// tools/timeline.py).  Thread 0 of every CTA stamps %globaltimer at phase boundaries into
// timeline[blockIdx.x * 8 + k].  Compiled out of the production library.
// ----------------------------------------------------------------------------
#ifdef SCC_TIMELINE
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define SCC_TL(ptr, k)                                                                              \
    do {                                                                                            \
        if (threadIdx.x == 0 && (ptr)) (ptr)[(size_t)blockIdx.x * 8 + (k)] = ::scc::globaltimer_ns(); \
    } while (0)
#else
#define SCC_TL(ptr, k) do { } while (0)
#endif

// ----------------------------------------------------------------------------
// Streaming ring of z tiles.
//   TILE points per tile, STAGES buffers, NT threads per CTA.
//   Protocol per CTA (tile_i = blockIdx.x + i * gridDim.x):
//     prologue : issue(s, tile_s) for s < STAGES ; __syncthreads()
//     iteration: wait(stage) ; read rows into registers ; __syncthreads() ;
//                issue(stage, tile_{i+STAGES}) ; compute
//   STAGES >= 2 so a cooperatively (non-TMA) filled stage is always separated
//   from its consumer by the __syncthreads() of an intervening iteration.
// ----------------------------------------------------------------------------
template <int D, int TILE, int STAGES, int NT>
struct ZRing {
    using L = RowLayout<D>;
    static constexpr int kTileFloats = TILE * L::LD;
    static constexpr uint32_t kTileBytes = TILE * D * sizeof(float);
    static_assert(STAGES >= 2, "ring needs two stages");
    static_assert((TILE * D) % 4 == 0, "tile must be a multiple of 16 bytes");

    float* buf;
    uint64_t* bar;
    const float* z;
    int num_tiles;      // tiles in the whole buffer (n <= 2^31 * TILE points)
    int last_points;    // points in the final tile (1..TILE)

    __device__ __forceinline__ void init(float* b, uint64_t* br, const float* z_, int64_t n_) {
        buf = b; bar = br; z = z_;
        num_tiles = (int)((n_ + TILE - 1) / TILE);
        last_points = (int)(n_ - (int64_t)(num_tiles - 1) * TILE);
        if (L::kDense && threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
            fence_mbar_init();
        }
    }
    __device__ __forceinline__ int points(int tile) const { return tile == num_tiles - 1 ? last_points : TILE; }
    __device__ __forceinline__ float* stage_ptr(int stage) const { return buf + stage * kTileFloats; }
    static __device__ __forceinline__ bool tma_ok(int np) { return L::kDense && np > 0 && ((np * D) & 3) == 0; }

    // padded float4 layouts are filled with cp.async (one commit group per issue() call per thread)
    static constexpr bool kCpAsync = L::kVec4 && !L::kDense;

    // Called by ALL threads at the same program point.
    __device__ __forceinline__ void issue(int stage, int tile) {
        if (tile < num_tiles) {
            float* dst = stage_ptr(stage);
            const float* src = z + (size_t)tile * (TILE * D);
            const int np = points(tile);
            if (tma_ok(np)) {                     // also a partial tile, when it is a whole number of 16-byte units
                if (threadIdx.x == 0) {
                    mbar_expect_tx(&bar[stage], (uint32_t)np * D * sizeof(float));
                    bulk_g2s(dst, src, (uint32_t)np * D * sizeof(float), &bar[stage]);
                }
            } else if constexpr (L::kVec4) {
                const int nvec = np * (D / 4);
                const float4* src4 = reinterpret_cast<const float4*>(src);
                for (int v = threadIdx.x; v < nvec; v += NT) {
                    const int row = v / (D / 4), c4 = v - row * (D / 4);
                    if constexpr (kCpAsync) cp_async16(dst + row * L::LD + 4 * c4, src4 + v);
                    else *reinterpret_cast<float4*>(dst + row * L::LD + 4 * c4) = ldg_stream4(src4 + v);
                }
            } else {
                const int nf = np * D;
                for (int f = threadIdx.x; f < nf; f += NT) {
                    const int row = f / D, c = f - row * D;
                    dst[row * L::LD + c] = ldg_stream(src + f);
                }
            }
        }
        if constexpr (kCpAsync) cp_async_commit();        // empty groups keep the per-thread count aligned
    }
    // use_index = how many times this stage has been consumed before (i / STAGES).
    // After wait() returns, every thread may read any row of the stage.
    __device__ __forceinline__ void wait(int stage, int tile, uint32_t use_index) {
        if constexpr (kCpAsync) {
            cp_async_wait<STAGES - 1>();                  // this thread's copies of the oldest stage landed
            __syncthreads();                              // ... and everybody else's
        } else {
            if (tma_ok(points(tile))) mbar_wait(&bar[stage], use_index & 1u);
        }
    }
};

// ----------------------------------------------------------------------------
// Peer-memory exchange window (see peer_exchange.cu).  Layout, identical on every rank:
//   [0, 512)      header {seq, ticket}
//   LL cells      uint64 [2 parity][16 src rank][2 * max_len]
// Every vector travels "flag in data" (LL): each double is shipped as two naturally aligned 8-byte words
// {32 data bits, 32-bit sequence number}, written with plain relaxed system-scope stores (an aligned 8-byte
// scalar store is single-copy atomic, also over NVLink).  The receiver polls the words themselves until they
// carry the sequence number of this exchange: no __threadfence_system, no separate flag round trip, and no
// ordering between elements — so a long vector can be exchanged by many CTAs, each owning a slice.
// Cells are double-buffered by sequence parity (a rank can be at most one exchange ahead of the slowest: to
// start exchange n+2 it needs everybody's n+1 data, which a rank only sends after it has consumed exchange n),
// and a stale cell carries seq - 2.
// `push` runs in ONE CTA (the last CTA of a statistics kernel) or slice-wise in every CTA of an exchange kernel;
// `pull` in the same CTA right after it (complete all-reduce inside the kernel) or in every CTA of the consumer.
// ----------------------------------------------------------------------------
constexpr int kPeerMaxWorld = 16;
constexpr size_t kPeerHeaderBytes = 512;
struct PeerHeader {
    unsigned int seq;
    unsigned int ticket;            // multi-CTA exchange kernels: CTAs that have finished their slice
    unsigned int pad[126];
};
struct PeerCtx {                    // by value inside kernel argument structs
    unsigned char* const* windows;  // device array of `world` window base pointers (nullptr = no exchange)
    int rank, world, max_len;
};

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long* peer_ll_slot(unsigned char* window, int parity, int src_rank, int max_len) {
    return reinterpret_cast<unsigned long long*>(window + kPeerHeaderBytes) +
           ((size_t)parity * kPeerMaxWorld + src_rank) * 2 * (size_t)max_len;
}

// Ship elements [lo, hi) of a vector (src[i - lo] holds element i; shared or global memory, visible to the CTA)
// to every rank's window (own window included) under sequence number seq.  All threads of the CTA; no barrier.
__device__ __forceinline__ void peer_push_slice(const PeerCtx& ex, const double* src, int lo, int hi, unsigned int seq) {
    const int parity = seq & 1u;
    const int nw = 2 * (hi - lo);
    for (int w = threadIdx.x; w < nw * ex.world; w += blockDim.x) {
        const int r = w / nw, e = w - r * nw;
        const double v = src[e >> 1];
        const unsigned int word = (e & 1) ? (unsigned int)__double2hiint(v) : (unsigned int)__double2loint(v);
        st_relaxed_sys_u64(peer_ll_slot(ex.windows[r], parity, ex.rank, ex.max_len) + 2 * lo + e,
                           ((unsigned long long)seq << 32) | word);
    }
}

// Wait for elements [lo, hi) of exchange `seq` from every rank and write their rank-ordered sums to
// dst[i - lo].  All threads of the CTA; no barrier.
__device__ __forceinline__ void peer_pull_slice(const PeerCtx& ex, double* dst, int lo, int hi, unsigned int seq) {
    const int parity = seq & 1u;
    const unsigned long long* base = peer_ll_slot(ex.windows[ex.rank], parity, 0, ex.max_len);
    const size_t rstride = 2 * (size_t)ex.max_len;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double acc = 0.0;
        for (int r0 = 0; r0 < ex.world; r0 += 4) {              // 8 polls in flight, summed in rank order
            unsigned long long wlo[4], whi[4];
            bool ok;
            do {
                ok = true;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (r0 + u < ex.world) {
                        const unsigned long long* c = base + (size_t)(r0 + u) * rstride + 2 * i;
                        wlo[u] = ld_relaxed_sys_u64(c);
                        whi[u] = ld_relaxed_sys_u64(c + 1);
                        ok = ok && (unsigned int)(wlo[u] >> 32) == seq && (unsigned int)(whi[u] >> 32) == seq;
                    }
                }
            } while (!ok);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + u < ex.world) acc += __hiloint2double((int)(unsigned int)whi[u], (int)(unsigned int)wlo[u]);
        }
        dst[i - lo] = acc;
    }
}

// All threads of ONE CTA: ship src[0..len) to every rank's window.  Advances the local sequence number and
// returns it.  Ends with a CTA barrier, so src may be overwritten afterwards.  Kept out of line: it runs once per
// kernel, in one CTA, and must not weigh on the register allocation of the streaming loops.
static __device__ __noinline__ unsigned int peer_push(const PeerCtx& ex, const double* src, int len) {
    PeerHeader* me = reinterpret_cast<PeerHeader*>(ex.windows[ex.rank]);
    const unsigned int seq = ld_relaxed_gpu_u32(&me->seq) + 1u;
    peer_push_slice(ex, src, 0, len, seq);
    __syncthreads();                                // every thread has read me->seq and src
    if (threadIdx.x == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&me->seq), "r"(seq) : "memory");
    return seq;
}

// All threads of a CTA: wait for exchange `seq` (0: the one most recently pushed on this GPU by a preceding
// kernel) and write the rank-ordered sum of the `len` doubles to dst (shared or global).  Ends with a CTA barrier.
static __device__ __noinline__ void peer_pull(const PeerCtx& ex, double* dst, int len, unsigned int seq = 0u) {
    PeerHeader* me = reinterpret_cast<PeerHeader*>(ex.windows[ex.rank]);
    if (seq == 0u) seq = ld_relaxed_gpu_u32(&me->seq);
    peer_pull_slice(ex, dst, 0, len, seq);
    __syncthreads();
}

// ----------------------------------------------------------------------------
// Reductions.
// ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp "reduce-scatter" of NVP (multiple of 32) per-lane floats: log2(32) halving steps, each lane
// keeps one half of the vector and ships the other half to its partner, so the whole reduction
// costs NVP - NVP/32 shuffles instead of 5 * NVP.  On return lane l holds in v[0..NVP/32) the
// warp totals of the original entries (NVP/32) * l + r.  Fixed tree order: deterministic.
template <int NVP>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[NVP]) {
    static_assert(NVP % 32 == 0, "pad the vector to a multiple of 32");
    const int lane = threadIdx.x & 31;
    int n = NVP;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int half = n / 2;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < NVP / 2; ++i) {
            if (i < half) {
                const float keep = up ? v[i + half] : v[i];
                const float send = up ? v[i] : v[i + half];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        n = half;
    }
}

// Single-instruction MUFU forms (no range fix-ups): callers guarantee normal, positive operands.
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ldcg_f64x2(const double* p, double& a, double& b) {
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}

// CTA-level statistics (float64, in shared memory) -> per-CTA slot of `partials`
// -> the last CTA to arrive sums all slots into `out` with ALL its threads:
// thread t owns statistic s = t % S and the CTA rows r, r+R, r+2R, ... (R = NT / S row groups),
// then the R partial sums are combined in row-group order.  The summation order depends only
// on (grid, S, NT), so the result is deterministic for a given device.
// `counter` must be 0 on entry and is reset.  scratch: NT doubles of shared memory.
// `partials` needs gridDim.x * ((S + 1) & ~1) doubles.
// Returns true in the CTA that arrived last (after `out` is complete).
// `push` (optional, multi-GPU): mode 1 = the last CTA ships the reduced vector to every rank's exchange
// window (a later kernel pulls); mode 2 = it also waits for the world's vectors and overwrites `out` with
// their rank-ordered sum — the all-reduce completes inside this kernel.
template <int NT, int U = 16>
__device__ __forceinline__ bool grid_publish(const double* cta_stats, int S, double* partials,
                                             unsigned int* counter, double* out, double* scratch,
                                             const PeerCtx* push = nullptr, int push_mode = 1) {
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int SP = (S + 1) & ~1;                  // slot stride: even, so slots stay 16-byte aligned
    double* mine = partials + (size_t)blockIdx.x * SP;
    // the slot is written by many threads; the CTA barrier orders those writes before thread 0's ticket,
    // whose acq_rel semantics at GPU scope publishes them (and acquires every earlier CTA's slot)
    for (int s = tid; s < SP; s += NT) __stcg(mine + s, (s < S) ? cta_stats[s] : 0.0);
    __syncthreads();
    if (tid == 0) {
        unsigned int prev;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(counter) : "memory");
        s_last = (prev == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return false;
    const int G = gridDim.x;
    const int C = SP / 2;                         // 128-bit columns (pairs of statistics)
    if (2 * C <= NT) {
        // thread t owns column pair c = t % C and the CTA rows r, r+R, r+2R, ... (R = NT / C row groups);
        // U independent 128-bit loads in flight per thread (the chain is L2-latency bound), summed in
        // a fixed order; the R partial sums are then combined in row-group order.
        const int R = NT / C;
        const int c = tid % C, r = tid / C;
        double a0 = 0.0, a1 = 0.0;
        if (r < R) {
            for (int b = r; b < G; b += U * R) {
                double v0[U], v1[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int bb = b + u * R;
                    v0[u] = 0.0; v1[u] = 0.0;
                    if (bb < G) ldcg_f64x2(partials + (size_t)bb * SP + 2 * c, v0[u], v1[u]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { a0 += v0[u]; a1 += v1[u]; }
            }
            scratch[(r * C + c) * 2] = a0;
            scratch[(r * C + c) * 2 + 1] = a1;
        }
        __syncthreads();
        double t = 0.0;
        if (tid < S) {
            for (int rr = 0; rr < R; ++rr) t += scratch[rr * 2 * C + tid];
            out[tid] = t;
        }
    } else {
        for (int s = tid; s < S; s += NT) {
            double acc = 0.0;
            for (int b = 0; b < G; b += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (b + u < G) ? __ldcg(partials + (size_t)(b + u) * SP + s) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += v[u];
            }
            out[s] = acc;
        }
    }
    if (push && push->windows && push_mode) {       // fused tail: ship the vector to every rank (any length)
        __syncthreads();                            // out[] (global) is complete and visible to the CTA
        const unsigned int seq = peer_push(*push, out, S);
        if (push_mode == 2) peer_pull(*push, out, S, seq);
    }
    if (tid == 0) *counter = 0u;
    return true;
}

// Grid-wide barrier + all-reduce of a short vector of NON-NEGATIVE, bounded statistics INSIDE a kernel whose
// CTAs are all co-resident (cooperative launch): the column sums f_j <= n and the label-change count of the
// one-kernel DEC step.  Every warp of the grid adds its S values as 64-bit FIXED-POINT integers to S global
// accumulators with integer atomics; every 64-bit accumulator word carries, above its fixed-point value, the NUMBER of
// contributions it has received, so a waiter learns that a sum is final from the sum's own word:
//   bits  0..41  value: sum of non-negative fixed-point contributions, total < 2^41
//   bits 42..52  poison: contributions that were negative, NaN or above the bound (at most 2047 contributors)
//   bits 53..63  count of contributions
// Every warp of the grid adds ONE word per statistic; every CTA then polls the S words until the count field
// reads `contributors` — a single L2 round trip per poll, against ticket -> poll -> read for the ticket barrier.
// Integer addition is associative: the sums are bit-reproducible whatever the arrival order.  A poisoned sum
// reads as NaN.  Multi-GPU: CTA 0 ships the local sums to every rank's window once it has them and EVERY CTA
// pulls the rank-ordered world sum from the (local) window itself — no relay through one CTA and a ready flag.
// The words must be 0 on entry; the caller resets them once no CTA can still be polling.
struct CountedFix {
    static constexpr int kValueBits = 42, kPoisonShift = 42, kCountShift = 53;
    static constexpr unsigned int kMaxContributors = 2047;
    __device__ static __forceinline__ unsigned long long word(float v, double scale, double bound) {
        const bool ok = v >= 0.f && (double)v <= bound;            // false for NaN
        return (1ull << kCountShift) + (ok ? (unsigned long long)__double2ll_rn((double)v * scale) : (1ull << kPoisonShift));
    }
};
template <int NT>
__device__ __forceinline__ void grid_barrier_counted(int S, const unsigned long long* fix, unsigned int contributors,
                                                     double inv_scale, double* out_s, double* scratch,
                                                     const PeerCtx* ex, unsigned int seq) {
    const int tid = threadIdx.x;
    const bool multi = ex && ex->windows;
    double* local = multi ? scratch : out_s;
    if (tid < S) {
        unsigned long long v;
        do {
            v = ld_relaxed_sys_u64(fix + tid);
        } while ((unsigned int)(v >> CountedFix::kCountShift) != contributors);
        const bool poisoned = ((v >> CountedFix::kPoisonShift) & 0x7ffull) != 0ull;
        const double val = (double)(v & ((1ull << CountedFix::kValueBits) - 1ull)) * inv_scale;
        local[tid] = poisoned ? __longlong_as_double(0x7ff8000000000000ll) : val;
    }
    __syncthreads();
    if (!multi) return;
    if (blockIdx.x == 0) {
        peer_push_slice(*ex, local, 0, S, seq);
        if (tid == 0) {
            PeerHeader* me = reinterpret_cast<PeerHeader*>(ex->windows[ex->rank]);
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&me->seq), "r"(seq) : "memory");
        }
    }
    peer_pull_slice(*ex, out_s, 0, S, seq);                       // rank-ordered sum of every GPU's vector
    __syncthreads();
}

// Stand-alone fixed-order reduction of per-CTA partial slots, for statistics vectors too long
// for one CTA to sum quickly (GMM: up to 8977 doubles x hundreds of CTAs).
static __global__ void __launch_bounds__(256)
reduce_partials_kernel(const double* __restrict__ partials, int S, int G, double* __restrict__ out,
                       const double* __restrict__ ctrl) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (ctrl && ctrl[5] != 0.0) return;          // frozen fit: the producer kernel did not run
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int b = 0;
    for (; b + 3 < G; b += 4) {
        a0 += partials[(size_t)b * S + s];
        a1 += partials[(size_t)(b + 1) * S + s];
        a2 += partials[(size_t)(b + 2) * S + s];
        a3 += partials[(size_t)(b + 3) * S + s];
    }
    for (; b < G; ++b) a0 += partials[(size_t)b * S + s];
    out[s] = (a0 + a1) + (a2 + a3);
}

__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }

__device__ __forceinline__ float round_dec5(float x) {
    // np.round(x, 5): rint (half-to-even) of x * 1e5, scaled back.
    // 0 <= x * 1e5 < 2^22: adding 1.5 * 2^23 rounds the EXACT product to an integer in the FMA itself
    // (one rounding, half-to-even) and keeps the conversion (XU) pipe free of an FRND per value
    const float y = fmaf(x, 100000.0f, 12582912.0f);
    return (y - 12582912.0f) * 1.0e-5f;
}

}  // namespace scc
