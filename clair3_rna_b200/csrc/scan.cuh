// Device-wide inclusive scan in ONE pass: decoupled look-back over tiles handed out by an atomic ticket.
//
// The pileup path needs a handful of small scans (CIGAR ops, bitmap words, rows,
// bins, candidates); each is expressed as an `Op` functor so the load (e.g. decode
// a CIGAR word) and the store (e.g. emit op geometry, mark coverage) are fused into
// the scan instead of being separate kernels.
//
//   struct Op {
//       typedef ... T;                                   // POD
//       __device__ T identity() const;
//       __device__ T combine(const T& a, const T& b) const;   // associative, a before b
//       __device__ T load(int64_t i) const;
//       __device__ void store(int64_t i, const T& inclusive, const T& own) const;
//       __device__ int64_t size() const;                 // element count (may read device memory)
//   };
//
// One kernel per scan, every element loaded once (the three-pass form this replaces launched three kernels and
// called load() twice per element):
//   * a block takes its tile from an atomic ticket, so the tiles before it are held by blocks that are running or
//     done - look-back never waits for a block that is not resident, whatever the grid and the device-side size;
//   * the tile's aggregate, then its inclusive prefix, are published as (value, status word) with a release store;
//     the status carries the scan's epoch, so the descriptors are never cleared between scans of a slot;
//   * one warp looks back 32 tiles at a time; the block scan is two levels of warp shuffles (2 __syncthreads);
//   * the last block to leave resets the ticket; the spins are bounded (a protocol bug sets the error flag instead
//     of hanging the GPU).
// The host only knows an upper bound of most sizes (row and candidate counts live on the device): the grid is sized
// from the bound, blocks whose ticket lies beyond size() leave at once.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace c3r {

constexpr int SCAN_BT = 256;      // threads per block
constexpr int SCAN_IPT = 8;       // items per thread
constexpr int SCAN_TILE = SCAN_BT * SCAN_IPT;
constexpr int SCAN_MAX_GRID = 148 * 8;
constexpr uint32_t SCAN_SPINS = 1u << 22;

// per-slot scan state: ticket / exit counters and the tile descriptors (sized for the largest scan of the slot)
struct ScanState {
    unsigned int* ctrl;           // [0] ticket, [1] blocks that have left
    uint32_t* status;             // [tiles]: epoch << 2 | {1: aggregate published, 2: inclusive prefix published}
    void* aggr;                   // [tiles] T
    void* incl;                   // [tiles] T
    int* err;                     // set to 900 when a look-back spin runs out
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <class T>
__device__ __forceinline__ T shfl_up_pod(const T& v, int delta) {
    static_assert(sizeof(T) % 4 == 0, "scan element must be a multiple of 4 bytes");
    T r;
    const uint32_t* s = (const uint32_t*)&v;
    uint32_t* d = (uint32_t*)&r;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); ++i) d[i] = __shfl_up_sync(0xffffffffu, s[i], delta);
    return r;
}
template <class T>
__device__ __forceinline__ T shfl_idx_pod(const T& v, int lane) {
    T r;
    const uint32_t* s = (const uint32_t*)&v;
    uint32_t* d = (uint32_t*)&r;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); ++i) d[i] = __shfl_sync(0xffffffffu, s[i], lane);
    return r;
}

// inclusive scan of one value per lane
template <class Op>
__device__ __forceinline__ typename Op::T warp_scan_inclusive(const Op& op, typename Op::T v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const typename Op::T u = shfl_up_pod(v, d);
        if (lane >= d) v = op.combine(u, v);
    }
    return v;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_BT) scan_lookback_kernel(Op op, ScanState st, uint32_t epoch, typename Op::T* total_out) {
    typedef typename Op::T T;
    __shared__ T warp_tot[SCAN_BT / 32];
    __shared__ T tile_prefix;
    __shared__ unsigned int tile_s;
    const int64_t n = op.size();
    const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* aggr = (T*)st.aggr;
    T* incl = (T*)st.incl;
    const uint32_t s_aggr = (epoch << 2) | 1u, s_incl = (epoch << 2) | 2u;
    for (;;) {
        __syncthreads();                                     // tile_s / tile_prefix of the previous round are consumed
        if (threadIdx.x == 0) tile_s = atomicAdd(&st.ctrl[0], 1u);
        __syncthreads();
        const int64_t tile = tile_s;
        if (tile >= n_tiles) break;
        const int64_t i0 = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
        T own[SCAN_IPT];
        T acc = op.identity();
#pragma unroll
        for (int k = 0; k < SCAN_IPT; ++k) {
            if (i0 + k < n) {
                own[k] = op.load(i0 + k);
                acc = op.combine(acc, own[k]);
            }
        }
        // block scan of the thread aggregates: warp scan, then the warp totals
        const T winc = warp_scan_inclusive(op, acc);
        if (lane == 31) warp_tot[warp] = winc;
        __syncthreads();
        T wbase = op.identity();
#pragma unroll
        for (int w = 0; w < SCAN_BT / 32; ++w)
            if (w < warp) wbase = op.combine(wbase, warp_tot[w]);
        T texcl = shfl_up_pod(winc, 1);                      // exclusive within the warp
        if (lane == 0) texcl = op.identity();
        const T excl_in_tile = op.combine(wbase, texcl);
        if (warp == SCAN_BT / 32 - 1) {
            // the last warp owns the tile aggregate (its lane 31's inclusive value) and does the look-back
            const T tile_aggr = shfl_idx_pod(op.combine(wbase, winc), 31);
            if (tile == 0) {
                if (lane == 0) {
                    incl[0] = tile_aggr;
                    __threadfence();
                    st_release_u32(&st.status[0], s_incl);
                    tile_prefix = op.identity();
                }
            } else {
                if (lane == 0) {
                    aggr[tile] = tile_aggr;
                    __threadfence();
                    st_release_u32(&st.status[tile], s_aggr);
                }
                // windows of 32 predecessors, nearest first: lane l looks at tile - 1 - l (- 32 per window)
                T prefix = op.identity();                    // combined aggregates of the tiles looked at so far
                int64_t base = tile - 1;
                bool done = false;
                uint32_t spins = 0;
                while (!done) {
                    const int64_t t = base - lane;
                    uint32_t s = 0;
                    if (t >= 0) {
                        s = ld_acquire_u32(&st.status[t]);
                        while ((s >> 2) != epoch || (s & 3u) == 0u) {          // not yet published in this scan
                            if (++spins > SCAN_SPINS) { if (st.err) atomicExch(st.err, 900); s = s_incl; break; }
                            s = ld_acquire_u32(&st.status[t]);
                        }
                    }
                    const bool has_incl = t >= 0 && (s & 3u) == 2u;
                    const uint32_t incl_mask = __ballot_sync(0xffffffffu, has_incl || t < 0);
                    // lanes up to (and including) the first one holding an inclusive prefix (or beyond tile 0) contribute
                    const int stop = __ffs(incl_mask) - 1;                      // -1: none in this window
                    const int last = stop < 0 ? 31 : stop;
                    T v = op.identity();
                    if (t >= 0 && lane <= last) v = has_incl ? incl[t] : aggr[t];
                    // combine in tile order: the farthest tile first
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const T u = shfl_idx_pod(v, (lane + d) & 31);
                        if (lane + d < 32) v = op.combine(u, v);
                    }
                    // lane 0 now holds tiles [base - last .. base] combined (lanes beyond `last` held the identity)
                    const T wsum = shfl_idx_pod(v, 0);
                    prefix = op.combine(wsum, prefix);
                    if (stop >= 0) done = true;
                    base -= 32;
                }
                if (lane == 0) {
                    incl[tile] = op.combine(prefix, tile_aggr);
                    __threadfence();
                    st_release_u32(&st.status[tile], s_incl);
                    tile_prefix = prefix;
                }
            }
            if (lane == 0 && tile == n_tiles - 1 && total_out) *total_out = op.combine(tile_prefix, tile_aggr);
        }
        __syncthreads();
        T run = op.combine(tile_prefix, excl_in_tile);
#pragma unroll
        for (int k = 0; k < SCAN_IPT; ++k) {
            if (i0 + k < n) {
                run = op.combine(run, own[k]);
                op.store(i0 + k, run, own[k]);
            }
        }
    }
    if (n_tiles == 0 && total_out && blockIdx.x == 0 && threadIdx.x == 0) *total_out = op.identity();
    // the last block to leave hands the ticket back
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&st.ctrl[1], 1u) == gridDim.x - 1) {
            st.ctrl[0] = 0u;
            st.ctrl[1] = 0u;
            __threadfence();
        }
    }
}

// Host helper.  `n_upper` bounds op.size(); the state's descriptors must hold ceil(n_upper/SCAN_TILE) entries of T.
// Consecutive scans of a slot run on one stream, so a scan's ticket reset is complete before the next starts.
// Returns the number of kernels launched.
template <class Op>
inline int device_scan(const Op& op, int64_t n_upper, const ScanState& st, uint32_t& epoch, typename Op::T* total_out,
                       cudaStream_t stream) {
    if (n_upper <= 0) n_upper = 1;
    int64_t nb64 = (n_upper + SCAN_TILE - 1) / SCAN_TILE;
    const unsigned nb = (unsigned)(nb64 < SCAN_MAX_GRID ? nb64 : SCAN_MAX_GRID);
    epoch = (epoch + 1u) & 0x3fffffffu;
    if (epoch == 0u) epoch = 1u;                         // 0 is the state of a cleared descriptor
    scan_lookback_kernel<Op><<<nb, SCAN_BT, 0, stream>>>(op, st, epoch, total_out);
    return 1;
}

}  // namespace c3r
