// Device-wide inclusive scan in ONE pass: co-resident blocks walk the tiles round robin, tile prefixes by direct summation.
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
//       struct Ctx { ... };  __device__ void ctx_init(Ctx&) const;   // thread-local state carried over a thread's
//       __device__ void store(int64_t i, const T& inclusive, const T& own, Ctx&) const;   // consecutive stores
//       __device__ int64_t size() const;                 // element count (may read device memory)
//   };
//
// One kernel per scan, every element loaded once (the three-pass form this replaces launched three kernels and
// called load() twice per element):
//   * the grid is no larger than what the device holds at once (occupancy query) and block b takes tiles b,
//     b + grid, ... in increasing order, so the tiles before any tile are held by blocks that are resident or done
//     (should another stream's kernel delay part of the grid, the waiting blocks spin - bounded - until it arrives).
//     An atomic ticket per tile, the usual way to get that guarantee, costs more than the scan itself here: ~27
//     cycles per atomic on ONE address, serialised, is 16 us for a grid of 1184 blocks;
//   * the tile's aggregate is published as (value, status word) with a release store; the status carries the
//     scan's epoch, so the descriptors are never cleared between scans of a slot;
//   * a tile's exclusive prefix is the ordered sum of the aggregates of the tiles before it, taken by the whole
//     block at once (thread i polls and adds a contiguous run of predecessors, then an ordered block reduction) -
//     no chain of look-back hops: these scans are a few hundred to a few thousand tiles that all start together,
//     where a look-back walks back window after window while the inclusive prefixes trickle forward (measured:
//     27-61 us per scan, as slow as the three kernels it replaced).  Every SCAN_CP-th tile also publishes its
//     exclusive prefix as a checkpoint, so a tile sums at most SCAN_CP aggregates;
//   * the block scan is two levels of warp shuffles (2 __syncthreads);
//   * the spins are bounded (a protocol bug sets the error flag instead of hanging the GPU).
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
constexpr int SCAN_CP = 1024;     // tiles between checkpoints (exclusive prefixes published for the tiles after them)

// per-slot scan state: the tile descriptors (sized for the largest scan of the slot)
struct ScanState {
    unsigned int* ctrl;           // (unused: kept for layout stability of the scratch buffer)
    uint32_t* status;             // [tiles]: epoch << 2 | {1: aggregate published, 3: and (checkpoint tiles) its prefix}
    void* aggr;                   // [tiles] T
    void* incl;                   // [tiles / SCAN_CP + 1] T: exclusive prefix of checkpoint tile k * SCAN_CP
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

// wait until tile t of this scan has published (state bits & want) == want; returns false when the spin ran out
__device__ __forceinline__ bool scan_wait(const ScanState& st, int64_t t, uint32_t epoch, uint32_t want) {
    for (uint32_t spins = 0; spins < SCAN_SPINS; ++spins) {
        const uint32_t s = ld_acquire_u32(&st.status[t]);
        if ((s >> 2) == epoch && (s & want) == want) return true;
    }
    if (st.err) atomicExch(st.err, 900);
    return false;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_BT) scan_lookback_kernel(Op op, ScanState st, uint32_t epoch, typename Op::T* total_out) {
    typedef typename Op::T T;
    __shared__ T warp_tot[SCAN_BT / 32];
    __shared__ T tile_prefix;
    const int64_t n = op.size();
    const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* aggr = (T*)st.aggr;
    T* cpx = (T*)st.incl;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();                                     // tile_prefix / warp_tot of the previous round are consumed
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
        T tile_aggr = op.identity();
        if (threadIdx.x == SCAN_BT - 1) {                    // the last thread's inclusive value is the tile aggregate
            tile_aggr = op.combine(wbase, winc);
            aggr[tile] = tile_aggr;
            __threadfence();
            st_release_u32(&st.status[tile], (epoch << 2) | 1u);
        }
        // ---- exclusive prefix of the tile: checkpoint prefix + the aggregates of the tiles since, in order
        const int64_t cp = tile == 0 ? 0 : ((tile - 1) / SCAN_CP) * SCAN_CP;     // nearest checkpoint tile below `tile`
        const int64_t cnt = tile - cp;                       // predecessors to add: tiles cp .. tile-1
        const int64_t per = (cnt + SCAN_BT - 1) / SCAN_BT;
        T part = op.identity();
        {
            const int64_t a = cp + (int64_t)threadIdx.x * per, b = a + per < tile ? a + per : tile;
            for (int64_t t = a; t < b; ++t) {
                if (!scan_wait(st, t, epoch, 1u)) break;
                part = op.combine(part, aggr[t]);
            }
        }
        __syncthreads();                                     // warp_tot is free again
        const T pinc = warp_scan_inclusive(op, part);
        if (lane == 31) warp_tot[warp] = pinc;
        __syncthreads();
        if (threadIdx.x == SCAN_BT - 1) {
            T prefix = op.identity();
            if (cp > 0) {                                    // the checkpoint tile's own exclusive prefix
                scan_wait(st, cp, epoch, 3u);
                prefix = cpx[cp / SCAN_CP];
            }
#pragma unroll
            for (int w = 0; w < SCAN_BT / 32; ++w) prefix = op.combine(prefix, warp_tot[w]);
            tile_prefix = prefix;
            if (tile > 0 && tile % SCAN_CP == 0) {           // this tile is a checkpoint for the tiles after it
                cpx[tile / SCAN_CP] = prefix;
                __threadfence();
                st_release_u32(&st.status[tile], (epoch << 2) | 3u);
            }
            if (tile == n_tiles - 1 && total_out) *total_out = op.combine(prefix, tile_aggr);
        }
        __syncthreads();
        T run = op.combine(tile_prefix, excl_in_tile);
        typename Op::Ctx ctx;
        op.ctx_init(ctx);
#pragma unroll
        for (int k = 0; k < SCAN_IPT; ++k) {
            if (i0 + k < n) {
                run = op.combine(run, own[k]);
                op.store(i0 + k, run, own[k], ctx);
            }
        }
    }
    if (n_tiles == 0 && total_out && blockIdx.x == 0 && threadIdx.x == 0) *total_out = op.identity();
}

// Host helper.  `n_upper` bounds op.size(); the state's descriptors must hold ceil(n_upper/SCAN_TILE) entries of T.
// Consecutive scans of a slot run on one stream, so a scan's ticket reset is complete before the next starts.
// Returns the number of kernels launched.
template <class Op>
inline int device_scan(const Op& op, int64_t n_upper, const ScanState& st, uint32_t& epoch, typename Op::T* total_out,
                       cudaStream_t stream) {
    if (n_upper <= 0) n_upper = 1;
    int64_t nb64 = (n_upper + SCAN_TILE - 1) / SCAN_TILE;
    // co-resident grid: blocks per SM of this instantiation x SMs (queried once)
    static int resident = 0;
    if (!resident) {
        int per_sm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_lookback_kernel<Op>, SCAN_BT, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        resident = per_sm * (sms > 0 ? sms : 1);
        if (resident > SCAN_MAX_GRID) resident = SCAN_MAX_GRID;
    }
    const unsigned nb = (unsigned)(nb64 < resident ? nb64 : resident);
    epoch = (epoch + 1u) & 0x3fffffffu;
    if (epoch == 0u) epoch = 1u;                         // 0 is the state of a cleared descriptor
    scan_lookback_kernel<Op><<<nb, SCAN_BT, 0, stream>>>(op, st, epoch, total_out);
    return 1;
}

}  // namespace c3r
