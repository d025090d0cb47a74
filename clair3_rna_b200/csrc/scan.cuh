// Device-wide inclusive scan, three passes (tile reduce / aggregate scan / tile apply).
//
// The pileup path needs a handful of small scans (CIGAR ops, bitmap words, rows,
// bins, candidates); each is expressed as an `Op` functor so the load (e.g. decode
// a CIGAR word) and the store (e.g. emit op geometry, mark coverage) are fused into
// the scan passes instead of being separate kernels.
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
// The host only knows an upper bound of most sizes (row and candidate counts live on the
// device), so the tile passes run as grid-stride loops over ceil(size()/SCAN_TILE) tiles on a
// grid of at most SCAN_MAX_GRID blocks: no block is launched just to find out it has nothing
// to do.  A three pass scan reads the input twice; it is used because every scan here is far
// smaller than the count/tensor traffic and a look-back scan can spin forever if a
// predecessor tile is not resident.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace c3r {

constexpr int SCAN_BT = 256;      // threads per block
constexpr int SCAN_IPT = 8;       // items per thread (1 per thread was measured slower: the block scan dominates)
constexpr int SCAN_TILE = SCAN_BT * SCAN_IPT;
constexpr int SCAN_MAX_GRID = 148 * 8;

template <class Op>
__device__ __forceinline__ typename Op::T block_scan_exclusive(const Op& op, typename Op::T v,
                                                               typename Op::T* smem, typename Op::T* total) {
    // Hillis-Steele over SCAN_BT thread aggregates in shared memory.
    typedef typename Op::T T;
    const int t = threadIdx.x;
    smem[t] = v;
    __syncthreads();
#pragma unroll
    for (int d = 1; d < SCAN_BT; d <<= 1) {
        T x = smem[t];
        if (t >= d) x = op.combine(smem[t - d], x);
        __syncthreads();
        smem[t] = x;
        __syncthreads();
    }
    T excl = t ? smem[t - 1] : op.identity();
    if (total) *total = smem[SCAN_BT - 1];
    return excl;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_BT) scan_reduce_kernel(Op op, typename Op::T* tile_aggr) {
    typedef typename Op::T T;
    __shared__ T smem[SCAN_BT];
    const int64_t n = op.size();
    for (int64_t tile = blockIdx.x; tile * SCAN_TILE < n; tile += gridDim.x) {
        const int64_t i0 = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
        T acc = op.identity();
#pragma unroll
        for (int k = 0; k < SCAN_IPT; ++k)
            if (i0 + k < n) acc = op.combine(acc, op.load(i0 + k));
        block_scan_exclusive(op, acc, smem, (T*)nullptr);
        // thread SCAN_BT-1's inclusive value is the tile aggregate
        if (threadIdx.x == SCAN_BT - 1) tile_aggr[tile] = smem[SCAN_BT - 1];
        __syncthreads();
    }
}

// single block: tile_aggr[0..nb) -> exclusive prefix in place; total to *total_out
template <class Op>
__global__ void __launch_bounds__(SCAN_BT) scan_aggr_kernel(Op op, typename Op::T* tile_aggr, typename Op::T* total_out) {
    typedef typename Op::T T;
    __shared__ T smem[SCAN_BT];
    __shared__ T carry_s;
    const int64_t n = op.size();
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (threadIdx.x == 0) carry_s = op.identity();
    __syncthreads();
    for (int64_t b0 = 0; b0 < nb; b0 += SCAN_BT) {
        const int64_t i = b0 + threadIdx.x;
        T v = i < nb ? tile_aggr[i] : op.identity();
        T excl = block_scan_exclusive(op, v, smem, (T*)nullptr);
        T carry = carry_s;
        T chunk_total = smem[SCAN_BT - 1];
        __syncthreads();
        if (i < nb) tile_aggr[i] = op.combine(carry, excl);
        if (threadIdx.x == 0) carry_s = op.combine(carry, chunk_total);
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_BT) scan_apply_kernel(Op op, const typename Op::T* tile_excl) {
    typedef typename Op::T T;
    __shared__ T smem[SCAN_BT];
    const int64_t n = op.size();
    for (int64_t tile = blockIdx.x; tile * SCAN_TILE < n; tile += gridDim.x) {
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
        T excl = block_scan_exclusive(op, acc, smem, (T*)nullptr);
        T run = op.combine(tile_excl[tile], excl);
#pragma unroll
        for (int k = 0; k < SCAN_IPT; ++k) {
            if (i0 + k < n) {
                run = op.combine(run, own[k]);
                op.store(i0 + k, run, own[k]);
            }
        }
        __syncthreads();
    }
}

// Host helper.  `n_upper` bounds op.size(); scratch must hold ceil(n_upper/SCAN_TILE) T's.
// Returns the number of kernels launched.
template <class Op>
inline int device_scan(const Op& op, int64_t n_upper, typename Op::T* scratch, typename Op::T* total_out,
                       cudaStream_t stream) {
    if (n_upper <= 0) n_upper = 1;
    int64_t nb64 = (n_upper + SCAN_TILE - 1) / SCAN_TILE;
    const unsigned nb = (unsigned)(nb64 < SCAN_MAX_GRID ? nb64 : SCAN_MAX_GRID);
    scan_reduce_kernel<Op><<<nb, SCAN_BT, 0, stream>>>(op, scratch);
    scan_aggr_kernel<Op><<<1, SCAN_BT, 0, stream>>>(op, scratch, total_out);
    scan_apply_kernel<Op><<<nb, SCAN_BT, 0, stream>>>(op, scratch);
    return 3;
}

}  // namespace c3r
