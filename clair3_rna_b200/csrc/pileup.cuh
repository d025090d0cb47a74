// Integer half of the pileup calling path: flat alignment records -> per-position
// count rows -> candidates -> 33 x C windows + alt_info.  sm_100a, HBM-bound work:
// no tensor cores here.
//
// Reference semantics restated (all under /root/reference):
//   samtools mpileup column membership / tokens   src/create_tensor_pileup.py:436-451 (external htslib)
//   generate_tensor                                src/create_tensor_pileup.py:85-302
//   candidate predicate, ring buffer, padding      src/create_tensor_pileup.py:463-611
//   depth rescale                                  clair3_rna/utils.py:85-92,120
//
// Data layout in HBM (one chunk):
//   position space  region [R0, R1), W = R1-R0 positions, bitmaps of ceil(W/32)+2 words
//       covA   1 bit/position: covered by >=1 admitted read (incl. D and N)  -> run contiguity
//       covE   1 bit/position: covered by an M/=/X/D op (a base or a `*`/`#`)
//       rowR   = dilate16(covE) & covA : positions that get a count row
//       word_base[w] = rank of the first row of word w  (row(p) = word_base + popc(bits below p))
//   row space  L rows (device-side count), rows of consecutive positions are consecutive
//       counts[L][C] int32, row_pos[L], row_depth[L], row_flag[L], max_skip[L]
//   bins   32-row tiles; tile_entries = M/D segments overlapping the tile (16 B each);
//          ev = indel events in CSR by row (16 B each)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "scan.cuh"

namespace c3r {

constexpr int WIN = 33;
constexpr int FLANK = 16;
constexpr int TILE_ROWS = 32;

struct SegEntry {              // an M/=/X or D op clipped to nothing: lanes test coverage themselves
    int32_t x;                 // reference start (0-based)
    int32_t len;               // reference length of an M/=/X op; 0 for a deletion
    uint32_t yx;               // M/=/X: query start - x (mod 2^32), base index of position p is yx + p; D: its length
    uint32_t info;             // rid << 4 | deletion (`*`/`#`) << 3 | hp(2 bits) << 1 | reverse
};

struct IndelEvent {
    uint32_t info;             // rid << 4 | hp << 2 | is_del << 1 | reverse
    int32_t len;
    uint32_t y;                // insertion: base index of inserted bases
    uint32_t pad;
};

struct AltEntry {              // mirrors c3r_alt_entry
    uint8_t kind, base;
    uint16_t len;
    int32_t count;
    uint32_t seq_off;
    uint32_t order;
};

struct ScanElem {              // segmented scan over CIGAR ops
    unsigned long long v;      // ref_len << 32 | qry_len (within the read so far)
    int32_t rid;
    int32_t flag;
};

struct Int2 { int32_t a, b; };

// Everything the kernels need, passed by value.
struct Dev {
    // ---- inputs
    int64_t n_reads, n_ops;
    const int32_t* pos; const uint16_t* flag; const uint8_t* mapq; const uint8_t* hp;
    const int32_t* cigar_off; const uint32_t* cigar; const int64_t* seq_off; const uint8_t* seq;
    int64_t n_seq_words;          // readable 32-bit words of the sequence buffer (incl. the slack after the last read)
    const uint8_t* ref; int64_t ref_start0; int64_t ref_len;
    int32_t R0, R1; int64_t W; int64_t NW;            // region (0-based half open), words
    // ---- params
    int32_t C; int32_t min_cov; int32_t min_mq; uint32_t excl; double snp_af, indel_af;
    int32_t padding; int32_t max_depth; double skip_prop;
    const uint16_t* thr_snp; const uint16_t* thr_indel;     // THR_N entries each (k_thr_table)
    // ---- per read / per op
    uint8_t* admit; int32_t* read_end; int32_t* op_head;
    int32_t* op_x; uint32_t* op_y; int32_t* op_rid;
    // ---- position space
    uint32_t* covA; uint32_t* covE; uint32_t* rowR; Int2* wdiff; int32_t* word_base;
    // ---- row space
    int64_t L_ub; int64_t NT_ub;
    int64_t* n_rows;            // device scalar L
    int32_t* row_pos; int32_t* counts; int32_t* row_depth; uint8_t* row_flag;
    int32_t* head_cnt; int32_t* tail_cnt; Int2* skipdiff; int32_t* max_skip;
    int32_t* row_inscnt; int32_t* row_delcnt;
    // ---- bins: binc = [tile counts (NT_ub+1)] [event counts (L_ub+1)]
    int32_t* binc; int32_t* bin_cur; SegEntry* entries; IndelEvent* events;
    int64_t entries_ub, events_ub;
    // ---- candidates
    int64_t* n_cand; int32_t* cand_row; int64_t cand_cap;
    int32_t* cand_pos; int32_t* cand_depth;
    int32_t* tensor;            // [n_cand][33][C]
    int64_t* alt_off; int32_t* alt_n; AltEntry* alt; int64_t alt_cap;
    // padding pass scratch
    Int2* cur_ref; uint8_t* deleted;
    int32_t* err;               // device error flag
};

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ bool op_consumes_ref(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
__device__ __forceinline__ bool op_consumes_qry(uint32_t op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }
__device__ __forceinline__ bool op_is_match(uint32_t op) { return op == 0 || op == 7 || op == 8; }

// rank of position p (0-based genome coordinate, inside the region) among rows: number of row bits below p
__device__ __forceinline__ int32_t row_lower_bound(const Dev& d, int32_t p) {
    if (p <= d.R0) return 0;
    int64_t o = (int64_t)p - d.R0;
    if (o >= d.W) o = d.W;
    const int64_t w = o >> 5;
    const uint32_t m = (1u << (o & 31)) - 1u;
    return d.word_base[w] + __popc(d.rowR[w] & m);
}
__device__ __forceinline__ bool bit_at(const uint32_t* bm, int64_t o) { return (bm[o >> 5] >> (o & 31)) & 1u; }

__device__ __forceinline__ uint32_t nib_at(const uint8_t* seq, uint32_t q) {
    const uint32_t b = seq[q >> 1];
    return (q & 1u) ? (b & 15u) : (b >> 4);
}
// reference base at genome position p -> channel index 0..3 (non-ACGT -> 0 like evc_base_from), acgt flag
__device__ __forceinline__ int ref_index(const Dev& d, int32_t p, bool* is_acgt) {
    const int64_t o = (int64_t)p - d.ref_start0;
    uint8_t c = (o >= 0 && o < d.ref_len) ? d.ref[o] : (uint8_t)'N';
    if (c >= 'a') c -= 32;
    int idx = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
    *is_acgt = idx >= 0;
    return idx < 0 ? 0 : idx;
}

// set bits [a, b) (region offsets, already clipped) in bitmap bm; interior whole words go
// through the word-level difference array (one +1/-1 pair instead of a store per word)
__device__ __forceinline__ void mark_range(uint32_t* bm, int32_t* wdiff_field, int stride, int64_t a, int64_t b) {
    if (b <= a) return;
    const int64_t wa = a >> 5, wb = (b - 1) >> 5;
    const uint32_t ma = 0xffffffffu << (a & 31);
    const uint32_t mb = 0xffffffffu >> (31 - ((b - 1) & 31));
    if (wa == wb) { atomicOr(&bm[wa], ma & mb); return; }
    atomicOr(&bm[wa], ma);
    atomicOr(&bm[wb], mb);
    if (wb > wa + 1) {
        atomicAdd(&wdiff_field[(wa + 1) * stride], 1);
        atomicAdd(&wdiff_field[wb * stride], -1);
    }
}

// ---------------------------------------------------------------- K0: reads
// admit flag per read (samtools mpileup filters) and segment heads
__global__ void k_read_prepare(Dev d) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= d.n_reads) return;
    const uint32_t f = d.flag[r];
    bool ok = !(f & (0x4u | 0x100u | 0x200u | 0x400u)) && !(f & d.excl) && (int)d.mapq[r] >= d.min_mq;
    if ((f & 0x1u) && !(f & 0x2u)) ok = false;
    const int32_t a = d.cigar_off[r], b = d.cigar_off[r + 1];
    if (b <= a) ok = false;
    d.admit[r] = ok ? 1 : 0;
    d.read_end[r] = d.pos[r];
    if (b > a) d.op_head[a] = (int32_t)r;
}

// ------------------------------------------------- K1: CIGAR segmented scan
// element k = CIGAR op k.  The inclusive segmented scan yields, per op, the reference
// and query lengths consumed by the read up to and including the op; subtracting the
// op's own lengths gives its start offsets.  The store marks coverage bitmaps.
struct OpCigar {
    typedef ScanElem T;
    Dev d;
    __device__ T identity() const { T t; t.v = 0; t.rid = -1; t.flag = 0; return t; }
    __device__ T combine(const T& a, const T& b) const {
        if (b.flag) return b;
        T t; t.v = a.v + b.v; t.rid = a.rid; t.flag = a.flag; return t;
    }
    __device__ int64_t size() const { return d.n_ops; }
    __device__ T load(int64_t k) const {
        const uint32_t c = d.cigar[k];
        const uint32_t op = c & 15u, len = c >> 4;
        T t;
        t.v = ((unsigned long long)(op_consumes_ref(op) ? len : 0u) << 32) | (op_consumes_qry(op) ? len : 0u);
        const int32_t h = d.op_head[k];
        t.rid = h; t.flag = h >= 0 ? 1 : 0;
        return t;
    }
    __device__ void store(int64_t k, const T& incl, const T& own) const {
        const int32_t r = incl.rid;
        const uint32_t rl = (uint32_t)(own.v >> 32), ql = (uint32_t)own.v;
        const uint32_t rx = (uint32_t)(incl.v >> 32) - rl, qy = (uint32_t)incl.v - ql;
        const int32_t x = d.pos[r] + (int32_t)rx;
        d.op_x[k] = x;
        d.op_y[k] = (uint32_t)d.seq_off[r] + qy;
        d.op_rid[k] = r;
        if (!d.admit[r]) return;
        const uint32_t op = d.cigar[k] & 15u;
        if (rl && (op_is_match(op) || op == 2)) {
            int64_t a = (int64_t)x - d.R0, b = a + rl;
            if (a < 0) a = 0;
            if (b > d.W) b = d.W;
            mark_range(d.covE, &d.wdiff[0].b, 2, a, b);
        }
        if (k + 1 == d.cigar_off[r + 1]) {          // last op: the read's whole span
            const int32_t end = d.pos[r] + (int32_t)(incl.v >> 32);
            d.read_end[r] = end;
            int64_t a = (int64_t)d.pos[r] - d.R0, b = (int64_t)end - d.R0;
            if (a < 0) a = 0;
            if (b > d.W) b = d.W;
            mark_range(d.covA, &d.wdiff[0].a, 2, a, b);
        }
    }
};

// ------------------------------------- S2: whole-word coverage from the diff
struct OpWords {
    typedef Int2 T;
    Dev d;
    __device__ T identity() const { T t; t.a = 0; t.b = 0; return t; }
    __device__ T combine(const T& x, const T& y) const { T t; t.a = x.a + y.a; t.b = x.b + y.b; return t; }
    __device__ int64_t size() const { return d.NW; }
    __device__ T load(int64_t w) const { return d.wdiff[w]; }
    __device__ void store(int64_t w, const T& incl, const T&) const {
        if (incl.a > 0) d.covA[w] = 0xffffffffu;
        if (incl.b > 0) d.covE[w] = 0xffffffffu;
    }
};

// -------------------------------------------- S3: rows = dilate16(E) & A, rank
__device__ __forceinline__ uint32_t row_word(const Dev& d, int64_t w) {
    const uint32_t e0 = w > 0 ? d.covE[w - 1] : 0u, e1 = d.covE[w], e2 = (w + 1 < d.NW) ? d.covE[w + 1] : 0u;
    // 96-bit smear by 16 to the left and to the right; only the middle word is needed
    // left smear (towards higher positions): bit i set if any of bits i-16..i set
    unsigned long long lo = ((unsigned long long)e1 << 32) | e0;      // positions of (w-1, w)
    unsigned long long hi = ((unsigned long long)e2 << 32) | e1;      // positions of (w, w+1)
    unsigned long long up = lo;                                       // smear towards higher bits
    up |= up << 1; up |= up << 2; up |= up << 4; up |= up << 8; up |= up << 1;   // 1+2+4+8+1 = 16
    unsigned long long dn = hi;                                       // smear towards lower bits
    dn |= dn >> 1; dn |= dn >> 2; dn |= dn >> 4; dn |= dn >> 8; dn |= dn >> 1;
    const uint32_t dil = (uint32_t)(up >> 32) | (uint32_t)dn;
    uint32_t a = d.covA[w];
    // clear bits beyond the region end
    const int64_t last = d.W - (w << 5);
    if (last < 32) a &= last <= 0 ? 0u : (0xffffffffu >> (32 - last));
    return dil & a;
}
struct OpRows {
    typedef int32_t T;
    Dev d;
    __device__ T identity() const { return 0; }
    __device__ T combine(const T& x, const T& y) const { return x + y; }
    __device__ int64_t size() const { return d.NW; }
    __device__ T load(int64_t w) const { return __popc(row_word(d, w)); }
    __device__ void store(int64_t w, const T& incl, const T& own) const {
        uint32_t bits = row_word(d, w);
        int32_t base = incl - own;
        d.rowR[w] = bits;
        d.word_base[w] = base;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (base < d.L_ub) d.row_pos[base] = d.R0 + (int32_t)(w << 5) + b;
            ++base;
        }
        if (w == d.NW - 1) {
            *d.n_rows = incl;
            if (incl > d.L_ub) atomicExch(d.err, 1);
        }
    }
};

// --------------------------------------------------- bins: count / fill pass
// One thread per CIGAR op.  FILL=false counts tile entries and events, FILL=true writes them.
template <bool FILL>
__global__ void k_bin(Dev d) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.n_ops) return;
    const int32_t r = d.op_rid[k];
    if (r < 0 || !d.admit[r]) return;
    const uint32_t c = d.cigar[k];
    const uint32_t op = c & 15u;
    const int32_t len = (int32_t)(c >> 4);
    if (!op_consumes_ref(op) || len == 0) return;
    const int32_t x = d.op_x[k];
    const uint32_t rev = (d.flag[r] >> 4) & 1u;
    uint32_t hp = d.hp[r];
    hp = hp == 1 ? 1u : hp == 2 ? 2u : 0u;
    int32_t a = x < d.R0 ? d.R0 : x;
    int32_t b = x + len > d.R1 ? d.R1 : x + len;
    int32_t* tile_cnt = d.binc;
    int32_t* ev_cnt = d.binc + d.NT_ub + 1;

    if (op == 3) {                                   // N: `>` / `<` over [a, b)
        if (d.padding && b > a) {
            const int32_t ra = row_lower_bound(d, a), rb = row_lower_bound(d, b);
            if (!FILL && rb > ra) {
                int32_t* f = rev ? &d.skipdiff[0].b : &d.skipdiff[0].a;
                atomicAdd(&f[2 * (int64_t)ra], 1);
                atomicAdd(&f[2 * (int64_t)rb], -1);
            }
        }
    } else if (b > a) {                              // M/=/X or D segment
        const int32_t ra = row_lower_bound(d, a);
        const int32_t rb = ra + (b - a) - 1;
        for (int32_t t = ra >> 5; t <= (rb >> 5); ++t) {
            if (!FILL) {
                atomicAdd(&tile_cnt[t], 1);
            } else {
                const int32_t slot = tile_cnt[t] + atomicAdd(&d.bin_cur[t], 1);
                SegEntry e;
                e.x = x;
                e.len = op == 2 ? 0 : len;
                e.yx = op == 2 ? (uint32_t)len : d.op_y[k] - (uint32_t)x;
                e.info = ((uint32_t)r << 4) | (op == 2 ? 8u : 0u) | (hp << 1) | rev;
                if (slot < d.entries_ub) d.entries[slot] = e; else atomicExch(d.err, 2);
            }
        }
    }
    // head / tail marks (only consumed by the padding rule)
    if (d.padding && !FILL) {
        if (x == d.pos[r] && x >= d.R0 && x < d.R1 && op != 3) {
            // first reference-consuming op of the read
            bool first = true;
            for (int32_t j = d.cigar_off[r]; j < k; ++j)
                if (op_consumes_ref(d.cigar[j] & 15u) && (d.cigar[j] >> 4)) first = false;
            if (first) atomicAdd(&d.head_cnt[row_lower_bound(d, x)], 1);
        }
        const int32_t end = d.read_end[r];
        if (x + len == end && end - 1 >= d.R0 && end - 1 < d.R1 && op != 3)
            atomicAdd(&d.tail_cnt[row_lower_bound(d, end - 1)], 1);
    }
    // indel tokens attach to the last column of an M/=/X/D op (htslib resolve_cigar2)
    if (op == 3) return;
    const int32_t last = x + len - 1;
    if (last < d.R0 || last >= d.R1) return;
    const int32_t kend = d.cigar_off[r + 1];
    int64_t j = k + 1;
    if (j >= kend) return;
    int32_t ins_len = 0, del_len = 0;
    uint32_t op2 = d.cigar[j] & 15u;
    if (op2 == 1) {
        for (; j < kend; ++j) {
            const uint32_t o = d.cigar[j] & 15u;
            if (o == 1) ins_len += (int32_t)(d.cigar[j] >> 4);
            else if (o != 6) break;
        }
    }
    if (j < kend && (d.cigar[j] & 15u) == 2 && op != 2) {
        for (; j < kend; ++j) {
            if ((d.cigar[j] & 15u) == 2) del_len += (int32_t)(d.cigar[j] >> 4);
            else break;
        }
    }
    if (!ins_len && !del_len) return;
    const int32_t row = row_lower_bound(d, last);
    for (int which = 0; which < 2; ++which) {
        const int32_t l = which ? del_len : ins_len;
        if (!l) continue;
        if (!FILL) {
            atomicAdd(&ev_cnt[row], 1);
        } else {
            const int64_t base = (int64_t)ev_cnt[row] - ev_cnt[0];
            const int32_t slot = (int32_t)base + atomicAdd(&d.bin_cur[d.NT_ub + 1 + row], 1);
            IndelEvent e;
            e.info = ((uint32_t)r << 4) | (hp << 2) | ((uint32_t)which << 1) | rev;
            e.len = l;
            // inserted bases start right after this op's query bases (M) or at its query position (D)
            e.y = d.op_y[k] + (op_is_match(op) ? (uint32_t)len : 0u);
            e.pad = 0;
            if (slot < d.events_ub) d.events[slot] = e; else atomicExch(d.err, 3);
        }
    }
}

// exclusive scans of the bin counters in place; sizes come from the device-side row count
struct OpTiles {
    typedef int32_t T;
    Dev d;
    __device__ T identity() const { return 0; }
    __device__ T combine(const T& x, const T& y) const { return x + y; }
    __device__ int64_t size() const { return (*d.n_rows + TILE_ROWS - 1) / TILE_ROWS + 1; }
    __device__ T load(int64_t i) const { return d.binc[i]; }
    __device__ void store(int64_t i, const T& incl, const T& own) const { d.binc[i] = incl - own; }
};
struct OpEvents {
    typedef int32_t T;
    Dev d;
    __device__ T identity() const { return 0; }
    __device__ T combine(const T& x, const T& y) const { return x + y; }
    __device__ int64_t size() const { return *d.n_rows + 1; }
    __device__ T load(int64_t i) const { return d.binc[d.NT_ub + 1 + i]; }
    __device__ void store(int64_t i, const T& incl, const T& own) const { d.binc[d.NT_ub + 1 + i] = incl - own; }
};
// zero the live part of the row-space accumulators once the row count is known
__global__ void k_clear_rows(Dev d) {
    const int64_t L = *d.n_rows;
    const int64_t nt = (L + TILE_ROWS - 1) / TILE_ROWS + 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= L + 1; i += stride) {
        if (i < nt) { d.binc[i] = 0; d.bin_cur[i] = 0; }
        d.binc[d.NT_ub + 1 + i] = 0;
        d.bin_cur[d.NT_ub + 1 + i] = 0;
        if (d.padding) {
            d.head_cnt[i] = 0; d.tail_cnt[i] = 0;
            Int2 z; z.a = 0; z.b = 0;
            d.skipdiff[i] = z;
            d.deleted[i] = 0;
        }
    }
}

struct OpSkip {
    typedef Int2 T;
    Dev d;
    __device__ T identity() const { T t; t.a = 0; t.b = 0; return t; }
    __device__ T combine(const T& x, const T& y) const { T t; t.a = x.a + y.a; t.b = x.b + y.b; return t; }
    __device__ int64_t size() const { return *d.n_rows; }
    __device__ T load(int64_t i) const { return d.skipdiff[i]; }
    __device__ void store(int64_t i, const T& incl, const T&) const {
        // '>' forward skips, '<' reverse skips, '^' heads, '$' tails (create_tensor_pileup.py:178)
        int32_t m = incl.a > incl.b ? incl.a : incl.b;
        const int32_t h = d.head_cnt[i], t = d.tail_cnt[i];
        m = m > h ? m : h;
        m = m > t ? m : t;
        d.max_skip[i] = m;
    }
};

// ------------------------------------------------------------- K2: counting
// One warp per 32-row tile, lane = row.  Each lane walks the tile's segment list and
// accumulates its own column in registers (four 16-bit fields per 64-bit word), so the
// histogram needs no atomics at all; rows leave through shared memory as coalesced
// 16-byte stores.
__device__ __forceinline__ bool ins_equal(const Dev& d, const IndelEvent& a, const IndelEvent& b, bool fold_strand) {
    if (a.len != b.len) return false;
    if (!fold_strand && ((a.info ^ b.info) & 1u)) return false;
    for (int32_t i = 0; i < a.len; ++i)
        if (nib_at(d.seq, a.y + i) != nib_at(d.seq, b.y + i)) return false;
    return true;
}

// first-occurrence key (read ordinal * 2 + is_indel_token) of each allele class at a row;
// only needed to break count ties the way Counter/dict insertion order does.
__device__ void first_keys(const Dev& d, int32_t row, int32_t p, uint32_t key[6]) {
    for (int i = 0; i < 6; ++i) key[i] = 0xffffffffu;
    const int32_t t = row >> 5;
    for (int32_t s = d.binc[t]; s < d.binc[t + 1]; ++s) {
        const SegEntry e = d.entries[s];
        if (e.info & 8u) continue;
        const uint32_t off = (uint32_t)(p - e.x);
        if (off >= (uint32_t)e.len) continue;
        const uint32_t nib = nib_at(d.seq, e.yx + (uint32_t)p);
        if (__popc(nib) != 1) continue;
        const int c = __ffs(nib) - 1;
        const uint32_t kk = (e.info >> 4) * 2u;
        if (kk < key[c]) key[c] = kk;
    }
    const int32_t* ev_off = d.binc + d.NT_ub + 1;
    const int32_t e0 = ev_off[row] - ev_off[0], e1 = ev_off[row + 1] - ev_off[0];
    for (int32_t s = e0; s < e1; ++s) {
        const IndelEvent e = d.events[s];
        const int c = (e.info & 2u) ? 5 : 4;
        const uint32_t kk = (e.info >> 4) * 2u + 1u;
        if (kk < key[c]) key[c] = kk;
    }
}

// allele-frequency thresholds as integer tables: thr[depth] = smallest count c >= 1 with
// (double)c / (double)max(depth, 1) >= af, the reference's float64 test (create_tensor_pileup.py:270-277)
// evaluated once per depth instead of once per row and class.  Depths >= THR_N use the division.
constexpr int THR_N = 4096;
__global__ void k_thr_table(uint16_t* thr_snp, uint16_t* thr_indel, double snp_af, double indel_af) {
    const int dpt = blockIdx.x * blockDim.x + threadIdx.x;
    if (dpt >= THR_N) return;
    const double den = dpt > 0 ? (double)dpt : 1.0;
    for (int which = 0; which < 2; ++which) {
        const double af = which ? indel_af : snp_af;
        int c = 1;
        const int lim = dpt > 1 ? dpt : 1;
        while (c <= lim && !((double)c / den >= af)) ++c;
        (which ? thr_indel : thr_snp)[dpt] = (uint16_t)(c <= lim ? c : 0xffff);
    }
}

// four 8-bit fields -> four 16-bit fields
__device__ __forceinline__ unsigned long long widen8(uint32_t v) {
    const uint32_t lo = __byte_perm(v, 0, 0x4140), hi = __byte_perm(v, 0, 0x4342);
    return ((unsigned long long)hi << 32) | lo;
}

// ------------------------------------------------------------- K2: counting (design)
// One warp per 32-row tile, lane = row, no atomics.  Per round of <= 32 segment entries:
//   staging  (lane = entry)  each lane cuts ITS read's 32-base window over the tile's positions out of the
//            4-bit sequence (5 word loads, nibble swap, funnel shift), clears bases outside the segment and
//            ambiguity codes, and drops the 16 bytes into a shared-memory slot; slots are grouped by
//            (strand, phase class) with each group padded to a multiple of 4 zero slots;
//   consume  (lane = row)    per slot: one shared load, shift, mask, multiply-spread of the one-hot nibble
//            into four 8-bit fields, add - 6 instructions per (row, read) instead of an address
//            computation and a dependent global load per (row, read).
// Tiles whose rows are not consecutive positions (a run ends inside the tile) are processed run by run.
constexpr int COUNT_WARPS = 8;
constexpr int SLOT_WORDS = 8;                // 4 data words, word 4 = 0 (lanes outside the run read it), 3 spare

__device__ __forceinline__ uint32_t nibmask(int n) {         // low n nibbles set, n clamped to 0..8
    uint32_t m;                                              // shl.b32 clamps the shift amount: 1 << 32 == 0
    asm("shl.b32 %0, 1, %1;" : "=r"(m) : "r"((uint32_t)(n < 0 ? 0 : 4 * n)));
    return m - 1u;
}
// keep only one-hot nibbles (A=1, C=2, G=4, T=8); N, other ambiguity codes and '=' count nothing
__device__ __forceinline__ uint32_t keep_onehot(uint32_t x) {
    const uint32_t c = x - ((x >> 1) & 0x77777777u) - ((x >> 2) & 0x33333333u) - ((x >> 3) & 0x11111111u);
    const uint32_t t = c ^ 0x11111111u;                      // nibble == 0  <=>  popcount == 1
    const uint32_t nz = (t | (t >> 1) | (t >> 2) | (t >> 3)) & 0x11111111u;
    return x & ((nz ^ 0x11111111u) * 15u);
}

template <int C>
__global__ void __launch_bounds__(COUNT_WARPS * 32) k_count(Dev d) {
    constexpr int NG = C == 30 ? 6 : 2;                      // slot groups: strand x {unphased, HP1, HP2}
    constexpr int NSLOT = 32 + 4 * NG;
    __shared__ __align__(16) int32_t stage[COUNT_WARPS][TILE_ROWS * C];
    __shared__ __align__(16) uint32_t slots[COUNT_WARPS][NSLOT][SLOT_WORDS];
    __shared__ uint2 dels[COUNT_WARPS][32];                  // deletion entries of the round: coverage mask, strand
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int64_t L = *d.n_rows;
    const int64_t n_tiles = (L + TILE_ROWS - 1) / TILE_ROWS;
    const int32_t* ev_off = d.binc + d.NT_ub + 1;
    const uint32_t* __restrict__ seqw = (const uint32_t*)d.seq;
    const int64_t n_seq_words = d.n_seq_words;
    for (int j = lane; j < NSLOT; j += 32) {                 // words 4..7 of every slot stay zero
        *(uint4*)&slots[warp][j][0] = make_uint4(0, 0, 0, 0);
        *(uint4*)&slots[warp][j][4] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    const int64_t t_stride = (int64_t)gridDim.x * COUNT_WARPS;
    int64_t t = (int64_t)blockIdx.x * COUNT_WARPS + warp;
    // software pipeline across tiles: the entry range of tile t+2 and the first round's entries of tile t+1
    // are requested while tile t is processed
    int32_t s0 = 0, s1 = 0, s0n = 0, s1n = 0;
    if (t < n_tiles) { s0 = d.binc[t]; s1 = d.binc[t + 1]; }
    if (t + t_stride < n_tiles) { s0n = d.binc[t + t_stride]; s1n = d.binc[t + t_stride + 1]; }
    SegEntry pre;
    pre.x = 0; pre.len = 0; pre.yx = 0; pre.info = 0;
    if (s0 + lane < s1) pre = d.entries[s0 + lane];
    for (; t < n_tiles; t += t_stride) {
        const int64_t row = t * TILE_ROWS + lane;
        const bool live = row < L;
        const int32_t p = live ? d.row_pos[row] : -0x40000000;
        // epilogue operands requested now, consumed after the rounds
        int32_t ev0 = 0, ev1 = 0;
        uint8_t ref_c = (uint8_t)'N';
        if (live) {
            ev0 = ev_off[row]; ev1 = ev_off[row + 1];
            const int64_t ro = (int64_t)p - d.ref_start0;
            if (ro >= 0 && ro < d.ref_len) ref_c = d.ref[ro];
        }
        const int32_t ev_base = ev_off[0];
        const SegEntry first = pre;
        pre.x = 0; pre.len = 0; pre.yx = 0; pre.info = 0;
        if (s0n + lane < s1n) pre = d.entries[s0n + lane];
        const int64_t tnn = t + 2 * t_stride;
        int32_t s0nn = 0, s1nn = 0;
        if (tnn < n_tiles) { s0nn = d.binc[tnn]; s1nn = d.binc[tnn + 1]; }
        // runs of consecutive positions inside the tile
        const int32_t p_prev = __shfl_up_sync(0xffffffffu, p, 1);
        const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        const uint32_t start_mask = __ballot_sync(0xffffffffu, live && (lane == 0 || p != p_prev + 1));
        const int n_live = __popc(live_mask);
        // wide accumulators: A,C,G,T as four 16-bit fields (forward, reverse, HP=1, HP=2); `*` / `#` counts
        unsigned long long wf = 0, wr = 0, wp = 0, wm = 0;
        uint32_t wst = 0;                                    // star_f | star_r << 16
        // narrow accumulators (four 8-bit fields), folded into the wide ones every 7 rounds (<= 224 per field)
        uint32_t nf = 0, nr = 0, np = 0, nm = 0, nst = 0;    // nst: star_f | star_r << 8
        int rounds = 0;
        for (uint32_t sm = start_mask; sm; sm &= sm - 1) {
            const int ra = __ffs(sm) - 1;
            const uint32_t rest = sm & (sm - 1);
            const int rb = rest ? __ffs(rest) - 1 : n_live;
            const int32_t P0 = __shfl_sync(0xffffffffu, p, ra);
            const int nrun = rb - ra;
            const bool in_run = lane >= ra && lane < rb;
            const int k = lane - ra;                         // this row's base index inside the window
            const uint32_t* lane_slot = &slots[warp][0][in_run ? (k >> 3) : 4];
            const uint32_t lane_sh = (uint32_t)(k & 7) * 4u;
            const uint32_t lane_bit = in_run ? (1u << k) : 0u;
            for (int32_t s = s0; s < s1; s += 32) {
                SegEntry e;
                if (s == s0 && sm == start_mask) e = first;
                else {
                    e.x = 0; e.len = 0; e.yx = 0; e.info = 0;
                    if (s + lane < s1) e = d.entries[s + lane];
                }
                // ---------------------------------------------------------------- staging (lane = entry)
                const bool is_del = (e.info & 8u) != 0u;
                const int32_t span = is_del ? (int32_t)e.yx : e.len;
                int k0 = e.x - P0, k1 = e.x + span - P0;     // window bases covered by the op: [k0, k1)
                k0 = k0 < 0 ? 0 : k0;
                k1 = k1 > nrun ? nrun : k1;
                const bool hit = k1 > k0;
                uint4 w = make_uint4(0, 0, 0, 0);
                if (hit && !is_del) {
                    const long long q0 = (long long)(uint32_t)(e.yx + (uint32_t)e.x) + ((long long)P0 - e.x);
                    const long long wb = q0 >> 3;            // floor: q0 is negative when the op starts inside the window
                    const uint32_t sh = ((uint32_t)q0 & 7u) * 4u;
                    uint32_t v[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const long long wi = wb + i;
                        uint32_t x = (wi >= 0 && wi < n_seq_words) ? seqw[wi] : 0u;
                        v[i] = ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);     // base j of the word at bits 4j..4j+3
                    }
                    w.x = keep_onehot(__funnelshift_r(v[0], v[1], sh) & nibmask(k1) & ~nibmask(k0));
                    w.y = keep_onehot(__funnelshift_r(v[1], v[2], sh) & nibmask(k1 - 8) & ~nibmask(k0 - 8));
                    w.z = keep_onehot(__funnelshift_r(v[2], v[3], sh) & nibmask(k1 - 16) & ~nibmask(k0 - 16));
                    w.w = keep_onehot(__funnelshift_r(v[3], v[4], sh) & nibmask(k1 - 24) & ~nibmask(k0 - 24));
                }
                const uint32_t rev = e.info & 1u;
                int grp = (int)rev;
                if (C == 30) {
                    const uint32_t hp = (e.info >> 1) & 3u;
                    grp = (int)rev * 3 + (hp == 1 ? 1 : hp == 2 ? 2 : 0);
                }
                int end_all = 0;
                int cnt4g[NG];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, hit && !is_del && grp == g);
                    const int cnt = __popc(bal), cnt4 = (cnt + 3) & ~3;
                    if (hit && !is_del && grp == g) *(uint4*)&slots[warp][end_all + __popc(bal & lt_mask)][0] = w;
                    if (lane < cnt4 - cnt) *(uint4*)&slots[warp][end_all + cnt + lane][0] = make_uint4(0, 0, 0, 0);
                    end_all += cnt4;
                    cnt4g[g] = cnt4;
                }
                const uint32_t dbal = __ballot_sync(0xffffffffu, hit && is_del);
                if (hit && is_del) {
                    const uint32_t cov = (k1 >= 32 ? 0xffffffffu : ((1u << k1) - 1u)) & ~((1u << k0) - 1u);
                    dels[warp][__popc(dbal & lt_mask)] = make_uint2(cov, rev);
                }
                __syncwarp();
                // ---------------------------------------------------------------- consume (lane = row)
                {
                    int j = 0;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        uint32_t acc = 0;
                        for (const int je = j + cnt4g[g]; j < je; j += 4) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const uint32_t nib = (lane_slot[(j + u) * SLOT_WORDS] >> lane_sh) & 15u;
                                acc += (nib * 0x204081u) & 0x01010101u;
                            }
                        }
                        if (C == 30) {
                            if (g < 3) nf += acc; else nr += acc;
                            if (g % 3 == 1) np += acc; else if (g % 3 == 2) nm += acc;
                        } else {
                            if (g == 0) nf += acc; else nr += acc;
                        }
                    }
                    const int nd = __popc(dbal);
                    for (int i = 0; i < nd; ++i) {
                        const uint2 dl = dels[warp][i];
                        if (dl.x & lane_bit) nst += 1u << (dl.y * 8u);
                    }
                }
                __syncwarp();
                if (++rounds == 7) {
                    wf += widen8(nf); wr += widen8(nr);
                    if (C == 30) { wp += widen8(np); wm += widen8(nm); }
                    wst += (nst & 0xffu) | ((nst & 0xff00u) << 8);
                    nf = nr = np = nm = nst = 0;
                    rounds = 0;
                }
            }
        }
        wf += widen8(nf); wr += widen8(nr);
        if (C == 30) { wp += widen8(np); wm += widen8(nm); }
        wst += (nst & 0xffu) | ((nst & 0xff00u) << 8);
        s0 = s0n; s1 = s1n; s0n = s0nn; s1n = s1nn;
        // the row vector lives in this lane's slice of the staging tile (dynamic channel
        // indices would otherwise force a register array into local memory)
        int32_t* v = &stage[warp][lane * C];
#pragma unroll
        for (int i = 0; i < C; ++i) v[i] = 0;
        if (live) {
            const int32_t star_f = (int32_t)(wst & 0xffffu), star_r = (int32_t)(wst >> 16);
            int32_t bf[4], br[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bf[i] = (int32_t)((wf >> (16 * i)) & 0xffff);
                br[i] = (int32_t)((wr >> (16 * i)) & 0xffff);
                v[i] = bf[i];
                v[9 + i] = br[i];
            }
            v[8] = star_f; v[17] = star_r;
            if (C == 30) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[18 + i] = (int32_t)((wp >> (16 * i)) & 0xffff);
                    v[24 + i] = (int32_t)((wm >> (16 * i)) & 0xffff);
                }
            }
            // indel events of this row: per strand totals and the largest distinct allele
            const int32_t e0 = ev0 - ev_base, e1 = ev1 - ev_base;
            int32_t ins_cnt = 0, del_cnt = 0;
            for (int32_t s = e0; s < e1; ++s) {
                const IndelEvent e = d.events[s];
                const bool is_del = e.info & 2u, rev = e.info & 1u;
                const uint32_t hp = (e.info >> 2) & 3u;
                if (is_del) { ++del_cnt; ++v[rev ? 15 : 6]; } else { ++ins_cnt; ++v[rev ? 13 : 4]; }
                if (C == 30) {
                    if (hp == 1) ++v[is_del ? 23 : 22]; else if (hp == 2) ++v[is_del ? 29 : 28];
                }
                bool seen = false;
                for (int32_t q = e0; q < s && !seen; ++q) {
                    const IndelEvent o = d.events[q];
                    if (((o.info ^ e.info) & 3u) == 0 && (is_del ? o.len == e.len : ins_equal(d, o, e, false))) seen = true;
                }
                if (seen) continue;
                int32_t same = 1;
                for (int32_t q = s + 1; q < e1; ++q) {
                    const IndelEvent o = d.events[q];
                    if (((o.info ^ e.info) & 3u) == 0 && (is_del ? o.len == e.len : ins_equal(d, o, e, false))) ++same;
                }
                const int ch = is_del ? (rev ? 16 : 7) : (rev ? 14 : 5);
                if (same > v[ch]) v[ch] = same;
            }
            const int32_t fsum = bf[0] + bf[1] + bf[2] + bf[3], rsum = br[0] + br[1] + br[2] + br[3];
            const int32_t depth = fsum + rsum + star_f + star_r;
            uint8_t rc = ref_c;
            if (rc >= 'a') rc -= 32;
            const int ri_raw = rc == 'A' ? 0 : rc == 'C' ? 1 : rc == 'G' ? 2 : rc == 'T' ? 3 : -1;
            const bool acgt = ri_raw >= 0;
            const int ri = acgt ? ri_raw : 0;
            // candidate predicate (create_tensor_pileup.py:268-299, 536, 555)
            int32_t cls[6];
#pragma unroll
            for (int i = 0; i < 4; ++i) cls[i] = bf[i] + br[i];
            cls[4] = ins_cnt; cls[5] = del_cnt;
            const int32_t cls_ref = ri == 0 ? cls[0] : ri == 1 ? cls[1] : ri == 2 ? cls[2] : cls[3];
            bool pass_snp = false, pass_indel = false;
            if (depth < THR_N) {
                const int32_t ts = d.thr_snp[depth], ti = d.thr_indel[depth];
#pragma unroll
                for (int i = 0; i < 4; ++i) if (i != ri && cls[i] >= ts) pass_snp = true;
                pass_indel = cls[4] >= ti || cls[5] >= ti;
            } else {
                const double den = (double)depth;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i != ri && cls[i] > 0 && (double)cls[i] / den >= d.snp_af) pass_snp = true;
                if (cls[4] > 0 && (double)cls[4] / den >= d.indel_af) pass_indel = true;
                if (cls[5] > 0 && (double)cls[5] / den >= d.indel_af) pass_indel = true;
            }
            bool pass = pass_snp || pass_indel;
            if (!pass) {
                int32_t top = 0;
#pragma unroll
                for (int i = 0; i < 6; ++i) top = cls[i] > top ? cls[i] : top;
                if (top > 0) {
                    if (cls_ref < top) pass = true;
                    else {
                        bool tie = false;
#pragma unroll
                        for (int i = 0; i < 6; ++i) if (i != ri && cls[i] == top) tie = true;
                        if (tie) {
                            uint32_t key[6];
                            first_keys(d, (int32_t)row, p, key);
                            for (int i = 0; i < 6; ++i) if (i != ri && cls[i] == top && key[i] < key[ri]) pass = true;
                        }
                    }
                }
            }
            if (depth > 0 && (d.snp_af == 0.0 || d.indel_af == 0.0)) pass = true;
            const uint8_t flag = (acgt && pass && depth >= d.min_cov) ? 1 : 0;
            v[ri] = -fsum;
            v[9 + ri] = -rsum;
            d.row_depth[row] = depth;
            d.row_flag[row] = flag;
            d.row_inscnt[row] = ins_cnt;
            d.row_delcnt[row] = del_cnt + star_f + star_r;
            if (d.padding) { Int2 cr; cr.a = -fsum; cr.b = -rsum; d.cur_ref[row] = cr; }
        }
        // rows of a tile are contiguous in HBM: write 16 B per lane per step from the staging tile
        __syncwarp();
        const int64_t rows_here = min((int64_t)TILE_ROWS, L - t * TILE_ROWS);
        const int n_int = (int)rows_here * C;
        int32_t* out = d.counts + t * TILE_ROWS * C;          // 32*C*4 bytes per tile: 16 B aligned
        for (int i = lane * 4; i < n_int; i += 128) {
            if (i + 4 <= n_int) *(int4*)(out + i) = *(const int4*)(&stage[warp][i]);
            else for (int q = i; q < n_int; ++q) out[q] = stage[warp][q];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------- K3: candidate list
// emitted iff eligible and the 33 positions c-16..c+16 are all covered (one gap-free run):
// restates "pop when pos - c == 16 and the ring has no empty slot" (create_tensor_pileup.py:565-568)
struct OpCand {
    typedef int32_t T;
    Dev d;
    __device__ T identity() const { return 0; }
    __device__ T combine(const T& x, const T& y) const { return x + y; }
    __device__ int64_t size() const { return *d.n_rows; }
    __device__ T load(int64_t row) const {
        if (!d.row_flag[row]) return 0;
        const int64_t o = (int64_t)d.row_pos[row] - d.R0;
        if (o - FLANK < 0 || o + FLANK >= d.W) return 0;
        // 33 consecutive bits of covA starting at o-16
        const int64_t s = o - FLANK;
        const int64_t w = s >> 5;
        const int sh = (int)(s & 31);
        unsigned long long lo = d.covA[w] | ((unsigned long long)d.covA[w + 1] << 32);
        const unsigned long long bits = lo >> sh;          // 64 - sh >= 33 bits are valid
        const unsigned long long need = (1ull << 33) - 1ull;
        return (bits & need) == need ? 1 : 0;
    }
    __device__ void store(int64_t row, const T& incl, const T& own) const {
        if (own) {
            const int64_t i = incl - 1;
            if (i < d.cand_cap) {
                d.cand_row[i] = (int32_t)row;
                d.cand_pos[i] = d.row_pos[row] + 1;
                d.cand_depth[i] = d.row_depth[row];
            }
        }
        if (row == *d.n_rows - 1) *d.n_cand = incl;
    }
};

// ------------------------------------------------------- K4: window assembly
// tensor[i][k][c] = counts[row_i - 16 + k][c]  (rows of consecutive positions are consecutive);
// with depth > 1.5*max_depth the window is divided by depth/max_depth in fp64 and truncated
// (clair3_rna/utils.py:85-92,120).  `scale` is skipped when the padding pass still has to patch.
__device__ __forceinline__ int32_t rescale(int32_t v, int32_t depth, int32_t max_depth) {
    const double sf = __ddiv_rn((double)depth, (double)max_depth);
    return (int32_t)__ddiv_rn((double)v, sf);
}
__global__ void k_window(Dev d, int apply_scale) {
    const int64_t n = *d.n_cand < d.cand_cap ? *d.n_cand : d.cand_cap;
    const int per = WIN * d.C;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const int32_t row = d.cand_row[i];
        const int32_t depth = d.cand_depth[i];
        const bool sc = apply_scale && depth > 0 && (double)depth > (double)d.max_depth * 1.5;
        const int32_t* src = d.counts + ((int64_t)row - FLANK) * d.C;
        int32_t* dst = d.tensor + i * per;
        for (int j = threadIdx.x; j < per; j += blockDim.x) {
            int32_t v = src[j];
            if (sc) v = rescale(v, depth, d.max_depth);
            dst[j] = v;
        }
    }
}
__global__ void k_rescale(Dev d) {
    const int64_t n = *d.n_cand < d.cand_cap ? *d.n_cand : d.cand_cap;
    const int per = WIN * d.C;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const int32_t depth = d.cand_depth[i];
        if (!(depth > 0 && (double)depth > (double)d.max_depth * 1.5)) continue;
        int32_t* dst = d.tensor + i * per;
        for (int j = threadIdx.x; j < per; j += blockDim.x) dst[j] = rescale(dst[j], depth, d.max_depth);
    }
}

// splice-junction padding (create_tensor_pileup.py:573-593, 611): sequential along chains of
// emitted candidates <= 32 bp apart because padded rows persist (shared row objects) and an
// emitted centre disappears from depth_dict.  One thread per chain head.
__global__ void k_padding(Dev d) {
    const int64_t n = *d.n_cand < d.cand_cap ? *d.n_cand : d.cand_cap;
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    if (h > 0 && d.cand_pos[h] - d.cand_pos[h - 1] <= 2 * FLANK && d.cand_row[h] - d.cand_row[h - 1] == d.cand_pos[h] - d.cand_pos[h - 1])
        return;                                      // not a chain head
    const int per = WIN * d.C;
    for (int64_t i = h; i < n; ++i) {
        if (i > h && !(d.cand_pos[i] - d.cand_pos[i - 1] <= 2 * FLANK &&
                       d.cand_row[i] - d.cand_row[i - 1] == d.cand_pos[i] - d.cand_pos[i - 1])) break;
        const int32_t row = d.cand_row[i];
        const int32_t depth = d.cand_depth[i];
        int32_t mx_depth = 0, mx_skip = 0;
        bool any = false;
        for (int k = 0; k < WIN; ++k) {
            const int32_t r = row - FLANK + k;
            if (!d.deleted[r]) { const int32_t dd = d.row_depth[r]; if (!any || dd > mx_depth) mx_depth = dd; any = true; }
            const int32_t ms = d.max_skip[r];
            mx_skip = ms > mx_skip ? ms : mx_skip;
        }
        if (any && mx_depth > 0 && __ddiv_rn((double)mx_skip, (double)mx_depth) > d.skip_prop) {
            const Int2 c = d.cur_ref[row];
            const int32_t sf = c.a < 0 ? -c.a : c.a, sr = c.b < 0 ? -c.b : c.b;
            const double pf = sf + sr > 0 ? __ddiv_rn((double)sf, (double)(sf + sr)) : 0.0;
            const double pr = 1.0 - pf;
            const int32_t vf = -(int32_t)__dmul_rn((double)depth, pf);
            const int32_t vr = -(int32_t)__dmul_rn((double)depth, pr);
            const double thr = __dmul_rn((double)depth, d.skip_prop);
            for (int k = 0; k < WIN; ++k) {
                if (k == FLANK) continue;
                const int32_t r = row - FLANK + k;
                const int32_t cd = d.deleted[r] ? 0 : d.row_depth[r];
                if ((double)cd < thr) {
                    bool acgt;
                    ref_index(d, d.row_pos[r], &acgt);
                    if (acgt) { Int2 nv; nv.a = vf; nv.b = vr; d.cur_ref[r] = nv; }
                }
            }
        }
        int32_t* dst = d.tensor + i * per;
        for (int k = 0; k < WIN; ++k) {
            const int32_t r = row - FLANK + k;
            bool acgt;
            const int ri = ref_index(d, d.row_pos[r], &acgt);
            const Int2 c = d.cur_ref[r];
            dst[k * d.C + ri] = c.a;
            dst[k * d.C + 9 + ri] = c.b;
        }
        d.deleted[row] = 1;
    }
}

// ------------------------------------------------------------- alt_info
// upper bound of alleles per candidate: 3 alt bases + its indel events + reference
struct OpAltOff {
    typedef long long T;
    Dev d;
    __device__ T identity() const { return 0; }
    __device__ T combine(const T& x, const T& y) const { return x + y; }
    __device__ int64_t size() const { return *d.n_cand < d.cand_cap ? *d.n_cand : d.cand_cap; }
    __device__ T load(int64_t i) const {
        const int32_t row = d.cand_row[i];
        const int32_t* ev_off = d.binc + d.NT_ub + 1;
        return 4 + (ev_off[row + 1] - ev_off[row]);
    }
    __device__ void store(int64_t i, const T& incl, const T& own) const { d.alt_off[i] = incl - own; }
};

// alleles in alt_dict insertion order (create_tensor_pileup.py:223,235,251,259-261): keys are
// case-folded, so the two strands of one allele merge and keep the earlier first occurrence.
// One warp per candidate: the lanes split the tile's segment list to find the first read
// showing each base; lane 0 then builds the (short) allele list.
__global__ void __launch_bounds__(256) k_altinfo(Dev d) {
    const int64_t n = *d.n_cand < d.cand_cap ? *d.n_cand : d.cand_cap;
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int32_t row = d.cand_row[i];
    const int32_t p = d.row_pos[row];
    const int64_t off = d.alt_off[i];
    const int32_t* ev_off = d.binc + d.NT_ub + 1;
    const int32_t e0 = ev_off[row] - ev_off[0], e1 = ev_off[row + 1] - ev_off[0];
    // first-occurrence key of each base among the reads covering p
    uint32_t key[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    {
        const int32_t t = row >> 5;
        for (int32_t s = d.binc[t] + lane; s < d.binc[t + 1]; s += 32) {
            const SegEntry e = d.entries[s];
            if (e.info & 8u) continue;
            const uint32_t o = (uint32_t)(p - e.x);
            if (o >= (uint32_t)e.len) continue;
            const uint32_t nib = nib_at(d.seq, e.yx + (uint32_t)p);
            if (__popc(nib) != 1) continue;
            const int c = __ffs(nib) - 1;
            const uint32_t kk = (e.info >> 4) * 2u;
#pragma unroll
            for (int b = 0; b < 4; ++b) if (b == c && kk < key[b]) key[b] = kk;
        }
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) key[b] = min(key[b], __shfl_xor_sync(0xffffffffu, key[b], sft));
    }
    if (lane != 0) return;
    if (off + 4 + (e1 - e0) > d.alt_cap) { atomicExch(d.err, 4); d.alt_n[i] = 0; return; }
    AltEntry* out = d.alt + off;
    int m = 0;
    bool acgt;
    const int ri = ref_index(d, p, &acgt);
    const char L4[4] = {'A', 'C', 'G', 'T'};
    const int32_t* v = d.counts + (int64_t)row * d.C;
    int32_t alt_cnt = 0;
    for (int b = 0; b < 4; ++b) {
        if (b == ri) continue;
        const int32_t c = v[b] + v[9 + b];
        if (c <= 0) continue;
        alt_cnt += c;
        AltEntry e; e.kind = 'X'; e.base = L4[b]; e.len = 0; e.count = c; e.seq_off = 0; e.order = key[b];
        out[m++] = e;
    }
    for (int32_t s = e0; s < e1; ++s) {
        const IndelEvent ev = d.events[s];
        const bool is_del = ev.info & 2u;
        bool seen = false;
        for (int32_t q = e0; q < s && !seen; ++q) {
            const IndelEvent o = d.events[q];
            if (((o.info ^ ev.info) & 2u) == 0 && (is_del ? o.len == ev.len : ins_equal(d, o, ev, true))) seen = true;
        }
        if (seen) continue;
        int32_t cnt = 0;
        uint32_t first = 0xffffffffu, first_y = ev.y;
        for (int32_t q = s; q < e1; ++q) {
            const IndelEvent o = d.events[q];
            if (((o.info ^ ev.info) & 2u) == 0 && (is_del ? o.len == ev.len : ins_equal(d, o, ev, true))) {
                ++cnt;
                const uint32_t kk = (o.info >> 4) * 2u + 1u;
                if (kk < first) { first = kk; first_y = o.y; }
            }
        }
        AltEntry e;
        e.kind = is_del ? 'D' : 'I'; e.base = L4[ri];
        e.len = (uint16_t)(ev.len > 65535 ? 65535 : ev.len);
        e.count = cnt; e.seq_off = first_y; e.order = first;
        out[m++] = e;
    }
    // insertion sort by first occurrence
    for (int a = 1; a < m; ++a) {
        AltEntry e = out[a];
        int b = a - 1;
        while (b >= 0 && out[b].order > e.order) { out[b + 1] = out[b]; --b; }
        out[b + 1] = e;
    }
    const int32_t depth = d.row_depth[row];
    const int32_t ref_cnt = depth - d.row_delcnt[row] - d.row_inscnt[row] - alt_cnt;
    if (ref_cnt > 0) {
        AltEntry e; e.kind = 'R'; e.base = L4[ri]; e.len = 0; e.count = ref_cnt; e.seq_off = 0; e.order = 0xffffffffu;
        out[m++] = e;
    }
    d.alt_n[i] = m;
}

}  // namespace c3r
