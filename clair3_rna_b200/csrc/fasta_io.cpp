// Indexed FASTA region fetch into a caller buffer (host): one pread of the lines that hold the region, line
// terminators dropped and letters upper-cased in one pass.  Replaces the per-chunk `samtools faidx` of
// /root/reference/shared/utils.py:168-194 (reference_sequence_from); the caller's buffer is typically pinned
// host memory that c3r_submit_chunk then reads by DMA.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <vector>
#include "../../include/c3r_b200.h"

extern "C" int c3r_fasta_fetch(const char* path, int64_t contig_len, int64_t offset, int64_t linebases, int64_t linewidth,
                               int64_t start1, int64_t end1, uint8_t* out, int64_t out_cap, int64_t* n_out) {
    if (!path || !out || !n_out || linebases <= 0 || linewidth < linebases) return C3R_ERR_ARG;
    if (start1 < 1) start1 = 1;
    if (end1 > contig_len) end1 = contig_len;
    *n_out = 0;
    if (end1 < start1) return C3R_OK;
    const int64_t n = end1 - start1 + 1;
    if (n > out_cap) return C3R_ERR_CAPACITY;
    const int64_t first_line = (start1 - 1) / linebases, last_line = (end1 - 1) / linebases;
    const int64_t file_off = offset + first_line * linewidth + (start1 - 1 - first_line * linebases);
    const int64_t span = (last_line - first_line) * linewidth + ((end1 - 1) % linebases) + 1 - ((start1 - 1) % linebases);
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return C3R_ERR_ARG;
    // read in pieces of whole lines into a small scratch and compact into `out`
    const int64_t term = linewidth - linebases;
    std::vector<uint8_t> buf((size_t)(1 << 20) + (size_t)linewidth);
    int64_t done = 0, pos_in_line = (start1 - 1) % linebases, fo = file_off, left = span;
    int rc = C3R_OK;
    while (left > 0 && rc == C3R_OK) {
        const int64_t want = left < (int64_t)(1 << 20) ? left : (int64_t)(1 << 20);
        int64_t got = 0;
        while (got < want) {
            const ssize_t r = pread(fd, buf.data() + got, (size_t)(want - got), (off_t)(fo + got));
            if (r <= 0) { rc = C3R_ERR_ARG; break; }
            got += r;
        }
        if (rc != C3R_OK) break;
        int64_t i = 0;
        while (i < got) {
            if (pos_in_line < linebases) {                       // inside the bases of a line
                int64_t take = linebases - pos_in_line;
                if (take > got - i) take = got - i;
                if (done + take > n) take = n - done;
                memcpy(out + done, buf.data() + i, (size_t)take);
                done += take; i += take; pos_in_line += take;
                if (done >= n) break;
            } else {                                             // the line terminator (may straddle two pieces)
                int64_t skip = linebases + term - pos_in_line;
                if (skip > got - i) skip = got - i;
                i += skip; pos_in_line += skip;
                if (pos_in_line >= linebases + term) pos_in_line = 0;
            }
        }
        fo += got; left -= got;
        if (done >= n) break;
    }
    close(fd);
    if (rc != C3R_OK) return rc;
    if (done != n) return C3R_ERR_ARG;
    for (int64_t i = 0; i < n; ++i) {                            // upper-case (auto-vectorised)
        const uint8_t c = out[i];
        out[i] = (uint8_t)((c >= 'a' && c <= 'z') ? c - 32 : c);
    }
    *n_out = n;
    return C3R_OK;
}
