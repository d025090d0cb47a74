"""Multi-GPU sharding of the (contig, chunk) list — the reference's unit of parallelism
(`parallel -j THREADS ... :::: CHUNK_LIST`, /root/reference/run_clair3_rna:681-706).

Shards are independent (no data-path collective): each rank owns one GPU and a subset of the
chunks; only per-shard VCF rows are gathered on rank 0, which merges them with sort_vcf
semantics (rows keyed by POS per contig, later chunks overwrite, /root/reference/src/sort_vcf.py:251).
"""
from __future__ import annotations


def assign(costs, world_size: int) -> list:
    """Greedy longest-processing-time partition: costs[i] (e.g. reads per chunk) -> rank of shard i."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world_size
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return owner


def merge_rows(rows_per_shard, contig_order) -> list:
    """rows_per_shard: list (in CHUNK_LIST order) of lists of VCF lines.  Dedup by (contig, POS) with
    later shards winning, output in contig order then position (sort_vcf.py:172-173,251)."""
    best = {}
    for rows in rows_per_shard:
        for row in rows:
            c = row.split("\t", 2)
            best[(c[0], int(c[1]))] = row
    rank = {c: i for i, c in enumerate(contig_order)}
    keys = sorted(best, key=lambda k: (rank.get(k[0], len(rank)), k[1]))
    return [best[k] for k in keys]


MAJOR_CONTIGS = ["chr" + str(a) for a in list(range(1, 23)) + ["X", "Y"]] + [str(a) for a in list(range(1, 23)) + ["X", "Y"]]


def contig_output_order(contigs) -> list:
    """sort_vcf.py:39-40,170-173: chr1..22,X,Y then 1..22,X,Y, every other contig in the order of the CONTIGS list"""
    order = MAJOR_CONTIGS + list(contigs)
    return sorted(contigs, key=order.index)


def read_rediportal(path: str, contigs=None, filter_tags: str = None) -> dict:
    """REDIportal TABLE1 rows -> {(contig, pos): (ref, alt, db filter)} (sort_vcf.py:175-207): the first row is a
    header, rows of other contigs, with an unreadable position or (filter_tags = 'A,D:A,R:...') with another
    database tag are left out; a later row of the same site replaces an earlier one."""
    import gzip
    tags = set(filter_tags.split(":")) if filter_tags is not None else None
    keep = set(contigs) if contigs else None
    out = {}
    with open(path, "rb") as fp:
        gz = fp.read(2) == b"\x1f\x8b"
    with (gzip.open(path, "rt") if gz else open(path, "rt")) as fp:
        for i, row in enumerate(fp):
            if i == 0:
                continue
            c = row.rstrip().split("\t", maxsplit=6)
            if keep is not None and c[0] not in keep:
                continue
            try:
                key = (c[0], int(c[1]))
            except ValueError:
                continue
            if tags is not None and c[5] not in tags:
                continue
            out[key] = (c[2], c[3], c[5])
    return out


def sort_vcf(rows_per_shard, contigs, qual=None, show_ref=True, rediportal=None):
    """The merge stage of the workflow, sort_vcf_from (src/sort_vcf.py:123-292), on rows instead of files.

    rows_per_shard: lists of VCF data lines in CHUNK_LIST order (a site called by two chunks keeps the later row);
    contigs: the CONTIGS list.  Reference calls (ALT '.' or REF == ALT) are dropped unless show_ref; other rows
    with QUAL <= qual get FILTER LowQual (qual None / 0: untouched); a row whose site, REF and ALT are in the
    REDIportal table gets FILTER RNAEditing unless it says Germline or RefCall.
    -> (rows, rows_without_tagging); the second list is None without a REDIportal table."""
    best = {}
    for rows in rows_per_shard:
        for row in rows:
            c = row.split(maxsplit=6)
            ref, alt = c[3], c[4]
            is_ref = alt == "." or ref == alt
            if is_ref and not show_ref:
                continue
            if not is_ref and qual and float(c[5]) <= qual:
                t = row.split("\t")
                t[6] = "LowQual"
                row = "\t".join(t)
            key = (c[0], int(c[1]))
            if rediportal is not None and key in rediportal and "Germline" not in row and "RefCall" not in row:
                t = row.split("\t", maxsplit=8)
                if rediportal[key][0] == t[3] and rediportal[key][1] == t[4]:
                    t[6] = "RNAEditing"
                    row = "\t".join(t)
            best[key] = row
    rank = {c: i for i, c in enumerate(contig_output_order(list(contigs)))}
    keys = sorted((k for k in best if k[0] in rank), key=lambda k: (rank[k[0]], k[1]))
    out = [best[k] for k in keys]
    return out, ([r.replace("RNAEditing", "PASS") for r in out] if rediportal is not None else None)


def run_sharded(shards, costs, runner, rank: int, world_size: int, gather=None) -> list | None:
    """Each rank runs `runner(shard)` (-> list of VCF rows) for its shards; rank 0 receives all rows
    ordered by shard index.  `gather(obj)` defaults to torch.distributed.gather_object."""
    owner = assign(costs, world_size)
    mine = {i: runner(s) for i, s in enumerate(shards) if owner[i] == rank}
    if world_size == 1:
        return [mine[i] for i in range(len(shards))]
    if gather is None:
        import torch.distributed as dist

        def gather(obj):
            out = [None] * world_size if rank == 0 else None
            dist.gather_object(obj, out, dst=0)
            return out
    parts = gather(mine)
    if rank != 0:
        return None
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(len(shards))]
