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
