"""Golden-vector cases: small instances of the five BASELINE.json configs plus
stress cases (count ties, dense indels, deep coverage next to junctions).

`build(name)` regenerates the exact reads/reference a golden .npz was made from,
so tests never need the reference repository at run time."""
import dataclasses
import numpy as np

from clair3_rna_b200 import synth

_BASE = dict(phased=False, padding=False, snp_af=0.08, indel_af=0.15, min_cov=4, min_mq=5)

CASES = {
    # name: config index, scale, SynthConfig overrides, caller options
    "cfg1_ont_drna": dict(_BASE, cfg=1, scale=0.08, over=dict(genes_per_mb=60), platform="ont"),
    "cfg2_ont_cdna": dict(_BASE, cfg=2, scale=0.0015, over=dict(genes_per_mb=50), platform="ont"),
    "cfg3_hifi_pad": dict(_BASE, cfg=3, scale=0.2, over=dict(genes_per_mb=50, hi_depth_genes=3, hi_depth=400),
                          platform="hifi", padding=True),
    "cfg4_hifi_phased": dict(_BASE, cfg=4, scale=0.4, over=dict(genes_per_mb=50), platform="hifi", phased=True),
    "ties_lowdepth": dict(_BASE, cfg=1, scale=0.05, over=dict(genes_per_mb=80, depth=5, sub=0.12, ins=0.06, dele=0.12,
                                                              seed=777001), platform="ont"),
    "pad_dense": dict(_BASE, cfg=3, scale=0.012, over=dict(genes_per_mb=100, hi_depth_genes=1, hi_depth=260, depth=25,
                                                           sub=0.02, ins=0.01, dele=0.02, seed=777002),
                      platform="hifi", padding=True),
    "phased_noisy": dict(_BASE, cfg=4, scale=0.01, over=dict(genes_per_mb=100, depth=18, sub=0.04, ins=0.03, dele=0.04,
                                                             seed=777003), platform="hifi", phased=True),
    # config 5 shape: a contig longer than one 5 Mb reference chunk, called chunk by chunk (--chunk_id i --chunk_num 2)
    # so that the reference's own chunk geometry (create_tensor_pileup.py:380-418) is part of the pin
    "cfg5_two_chunks": dict(_BASE, cfg=5, scale=0.0245, over=dict(genes_per_mb=6, seed=777052), platform="ont", chunks=2),
    # region modes: --bed_fn + --extend_bed (confident BED, mpileup -l) and --vcf_fn
    # (genotyping at known sites), both chunked so that the BED-span / site-count chunk geometry is pinned too
    "bed_regions": dict(_BASE, cfg=1, scale=0.06, over=dict(genes_per_mb=70, depth=14, sub=0.05, ins=0.03, dele=0.08,
                                                            seed=777061), platform="ont", chunks=2,
                        bed=dict(n=90, seed=11)),
    "bed_pad_hifi": dict(_BASE, cfg=3, scale=0.012, over=dict(genes_per_mb=100, hi_depth_genes=1, hi_depth=200, depth=20,
                                                              sub=0.02, ins=0.01, dele=0.03, seed=777063),
                         platform="hifi", padding=True, chunks=1, bed=dict(n=60, seed=13)),
    "known_sites": dict(_BASE, cfg=1, scale=0.05, over=dict(genes_per_mb=70, depth=10, seed=777062), platform="ont",
                        chunks=3, known=dict(n=400, seed=12)),
    # --enable_variant_calling_at_sequence_head_and_tail: zero rows before a run starts and
    # after the stream ends; with the padding rule the shared zero row collects the padding writes
    "headtail_ont": dict(_BASE, cfg=1, scale=0.05, over=dict(genes_per_mb=80, depth=9, sub=0.06, ins=0.03, dele=0.06,
                                                             seed=777071), platform="ont", head_tail=True),
    "headtail_pad": dict(_BASE, cfg=3, scale=0.012, over=dict(genes_per_mb=100, hi_depth_genes=1, hi_depth=200, depth=20,
                                                              sub=0.03, ins=0.01, dele=0.03, seed=777072),
                         platform="hifi", padding=True, head_tail=True),
    "headtail_phased_bed": dict(_BASE, cfg=4, scale=0.01, over=dict(genes_per_mb=100, depth=15, sub=0.04, ins=0.03,
                                                                    dele=0.04, seed=777073),
                                platform="hifi", phased=True, head_tail=True, chunks=2, bed=dict(n=60, seed=14)),
    # --ctgStart/--ctgEnd instead of a chunk (create_tensor_pileup.py:407-415), and head/tail calling at known sites
    "range_mode": dict(_BASE, cfg=1, scale=0.06, over=dict(genes_per_mb=70, depth=9, seed=777081), platform="ont",
                       ctg_range=(9000, 41000)),
    "headtail_known": dict(_BASE, cfg=1, scale=0.05, over=dict(genes_per_mb=70, depth=9, sub=0.05, ins=0.03, dele=0.05,
                                                               seed=777082), platform="ont", head_tail=True, chunks=2,
                           known=dict(n=300, seed=15)),
    "af_zero": dict(_BASE, cfg=1, scale=0.03, over=dict(genes_per_mb=70, depth=8, seed=777004), platform="ont",
                    snp_af=0.0, min_cov=2),
}


def synth_config(name):
    c = CASES[name]
    cfg = synth.config(c["cfg"], scale=c["scale"])
    return dataclasses.replace(cfg, **c["over"])


def build(name):
    """-> (ReadBatch of the whole contig, reference bytes of the whole contig, contig name)"""
    cfg = synth_config(name)
    ref = synth.Reference(cfg)
    batch = synth.make_contig_reads(cfg, 0, ref)
    contig, length = cfg.contigs[0]
    return batch, ref.fetch(contig, 0, length).tobytes(), contig


def bed_rows(name):
    """confident BED rows (start0, end0) of a case, in file order: unsorted, some overlapping or touching, one in
    twenty empty; anchored on read starts so that their edges fall inside covered exons."""
    spec = CASES[name].get("bed")
    if spec is None:
        return None
    batch, ref, _ = build(name)
    rng = np.random.default_rng(spec["seed"])
    rows = []
    for _ in range(spec["n"]):
        r = int(rng.integers(batch.n_reads))
        a = max(0, int(batch.pos[r]) + int(rng.integers(-60, 900)))
        ln = 0 if rng.random() < 0.05 else int(rng.choice([1, 7, 40, 90, 250, 700]))
        rows.append((a, min(len(ref), a + ln)))
        if rng.random() < 0.15:                      # a touching neighbour
            rows.append((rows[-1][1], min(len(ref), rows[-1][1] + 30)))
    return rows


def known_positions(name):
    """sorted distinct 1-based sites of a --vcf_fn case: positions inside reads (mostly not candidates by AF), a few
    within 33 bp of the contig start and a few far from any read."""
    spec = CASES[name].get("known")
    if spec is None:
        return None
    batch, ref, _ = build(name)
    rng = np.random.default_rng(spec["seed"])
    out = {5, 20, 34, 35}
    for _ in range(spec["n"]):
        r = int(rng.integers(batch.n_reads))
        out.add(int(batch.pos[r]) + 1 + int(rng.integers(0, 400)))
    out |= {int(x) for x in rng.integers(1, len(ref), 12)}
    return sorted(p for p in out if 1 <= p <= len(ref))


def chunk_plans(name):
    """[ChunkPlan | None] of a case, one per producer call, through the product's own host logic
    (clair3_rna_b200/regions.py) - the goldens made by the reference pin it."""
    from clair3_rna_b200 import regions
    case = CASES[name]
    _, ref, _ = build(name)
    n = case.get("chunks", 1)
    conf, known = bed_rows(name), known_positions(name)
    ext = regions.extend_bed_rows(conf) if conf is not None else regions.extend_known_rows(known) if known is not None else None
    if "ctg_range" in case:
        a, b = case["ctg_range"]
        return [regions.plan_chunk(len(ref), ctg_start=a, ctg_end=b, extend_rows=ext, confident_rows=conf,
                                   known_positions=known)]
    if n == 1 and conf is None and known is None:
        return [regions.ChunkPlan(1, len(ref) + 33, 1, len(ref))]
    return [regions.plan_chunk(len(ref), chunk_id=cid, chunk_num=n, extend_rows=ext, confident_rows=conf,
                               known_positions=known) for cid in range(1, n + 1)]
