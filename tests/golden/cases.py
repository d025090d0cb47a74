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
    "cfg3_hifi_pad": dict(_BASE, cfg=3, scale=0.02, over=dict(genes_per_mb=50, hi_depth_genes=2, hi_depth=400),
                          platform="hifi", padding=True),
    "cfg4_hifi_phased": dict(_BASE, cfg=4, scale=0.02, over=dict(genes_per_mb=50), platform="hifi", phased=True),
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
