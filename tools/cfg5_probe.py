"""Host-side timing of the sharded whole-genome leg on one GPU (dataset from bench.py's cache)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from clair3_rna_b200 import run_chunks
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    os.environ["DEV"] = os.environ.get("LOCAL_RANK", "0")
    if os.environ.get("WITH_NCCL"):
        import torch, torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["DEV"])))
        torch.cuda.set_device(int(os.environ["DEV"]))
        dist.barrier()
cfg, bam, fa, wnpz, prep = bench.cfg5_dataset(scale, 0, 1, len(os.sched_getaffinity(0)), lambda: None)
from clair3_rna_b200 import weights
from clair3_rna_b200.engine import Engine
eng = Engine(int(os.environ.get("DEV", "0")), 18); eng.set_weights(weights.load(wnpz))
for rep in range(3):
    st = {}
    t = time.time()
    run_chunks.run(bam, fa, wnpz, "/tmp/cfg5_probe%s.vcf" % os.environ.get("DEV", "0"), stats=st, device=int(os.environ.get("DEV", "0")), rank=rank, world=world, merge=False, engine=eng, loader_threads=int(os.environ.get("LOADERS", "2")), native_threads=int(os.environ.get("NATIVE", "0")))
    ms = np.array(st["submit_ms"])
    print("[rank %d] OMP=%s" % (rank, os.environ.get("OMP_NUM_THREADS")), "run %d: loop %.3f s total %.3f s, %d candidates, %.0f sites/s; host %s" % (
        rep, st["seconds"], time.time() - t, st["candidates"], st["candidates"] / st["seconds"],
        {k: round(v, 3) for k, v in st["host_seconds"].items()}))
    print("   submit ms: first 8", np.round(ms[:8], 2), "median %.2f p90 %.2f max %.2f sum %.1f" % (np.median(ms), np.percentile(ms, 90), ms.max(), ms.sum()))
