"""Host logic of the multi-GPU path on CPU: partition, gather over gloo (world size 2), merge."""
import os
import socket
import subprocess
import sys
import textwrap

from clair3_rna_b200 import sharder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_is_balanced_and_deterministic():
    costs = [9, 7, 6, 5, 4, 3, 1]
    owner = sharder.assign(costs, 2)
    assert owner == sharder.assign(costs, 2)
    loads = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(2)]
    assert abs(loads[0] - loads[1]) <= max(costs)
    assert sharder.assign(costs, 1) == [0] * len(costs)


def test_merge_rows_dedups_chunk_overlap():
    a = ["chr1\t100\t.\tA\tG\t10.00\tPASS", "chr1\t4999990\t.\tC\tT\t5.00\tPASS"]
    b = ["chr1\t4999990\t.\tC\tT\t5.00\tPASS", "chr1\t5000100\t.\tG\tA\t7.00\tPASS"]
    c = ["chr2\t50\t.\tT\tC\t3.00\tPASS"]
    out = sharder.merge_rows([c, a, b], ["chr1", "chr2"])
    assert [r.split("\t")[:2] for r in out] == [["chr1", "100"], ["chr1", "4999990"], ["chr1", "5000100"], ["chr2", "50"]]


def test_two_rank_gather_over_gloo(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        sys.path.insert(0, %r)
        import torch.distributed as dist
        from clair3_rna_b200 import sharder
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        shards = [("chr1", i + 1, 6) for i in range(6)]
        costs = [10, 1, 8, 2, 7, 3]
        runner = lambda sh: ["chr1\\t%%d\\t.\\tA\\tG\\t9.00\\tPASS\\trank%%d" %% (sh[1] * 1000, rank)]
        rows = sharder.run_sharded(shards, costs, runner, rank, world)
        if rank == 0:
            merged = sharder.merge_rows(rows, ["chr1"])
            print(json.dumps({"n": len(merged), "pos": [int(r.split("\\t")[1]) for r in merged],
                              "ranks": sorted({r.split("\\t")[-1] for r in merged})}))
        dist.destroy_process_group()
    """ % ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["n"] == 6 and d["pos"] == [1000, 2000, 3000, 4000, 5000, 6000]
    assert d["ranks"] == ["rank0", "rank1"]
