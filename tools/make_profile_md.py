"""profiles/<TAG>_summary.md + profiles/<TAG>_traffic.json from the artefacts of tools/profile_round.sh:
    gpurun_out/launches_<TAG>.csv   (ncu --metrics gpu__time_duration.sum, whole bench run)
    gpurun_out/prof_<TAG>.ncu-rep   (ncu --set full of every launch of one timed pass)
usage: python tools/make_profile_md.py TAG [OUT_TAG]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TUNIT = {"msecond": 1e3, "ms": 1e3, "usecond": 1.0, "us": 1.0, "nsecond": 1e-3, "ns": 1e-3, "second": 1e6, "s": 1e6}
BUNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def short(name):
    n = name.replace("c3r::", "").replace("void ", "")
    return n.split("(")[0]


def launch_rows(path, which=2):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    H, body = rows[h], rows[h + 1:]
    ik, iv, iu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    starts = [i for i, r in enumerate(body) if 'k_cigar' in r[ik]] + [len(body)]
    seg = body[starts[which]:starts[which + 1]]
    return [(short(r[ik]), float(r[iv].replace(',', '')) * TUNIT.get(r[iu], 1.0)) for r in seg]


def full_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, U = rows[0], rows[1]

    def get(r, key):
        if key not in H:
            return None
        i = H.index(key)
        try:
            v = float(r[i].replace(',', ''))
        except ValueError:
            return None
        if key.startswith("dram__bytes"):
            v *= BUNIT.get(U[i], 1.0)
        if key == "gpu__time_duration.sum":
            v *= TUNIT.get(U[i], 1.0)
        return v
    stall_keys = [h for h in H if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    out = []
    for r in rows[2:]:
        st = sorted(((get(r, k) or 0.0, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for k in stall_keys), reverse=True)[:3]
        out.append(dict(
            name=short(r[H.index("Kernel Name")]), grid=r[H.index("Grid Size")], block=r[H.index("Block Size")],
            us=get(r, "gpu__time_duration.sum"), rd=get(r, "dram__bytes_read.sum"), wr=get(r, "dram__bytes_write.sum"),
            dram=get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            tensor=get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            xu=get(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            alu=get(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            issue=get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            inst=get(r, "smsp__inst_executed.sum"), regs=get(r, "launch__registers_per_thread"),
            warps=get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            stalls=", ".join("%s %.1f" % (n, v) for v, n in st)))
    return out


def stage_of(name):
    if name.startswith("k_lstm") or name.startswith("k_gemm") or name.startswith("k_l4") or name.startswith("k_heads") or name.startswith("k_xop"):
        return "k5"
    if "OpCand" in name or name.startswith("k_cand_emit"):
        return "k3"
    if name.startswith("k_window") or name.startswith("k_padding") or name.startswith("k_rescale") or "OpAltOff" in name or name.startswith("k_altinfo"):
        return "k4"
    if name.startswith("k_cmp") or "OpEvents" in name or name.startswith("k_scatter") or name.startswith("k_cov_aggr") or name.startswith("k_rows") or "OpSkip" in name:
        return "k2"
    return "k1"


def main():
    tag = sys.argv[1]
    out_tag = sys.argv[2] if len(sys.argv) > 2 else tag
    lcsv = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
    rep = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
    L = launch_rows(lcsv)
    F = full_rows(rep)
    agg = collections.OrderedDict()
    for n, t in L:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(t for _, t in L)
    md = ["# %s - ncu evidence of one hot-path pass (config 2: chr20-sized contig, 30 844 reads, 996 746 rows, 14 617 candidates)" % out_tag, "",
          "Made by `tools/profile_round.sh %s` under `gpurun` (one B200) and `tools/make_profile_md.py`:" % tag, "",
          "    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_%s.csv \\" % tag,
          "        python bench.py --steps 1 --warmup 1 --no_cpu_baseline --cfg5_scale 0",
          "    ncu --set full --clock-control none --import-source on -s <first launch of the timed pass> -c <launches of the pass> \\",
          "        -o gpurun_out/prof_%s python bench.py --steps 1 --warmup 1 --no_cpu_baseline --cfg5_scale 0" % tag, "",
          "Raw launch list of the whole run: `%s_launches.csv`.  ncu times are cold-cache and serialised (the two-stream overlap of" % out_tag,
          "the network tail does not exist under ncu): compare SHARES with `bench.py`'s `stage_ms`, not absolutes.", "",
          "## Launch list of the timed pass (`c3r_rerun_resident`)", "",
          "| kernel | launches | us | share |", "|---|---:|---:|---:|"]
    for k, (n, t) in agg.items():
        md.append("| `%s` | %d | %.1f | %.1f%% |" % (k, n, t, 100 * t / tot))
    md.append("| **total** | %d | %.1f | 100%% |" % (len(L), tot))
    st = collections.OrderedDict((k, 0.0) for k in ("k1", "k2", "k3", "k4", "k5"))
    for n, t in L:
        st[stage_of(n)] += t
    md += ["", "Per stage: " + ", ".join("%s %.1f us (%.1f %%)" % (k.upper(), v, 100 * v / tot) for k, v in st.items()), "",
           "## `ncu --set full`, every launch of the pass", "",
           "| kernel | grid x block | us | dram rd MB | dram wr MB | dram % | tensor % | XU % | ALU % | issue % | warp inst (M) | regs | warps active % | top stalls (cycles per issue) |",
           "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
    f = lambda v, p="%.1f": "-" if v is None else p % v
    for r in F:
        md.append("| `%s` | %s x %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
            r["name"], r["grid"].replace(", 1, 1", "").strip("()"), r["block"].replace(", 1, 1", "").strip("()"), f(r["us"]),
            f(None if r["rd"] is None else r["rd"] / 1e6), f(None if r["wr"] is None else r["wr"] / 1e6), f(r["dram"]), f(r["tensor"]),
            f(r["xu"]), f(r["alu"]), f(r["issue"]), f(None if r["inst"] is None else r["inst"] / 1e6, "%.2f"), f(r["regs"], "%.0f"),
            f(r["warps"]), r["stalls"]))
    tr = collections.OrderedDict((k, 0.0) for k in ("k1", "k2", "k3", "k4", "k5"))
    per = collections.OrderedDict()
    for r in F:
        b = (r["rd"] or 0.0) + (r["wr"] or 0.0)
        tr[stage_of(r["name"])] += b
        per[r["name"]] = per.get(r["name"], 0.0) + b
    md += ["", "DRAM traffic per pass (read + write, sums over the launches): " + ", ".join("%s %.1f MB" % (k.upper(), v / 1e6) for k, v in tr.items()), ""]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", "%s_summary.md" % out_tag), "w").write("\n".join(md))
    shutil.copy(lcsv, os.path.join(ROOT, "profiles", "%s_launches.csv" % out_tag))
    json.dump({"source": "ncu --set full --clock-control none, gpurun_out/prof_%s.ncu-rep, bench.py --steps 1 --warmup 1 (config 2), every launch of the timed pass" % tag,
               "workload": "cfg2_ont_r10_cdna_chr20", "candidates": 14617, "rows": 996746,
               "dram_bytes_per_pass": {k: int(v) for k, v in per.items()},
               "dram_bytes_per_stage": {k: int(v) for k, v in tr.items()},
               "ncu_us_per_stage": {k: round(v, 1) for k, v in st.items()}},
              open(os.path.join(ROOT, "profiles", "%s_traffic.json" % out_tag), "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
