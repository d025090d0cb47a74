"""Turns an ncu launch list (--metrics gpu__time_duration.sum --csv) and an `ncu --set full` report into the
markdown tables kept under profiles/.  usage: summarize_profile.py launches.csv prof.ncu-rep [pass_index]"""
import csv, subprocess, sys, collections

def launch_table(path, which):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    H, body = rows[h], rows[h + 1:]
    ik, iv, iu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    starts = [i for i, r in enumerate(body) if 'k_cigar' in r[ik]] + [len(body)]
    seg = body[starts[which]:starts[which + 1]]
    agg = collections.OrderedDict()
    for r in seg:
        t = float(r[iv].replace(',', ''))
        t = t / 1000.0 if r[iu] in ('nsecond', 'ns') else t * (1000.0 if r[iu] in ('msecond', 'ms') else 1.0)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | us | share |", "|---|---:|---:|---:|"]
    for k, (n, t) in agg.items():
        out.append("| `%s` | %d | %.1f | %.1f%% |" % (k.replace('|', '\\|'), n, t, 100 * t / tot))
    out.append("| **total** | %d | %.1f | 100%% |" % (sum(v[0] for v in agg.values()), tot))
    return "\n".join(out)

def full_table(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H = rows[0]
    cols = [("gpu__time_duration.sum", "time us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) %"),
            ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("dram__bytes_read.sum", "dram read MB"), ("dram__bytes_write.sum", "dram write MB"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
    out = ["| kernel | " + " | ".join(c[1] for c in cols) + " |", "|---|" + "---:|" * len(cols)]
    seen = {}
    units = rows[1]
    for r in rows[2:]:
        name = r[H.index("Kernel Name")]
        if name in seen:
            continue
        seen[name] = 1
        vals = []
        for key, label in cols:
            if key not in H:
                vals.append("-"); continue
            v, u = r[H.index(key)], units[H.index(key)]
            try:
                f = float(v.replace(',', ''))
                if key.startswith("dram__bytes"):
                    f *= {"Gbyte": 1000.0, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
                if key == "gpu__time_duration.sum":
                    f *= {"msecond": 1000.0, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
                vals.append("%.1f" % f)
            except ValueError:
                vals.append(v)
        out.append("| `%s` | %s |" % (name, " | ".join(vals)))
    return "\n".join(out)

if __name__ == "__main__":
    print(launch_table(sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 2))
    print()
    print(full_table(sys.argv[2]))
