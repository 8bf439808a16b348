"""Turn ncu outputs into the markdown summaries kept under profiles/.

    python scripts/summarize_ncu.py launches <launches.csv> <title> > profiles/xxx.md
    python scripts/summarize_ncu.py full <report.ncu-rep> <title>   > profiles/yyy.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(CUtensor.*", "", name)
    name = re.sub(r"\((?!int\)).*", "", name)
    name = name.replace("digs::", "").replace("(int)", "")
    return name.strip()


def launches(path, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        t = float(r[mv].replace(",", ""))
        t = t / 1e3 if r[mu] == "ns" else (t * 1e3 if r[mu] == "ms" else t)
        a = agg[short(r[kn])[:90]]
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print("# %s\n" % title)
    print("Per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


def full(path, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    seen = collections.OrderedDict()
    for r in rows[2:]:
        seen.setdefault(short(r[kn]), []).append(r)
    print("# %s\n" % title)
    print("`ncu --set full --clock-control none --import-source on`, one GPU.  First captured instance of each kernel; "
          "n = instances captured.\n")
    names = list(seen)
    print("| metric | " + " | ".join("`%s` (n=%d)" % (n, len(seen[n])) for n in names) + " |")
    print("|---|" + "---:|" * len(names))
    for key, label in METRICS:
        idx = [i for i, h in enumerate(hdr) if h == key]
        if not idx:
            continue
        i = idx[0]
        print("| %s [%s] | " % (label, units[i]) + " | ".join(seen[n][-1][i] for n in names) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
