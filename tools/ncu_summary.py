"""Turn ncu reports into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md>       # per-kernel share of a launch list
    python tools/ncu_summary.py full <report.ncu-rep> <out.md> [title]   # key metrics of one --set full capture
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    return re.sub(r"sdnq::<unnamed>::|sdnq::\(anonymous namespace\)::|at::native::|<unnamed>::", "", name)[:90]


def launches(path, out):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.reader(io.StringIO(text[start:])))
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        us = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
        k = short(r[ci["Kernel Name"]])
        tot[k] += us
        cnt[k] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# kernel launch list summary ({path})\n\ncold-cache, serialised per-launch times (ncu): compare SHARES, not absolutes\n\n")
        f.write(f"total {total:.1f} us over {sum(cnt.values())} launches\n\n| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in tot.most_common(25):
            f.write(f"| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}% | {v / cnt[k]:.2f} |\n")
    print(open(out).read())


def full(rep, out, title=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary: {title or rep}\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## `{short(name)}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
