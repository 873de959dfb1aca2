"""Turn ncu reports into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md>       # per-kernel share of a launch list
    python tools/ncu_summary.py full <report.ncu-rep> <out.md> [title]   # key metrics of one --set full capture
    python tools/ncu_summary.py traffic <dram.csv> <workload> <regex>    # per-launch DRAM bytes of the kernels matching <regex>
                                                                         # -> merged into profiles/traffic.json (bench.py reads it)
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
    # a .csv argument is the output of `ncu -i X.ncu-rep --page raw --csv` made on the GPU box (reports can exceed the transfer limit)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
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


def traffic(path, workload, pattern):
    """csv of `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` over one step of bench.py"""
    import json
    import os
    text = open(path, errors="replace").read()
    rows = list(csv.reader(io.StringIO(text[text.find('"ID"'):])))
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_id = collections.defaultdict(float)
    names = {}
    for r in rows[1:]:
        if len(r) < len(hdr) or not r[ci["Metric Name"]].startswith("dram__bytes"):
            continue
        per_id[r[ci["ID"]]] += float(r[ci["Metric Value"]].replace(",", "")) * mult.get(r[ci["Metric Unit"]], 1.0)
        names[r[ci["ID"]]] = r[ci["Kernel Name"]]
    sel = [v for k, v in per_id.items() if re.search(pattern, names[k])]
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[workload] = {"dominant_kernel_dram_bytes_per_launch": sum(sel) / max(len(sel), 1), "launches": len(sel), "kernel_regex": pattern,
                      "all_kernels_dram_bytes_per_step": sum(per_id.values()),
                      "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one graph-replayed step of bench.py --workload {workload} "
                                f"({os.path.basename(path)}); mean over the {len(sel)} launches matching /{pattern}/"}
    json.dump(data, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[workload], indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
