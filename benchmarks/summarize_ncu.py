#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries committed under profiles/.

  launches  <launches.csv> <n_last> <out.md>   per-kernel time share of the last n launches
                                               (ncu --metrics gpu__time_duration.sum --csv log)
  report    <file.ncu-rep> <out.md>            key metrics + top stall instructions per launch
"""
import collections
import csv
import re
import subprocess
import sys


def _short(name: str) -> str:
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return name[:80]


def launches(path: str, n_last: int, out: str) -> None:
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    rows = rows[-n_last:] if n_last > 0 else rows
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        k = _short(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    with open(out, "w") as fh:
        fh.write(f"# ncu launch list: {path} (last {len(rows)} launches = one forward step)\n\n")
        fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and "
                 "serialised: compare SHARES, not absolutes.\n\n")
        fh.write(f"total {tot / 1e6:.3f} ms over {len(rows)} launches\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| {v / 1e6:.3f} | {100 * v / tot:.1f}% | {c} | `{k}` |\n")
    print(open(out).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def report(rep: str, out: str) -> None:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full summary of {rep}\n\n")
        for li, r in enumerate(rows[2:]):
            fh.write(f"## launch {li}: `{_short(r[hdr.index('Kernel Name')])}` grid {r[hdr.index('Grid Size')]} "
                     f"block {r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    fh.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                                 capture_output=True, text=True).stdout
            srows = list(csv.reader(src.splitlines()))
            if len(srows) > 2:
                h = srows[1]
                ci = {n: i for i, n in enumerate(h)}
                data = [x for x in srows[2:] if len(x) == len(h) and x[ci["# Samples"]] not in ("# Samples", "")]
                data = data[: len(data) // 2] if len(data) > 1 else data          # the page lists SASS twice
                def fl(x):
                    try:
                        return float(x)
                    except ValueError:
                        return 0.0
                tot = sum(fl(x[ci["# Samples"]]) for x in data) or 1.0
                stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
                agg = {n: sum(fl(x[ci[n]]) for x in data) for n in stalls}
                fh.write("\nstall reasons (share of warp samples): " + ", ".join(
                    f"{n[6:]} {100 * v / tot:.0f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]) + "\n\n")
                fh.write("top instructions by samples:\n\n| samples | instruction | main stall |\n|---:|---|---|\n")
                for x in sorted(data, key=lambda x: -fl(x[ci["# Samples"]]))[:12]:
                    st = sorted(((n, fl(x[ci[n]])) for n in stalls), key=lambda kv: -kv[1])[0]
                    fh.write(f"| {x[ci['# Samples']]} | `{x[ci['Source']].strip()[:70]}` | {st[0][6:]} |\n")
            fh.write("\n")
    print(open(out).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]), sys.argv[4])
    else:
        report(sys.argv[2], sys.argv[3])
