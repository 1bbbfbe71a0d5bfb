"""Turn an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python
bench.py ...`) into the per-kernel table kept under profiles/ (B200_PROFILING.md's launch-list pass).

    python tests/summarise_launches.py gpurun_out/launches.csv [--title "..."] [--command "..."] > profiles/rN_launch_list.md

The table covers the LAST complete training step of the capture: steps are delimited by the stem's layout pass
(`video_prep*`), which runs once per step.  Durations are cold-cache and serialised (ncu replays every kernel alone):
compare SHARES with the bench line, not absolute times.
"""
import argparse
import csv
import re
import sys
from collections import OrderedDict


def read_launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r and "Metric Value" in r:
                header = r
            continue
        if len(r) != len(header):
            continue
        d = dict(zip(header, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        rows.append((d["Kernel Name"], val * scale))
    return rows


def short(name):
    """Kernel family: the function name without template arguments and parameter list."""
    n = re.sub(r"\(.*$", "", name)
    n = re.sub(r"<.*$", "", n)
    return n.strip()


def last_step(rows, marker="video_prep"):
    idx = [i for i, (k, _) in enumerate(rows) if marker in k]
    if len(idx) < 2:
        return rows
    return rows[idx[-2]:idx[-1]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--title", default="ncu launch list")
    ap.add_argument("--command", default="")
    ap.add_argument("--marker", default="video_prep")
    a = ap.parse_args()
    rows = read_launches(a.csv)
    step = last_step(rows, a.marker)
    fam = OrderedDict()
    for k, ms in step:
        s = short(k)
        f = fam.setdefault(s, [0, 0.0])
        f[0] += 1
        f[1] += ms
    total = sum(v[1] for v in fam.values())
    out = sys.stdout
    out.write("# %s\n\n" % a.title)
    if a.command:
        out.write("Command: `%s`\n\n" % a.command)
    out.write("Last complete training step of the capture (%d launches in the file, %d in the step, %.2f ms of kernel "
              "time). Cold-cache, serialised durations: compare SHARES with the bench line.\n\n"
              % (len(rows), len(step), total))
    out.write("| kernel | launches/step | ms/step | share |\n|---|---:|---:|---:|\n")
    for k, (n, ms) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        out.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k, n, ms, 100.0 * ms / total if total else 0.0))
    mine = sum(v[0] for k, v in fam.items() if "m3t::" in k)
    out.write("\nLibrary kernels: %d launches; other (ATen / NCCL / memset): %d.\n" % (mine, len(step) - mine))


if __name__ == "__main__":
    main()
