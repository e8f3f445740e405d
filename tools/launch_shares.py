"""Shares of the kernels of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_shares.py gpurun_out/launches.csv <first kernel of a step (substring)> > profiles/<name>.txt

The step is everything from the last launch whose name contains the marker to the end of the list."""
import csv
import re
import sys


def main():
    path, marker = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    data = [(r[ki], float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1.0))
            for r in rows if r is not hdr and len(r) > max(ki, vi) and r[vi].replace(",", "").replace(".", "").isdigit()]
    starts = [i for i, (n, _) in enumerate(data) if marker in n]
    step = data[starts[-1]:] if starts else data
    agg = {}
    for n, us in step:
        short = re.sub(r"\(.*$", "", n)
        short = re.sub(r"^void ", "", short)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: last step = {len(step)} launches, {total / 1000:.2f} ms of kernel time (cold-cache, serialised: compare shares)")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * us / total:5.1f} %  {us / 1000:8.3f} ms  {c:4d} x  {us / c:8.1f} us  {n[:110]}")


if __name__ == "__main__":
    main()
