"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py: per-kernel launches, total and mean
time, share — for the whole capture and for the LAST window that looks like one graph-replayed step (the launches between
two consecutive `unpack_targets` launches, or the last N launches with --tail N).

    python tools/launchlist_summary.py gpurun_out/r2_launches_bench.csv.gz [--tail N] [--from ID --to ID]
"""
import csv
import gzip
import io
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(?:hgs::)?([A-Za-z0-9_:]+)", name)
    s = m.group(1) if m else name
    if s.startswith("native::") or s.startswith("at::"):
        inner = re.search(r"native::(\w+?)(?:_kernel|Functor|Ops)", name)
        s = "torch:" + (inner.group(1) if inner else s.split("::")[-1])
    return s


def load(path):
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    lines = [l for l in raw.splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    return [(int(r["ID"]), short(r["Kernel Name"]), float(r["Metric Value"]) / 1000.0, r["Stream"]) for r in rows]


def table(rows, title):
    agg = OrderedDict()
    for _, k, us, _ in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"\n{title}: {len(rows)} launches, {tot:.1f} us serialised")
    print("| kernel | launches | us total | us/launch | share |")
    print("|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {us:.1f} | {us / n:.2f} | {us / tot:.3f} |")


def main():
    path = sys.argv[1]
    rows = load(path)
    args = sys.argv[2:]
    lo = int(args[args.index("--from") + 1]) if "--from" in args else None
    hi = int(args[args.index("--to") + 1]) if "--to" in args else None
    if lo is not None:
        table([r for r in rows if lo <= r[0] < hi], f"launches {lo}..{hi}")
        return
    if "--tail" in args:
        n = int(args[args.index("--tail") + 1])
        table(rows[-n:], f"last {n} launches")
        return
    if "--seq" in args:
        prev, cnt = None, 0
        for i, k, us, st in rows:
            if k == prev:
                cnt += 1
                continue
            if prev is not None:
                print(f"{start}: {prev} x{cnt}")
            prev, cnt, start = k, 1, i
        print(f"{start}: {prev} x{cnt}")
        return
    table(rows, "whole capture")


if __name__ == "__main__":
    main()
