"""Per-kernel share of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`): which kernels the step is made of.
usage: python tools/launch_share.py LAUNCHES.csv OUT.txt "command that produced the list" """
import collections
import csv
import re
import sys


def main():
    src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    tot = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"s1s2::|<unnamed>::", "", re.sub(r"\(.*", "", r[4]))
        v, u = float(r[14]), r[13]
        ms = v / 1e6 if u in ("nsecond", "ns") else (v / 1e3 if u in ("usecond", "us") else v)
        d = tot.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ms
    al = sum(v[1] for v in tot.values())
    conv = sum(v[1] for k, v in tot.items() if "conv_" in k)
    lines = [f"# {cmd}", f"# {len(rows)} launches: per-kernel share of the launch-list time (cold, serialised under the profiler)",
             f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:70]:70s} {v[0]:8d} {v[1]:10.3f} {v[1] / al:7.4f}")
    lines.append(f"# conv kernel family: {conv / al:.4f} of the listed time; the rest is the one-off weight repack and input assembly "
                 "(bench.py's timed region launches only the conv kernels: gpu_launches = 16 x 50 x steps)")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
