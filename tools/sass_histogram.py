"""Instruction histogram of the in-tree library (runs wherever cuobjdump is installed; no GPU needed).

usage: python tools/sass_histogram.py [OUT.txt]

For every kernel in libs1s2_b200.so: total SASS instructions and the counts of the mnemonics that show which hardware
paths it uses -- UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor loads /
stores), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (would be a legacy mma.sync path: must be 0).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "s1-to-s2_super-resolution_project-code_b200", "s1s2_b200", "libs1s2_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UTCATOMSWS", "HMMA", "FFMA", "STS", "LDS"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    lines = [f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes): mnemonic counts per kernel",
             f"{'kernel':110s} {'instrs':>7s} " + " ".join(f"{k:>12s}" for k in KEYS)]
    total = collections.Counter()
    for name, blk in zip(names, blocks):
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk, flags=re.M)
        c = collections.Counter()
        for op in ops:
            for k in KEYS:
                if k == "UTCHMMA":
                    hit = op.startswith("UTCHMMA") and ".2CTA" not in op
                elif k == "UTCHMMA.2CTA":
                    hit = op.startswith("UTCHMMA") and ".2CTA" in op
                else:
                    hit = op.startswith(k)
                if hit:
                    c[k] += 1
        total.update(c)
        short = re.sub(r"s1s2::|\(anonymous namespace\)::|\(s1s2::ConvParams\)", "", name)
        lines.append(f"{short[:110]:110s} {len(ops):7d} " + " ".join(f"{c[k]:12d}" for k in KEYS))
    lines.append(f"{'TOTAL':110s} {'':7s} " + " ".join(f"{total[k]:12d}" for k in KEYS))
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
