#!/usr/bin/env python
"""SASS evidence for profiles/: the hottest FP64 loop of a kernel with its instruction mix, plus the bulk-TMA /
mbarrier instructions of the whole kernel.

    python tools/sass_excerpt.py <so> <mangled-kernel-substring> > profiles/rN_<kernel>.sass
"""
import collections
import re
import subprocess
import sys


def main():
    obj, key = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    body = next(b for b in blocks if key in b.split("\n", 1)[0])
    name = body.split("\n", 1)[0].strip()
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    index = {a: i for i, (a, _) in enumerate(ins)}
    fp64 = lambda t: re.match(r"(@!?U?P\d+\s+)?D(FMA|MUL|ADD|SETP)", t) is not None
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"0x([0-9a-f]+)", t) if "BRA" in t else None
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in index:
                lo = index[tgt]
                n = sum(1 for _, x in ins[lo:i + 1] if fp64(x))
                if best is None or n > best[0]:
                    best = (n, lo, i)
    whole = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
    print(f"# {name}")
    print(f"# whole kernel: {len(ins)} instructions; "
          + ", ".join(f"{k} {whole[k]}" for k in ("UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "SHFL", "LDS", "STS",
                                                   "LDG", "STG", "BAR") if whole[k]))
    tma = [t for _, t in ins if t.split()[0].startswith(("UBLKCP", "SYNCS")) or " UBLKCP" in t or " SYNCS" in t]
    print("# bulk-TMA / mbarrier instructions (UBLKCP = cp.async.bulk, SYNCS = mbarrier):")
    for t in sorted(set(tma)):
        print(f"#   {t}")
    if best is None:
        return
    n, lo, hi = best
    loop = ins[lo:hi + 1]
    mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in loop)
    reuse = sum(t.count(".reuse") for _, t in loop)
    print(f"# hottest loop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {len(loop)} instructions, FP64 {n}; mix: "
          + ", ".join(f"{k} {v}" for k, v in mix.most_common()) + f"; .reuse operand flags: {reuse}")
    for a, t in loop:
        print(f"        /*{a:04x}*/  {t} ;")


if __name__ == "__main__":
    main()
