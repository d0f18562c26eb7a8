#!/usr/bin/env python
"""Static register-file cost model of the FP64 instructions in a kernel's hottest loop (development aid).

Measured on B200 (tools/dfma_probe*.cu, profiles/r1_dfma_probe.txt): the FP64 pipe takes max(2, R) cycles per warp
instruction, R = number of distinct 64-bit register operands that have to be read from the register file; an
operand is free when the previous instruction kept the same register in the same slot with `.reuse`.

    python tools/sass_rf_model.py <object-or-so> <mangled-kernel-substring> [pairs_per_iteration]
"""
import re
import subprocess
import sys


def main():
    obj, key = sys.argv[1], sys.argv[2]
    pairs = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    body = next(b for b in blocks if b.split("\n", 1)[0].find(key) >= 0)
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    # loops = backward branches; pick the one holding the most FP64 instructions
    best = None
    want = int(sys.argv[4]) if len(sys.argv) > 4 else -1     # optional: pick the loop with exactly this many FP64
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\w*\)?\s*$", t)
        mt = re.search(r"0x([0-9a-f]+)", t) if "BRA" in t else None
        if "BRA" in t and mt:
            tgt = int(mt.group(1), 16)
            if tgt < a and tgt in addr_index:
                lo = addr_index[tgt]
                n = sum(1 for _, x in ins[lo:i + 1] if re.match(r"(@!?U?P\d+\s+)?D(FMA|MUL|ADD)", x))
                if n:
                    print(f"  loop 0x{tgt:x}..0x{a:x}: {i + 1 - lo} instructions, {n} FP64")
                if (want < 0 and (best is None or n > best[0])) or n == want:
                    best = (n, lo, i)
    if best is None:
        print("no loop found (labels instead of addresses?)"); return
    n, lo, hi = best
    loop = [t for _, t in ins[lo:hi + 1]]
    cyc = 0.0
    nfp = 0
    hist = {}
    prev_slots = {}
    for t in loop:
        t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
        m = re.match(r"D(FMA|MUL|ADD)\S*\s+(.*)", t2)
        if not m:
            prev_slots = {} if not t2.startswith("NOP") else prev_slots
            continue
        ops = [o.strip() for o in m.group(2).split(",")][1:]
        reads = set()
        slots = {}
        for k, o in enumerate(ops):
            r = re.match(r"[-|]*\s*(R\d+)(\.reuse)?", o)
            if not r or r.group(1) == "RZ":
                continue
            if prev_slots.get(k) != r.group(1):
                reads.add(r.group(1))
            if r.group(2):
                slots[k] = r.group(1)
        prev_slots = slots
        c = max(2, len(reads))
        hist[len(reads)] = hist.get(len(reads), 0) + 1
        cyc += c
        nfp += 1
    print(f"loop: {len(loop)} instructions, {nfp} FP64, reads histogram {dict(sorted(hist.items()))}, "
          f"model {cyc:.0f} cycles ({cyc / nfp:.3f} per FP64 instr)" + (f", {cyc / pairs:.2f} cycles/pair" if pairs else ""))


if __name__ == "__main__":
    main()
