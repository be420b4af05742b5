#!/usr/bin/env python
"""Row-loop instruction census of a kernel in libqcat_b200.so, read off `cuobjdump -sass` (no GPU needed).

    python tools/sass_loop.py k_barcode_fast [--listing out.txt] [--min-dp 24]

Finds every backward branch of the kernel, takes the innermost loop body that holds at least --min-dp VIMNMX3.U16x2
instructions, and prints its instruction count per opcode; --listing also writes the kernel's whole SASS."""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
LIB = os.path.join(ROOT, "qcat_b200", "libqcat_b200.so")
LINE = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*(/\*.*\*/)?\s*$")


def kernel_sass(name):
    text = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    out, keep = [], False
    for line in text.splitlines():
        if "Function :" in line:
            keep = name in line
        if keep:
            out.append(line)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--listing")
    ap.add_argument("--min-dp", type=int, default=24)
    ap.add_argument("--all", action="store_true", help="every innermost loop that qualifies, not only the shortest")
    args = ap.parse_args()
    lines = kernel_sass(args.kernel)
    if not lines:
        sys.exit("kernel %s not found in %s" % (args.kernel, LIB))
    if args.listing:
        with open(args.listing, "w") as fh:
            fh.write("\n".join(lines) + "\n")
    inst = []
    for line in lines:
        m = LINE.match(line)
        if m:
            inst.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, text in inst:
        m = re.search(r"\bBRA\s+(0x[0-9a-f]+)", text)
        if not m:
            continue
        target = int(m.group(1), 16)
        if target >= addr:
            continue
        body = [t for a, t in inst if target <= a <= addr]
        dp = sum("VIMNMX3" in t for t in body)
        if dp >= args.min_dp:
            loops.append((target, addr, body))
    # innermost only: drop loops that contain another qualifying loop
    loops = [l for l in loops if not any(o is not l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    if not loops:
        sys.exit("no loop with >= %d VIMNMX3 found" % args.min_dp)
    if not args.all:
        loops = [min(loops, key=lambda l: len(l[2]))]
    for target, addr, body in loops:
        ops = collections.Counter()
        for t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            ops[t.split()[0]] += 1
        print("%s: loop 0x%04x..0x%04x, %d instructions" % (args.kernel, target, addr, len(body)))
        for op, n in sorted(ops.items(), key=lambda kv: -kv[1]):
            print("  %-22s %3d" % (op, n))


if __name__ == "__main__":
    main()
