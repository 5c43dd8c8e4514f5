#!/usr/bin/env python
"""SASS evidence for profiles/: mnemonic histogram of every kernel in libepivo_b200.so (sm_100a cubin) and the
hot loop of the headline matcher instantiation.

  python tools/sass_excerpt.py > profiles/r2_sass_match.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "epivo_b200", "libepivo_b200.so")
WATCH = ["POPC", "LOP3", "VIMNMX", "IADD3", "UBLKCP", "SYNCS", "REDUX", "UCGABAR", "DFMA", "DMUL", "DADD", "DMMA",
         "MUFU", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "BAR", "SHFL", "UTMALDG", "HMMA", "UTCHMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    out = [o.replace("(anonymous namespace)::", "").replace("void ", "") for o in out]
    return [o[:o.rfind(">(") + 1] if ">(" in o else o.split("(")[0] for o in out]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    funcs = []          # (mangled, [instruction lines])
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = (m.group(1), [])
            funcs.append(cur)
        elif cur is not None and re.search(r"/\*[0-9a-f]{4,5}\*/", line):
            cur[1].append(line)
    names = demangle([f[0] for f in funcs])
    print("# cuobjdump -sass epivo_b200/libepivo_b200.so   (archs in the fat binary: %s)" % ", ".join(arch))
    print("# per kernel: instruction count and the mnemonics that matter on this path")
    print("kernel,instructions," + ",".join(WATCH))
    head = None
    for (mangled, ins), name in zip(funcs, names):
        ops = collections.Counter()
        for l in ins:
            m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
            if m:
                ops[m.group(1)] += 1
        tot = sum(ops.values())
        row = [sum(v for k, v in ops.items() if k == w or k.startswith(w)) for w in WATCH]
        print("%s,%d,%s" % (name, tot, ",".join(map(str, row))))
        if name.startswith("match_tile_kernel<8, 8, true, false, false>"):
            head = (name, ins)
    if head:
        name, ins = head
        print("\n# hot loop of %s (the headline instantiation: 8 words = 256 bits, 8 query rows per thread, HAMMING2,"
              "\n# best-1, consecutive pairs): the block with the highest POPC density between two backward branches" % name)
        # find backward branches: BRA to an earlier address
        addr = lambda l: int(re.search(r"/\*([0-9a-f]{4,5})\*/", l).group(1), 16)
        best = None
        for i, l in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA\s+0x([0-9a-f]+)", l)
            if "BRA" in l:
                t = re.search(r"0x([0-9a-f]+)", l.split("BRA")[1])
                if t and int(t.group(1), 16) < addr(l):
                    lo = next((j for j, x in enumerate(ins) if addr(x) >= int(t.group(1), 16)), None)
                    if lo is not None:
                        pop = sum("POPC" in x for x in ins[lo:i + 1])
                        if best is None or pop > best[0]:
                            best = (pop, lo, i)
        if best:
            pop, lo, hi = best
            body = ins[lo:hi + 1]
            c = collections.Counter(re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l).group(1) for l in body)
            print("# loop body: %d instructions; %s" % (len(body), ", ".join("%s x%d" % kv for kv in c.most_common(12))))
            for l in body[:60]:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip())
            if len(body) > 60:
                print("        ... (%d more instructions of the same pattern)" % (len(body) - 60))
    # bulk-copy / mbarrier lines of the matcher prologue
    if head:
        print("\n# TMA-engine bulk copies + mbarrier of the same kernel (train tile staging):")
        for l in head[1]:
            if any(k in l for k in ("UBLKCP", "SYNCS", "REDUX")):
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip())


if __name__ == "__main__":
    sys.exit(main())
