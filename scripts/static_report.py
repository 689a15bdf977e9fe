#!/usr/bin/env python3
"""profiles/r2_static.md: what can be read off the built library WITHOUT a GPU -- registers, shared memory, local-memory
stack per kernel (cuobjdump -res-usage) and the SASS mnemonics that show which hardware paths the kernels use
(bulk async copies + mbarrier, match/vote/redux warp primitives, 64-bit CAS, f64 reductions, 128-bit loads).
    python scripts/static_report.py > profiles/r2_static.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "quids_b200", "libquids_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name, limit=150):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)  # drop the parameter list
    return name if len(name) <= limit else name[:limit - 3] + "..."


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    kernels = []
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            kernels.append((fn, *map(int, m.groups())))
            fn = None
    names = demangle([k[0] for k in kernels])
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per_fn = collections.defaultdict(collections.Counter)
    cur = None
    wanted = ("UBLKCP", "SYNCS", "MATCH", "VOTE", "REDUX", "ATOMG", "REDG", "RED.E", "ATOMS", "LDG.E.128", "STG.E.128", "LDGSTS", "SHFL", "LDL", "STL", "DMUL", "DFMA", "DADD", "IMAD.WIDE", "LOP3", "SHF")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per_fn[cur]["_all"] += 1
            for w in wanted:
                if op.startswith(w):
                    per_fn[cur][w] += 1
    print("# Static report of `quids_b200/libquids_b200.so` (sm_100a) -- no GPU needed\n")
    print("`python scripts/static_report.py > profiles/r2_static.md`; source: `cuobjdump -res-usage` and `cuobjdump -sass` of the library the tests and the bench load.\n")
    print(f"{len(kernels)} kernels. STACK is the per-thread local-memory frame (0 = no spills, no local arrays).\n")
    print("| kernel | regs | smem (static) | stack | SASS instr. | bulk copy / mbarrier | match / vote / redux | global atomics (CAS, RED) | 128-bit LDG / STG | f64 (DMUL+DFMA+DADD) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    shown = collections.Counter(short(names[k[0]]) for k in kernels)
    for fn, reg, stack, shared, local in sorted(kernels, key=lambda k: -per_fn[k[0]]["_all"]):
        c = per_fn[fn]
        label = short(names[fn])
        if shown[label] > 1:  # overloads: keep the first parameter to tell them apart
            label += "(" + re.sub(r"^.*?\((.*)\)$", r"\1", re.sub(r"\(anonymous namespace\)::", "", names[fn])).split(",")[0][:60] + ", ...)"
        print(f"| `{label}` | {reg} | {shared} | {stack} | {c['_all']} | {c['UBLKCP']} / {c['SYNCS']} | {c['MATCH']} / {c['VOTE']} / {c['REDUX']} | "
              f"{c['ATOMG']} / {c['REDG'] + c['RED.E']} | {c['LDG.E.128']} / {c['STG.E.128']} | {c['DMUL'] + c['DFMA'] + c['DADD']} |")
    total = collections.Counter()
    for c in per_fn.values():
        total.update(c)
    print("\nTotals over the library: " + ", ".join(f"{w} {total[w]}" for w in wanted if total[w]) + f"; {total['_all']} SASS instructions.")
    spill = [short(names[k[0]], 90) for k in kernels if k[2] > 0]
    print(f"\nKernels with a local-memory frame: {len(spill)}" + (": " + "; ".join(f"`{s}`" for s in spill) if spill else "") + ".")


if __name__ == "__main__":
    main()
