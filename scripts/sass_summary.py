#!/usr/bin/env python
"""SASS evidence for DESIGN.md: per-kernel opcode counts of jubjub_b200/libjubjub_b200.so.

    python scripts/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_summary.txt

For every function in the cubin: static instruction count, IMAD.WIDE (the multiplier instruction the integer
roofline is counted in), other IMAD forms (IMAD.MOV / IMAD.X / IMAD.IADD ... -- they share the FMA-heavy pipe),
IADD3, SEL, LOP3/SHF, 256-bit global accesses with their L2 hints, local-memory spill traffic (STL / LDL), TMA bulk
copies (UBLKCP) and mbarrier waits (SYNCS), calls.  For the variable-base scalar-mul kernel the innermost loops are
also cut out by their backward branches, so the counts per point doubling / per window iteration can be read off
(spills inside them would show up there).  Runs on the CPU box: needs cuobjdump only.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSN = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)\s*(.*?);")


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
        return dict(zip(names, out.splitlines()))
    except Exception:  # noqa: BLE001
        return {n: n for n in names}


def parse(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = INSN.match(line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2), m.group(3)))
    return funcs


def classify(op):
    if op.startswith("IMAD.WIDE"):
        return "imad_wide"
    if op.startswith("IMAD") or op.startswith("IMUL"):
        return "imad_other"
    if op.startswith("IADD3") or op.startswith("IADD"):
        return "iadd3"
    if op.startswith("SEL"):
        return "sel"
    if op.startswith(("LOP3", "SHF", "PLOP3", "ISETP", "LEA", "PRMT", "MOV", "UMOV", "R2UR", "S2R", "CS2R", "S2UR")):
        return "alu_other"
    if op.startswith(("LDG", "STG")):
        return "global"
    if op.startswith(("LDL", "STL")):
        return "local"
    if op.startswith(("LDS", "STS")):
        return "shared"
    if op.startswith(("CALL", "RET", "BRA", "EXIT", "BSSY", "BSYNC", "WARPSYNC", "BAR")):
        return "control"
    return "other"


def counts(insns):
    c = collections.Counter(classify(op) for _, op, _ in insns)
    c["total"] = len(insns)
    return c


def detail(insns):
    d = collections.Counter()
    for _, op, _ in insns:
        if op.startswith(("LDG", "STG", "LDL", "STL", "UBLKCP", "SYNCS", "CALL", "LDS", "STS", "ATOM", "RED", "UTMA")):
            d[op] += 1
    return d


def loops(insns):
    """Backward branches -> (start, end) address ranges, innermost first."""
    out = []
    for addr, op, args in insns:
        if op.startswith("BRA") and not op.startswith("BRA.U"):
            m = re.search(r"0x([0-9a-f]+)", args)  # also the `BRA P3, 0x...` form
            if m and int(m.group(1), 16) < addr:
                out.append((int(m.group(1), 16), addr))
    return sorted(out, key=lambda r: r[1] - r[0])


def fmt(c):
    keys = ["total", "imad_wide", "imad_other", "iadd3", "sel", "alu_other", "global", "local", "shared", "control", "other"]
    return "  ".join(f"{k}={c.get(k, 0)}" for k in keys)


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "jubjub_b200", "libjubjub_b200.so")
    funcs = parse(lib)
    names = demangle(list(funcs))
    print(f"# {os.path.relpath(lib, ROOT)}: {os.path.getsize(lib)} bytes, {len(funcs)} SASS functions")
    print(f"# nvcc: {subprocess.run(['nvcc', '--version'], capture_output=True, text=True).stdout.strip().splitlines()[-2]}")
    print("# cost model of DESIGN.md section 5: time ~ 2 x imad_wide + (everything else)\n")
    for mangled, insns in funcs.items():
        c = counts(insns)
        name = names[mangled]
        print(f"== {name}")
        print(f"   {fmt(c)}")
        d = detail(insns)
        if d:
            print("   " + "  ".join(f"{k} x{v}" for k, v in sorted(d.items())))
        if "k_scalar_mul<" in name or "k_scalar_mul_const" in name:
            for lo, hi in loops(insns):
                body = [i for i in insns if lo <= i[0] <= hi]
                if len(body) < 200:
                    continue
                cb = counts(body)
                calls = sum(1 for _, op, _ in body if op.startswith("CALL"))
                print(f"   loop 0x{lo:04x}-0x{hi:04x} ({(hi - lo + 16)} B): {fmt(cb)}  calls={calls}")
        print()
    # the shared Fq product / square bodies are device functions inside the kernels' text: report them by CALL target
    for mangled, insns in funcs.items():
        if "k_scalar_mul" not in names[mangled] or "fixed" in names[mangled]:
            continue
        targets = collections.Counter()
        for _, op, args in insns:
            if op.startswith("CALL"):
                m = re.search(r"0x([0-9a-f]+)", args)
                if m:
                    targets[int(m.group(1), 16)] += 1
        for t, ncalls in sorted(targets.items()):
            body = []
            for i in insns:
                if i[0] >= t:
                    body.append(i)
                    if i[1].startswith("RET"):
                        break
            print(f"== {names[mangled]} :: device function at 0x{t:04x} ({ncalls} static call sites)")
            print(f"   {fmt(counts(body))}")
        print()


if __name__ == "__main__":
    main()
