#!/usr/bin/env python
"""Per-kernel SASS evidence of libsparse_b200.so (no GPU needed): architecture, registers, shared memory, and the
counts of the mnemonics that show what each kernel is built from — UBLKCP (cp.async.bulk, the TMA engine),
SYNCS (mbarrier), LDGSTS (cp.async), REDG (red.global), ATOMS (shared-memory atomics), MATCH, SHFL, BAR, LDG/STG widths.

    python tools/sass_summary.py > profiles/r02/sass_summary.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rcppsparse_b200", "libsparse_b200.so")
WANT = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "REDG", "ATOMS", "ATOMG", "MATCH", "SHFL", "BAR", "LDG.E.128", "LDG.E.64", "LDG.E",
        "STG.E.128", "STG.E.64", "STG.E", "LDS", "STS", "DFMA", "DADD", "NANOSLEEP", "ERRBAR", "CCTL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# {os.path.relpath(LIB, ROOT)}: arch = {', '.join(arch)}")
    print("# per kernel: resource usage (cuobjdump -res-usage), then mnemonic counts (cuobjdump -sass)\n")
    blocks = re.split(r"\n\s*Function : ", sass)[1:]
    names = [b.split("\n", 1)[0].strip() for b in blocks]
    dm = demangle(names)
    for name, body in sorted(zip(names, blocks), key=lambda t: dm[t[0]]):
        ops = Counter()
        n_instr = 0
        for line in body.splitlines():
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            n_instr += 1
            op = m.group(1)
            for w in WANT:
                if op.startswith(w):
                    ops[w] += 1
                    break
        short = re.sub(r"\(anonymous namespace\)::", "", dm[name])
        short = re.sub(r"sb200::", "", short)
        print(f"{short}")
        print(f"    {usage.get(name, '')}")
        print(f"    instructions {n_instr}: " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda t: -t[1])))


if __name__ == "__main__":
    main()
