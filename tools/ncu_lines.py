#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep captured with --import-source on (read here, no GPU needed): instructions
executed, thread instructions per unit of work, stall samples — the lines that matter, per kernel launch."""
import csv
import subprocess
import sys


def main(path, units=None, thresh=0.004):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fn, kern, cur = None, None, {}

    def flush():
        if not cur:
            return
        tot = sum(v[1] for v in cur.values())
        tots = sum(v[0] for v in cur.values())
        tt = sum(v[2] for v in cur.values())
        print(f"\n## {kern}\nwarp instructions {tot}, thread instructions {tt}, samples {tots}" + (f", thread instructions per unit {tt / units:.1f}" if units else ""))
        for (f, ln, src), v in cur.items():
            if v[1] > thresh * tot or v[0] > thresh * tots:
                per = f"{v[2] / units:6.1f}/unit " if units else ""
                print(f"{f[:18]:18s} {ln:5d} inst {v[1] / tot * 100:5.1f}% {per}samples {v[0] / tots * 100:5.1f}%  {src[:100]}")

    def num(s):
        try:
            return int(s)
        except ValueError:
            return 0

    for r in rows:
        if r and r[0] == "File Path":
            fn = r[1].split("/")[-1]
            continue
        if r and r[0] == "Function Name":
            if kern is not None and r[1] != kern:
                flush()
                cur = {}
            kern = r[1]
            continue
        if len(r) > 8 and r[0].isdigit():
            key = (fn, int(r[0]), r[1].strip())
            v = cur.setdefault(key, [0, 0, 0])
            v[0] += num(r[6])
            v[1] += num(r[7])
            v[2] += num(r[8])
    flush()


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
