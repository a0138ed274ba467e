#!/usr/bin/env python
"""Per-call wall time of repeated transposes of one matrix (allocator behaviour shows up as spikes)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from rcppsparse_b200 import DeviceMatrix, _lib, synth

    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    dirty = len(sys.argv) > 2 and sys.argv[2] == "dirty"
    if dirty:  # leave blocks of other sizes in the library's pool first
        with DeviceMatrix.synth(synth.config("C3", 0.3)) as X:
            t = X.transpose_dev()
            t.close()
    D = DeviceMatrix.synth(synth.config(wl))
    D.set_stream(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    keep, ms = [], []
    for _ in range(30):
        t0 = time.perf_counter()
        keep.clear()
        keep.append(D.transpose_dev())
        torch.cuda.synchronize()
        ms.append(round((time.perf_counter() - t0) * 1e3, 2))
    print(json.dumps({"workload": wl, "dirty_pool": dirty, "ms": ms}))


if __name__ == "__main__":
    main()
