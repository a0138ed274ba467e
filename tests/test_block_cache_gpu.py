"""The library's device memory (csrc/blockcache.cu): freed blocks are kept by capacity and handed to the next
create -> op -> destroy cycle (the reference builds a fresh Matrix per .Call, src/RcppExports.cpp:20), ordered after
the previous owner's work by an event.  A block reused too early shows up as a wrong result, so every cycle is checked
against the oracle; the same with the cache switched off, after a trim, and with a cache too small to hold the blocks."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import oracle
from rcppsparse_b200 import DeviceMatrix, _lib, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cycles(n):
    spec = synth.config("C2", 0.04)
    i, p, x = synth.generate_host(spec)
    a = (i, p, x, spec.nrow, spec.ncol)
    chk = oracle.best()
    wi, wp, wx = chk.transpose(*a)
    want_rows = chk.rowSums(*a)
    for rep in range(n):
        xr = x * (rep + 1.0)  # other values in the same block sizes: a stale block would give the previous cycle's sums
        with DeviceMatrix.from_host(i, p, xr, spec.nrow, spec.ncol, device=0, validate=True) as M:
            ti, tp, tx = M.transpose_host()
            assert np.array_equal(tp, wp) and np.array_equal(ti, wi), f"cycle {rep}: transposed structure"
            assert np.array_equal(tx.view(np.uint64), (wx * (rep + 1.0)).view(np.uint64)), f"cycle {rep}: transposed values"
            oracle.assert_within("rowSums", M.row_sums(), want_rows * (rep + 1.0), i, p, xr, spec.nrow, spec.ncol, tol=1e-12)


def test_one_shot_cycles_reuse_blocks_correctly():
    _cycles(6)


def test_cycles_after_trim():
    _cycles(2)
    _lib.check(_lib.lib().sb200_trim(0))
    _cycles(2)


@pytest.mark.parametrize("cache_mb", ["0", "8"])
def test_cycles_with_the_cache_off_or_too_small(cache_mb):
    """SB200_CACHE_MB is read once per process: a child process with the cache off (every block straight back to the
    driver pool) and with room for the small blocks only (the large ones are evicted at once)."""
    code = textwrap.dedent("""
        import importlib.util, sys
        sys.path.insert(0, %r)
        spec = importlib.util.spec_from_file_location("block_cache_cycles", %r)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod._cycles(3)
        print("ok")
    """ % (ROOT, os.path.abspath(__file__)))
    env = dict(os.environ, SB200_CACHE_MB=cache_mb)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
