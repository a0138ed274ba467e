"""Builds tests/dropin/libdropin_shim.so: user code against include/RcppSparse.h, compiled with the Rcpp
stand-in from oracle/stub (test infrastructure) and linked to libsparse_b200."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "rcppsparse_b200")
OUT = os.path.join(HERE, "libdropin_shim.so")


def build(force: bool = False) -> str:
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from rcppsparse_b200 import build as libbuild

    lib = libbuild.build_library()
    src = os.path.join(HERE, "dropin_shim.cpp")
    deps = [src, lib, os.path.join(ROOT, "include", "RcppSparse.h"), os.path.join(ROOT, "include", "sparse_b200.h"),
            os.path.join(ROOT, "oracle", "stub", "Rcpp.h")]
    if force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        cxx = shutil.which("g++") or "g++"
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I" + os.path.join(ROOT, "include"),
               "-I" + os.path.join(ROOT, "oracle", "stub"), src, "-o", OUT, "-L" + PKG, "-lsparse_b200",
               "-Wl,-rpath," + PKG, "-Wl,-rpath,$ORIGIN/../../rcppsparse_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on the drop-in header test:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
