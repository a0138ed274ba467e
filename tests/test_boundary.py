"""CPU-side checks of the drop-in boundary (no GPU, no compute calls):
the shared library loads, exports exactly what include/sparse_b200.h declares, fails loudly
without a device, and the product never touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sparse_b200.h")


@pytest.fixture(scope="module")
def built_lib():
    from rcppsparse_b200 import build

    return build.build_library()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_the_binding_binds():
    from rcppsparse_b200 import _lib

    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_status_codes_and_create_flags_agree_with_the_header():
    """The ctypes binding repeats the header's constants: status codes and the flags of sb200_matrix_create."""
    from rcppsparse_b200 import _lib

    src = open(HEADER).read()
    consts = {k: v for k, v in re.findall(r"#define\s+SB200_([A-Z_]+)\s+\(?(-?\d+)u?\)?", src)}
    for name in ("E_INVALID", "E_CUDA", "E_NOMEM", "E_STRUCTURE", "E_NODEVICE", "E_UNSUPPORTED",
                 "PIN_HOST", "NO_VALIDATE", "NO_ROW_PLAN", "LAZY_ROWS"):
        assert name in consts, name
        assert int(consts[name]) == getattr(_lib, name), name


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (sb200_[a-z0-9_]+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    lib = ctypes.CDLL(built_lib)
    for s in declared_symbols():
        getattr(lib, s)


def test_library_loads_and_reports_abi(built_lib):
    from rcppsparse_b200 import _lib

    L = _lib.lib()
    assert L.sb200_abi_version() == 1
    assert isinstance(L.sb200_last_error(), bytes)
    assert L.sb200_launch_count() >= 0


def test_no_device_is_an_error_not_a_fallback(built_lib):
    """On a box without a GPU every entry that would compute must fail with SB200_E_NODEVICE."""
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rcppsparse_b200 import Matrix, SparseB200Error, _lib

    assert _lib.device_count() == 0
    m = Matrix(np.array([1.0]), np.array([0], np.int32), np.array([0, 1], np.int32), np.array([1, 1], np.int32))
    with pytest.raises(SparseB200Error) as ei:
        m.colSums()
    assert ei.value.code == _lib.E_NODEVICE
    assert "no CPU fallback" in ei.value.message or "no CUDA device" in ei.value.message


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the package, the header or the C sources may name it."""
    pat = re.compile(r"\boracle\b|liboracle|oport_|oref_")
    offenders = []
    for base in (os.path.join(ROOT, "rcppsparse_b200"), os.path.join(ROOT, "include")):
        for dp, _, files in os.walk(base):
            if os.path.basename(dp) in ("build", "__pycache__"):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")):
                    text = open(os.path.join(dp, f), errors="replace").read()
                    for ln, line in enumerate(text.splitlines(), 1):
                        if pat.search(line) and "the CPU oracle can be fed" not in line:
                            offenders.append(f"{os.path.relpath(os.path.join(dp, f), ROOT)}:{ln}: {line.strip()}")
    assert not offenders, "\n".join(offenders)


def test_missing_library_raises(monkeypatch, tmp_path):
    from rcppsparse_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.lib()


# ---- the R package around the drop-in: R is not installed here, so what can be checked is that
# every layer names the same entry points -------------------------------------------------------------------------
def test_r_package_entry_points_consistent():
    pkg = os.path.join(ROOT, "r-package")
    exports_cpp = open(os.path.join(pkg, "src", "exports.cpp")).read()
    rcpp_exports = open(os.path.join(pkg, "src", "RcppExports.cpp")).read()
    r_code = "".join(open(os.path.join(pkg, "R", f)).read() for f in sorted(os.listdir(os.path.join(pkg, "R"))))
    namespace = open(os.path.join(pkg, "NAMESPACE")).read()
    # functions tagged for export in C++ (name and arity)
    tagged = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
              for m in re.finditer(r"//\[\[Rcpp::export\]\]\s*\n[\w:<>&\s]+?\b(\w+)\(([^)]*)\)", exports_cpp)}
    assert "columnSums" in tagged and tagged["columnSums"] == 1  # the reference's exported function, same arity
    # the registration table: one entry per tagged function, arity registered (reference src/RcppExports.cpp:26-30)
    table = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"_RcppSparse_(\w+)", \(DL_FUNC\) &_RcppSparse_\1, (\d+)\}', rcpp_exports)}
    assert table == tagged
    assert "R_useDynamicSymbols(dll, FALSE)" in rcpp_exports
    # every .Call in the R code hits a registered entry, and every registered entry is reachable from R
    called = set(re.findall(r"\.Call\(`_RcppSparse_(\w+)`", r_code))
    assert called == set(table)
    # NAMESPACE exports exist as R functions
    exported = set(sum((re.split(r"\s*,\s*", m) for m in re.findall(r"export\(([^)]*)\)", namespace)), []))
    defined = set(re.findall(r"^(\w+)\s*<-\s*function", r_code, flags=re.M))
    assert exported and exported <= defined, exported - defined
    assert "useDynLib(RcppSparse, .registration = TRUE)" in namespace


def test_r_package_call_table_is_what_the_generator_writes(tmp_path, monkeypatch):
    """src/RcppExports.cpp is generated from the //[[Rcpp::export]] tags (tools/gen_rcpp_exports.py stands in for
    Rcpp::compileAttributes()); the committed file must be up to date."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tools", "gen_rcpp_exports.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    committed = open(os.path.join(ROOT, "r-package", "src", "RcppExports.cpp")).read()
    src = tmp_path / "src"
    src.mkdir()
    (src / "exports.cpp").write_text(open(os.path.join(ROOT, "r-package", "src", "exports.cpp")).read())
    monkeypatch.setattr(gen, "SRC", str(src))
    gen.main()
    assert (src / "RcppExports.cpp").read_text() == committed
