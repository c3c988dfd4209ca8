"""The product (eigen_b200/, include/) never imports, links or mentions the CPU oracle: oracle/ is test infrastructure
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may touch it), and there is no
CPU fallback to fall back to."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _product_files():
    for base in ("eigen_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            if os.sep + "build" in d or "__pycache__" in d:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                    yield os.path.join(d, f)


def test_no_reference_to_the_oracle_in_product_sources():
    offenders = [p for p in _product_files() if re.search(r"\boracle\b", open(p, errors="replace").read())]
    assert not offenders, offenders


def test_library_does_not_link_the_oracle_or_a_cpu_blas():
    so = os.path.join(ROOT, "eigen_b200", "libb200blas.so")
    if not os.path.exists(so):   # *.so is git-ignored: build it rather than pass vacuously on an empty ldd
        import eigen_b200
        eigen_b200.build()
    assert os.path.exists(so), "libb200blas.so is missing: nothing was checked"
    out = subprocess.run(["ldd", so], stdout=subprocess.PIPE, text=True).stdout
    assert "libc.so" in out, out   # ldd really resolved the library
    for banned in ("liboracle", "eigen_blas", "openblas", "libblas", "mkl", "cublas", "nccl"):
        assert banned not in out, (banned, out)
