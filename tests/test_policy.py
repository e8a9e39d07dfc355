"""The oracle is test infrastructure: nothing in the product tree may import, link or execute it (a product path that
routes through the oracle, or any CPU fallback, would void every parity claim).  Also: every source kernel file is built
for sm_100a only."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _files(top, exts):
    for d, _, fs in os.walk(os.path.join(ROOT, top)):
        if "build" in d or "__pycache__" in d:
            continue
        for f in fs:
            if f.endswith(exts):
                yield os.path.join(d, f)


def test_product_tree_never_touches_the_oracle():
    pat = re.compile(r"pyoracle|pytypes|sqt_oracle|libsqt_oracle|oracle/_ref|liblmp2sqt_ref")
    bad = [p for top in ("lmp2sqt_b200", "include") for p in _files(top, (".py", ".c", ".cpp", ".cu", ".cuh", ".h", "Makefile"))
           if pat.search(open(p, errors="replace").read())]
    assert not bad, bad


def test_bench_uses_the_oracle_only_as_checker_and_cpu_baseline():
    """bench.py touches oracle/ in two functions only: reference_sample (the cpu_baseline key and --impl reference: the
    reference's CPU code is the thing timed there) and parity_check (the oracle as the CHECKER of a small sharded run,
    before anything is timed).  Nothing that produces `value` or `e2e` may."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    funcs = re.split(r"\ndef ", src)
    users = [f.split("(")[0] for f in funcs if "pyoracle" in f]
    assert users and set(users) <= {"reference_sample", "parity_check"}, users
    timed = [f for f in funcs if f.startswith("time_workload(")]
    assert len(timed) == 1 and "oracle" not in timed[0]


def test_kernels_are_built_for_sm_100a_only():
    mk = open(os.path.join(ROOT, "lmp2sqt_b200", "csrc", "Makefile")).read()
    assert "arch=compute_100a,code=sm_100a" in mk and "-lineinfo" in mk
    assert not re.search(r"sm_(?!100a)\d+", mk)
