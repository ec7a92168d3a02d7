"""Copy recipe for the real reference as the CPU arm (test / bench infrastructure only).

helicon is pure Python, so "building" the reference for this path means making its package importable where
/root/reference does not exist (the GPU box): this script copies the UNMODIFIED package tree
/root/reference/src/helicon -> oracle/_ref/helicon.  oracle/_ref/ is git-ignored (no reference source enters the
history) but NOT gpurun-ignored, so it travels with the snapshot like the built .so files.  bench.py --impl reference and
bench.py's cpu_baseline leg import it from there (kind "reference"); when it is absent they fall back to the oracle port
(kind "port").  Nothing under helicon_b200/ imports it.

Usage: python oracle/build_ref.py        (run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

SRC = "/root/reference/src/helicon"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "helicon")
# the reference's own tests of the path: run against helicon_b200 through the module alias of
# INTEGRATION.md by tests/test_reference_contract.py; copied next to the package, git-ignored like it
TESTS = ("test_denovo3D_solver.py", "test_denovo3D_pipeline.py")
TSRC = "/root/reference/tests"
TDST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_tests")


def build_ref(force=False):
    if os.path.isdir(TSRC) and (force or not os.path.isdir(TDST)):
        os.makedirs(TDST, exist_ok=True)
        for t in TESTS:
            shutil.copyfile(os.path.join(TSRC, t), os.path.join(TDST, t))
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    if os.path.isdir(DST) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    ok = build_ref(force="--force" in sys.argv)
    print("oracle/_ref/helicon", "present" if ok else "absent (no /root/reference here)")
