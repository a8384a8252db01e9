"""Recipe for oracle/_ref: the UNMODIFIED reference package, staged for the CPU arm of bench.py (test infrastructure).

The reference is pure Python (numba-jitted), so "building" it means making it importable where /root/reference does not
exist (the GPU box): the package directory is copied from where it lies into oracle/_ref/ (git-ignored, so no reference
source enters the history; not gpurun-ignored, so it travels with the snapshot), and the cache file the reference writes
on first use -- pttools/bubble/fluid_reference.hdf5, bubble/fluid_reference.py:105-145 -- is put in place from the copy
of that very table the reference produced in the build container (tests/golden/ref_fluid_reference.npz, provenance:
tools/make_golden.py::fluid_reference_table), in the npz form the h5py stand-in of tools/refenv/stubs reads.  Only
bench.py's CPU arm and its cpu_baseline leg import oracle/_ref; the product never does (tests/test_abi.py).
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference/pttools"
DST = os.path.join(HERE, "_ref")


def build(force: bool = False) -> str | None:
    """Returns the directory to put on sys.path, or None when the reference is not available here and was not staged
    before."""
    pkg = os.path.join(DST, "pttools")
    if os.path.isdir(REF_SRC) and (force or not os.path.isdir(pkg)):
        if os.path.isdir(pkg):
            shutil.rmtree(pkg)
        shutil.copytree(REF_SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for base, dirs, files in os.walk(pkg):
            for name in dirs + files:
                p = os.path.join(base, name)
                os.chmod(p, os.stat(p).st_mode | 0o200)
    if not os.path.isdir(pkg):
        return None
    cache = os.path.join(pkg, "bubble", "fluid_reference.hdf5")
    if not os.path.exists(cache):
        shutil.copyfile(os.path.join(ROOT, "tests", "golden", "ref_fluid_reference.npz"), cache)
    return DST


if __name__ == "__main__":
    print(build(force=True))
