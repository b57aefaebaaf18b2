"""Stage the UNMODIFIED reference package for the GPU box: /root/reference/gym_anm -> oracle/_ref/gym_anm.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is a copy of its package
directory (the `.py` files, nothing else) into the git-ignored output directory `oracle/_ref/`, which travels to the
GPU box with the snapshot like the built `.so` files do.  There, `oracle/ref_loader.py` imports it (with the
stand-ins of oracle/shims/ for the uninstalled cvxpy / gymnasium), so `bench.py --impl reference` and `cpu_baseline`
time the reference's own `ANMEnv.step` -> `Simulator.transition` code (kind "reference"), not a port.
Nothing under oracle/_ref is committed; the product never imports it.

    python oracle/build_ref.py        (run by __graft_entry__.build() when /root/reference exists)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ANM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(force=False):
    src_pkg = os.path.join(SRC, "gym_anm")
    if not os.path.isdir(src_pkg):
        return None  # GPU box: use what was staged here
    dst_pkg = os.path.join(DST, "gym_anm")
    if os.path.isdir(dst_pkg) and not force:
        newest = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(src_pkg) for f in fs if f.endswith(".py"))
        if os.path.getmtime(dst_pkg) >= newest:
            return dst_pkg
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src_pkg, dst_pkg, ignore=lambda d, names: [n for n in names if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))])
    os.utime(dst_pkg, None)
    return dst_pkg


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
