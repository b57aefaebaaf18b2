"""Stage the UNMODIFIED reference package for the GPU box: /root/reference/gym_anm -> oracle/_ref/gym_anm_ref.zip.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it means packing its package
directory (the `.py` files, nothing else, byte for byte) into ONE archive in the git-ignored output directory
`oracle/_ref/`, which travels to the GPU box with the snapshot like the built `.so` files do.  There,
`oracle/ref_loader.py` puts the archive on `sys.path` (zipimport; with the stand-ins of oracle/shims/ for the
uninstalled cvxpy / gymnasium), so `bench.py --impl reference` and `cpu_baseline` time the reference's own
`ANMEnv.step` -> `Simulator.transition` code (kind "reference"), not a port.  Nothing under oracle/_ref is committed,
no reference source file is copied into the tree, and the product never imports it.

    python oracle/build_ref.py        (run by __graft_entry__.build() when /root/reference exists)
"""
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ANM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "gym_anm_ref.zip")


def build(force=False):
    src_pkg = os.path.join(SRC, "gym_anm")
    if not os.path.isdir(src_pkg):
        return DST if os.path.exists(DST) else None  # GPU box: use what was staged
    files = sorted(os.path.join(r, f) for r, _, fs in os.walk(src_pkg) for f in fs if f.endswith(".py"))
    if os.path.exists(DST) and not force and os.path.getmtime(DST) >= max(os.path.getmtime(f) for f in files):
        return DST
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    tmp = DST + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for f in files:
            z.write(f, os.path.relpath(f, SRC))
    os.replace(tmp, DST)
    return DST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
