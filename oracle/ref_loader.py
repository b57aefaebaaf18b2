"""Import the UNMODIFIED reference package from /root/reference (this container) or from the archive that
oracle/build_ref.py staged under oracle/_ref/ (the GPU box, where /root/reference does not exist; zipimport).

TEST INFRASTRUCTURE ONLY.  Nothing in the product imports this.  The GPU box has no
/root/reference, so this loader is used (a) by oracle/gen_golden.py to generate the
committed fixtures in tests/golden/, and (b) by the `not gpu` tests that pin the
restatements (oracle/anm_numpy.py, oracle/anm_oracle.c) against the real thing when
the reference tree is present.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# written by oracle/build_ref.py: the unmodified package as one archive (git-ignored; what the GPU box has)
_STAGED = os.path.join(_HERE, "_ref", "gym_anm_ref.zip")
REFERENCE_ROOT = os.environ.get("ANM_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/gym_anm") else _STAGED)
_SHIMS = os.path.join(_HERE, "shims")


def reference_available():
    """A source tree (this container) or the staged archive (zipimport)."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_anm")) or (
        REFERENCE_ROOT.endswith(".zip") and os.path.isfile(REFERENCE_ROOT))


def _needs_shim(name):
    try:
        __import__(name)
        return False
    except ImportError:
        return True


def load_reference():
    """Return the reference `gym_anm` module (stand-ins only for what is not installed)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "gym_anm" in sys.modules:
        return sys.modules["gym_anm"]
    missing = [m for m in ("cvxpy", "gymnasium", "websocket", "websocket_server") if _needs_shim(m)]
    if missing and _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    import gym_anm  # noqa: E402

    return gym_anm
