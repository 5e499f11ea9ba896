"""Recipe that stages the UNMODIFIED reference into oracle/_ref/ (git-ignored; travels to the GPU box with the
snapshot like a built .so).  TEST / BENCH INFRASTRUCTURE ONLY -- nothing in pointcloud_rl_b200/ reads it.

The reference is pure Python (no native code), so "building" it is copying the two trees its
update path imports -- pyrl/ and configs/ (1.3 MB) -- from the read-only mount; missing third-party modules are
stubbed by oracle/refshim.  No reference file is edited, and none enters git history.

    python oracle/build_ref.py        # also run by __graft_entry__.build() when /root/reference is present
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PCRL_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build_ref(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "pyrl")):
        if verbose:
            print(f"[oracle/_ref] {SRC} not present: keeping the prebuilt copy" if os.path.isdir(DST) else
                  f"[oracle/_ref] {SRC} not present and no prebuilt copy")
        return os.path.isdir(os.path.join(DST, "pyrl"))
    for sub in ("pyrl", "configs"):
        dst = os.path.join(DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for root, dirs, files in os.walk(DST):  # the mount is read-only; the copy must be readable and traversable
        for d in dirs:
            os.chmod(os.path.join(root, d), 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"[oracle/_ref] staged {n} files from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build_ref() else 1)
