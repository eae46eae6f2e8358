"""ORACLE (test infrastructure) -- stages the UNMODIFIED reference package for the GPU box.

The reference is pure Python, so "building" it is a copy: /root/reference/graphslim (sources + its JSON configs) ->
oracle/_ref/graphslim.  oracle/_ref/ is git-ignored (no reference source ever enters the history) but not
gpurun-ignored, so it travels to the GPU box with the snapshot, where `bench.py --impl reference` runs the reference's own
`GCond(...).reduce` through oracle/ref_shim (CPU stand-ins for the third-party wheels that are not installed) on the
box's host cores (`cpu_baseline.kind == "reference"`).  Without a staged copy the arm falls back to the oracle
restatement (`kind == "port"`).

    python -m oracle.stage_ref            # called by __graft_entry__.build() when /root/reference exists
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("GRAPHSLIM_REFERENCE_SOURCE", "/root/reference")


def staged():
    return os.path.isfile(os.path.join(REF_DIR, "graphslim", "condensation", "gcond.py"))


def stage(force=False):
    """Returns the staged root (oracle/_ref) or None when there is nothing to stage from."""
    src = os.path.join(SOURCE, "graphslim")
    if not os.path.isdir(src):
        return REF_DIR if staged() else None
    stamp = os.path.join(REF_DIR, "STAGED_FROM")
    if staged() and not force and os.path.exists(stamp):
        return REF_DIR
    dst = os.path.join(REF_DIR, "graphslim")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(REF_DIR, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.sh", "scripts"))
    n = sum(len(files) for _, _, files in os.walk(dst))
    with open(stamp, "w") as f:
        f.write(f"{src}\n{n} files copied verbatim by oracle/stage_ref.py\n")
    return REF_DIR


if __name__ == "__main__":
    print(stage(force=True))
