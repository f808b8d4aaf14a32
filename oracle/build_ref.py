"""ORACLE - TEST INFRASTRUCTURE ONLY.  Recipe for oracle/_ref/: a git-ignored, build-time copy of the reference's own
`models/` package (Python sources only) taken from /root/reference where it lies, so that the GPU box - which has no
/root/reference - can time the REFERENCE's own CPU implementation of the hot path (bench.py --impl reference,
cpu_baseline.kind = "reference") and so that oracle/gen_golden.py style checks can run there.

  python oracle/build_ref.py            # in the build container; no-op (exit 0) when /root/reference is absent

Nothing under oracle/_ref/ is tracked (.gitignore) and the product package never imports it; it travels with the
gpurun snapshot like the built .so files.  The reference is a pure-Python/PyTorch repository: there is nothing to
compile, "building" the reference arm is this copy.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MVSF_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build_ref(verbose=True):
    src_models = os.path.join(SRC, "models")
    if not os.path.isdir(src_models):
        if verbose:
            print(f"oracle/build_ref: {SRC} not present - keeping whatever oracle/_ref already holds")
        return os.path.isdir(os.path.join(DST, "models"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for root, dirs, files in os.walk(src_models):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, SRC)
        os.makedirs(os.path.join(DST, rel), exist_ok=True)
        for f in files:
            if f.endswith(".py"):
                shutil.copy2(os.path.join(root, f), os.path.join(DST, rel, f))
                n += 1
    cfg = os.path.join(SRC, "config", "mvsformer++.json")
    if os.path.isfile(cfg):
        os.makedirs(os.path.join(DST, "config"), exist_ok=True)
        shutil.copy2(cfg, os.path.join(DST, "config", "mvsformer++.json"))
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Build output of oracle/build_ref.py: untracked copy of the reference's models/ package for the CPU "
                "reference arm. Do not edit, do not commit.\n")
    if verbose:
        print(f"oracle/build_ref: copied {n} python files to {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build_ref() or True else 1)
