#!/usr/bin/env python
"""Recipe for oracle/_ref/: the REFERENCE's own implementation of the hot path, made importable where /root/reference is
absent (the GPU box).  TEST / BASELINE INFRASTRUCTURE -- never imported by the product.

    python oracle/make_ref.py        # build container only (needs /root/reference, read-only)

The reference is a Python source tree (no compiled code); its path lives in src/models/dicow/*.py, which depends only on
torch, transformers, numpy and pandas -- all in this image.  This script copies those files, unmodified, from where they
lie under /root/reference into oracle/_ref/models/dicow/ .  oracle/_ref/ is listed in .gitignore (it stays out of the
history: the repo contains no reference sources) but not in .gpurunignore, so it travels to the GPU box like a built .so;
__graft_entry__.build() runs this when /root/reference is present.  oracle/ref_loader.py imports the copy with the
transformers 4.55 -> 5.x compatibility shim of SURVEY.md section 8c; bench.py --impl reference times its
DiCoWEncoder.forward (cpu_baseline.kind = "reference"), and falls back to the oracle port where the copy is missing."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/models/dicow"
DST = os.path.join(HERE, "_ref", "models", "dicow")


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: oracle/_ref left as it is")
        return False
    os.makedirs(DST, exist_ok=True)
    names = sorted(n for n in os.listdir(SRC) if n.endswith(".py"))
    for n in names:
        shutil.copyfile(os.path.join(SRC, n), os.path.join(DST, n))
    init = os.path.join(os.path.dirname(DST), "__init__.py")
    if not os.path.exists(init):
        open(init, "w").close()
    if verbose:
        print(f"oracle/_ref: {len(names)} reference modules from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
