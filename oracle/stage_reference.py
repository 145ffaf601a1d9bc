"""Stage the unmodified reference package where the GPU box can import it (TEST / BASELINE INFRASTRUCTURE).

    python oracle/stage_reference.py [--force]

Copies ``/root/reference/dro_sfm`` (Python sources only), ``configs/`` and the licence into the git-ignored
``baseline/_ref/`` -- the pure-Python counterpart of compiling a C reference into ``oracle/_ref``: nothing is edited,
nothing is committed, and the copy travels with the repo snapshot so that the `-m gpu` end-to-end test and
``bench.py``'s reference legs can run the reference's own code on the B200 box (where ``/root/reference`` does not
exist).  A no-op when the source tree is absent (the GPU box uses the staged copy as it arrived).
"""
import os
import shutil
import sys

SRC = os.environ.get("DROSFM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def stage(force=False):
    if not os.path.isdir(os.path.join(SRC, "dro_sfm")):
        return None
    marker = os.path.join(DST, ".staged_from")
    if not force and os.path.exists(marker) and os.path.isdir(os.path.join(DST, "dro_sfm")):
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    keep = shutil.ignore_patterns("__pycache__", "*.pyc", "*.png", "*.jpg", "*.gif", "*.npz", "*.ckpt")
    shutil.copytree(os.path.join(SRC, "dro_sfm"), os.path.join(DST, "dro_sfm"), ignore=keep)
    if os.path.isdir(os.path.join(SRC, "configs")):
        shutil.copytree(os.path.join(SRC, "configs"), os.path.join(DST, "configs"), ignore=keep)
    for name in ("LICENSE", "README.md"):
        if os.path.exists(os.path.join(SRC, name)):
            shutil.copy2(os.path.join(SRC, name), os.path.join(DST, name))
    with open(marker, "w") as f:
        f.write(SRC + "\n")
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
