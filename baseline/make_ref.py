"""Copy the reference's hot-path sources into baseline/_ref/ so that the reference itself can run on the GPU box.

    python baseline/make_ref.py            (in the build container, where /root/reference exists)

/root/reference is not part of the snapshot gpurun ships; baseline/_ref/ is (it is git-ignored, NOT gpurun-ignored).  Only the
packages the rendering path imports are copied, unmodified: src/model, src/render, src/util and conf/ (SURVEY.md section 8c).
The three import shims the reference needs here (pyhocon, dotmap, the un-vendored models.yolo) live in baseline/ref_shims.py.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PNR_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make_ref(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(REF, "src")):
        raise FileNotFoundError(f"{REF}/src not found: the reference is only mounted in the build container")
    stamp = os.path.join(DST, ".copied_from")
    if os.path.exists(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, "src"))
    for pkg in ("model", "render", "util"):
        shutil.copytree(os.path.join(REF, "src", pkg), os.path.join(DST, "src", pkg),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copytree(os.path.join(REF, "conf"), os.path.join(DST, "conf"))
    with open(stamp, "w") as fh:
        fh.write(REF + "\n")
    return DST


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
