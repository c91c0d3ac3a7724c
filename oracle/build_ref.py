"""TEST INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference (masabdi/LSPS, pure Python) under oracle/_ref/.

The reference has no native sources, so there is nothing to compile; what `oracle/_ref/` carries is the reference's own
`src/` and `exps/` trees with ONE mechanical change: every `.py` file is rewritten with `str.expandtabs(8)` (python-2
tab semantics; `src/trainers/lsps_trainer.py:58` mixes tabs and spaces and does not compile under python 3 otherwise).
No arithmetic is edited.  `oracle/_ref/` is git-ignored (reference sources never enter this repo's history) but not
gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not exist; there it serves
  * `bench.py --impl reference` / `cpu_baseline` (kind "reference": the reference's own LSPSTrainer on the host cores),
  * `tools/library_line.py` (the same trainer on the B200 through torch/cuDNN -- the "library line").

Run:  python oracle/build_ref.py        (only where /root/reference is mounted; a no-op elsewhere)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LSPS_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build_ref(force=False):
    if not os.path.isdir(os.path.join(REF, "src", "trainers")):
        return None
    stamp = os.path.join(DST, ".stamp")
    if os.path.exists(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(REF, "src"), os.path.join(DST, "src"), ignore=shutil.ignore_patterns("*.pyc"))
    shutil.copytree(os.path.join(REF, "exps"), os.path.join(DST, "exps"))
    for root, _, files in os.walk(os.path.join(DST, "src")):
        for f in files:
            if f.endswith(".py"):
                p = os.path.join(root, f)
                with open(p, "r", encoding="utf-8", errors="replace") as fh:
                    text = fh.read()
                with open(p, "w", encoding="utf-8") as fh:
                    fh.write(text.expandtabs(8))
    with open(stamp, "w") as fh:
        fh.write("tab-expanded copy of %s/src and exps; see oracle/build_ref.py\n" % REF)
    return DST


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
