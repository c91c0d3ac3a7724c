"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (masabdi/LSPS).

Only usable where /root/reference is mounted (the build container).  It is used by
`oracle/make_golden.py` to pin the oracle port (`oracle/lsps_oracle.py`) and to
generate the committed fixtures under tests/golden/.  Nothing in the product
package (`lsps_b200/`) may import this file.

Recipe: copy /root/reference/src to a scratch dir, expand tabs
(python-2 tab semantics, tab stop 8 -- `src/trainers/lsps_trainer.py:58` mixes tabs
and spaces), stub `matplotlib`, and make `.cuda()` a no-op on CPU.  No arithmetic
is edited.
"""
import os
import shutil
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("LSPS_REFERENCE_ROOT", "/root/reference")
# tab-expanded copy staged by oracle/build_ref.py (git-ignored; travels to the GPU box, where /root/reference is absent)
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "trainers"))


def staged_available():
    return os.path.isdir(os.path.join(STAGED_ROOT, "src", "trainers"))


_scratch = None


def load_reference(cpu_shim=True):
    """Returns the imported `trainers` package of the reference.  cpu_shim: True = make `.cuda()` a no-op when no GPU is
    present; "force" = always (the CPU arm of bench.py on a GPU box); False = never (tools/library_line.py on the GPU)."""
    global _scratch
    import torch

    if "trainers" in sys.modules and getattr(sys.modules["trainers"], "_lsps_ref", False):
        return sys.modules["trainers"]
    if not reference_available():
        if not staged_available():
            raise RuntimeError("reference sources neither mounted at %s nor staged at %s" % (REFERENCE_ROOT, STAGED_ROOT))
        # the staged copy is already tab-expanded: import it in place
        _scratch = STAGED_ROOT
        sys.path.insert(0, os.path.join(STAGED_ROOT, "src"))
        for m in ("matplotlib", "matplotlib.pyplot"):
            if m not in sys.modules:
                sys.modules[m] = types.ModuleType(m)
        if cpu_shim == "force" or (cpu_shim and not torch.cuda.is_available()):
            torch.Tensor.cuda = lambda self, *a, **k: self
            torch.nn.Module.cuda = lambda self, *a, **k: self
        import trainers  # noqa: the reference package
        trainers._lsps_ref = True
        return trainers
    _scratch = tempfile.mkdtemp(prefix="lsps_ref_")
    dst = os.path.join(_scratch, "src")
    shutil.copytree(os.path.join(REFERENCE_ROOT, "src"), dst,
                    ignore=shutil.ignore_patterns("*.pyc"))
    for root, _, files in os.walk(dst):
        for f in files:
            if f.endswith(".py"):
                p = os.path.join(root, f)
                with open(p, "r", encoding="utf-8", errors="replace") as fh:
                    text = fh.read()
                with open(p, "w", encoding="utf-8") as fh:
                    fh.write(text.expandtabs(8))
    sys.path.insert(0, dst)
    for m in ("matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    if cpu_shim == "force" or (cpu_shim and not torch.cuda.is_available()):
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    import trainers  # noqa: the reference package

    trainers._lsps_ref = True
    return trainers


def load_reference_augment():
    """The reference's input pipeline pieces for the section-8f n3 row: (normalize, augmentCrop, HandDetector,
    DepthImporter).  Extra shims on top of load_reference(): `utils/handdetector.py` has ONE python-2 print statement
    (:214, a warning) which is rewritten as a call, `xrange` is aliased to `range`, and `pylab` is stubbed like
    matplotlib.  No arithmetic is edited."""
    import builtins
    import re
    load_reference()
    src = os.path.join(_scratch, "src")
    hd = os.path.join(src, "utils", "handdetector.py")
    with open(hd) as fh:
        text = fh.read()
    text = re.sub(r'^(\s*)print "([^"]*)"\s*$', r'\1print("\2")', text, flags=re.M)
    with open(hd, "w") as fh:
        fh.write(text)
    builtins.xrange = range
    for m in ("pylab", "progressbar"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    if "cPickle" not in sys.modules:                           # python-2 module name (data/importers.py:34)
        import pickle
        sys.modules["cPickle"] = pickle
    from data.dataset_hand2 import normalize, augmentCrop      # noqa
    from utils.handdetector import HandDetector                # noqa
    from data.importers import DepthImporter                   # noqa
    return normalize, augmentCrop, HandDetector, DepthImporter


def load_hyperparameters(name="nnyu"):
    import yaml

    root = REFERENCE_ROOT if reference_available() else STAGED_ROOT
    with open(os.path.join(root, "exps", name + ".yaml")) as fh:
        return yaml.safe_load(fh)["train"]["hyperparameters"]
