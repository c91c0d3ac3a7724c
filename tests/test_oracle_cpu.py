"""The CPU oracle (oracle/lsps_oracle.py) must reproduce the committed outputs of the UNMODIFIED reference
(tests/golden/*.npz, written by oracle/make_golden.py in a container where /root/reference is mounted)."""
import os

import numpy as np
import pytest
import torch
import yaml

import lsps_oracle as O
from common import GOLDEN_CASES, run_schedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _hp(name):
    from common import load_hp
    return load_hp(name)


@pytest.mark.parametrize("case", sorted(GOLDEN_CASES))
def test_oracle_reproduces_reference(case, golden_dir):
    cfg, schedule, batch, steps, kind = GOLDEN_CASES[case]
    hp = _hp(cfg)
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    assert int(gold["meta_batch"]) == batch and int(gold["meta_steps"]) == steps and str(gold["meta_kind"]) == kind
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    oracle = O.OracleTrainer(hp, seed=int(gold["meta_seed"]))
    rec = run_schedule(oracle, hp, schedule, batch, steps, kind)
    sd = oracle.state_dict
    from common import sample
    for k in gold.files:
        if k.startswith("meta_"):
            continue
        if k.startswith("w_"):
            net, key = k[2:].split("_", 1)
            got = sample(sd(net)[key])
        else:
            got = rec[k]
        ref = gold[k]
        if np.ndim(ref) == 0:
            # fp32 thread-count noise through a GAN step: SURVEY appendix A (2e-4 at step 3)
            assert abs(float(got) - float(ref)) <= 5e-4 * abs(float(ref)) + 1e-6, (k, float(ref), float(got))
        else:
            assert np.max(np.abs(ref - got)) <= 5e-4 * (np.max(np.abs(ref)) + 1e-6) + 1e-6, k


def test_oracle_header_says_test_infrastructure():
    for f in ("lsps_oracle.py", "ref_loader.py", "make_golden.py", "augment_oracle.py", "make_augment_golden.py"):
        with open(os.path.join(ROOT, "oracle", f)) as fh:
            assert "TEST INFRASTRUCTURE ONLY" in fh.read(400)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lsps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    txt = fh.read()
                for bad in ("import lsps_oracle", "from lsps_oracle", "ref_loader", "oracle/", "oracle."):
                    assert bad not in txt, (f, bad)
