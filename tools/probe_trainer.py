"""GPU probe: run golden cases through LSPSTrainerB200 and print every scalar next to the reference value."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lsps_oracle as O  # noqa
from common import GOLDEN_CASES, load_from_oracle, run_schedule  # noqa
import lsps_b200  # noqa

names = sys.argv[1:] or list(GOLDEN_CASES)
for case in names:
    cfg, schedule, batch, steps, kind = GOLDEN_CASES[case]
    hp = lsps_b200.load_hyperparameters(cfg)
    gold = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
    oracle = O.OracleTrainer(hp, seed=0)
    tr = lsps_b200.LSPSTrainerB200(hp, device=0, noise="host")
    load_from_oracle(tr, oracle)
    t0 = time.time()
    rec = run_schedule(tr, hp, schedule, batch, steps, kind, device="cuda")
    torch.cuda.synchronize()
    print("== %s (%.1fs)" % (case, time.time() - t0), flush=True)
    for k in gold.files:
        if k.startswith("meta_") or k.startswith("w_"):
            continue
        ref, got = gold[k], rec[k]
        if np.ndim(ref) == 0:
            print("   %-22s ref %.6f got %.6f rel %.2e" % (k, ref, got, abs(got - ref) / (abs(ref) + 1e-12)))
        else:
            print("   %-22s max abs err %.3e (ref absmean %.3f)" % (k, np.max(np.abs(ref - got)), ref[-2]))
