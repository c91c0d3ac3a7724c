"""Run one golden case REPS times in one process and print, per repetition, the relative difference of every scalar
loss to the reference's value -- the spread is the run-to-run nondeterminism of the fp32-atomic statistics amplified by
the free-running second step.  Test-side tool (uses the oracle to seed the weights, like tests/test_trainer_gpu.py).

    python tools/repeat_golden.py pretrain_resx_nnyu_b1 20
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import lsps_oracle as O                                  # noqa: E402
from common import GOLDEN_CASES, load_from_oracle, load_hp, run_schedule   # noqa: E402


def main():
    import lsps_b200
    case, reps = sys.argv[1], int(sys.argv[2])
    cfg, schedule, batch, steps, kind = GOLDEN_CASES[case]
    gold = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
    keys = [k for k in gold.files if np.ndim(gold[k]) == 0 and not k.startswith(("meta_", "w_")) and not k.endswith("_acc")]
    worst = []
    for r in range(reps):
        hp = load_hp(cfg)
        oracle = O.OracleTrainer(hp, seed=int(gold["meta_seed"]))
        tr = lsps_b200.LSPSTrainerB200(hp, device=0, noise="host")
        load_from_oracle(tr, oracle)
        rec = run_schedule(tr, hp, schedule, batch, steps, kind, device="cuda")
        d = {k: abs(float(rec[k]) - float(gold[k])) / (abs(float(gold[k])) + 1e-5) for k in keys}
        k = max(d, key=d.get)
        worst.append(d[k])
        print("rep %2d  worst %.2e (%s)   s1_gen_enc_loss2 %.2e  s1_gen_total_loss %.2e" % (
            r, d[k], k, d.get("s1_gen_enc_loss2", 0), d.get("s1_gen_total_loss", 0)), flush=True)
    w = np.array(worst)
    print("%s x%d: worst-loss relative difference min %.2e median %.2e max %.2e" % (case, reps, w.min(), np.median(w), w.max()))


if __name__ == "__main__":
    main()
