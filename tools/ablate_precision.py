#!/usr/bin/env python
"""Which network's bf16 operands cause the adversarial-loss deviation?  ("weak" item 1 / next-round 1b)

CPU experiment on the oracle (test infrastructure): a master fp32 oracle trains `--steps` pretrain steps at batch 1;
before every step a second oracle takes the master's weights (teacher forcing) and runs the SAME step on the same batch
and host noise with the conv operands (input, weight, incoming gradient) and the conv outputs rounded to bf16 (RNE) --
in the generator only, in the discriminator only, or in both.  The losses of the emulated step are compared with the
master's.  This is the same protocol as tests/test_trainer_gpu.py::test_100_steps_teacher_forced..., so the curves are
directly comparable with the CUDA path's.

  python tools/ablate_precision.py --steps 100 --out profiles/r02_precision_ablation.json
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lsps_oracle as O  # noqa: E402

SCOPE = {"cur": None, "on": set(), "mode": "bf16", "what": "xwgo"}   # what: x inputs, w weights, g gradients, o outputs


def _rnd(t, tag="x"):
    if tag not in SCOPE["what"]:
        return t
    if SCOPE["mode"] == "bf16":
        return t.to(torch.bfloat16).to(torch.float32)
    # split-bf16 (hi + lo): 16 mantissa bits
    hi = t.to(torch.bfloat16).to(torch.float32)
    lo = (t - hi).to(torch.bfloat16).to(torch.float32)
    return hi + lo


class _EmuConv(torch.autograd.Function):
    """conv / transposed conv with rounded operands; the fp32 accumulation is torch's own."""

    @staticmethod
    def forward(ctx, x, w, b, transposed, stride, padding, output_padding):
        xr, wr = _rnd(x, "x"), _rnd(w, "w")
        ctx.save_for_backward(xr, wr)
        ctx.cfg = (transposed, stride, padding, output_padding, b is not None)
        if transposed:
            y = F.conv_transpose2d(xr, wr, b, stride=stride, padding=padding, output_padding=output_padding)
        else:
            y = F.conv2d(xr, wr, b, stride=stride, padding=padding)
        return y

    @staticmethod
    def backward(ctx, gy):
        xr, wr = ctx.saved_tensors
        transposed, stride, padding, output_padding, has_b = ctx.cfg
        gyr = _rnd(gy, "g")
        with torch.enable_grad():
            xd, wd = xr.detach().requires_grad_(True), wr.detach().requires_grad_(True)
            if transposed:
                y = F.conv_transpose2d(xd, wd, None, stride=stride, padding=padding, output_padding=output_padding)
            else:
                y = F.conv2d(xd, wd, None, stride=stride, padding=padding)
            gx, gw = torch.autograd.grad(y, (xd, wd), gyr)
        gb = gy.sum((0, 2, 3)) if has_b else None
        return gx, gw, gb, None, None, None, None


class _FProxy:
    """Stands in for torch.nn.functional inside the oracle module: convs are emulated when the current scope is on."""

    def __getattr__(self, name):
        return getattr(F, name)

    @staticmethod
    def conv2d(x, w, b=None, stride=1, padding=0):
        if SCOPE["cur"] in SCOPE["on"]:
            y = _EmuConv.apply(x, w, b, False, stride, padding, 0)
            return _StoreRound.apply(y) if SCOPE.get("store", True) else y
        return F.conv2d(x, w, b, stride=stride, padding=padding)

    @staticmethod
    def conv_transpose2d(x, w, b=None, stride=1, padding=0, output_padding=0):
        if SCOPE["cur"] in SCOPE["on"]:
            y = _EmuConv.apply(x, w, b, True, stride, padding, output_padding)
            return _StoreRound.apply(y) if SCOPE.get("store", True) else y
        return F.conv_transpose2d(x, w, b, stride=stride, padding=padding, output_padding=output_padding)


class _StoreRound(torch.autograd.Function):
    """activation stored rounded; straight-through gradient"""

    @staticmethod
    def forward(ctx, y):
        return _rnd(y, "o")

    @staticmethod
    def backward(ctx, g):
        return g


def _scoped(cls, names, scope):
    for nm in names:
        orig = getattr(cls, nm)

        def wrapper(self, *a, __orig=orig, **kw):
            prev = SCOPE["cur"]
            SCOPE["cur"] = scope
            try:
                return __orig(self, *a, **kw)
            finally:
                SCOPE["cur"] = prev
        setattr(cls, nm, wrapper)


def install():
    O.F = _FProxy()
    _scoped(O.Gen, ["forward", "forward_a2b", "forward_b2a", "decode"], "gen")
    _scoped(O.Dis, ["forward", "feats", "regress"], "dis")


def _copy_adam(src, dst):
    """teacher forcing includes the optimiser state (as tests/test_trainer_gpu.py::_load_adam does)"""
    for net, so, do in (("gen", src.gen_opt, dst.gen_opt), ("dis", src.dis_opt, dst.dis_opt)):
        for k, p in src.params[net].items():
            st = so.state.get(p)
            q = dst.params[net][k]
            if not st:
                do.state.pop(q, None)
                continue
            do.state[q] = {"step": st["step"].clone(), "exp_avg": st["exp_avg"].clone(),
                           "exp_avg_sq": st["exp_avg_sq"].clone()}


KEYS = ("dis_loss", "dis_ad_loss", "gen_total_loss", "gen_ad_loss", "gen_ll_loss", "gen_ll_loss2", "gen_enc_loss",
        "gen_enc_loss2")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_precision_ablation.json"))
    ap.add_argument("--variants", default="gen,dis,gen+dis")
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--what", default="xwgo", help="which conv tensors are rounded: x inputs, w weights, g gradients, o outputs")
    args = ap.parse_args()
    install()
    import yaml
    with open(os.path.join(ROOT, "exps", "nnyu.yaml")) as fh:
        hp = yaml.safe_load(fh)["train"]["hyperparameters"]
    torch.set_num_threads(os.cpu_count() or 1)
    SCOPE["mode"] = args.mode
    SCOPE["what"] = args.what
    master = O.OracleTrainer(hp, seed=0)
    emu = O.OracleTrainer(hp, seed=0)
    variants = [tuple(v.split("+")) for v in args.variants.split(",")]
    g = torch.Generator().manual_seed(1234)
    torch.manual_seed(42)
    curves = {"+".join(v): {k: [] for k in KEYS} for v in variants}
    ref = {k: [] for k in KEYS}
    for s in range(args.steps):
        ia, ib, la, lb = O.synthetic_batch(args.batch, 108, g, "uniform")
        for net in ("gen", "dis"):
            for k, v in master.params[net].items():
                emu.params[net][k].data.copy_(v.data)
        rng = torch.get_rng_state()
        for v in variants:
            # the emulated oracle must not keep its own update: weights are re-loaded per variant
            for net in ("gen", "dis"):
                for k, p in master.params[net].items():
                    emu.params[net][k].data.copy_(p.data)
            _copy_adam(master, emu)
            SCOPE["on"] = set(v)
            torch.set_rng_state(rng)
            emu.dis_update(ia, la, ib, lb, None, None, hp)
            # gen_update sees the master's post-dis_update discriminator in the real protocol too: the CUDA trainer
            # steps its own dis; mirror that (emu keeps its own dis step)
            emu.gen_update(ia, la, ib, lb, hp)
            for k in KEYS:
                curves["+".join(v)][k].append(float(getattr(emu, k)))
        SCOPE["on"] = set()
        torch.set_rng_state(rng)
        master.dis_update(ia, la, ib, lb, None, None, hp)
        master.gen_update(ia, la, ib, lb, hp)
        for k in KEYS:
            ref[k].append(float(getattr(master, k)))
        if s % 10 == 9 or s == args.steps - 1:
            msg = []
            for v in curves:
                w = {k: max(abs(a - b) / (abs(b) + 1e-12) for a, b in zip(curves[v][k], ref[k])) for k in
                     ("dis_ad_loss", "gen_ad_loss", "gen_ll_loss")}
                msg.append("%s: %s" % (v, {k: "%.2e" % x for k, x in w.items()}))
            print("step %d  max rel dev so far  %s" % (s, " | ".join(msg)), flush=True)
    out = {"protocol": "teacher-forced, batch %d, %d steps, uniform inputs, seeds 0/1234/42; %s rounding of conv "
                       "tensors [%s] (x inputs, w weights, g gradients, o outputs) in the named networks of the CPU oracle" %
                       (args.batch, args.steps, args.mode, args.what),
           "reference": ref, "emulated": curves, "max_rel_dev": {}, "per_step_rel_dev": {}}
    for v in curves:
        out["max_rel_dev"][v] = {k: max(abs(a - b) / (abs(b) + 1e-12) for a, b in zip(curves[v][k], ref[k])) for k in KEYS}
        out["per_step_rel_dev"][v] = {k: [abs(a - b) / (abs(b) + 1e-12) for a, b in zip(curves[v][k], ref[k])] for k in
                                      ("dis_ad_loss", "gen_ad_loss")}
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out["max_rel_dev"], indent=1))


if __name__ == "__main__":
    main()
