#!/usr/bin/env python
"""The "library line" (SURVEY.md section 7 step 0 / 8d): the UNMODIFIED reference LSPSTrainer on ONE B200 through stock
PyTorch / cuDNN -- the bar a hand-written path has to beat on the same box.

Runs the staged reference copy (oracle/_ref, built by oracle/build_ref.py; test infrastructure, not product code) with
`trainer.cuda(0)` and times the BASELINE config-2 step (dis_update + gen_update, batch 64 per domain) and the config-3
step (post_update mode 3, batch 256) under
    fp32 (TF32 off) | TF32 | bf16 autocast + channels_last       each with cudnn.benchmark = True
and with the latent noise drawn on the host as written (common_net.py:39) or on the device (`randn_like`, the only
patch; it is what lsps_b200's timed mode does too).  Prints one JSON object; nothing of lsps_b200 is imported.

  python tools/library_line.py --batch 64 --steps 5 > gpurun_out/library_line.json
"""
import argparse
import contextlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--est-batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--variants", default="fp32,tf32,bf16_cl")
    ap.add_argument("--host-noise", action="store_true", help="also time the as-written host-RNG noise (slow)")
    args = ap.parse_args()
    import torch
    import ref_loader
    trainers = ref_loader.load_reference(cpu_shim=False)
    from trainers import common_net
    hp = ref_loader.load_hyperparameters("nnyu")
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True

    orig_noise = common_net.GaussianNoiseLayer.forward

    def device_noise(self, x):
        if self.training is False:
            return x
        return x + torch.randn_like(x)

    g = torch.Generator().manual_seed(1234)

    def batch(b):
        ia = (torch.rand(b, 1, 128, 128, generator=g) * 2 - 1).to(dev)
        ib = (torch.rand(b, 1, 128, 128, generator=g) * 2 - 1).to(dev)
        la, lb = (torch.randn(b, 108, generator=g) * 0.3).to(dev), (torch.randn(b, 108, generator=g) * 0.3).to(dev)
        return ia, la, ib, lb, torch.zeros(b, 3, device=dev), torch.zeros(b, 3, device=dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps

    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "what": "unmodified reference LSPSTrainer (oracle/_ref) on one GPU through torch/cuDNN", "runs": []}
    for variant in args.variants.split(","):
        for noise in (["device"] + (["host"] if args.host_noise else [])):
            common_net.GaussianNoiseLayer.forward = device_noise if noise == "device" else orig_noise
            tf32 = variant != "fp32"
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            torch.manual_seed(0)
            tr = trainers.LSPSTrainer(hp)
            tr.cuda(0)
            cl = variant == "bf16_cl"
            if cl:
                for net in (tr.gen, tr.dis, tr.map):
                    net.to(memory_format=torch.channels_last)
            ctx = contextlib.nullcontext
            if cl:
                # autocast only around the network calls (F.binary_cross_entropy refuses to run inside an autocast
                # region); network outputs are handed back to the trainer's loss code as fp32
                def _f32(o):
                    if torch.is_tensor(o):
                        return o.float() if o.is_floating_point() else o
                    if isinstance(o, (tuple, list)):
                        return type(o)(_f32(x) for x in o)
                    return o

                def _wrap(obj, name):
                    fn = getattr(obj, name)

                    def call(*a, **k):
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            return _f32(fn(*a, **k))
                    setattr(obj, name, call)
                for name in ("forward", "forward_a2b", "forward_b2a", "decode"):
                    _wrap(tr.gen, name)
                for name in ("forward", "feats", "regress_a", "regress_b"):
                    _wrap(tr.dis, name)
            ia, la, ib, lb, ca, cb = batch(args.batch)
            if cl:
                ia, ib = ia.contiguous(memory_format=torch.channels_last), ib.contiguous(memory_format=torch.channels_last)

            def pretrain():
                with ctx():
                    tr.dis_update(ia, la, ib, lb, ca, cb, hp)
                    tr.gen_update(ia, la, ib, lb, hp)
            rec = {"variant": variant, "noise": noise}
            try:
                ms, wall = timed(pretrain, args.steps, args.warmup)
                rec.update(pretrain_batch=args.batch, pretrain_ms_per_step=ms, pretrain_wall_ms=wall,
                           pretrain_images_per_s=2 * args.batch / (ms / 1e3),
                           losses={k: float(getattr(tr, k)) for k in ("dis_loss", "gen_total_loss", "gen_ad_loss")})
            except Exception as e:  # noqa
                rec["pretrain_error"] = repr(e)[:300]
            ea, ela, eb, elb, eca, ecb = batch(args.est_batch)

            def estimate3():
                with ctx():
                    tr.post_update(ea, ela, eb, elb, eca, ecb, 3, hp)
            try:
                ms, wall = timed(estimate3, args.steps, args.warmup)
                rec.update(estimate3_batch=args.est_batch, estimate3_ms_per_step=ms, estimate3_wall_ms=wall,
                           estimate3_images_per_s=2 * args.est_batch / (ms / 1e3))
            except Exception as e:  # noqa
                rec["estimate3_error"] = repr(e)[:300]
            rec["max_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
            out["runs"].append(rec)
            sys.stderr.write(json.dumps(rec) + "\n")
            del tr
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
