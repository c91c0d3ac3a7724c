#!/usr/bin/env python
"""bench.py -- train-step depth-images/sec of the LSPS pretrain step (dis_update + gen_update) on B200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                   (the reference's own CPU path: oracle/_ref, else the oracle port)

Workload (BASELINE.json configs[1]): depth_train.py --mode pretrain, exps/nnyu.yaml, synthetic 128x128 depth
crops, batch 64 per domain per GPU (weak scaling: the per-GPU batch is fixed as N grows).  One step = one
dis_update + one gen_update (src/depth_train.py:158-161) = 2*64 depth images per GPU.

Prints ONE JSON line.  `value`: inputs already resident in HBM.  `e2e`: same calls with pinned HOST inputs, the
host->device copies and the per-update device->host loss read inside the timed region.  `roofline`: the dominant
kernel (3x3 s1 256->256 implicit GEMM on tcgen05, the K1 shape of SURVEY.md 2.2) timed per launch with CUDA events
in an extra instrumented step.  `cpu_baseline`: the unmodified reference (oracle/_ref; else the oracle port) timed on this box's host cores on a
bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train-step depth-images/sec at 128x128 (pretrain: dis_update+gen_update, nnyu)"
UNIT = "images/s"
BATCH = int(os.environ.get("LSPS_BENCH_BATCH", "64"))          # per domain per GPU
CPU_SAMPLE_BATCH = int(os.environ.get("LSPS_BENCH_CPU_BATCH", "2"))
# algorithmic work, SURVEY.md section 8d: 195.4 GMAC per (a,b) image pair per pretrain step
GFLOP_PER_PAIR = 390.7


WORKLOAD = "depth_train.py --mode pretrain, exps/nnyu.yaml, 128x128 synthetic depth, batch %d per domain per GPU " \
           "(dis_update + gen_update)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))
                if self.stop_flag:
                    break
        except Exception:  # noqa
            pass

    def finish(self, t0=None, t1=None):
        """Samples taken inside the timed region [t0, t1]; the sampler runs from before the warm-up steps (nvidia-smi needs
        a few hundred ms to deliver its first line), so when a short timed region caught fewer than two samples the ones
        of the warm-up steps -- the same step, the same load -- are used as well and `window` says so."""
        self.stop_flag = True
        time.sleep(0.15)
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 is None or t0 <= t <= t1 + 0.05]
        window = "timed"
        if len(rows) < 2:
            rows, window = [r for _, r in self.rows], "warmup+timed"
        sm, reasons, smax = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[0])); smax = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def cpu_reference_rate(steps, warmup, batch):
    """The reference's own CPU path of this workload on all host cores: the UNMODIFIED reference LSPSTrainer from the
    staged copy oracle/_ref (kind "reference") when it travelled to this box, else the oracle port (kind "port")."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "port"
    try:
        import ref_loader
        if ref_loader.reference_available() or ref_loader.staged_available():
            trainers = ref_loader.load_reference(cpu_shim="force")
            hp = ref_loader.load_hyperparameters("nnyu")
            torch.manual_seed(0)
            tr = trainers.LSPSTrainer(hp)
            tr.gpu = 0
            kind = "reference"
    except Exception as e:  # noqa
        sys.stderr.write("bench: staged reference not usable (%r); timing the oracle port\n" % (e,))
        kind = "port"
    if kind == "port":
        import lsps_oracle as O
        import yaml
        with open(os.path.join(ROOT, "exps", "nnyu.yaml")) as fh:
            hp = yaml.safe_load(fh)["train"]["hyperparameters"]
        tr = O.OracleTrainer(hp, seed=0)
    # synthetic crops of the same kind as the B200 arm's (background +1, hand pixels in [-1,1])
    g = torch.Generator().manual_seed(1234)
    ia = (torch.randn(batch, 1, 128, 128, generator=g) * 0.35).clamp(-1, 1)
    ib = (torch.randn(batch, 1, 128, 128, generator=g) * 0.35).clamp(-1, 1)
    yy, xx = torch.meshgrid(torch.arange(128.0), torch.arange(128.0), indexing="ij")
    bg = (((yy - 63.5) / 40.0) ** 2 + ((xx - 63.5) / 40.0) ** 2 > 1.0)
    ia[:, 0][:, bg] = 1.0
    ib[:, 0][:, bg] = 1.0
    la, lb = torch.randn(batch, 108, generator=g) * 0.3, torch.randn(batch, 108, generator=g) * 0.3
    com = torch.zeros(batch, 3)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        tr.dis_update(ia, la, ib, lb, com, com, hp)
        tr.gen_update(ia, la, ib, lb, hp)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return 2 * batch / sec, sec, cores, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    rate, sec, cores, kind = cpu_reference_rate(steps, warmup, CPU_SAMPLE_BATCH)
    sample = "pretrain step (dis_update+gen_update) at batch %d per domain, %d timed steps, %s on torch CPU fp32" % (
        CPU_SAMPLE_BATCH, steps, "the unmodified reference LSPSTrainer (oracle/_ref)" if kind == "reference"
        else "oracle port of LSPSTrainer")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the workload is the B200 arm's; each CPU step is a bounded sample of it (see cpu_baseline.sample)
        "config": {"workload": WORKLOAD % BATCH, "sample_batch_per_domain": CPU_SAMPLE_BATCH,
                   "noise": "host RNG (the reference's own draws)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # keep stdout clean for the ONE JSON line: NCCL / torch.distributed print banners ("NCCL version ...") to fd 1
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lsps_b200
    from lsps_b200 import engine as _engine
    hp = lsps_b200.load_hyperparameters("nnyu")
    tr = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="device")
    B = BATCH
    g = torch.Generator().manual_seed(1234 + rank)
    ia_h, ib_h, la_h, lb_h = (t.pin_memory() for t in lsps_b200.synthetic_batch(B, 108, g, "hand"))
    ia, ib, la, lb = (t.cuda(non_blocking=True) for t in (ia_h, ib_h, la_h, lb_h))

    def step_dev():
        tr.dis_update(ia, la, ib, lb, None, None, hp)
        tr.gen_update(ia, la, ib, lb, hp)

    def step_e2e():
        a, b_, c, d = (t.cuda(non_blocking=True) for t in (ia_h, ib_h, la_h, lb_h))
        tr.dis_update(a, c, b_, d, None, None, hp)
        tr.gen_update(a, c, b_, d, hp)

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return ms.item()

    W, K = max(3, args.warmup), args.steps
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(W):
        step_dev()
    l0 = tr.ops.ctx.launch_count()
    t_begin = time.monotonic()
    ms = timed(step_dev, K)
    t_end = time.monotonic()
    launches = tr.ops.ctx.launch_count() - l0
    clocks = sampler.finish(t_begin, t_end) if sampler else None
    value = world * 2 * B * K / (ms / 1e3)
    light = os.environ.get("LSPS_BENCH_LIGHT") == "1"     # profiler runs: timed loop only
    if light:
        if rank == 0:
            emit({"light": True, "ms_per_step": ms / K, "gpu_launches": launches, "value": value})
        return
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, K)
    e2e = world * 2 * B * K / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in (ia_h, ib_h, la_h, lb_h))
    d2h = (8 + 16) * 4

    # ---- roofline of the dominant kernel: every launch of the K1 shape in one instrumented step
    # (every rank runs the step -- it contains the gradient allreduce -- but only rank 0 instruments and reports)
    roof = None
    ev = []
    orig_f, orig_d = _engine.Ops.conv_fwd, _engine.Ops.conv_dgrad
    if rank == 0:
        def wrap(orig, is_fwd):
            def f(self, S, key, kind, x, *a, **kw):
                shp = x.shape if is_fwd else a[0]
                k1 = kind == 0 and shp[1] == 32 and shp[3] == 256
                if k1:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                out = orig(self, S, key, kind, x, *a, **kw)
                if k1:
                    e1.record()
                    masked = (not is_fwd) and (kw.get("mask") is not None or kw.get("add") is not None)
                    ev.append((e0, e1, shp[0], ("fwd" if is_fwd else ("dgrad+ep" if masked else "dgrad")) + "_n%d" % shp[0]))
                return out
            return f
        _engine.Ops.conv_fwd, _engine.Ops.conv_dgrad = wrap(orig_f, True), wrap(orig_d, False)
    # per-launch durations are only meaningful without the concurrent side-stream wgrad kernels: this one extra step
    # runs them inline (the timed region above uses the side stream)
    tr.ops.use_side = False
    for _ in range(2):      # the first pass creates ~300 CUDA events (slow: the host falls behind the GPU and the
        ev.clear()          # launch gap would be billed to the kernel); the second pass is the measurement
        step_dev()
        torch.cuda.synchronize()
    tr.ops.use_side = os.environ.get("LSPS_NO_SIDE", "0") != "1"
    _engine.Ops.conv_fwd, _engine.Ops.conv_dgrad = orig_f, orig_d
    if rank == 0:
        pk = _peaks()
        tot_ms = sum(a.elapsed_time(b_) for a, b_, _, _ in ev)
        tot_flop = sum(2.0 * n * 1024 * 256 * 2304 for _, _, n, _ in ev)
        classes = {}
        for a, b_, n, cls in ev:
            c = classes.setdefault(cls, {"launches": 0, "ms": 0.0, "flop": 0.0})
            c["launches"] += 1
            c["ms"] += a.elapsed_time(b_)
            c["flop"] += 2.0 * n * 1024 * 256 * 2304
        classes = {k: {"launches": c["launches"], "avg_ms": round(c["ms"] / c["launches"], 4),
                       "tflops": round(c["flop"] / (c["ms"] * 1e-3) / 1e12, 1)} for k, c in sorted(classes.items())}
        ach = tot_flop / (tot_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel<256> (3x3 s1 256->256 @32x32, fwd+dgrad)",
                "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                "traffic": 87.6e6, "traffic_note": "dram read+write bytes per launch at N=128 images, forward with the "
                "statistics epilogue, profiles/r02_ncu_targets_v2.md row 19 (68.6 MB read + 19.0 MB written; algorithmic: "
                "64 MB in + 64 MB out + 1.2 MB weights; most of the output is still in L2 at kernel end)",
                "launches": len(ev), "avg_launch_ms": tot_ms / max(1, len(ev)),
                "share_of_step": tot_ms / (ms / K), "peak_source": pk["src"] + ", sustained bf16",
                "peak_burst": pk["tf"], "frac_of_burst_peak": ach / pk["tf"],
                "regime_note": "this bench runs for ~1 s, between burst and power-limited steady state; over a 4 s "
                "loop the same kernel holds 0.88 of cuBLAS' sustained rate (profiles/r01_probe_sustained.log)",
                "note": "launch durations from one extra step with the wgrad side stream disabled (no concurrent kernels)",
                "by_class": classes}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores, kind = cpu_reference_rate(4, 1, CPU_SAMPLE_BATCH)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "pretrain step at batch %d per domain, 1 warm-up + 4 timed steps (%.1f s/step), %s on torch "
                         "CPU fp32" % (CPU_SAMPLE_BATCH, sec, "the unmodified reference LSPSTrainer (oracle/_ref)"
                                       if kind == "reference" else "oracle port of the reference LSPSTrainer")}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD % B,
                   "global_batch_per_domain": B * world, "parallelism": "dp%d" % world,
                   "noise": "device Philox (host-RNG parity mode is not the timed mode)",
                   "l2": "working set >> 126 MB L2 (activations of one step are several GB); no explicit flush",
                   "algorithmic_tflop_per_step_per_gpu": GFLOP_PER_PAIR * B / 1e3},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "step_tflops": GFLOP_PER_PAIR * B / 1e3 / (ms / K / 1e3) * world,
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
