#!/usr/bin/env python
"""BASELINE.json configs 3 and 4 on their GPU counts (secondary to bench.py's contract line).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/bench_configs.py --yaml nnyu --mode estimate3 --batch 32            # weak: 32 per domain per rank
  ... --global-batch 256                                                        # strong: 256 per domain over all ranks

One step = post_update(mode 3) (estimate3) or dis_update + gen_update (pretrain) on this rank's shard; the only exchange
is the sum-allreduce of the flat gradient buffer.  Timed with CUDA events after warm-up, barrier + synchronize on both
sides, MAX over ranks; the allreduce's share is timed with its own events on the same stream.  One JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--yaml", default="nnyu")
    ap.add_argument("--mode", default="estimate3")
    ap.add_argument("--batch", type=int, default=0, help="per domain per rank (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0, help="per domain over all ranks (strong scaling)")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--graphs", type=int, default=0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    saved = os.dup(1)
    os.dup2(2, 1)                                   # NCCL banners go to stderr; stdout carries the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lsps_b200
    hp = lsps_b200.load_hyperparameters(args.yaml)
    B = args.batch or max(1, args.global_batch // world)
    tr = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="device", graphs=bool(args.graphs))
    g = torch.Generator().manual_seed(1234 + rank)
    ia, ib, la, lb = (t.cuda() for t in lsps_b200.synthetic_batch(B, hp["vae"]["input_dim"], g, "hand"))

    ar_events = []
    orig_ar = tr._allreduce

    def timed_allreduce(store):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_ar(store)
        e1.record()
        ar_events.append((e0, e1))
    tr._allreduce = timed_allreduce

    if args.mode == "pretrain":
        def step():
            tr.dis_update(ia, la, ib, lb, None, None, hp)
            tr.gen_update(ia, la, ib, lb, hp)
    else:
        mode = int(args.mode[len("estimate"):])

        def step():
            tr.post_update(ia, la, ib, lb, None, None, mode, hp)
    for _ in range(args.warmup):
        step()
    ar_events.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = tr.ops.ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / args.steps
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps, sum(a.elapsed_time(b) for a, b in ar_events) / args.steps, wall],
                      device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank == 0:
        os.dup2(saved, 1)
        print(json.dumps({"yaml": args.yaml, "mode": args.mode, "n_gpus": world, "batch_per_domain_per_rank": B,
                          "global_batch_per_domain": B * world, "scaling": "strong" if args.global_batch else "weak",
                          "ms_per_step": ms[0].item(), "allreduce_ms_per_step": ms[1].item(), "wall_ms_per_step": ms[2].item(),
                          "images_per_s": 2 * B * world / (ms[0].item() / 1e3), "graphs": bool(args.graphs),
                          "precision": tr.precision, "launches_per_step": (tr.ops.ctx.launch_count() - l0) // args.steps}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
