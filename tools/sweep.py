#!/usr/bin/env python3
"""Config 5 of BASELINE.json: mixed-curve sweep over batch sizes 2^10 .. 2^24 (device-resident timing).

    python tools/sweep.py [--max-log2 24] > profiles/rNN_sweep.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sweep.py ...
        (N GPUs: every batch is sliced over the ranks — the sweep "across 1/2/4/8 B200" of config 5)

One JSON line per (workload, batch size): scalar-mults/s, ms per batch, fraction of the IMAD roofline.
Same measurement as bench.py's `value` (CUDA events on the launching stream, inputs resident in HBM,
3 warm-up batches); small batches are latency-bound (one thread per scalar multiplication).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=24)
    ap.add_argument("--workloads", default="ed25519_mul_base,p256_mul,p384_mul,x448")
    args = ap.parse_args()
    import torch

    from eccoxide_b200 import Context

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:   # torchrun: the batch of every line is the WHOLE job, each rank takes its contiguous 1/world slice
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(devices=[local])
    cand = [ctx.imad_probe(2, 2048)[0] / 1e12, ctx.imad_probe(0, 2048)[0] / 2e12, ctx.imad_probe(3, 2048)[0] / 2e12]
    peak = max(cand)
    for name in args.workloads.split(","):
        for lg in range(args.min_log2, args.max_log2 + 1, 2):
            n = (1 << lg) // world
            bench.WORKLOADS[name] = (lg,) + bench.WORKLOADS[name][1:]
            try:
                r = bench.measure_device(torch, ctx, name, n, 5, 3, 0xECC00005 + rank, dist, min_s=0.25)
                v = world * n / (r["ms_per_batch"] * 1e-3)
                if rank == 0:
                    print(json.dumps({"workload": name, "log2_n": lg, "n_gpus": world, "batch_per_gpu": n, "value": v, "unit": "scalar-mults/s",
                                      "ms_per_batch": r["ms_per_batch"], "passes_timed": 5 * r["inner"],
                                      "roofline_frac": world * n * bench.work_of(name) / (r["ms_per_batch"] * 1e-3) / 1e12 / (peak * world),
                                      "kernels_ms": {"scalar_mult": r["main_ms"], "batch_inversion_encode": r["fin_ms"]}}), flush=True)
                del r
                torch.cuda.empty_cache()
            except Exception as e:
                if rank == 0:
                    print(json.dumps({"workload": name, "log2_n": lg, "n_gpus": world, "error": repr(e)}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
