#!/usr/bin/env python3
"""Config 5 of BASELINE.json: mixed-curve sweep over batch sizes 2^10 .. 2^24 (device-resident timing).

    python tools/sweep.py [--max-log2 24] > profiles/rNN_sweep.jsonl

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

    ctx = Context()
    cand = [ctx.imad_probe(2, 2048)[0] / 1e12, ctx.imad_probe(0, 2048)[0] / 2e12, ctx.imad_probe(3, 2048)[0] / 2e12]
    peak = max(cand)
    for name in args.workloads.split(","):
        for lg in range(args.min_log2, args.max_log2 + 1, 2):
            n = 1 << lg
            steps = 3 if lg >= 20 else 10
            bench.WORKLOADS[name] = (lg,) + bench.WORKLOADS[name][1:]
            try:
                r = bench.measure_device(torch, ctx, name, n, steps, 3, 0xECC00005, None, nbuf=2 if lg >= 22 else None)
                v = n / (r["ms_per_step"] * 1e-3)
                print(json.dumps({"workload": name, "log2_n": lg, "value": v, "unit": "scalar-mults/s", "ms_per_batch": r["ms_per_step"],
                                  "roofline_frac": n * bench.work_of(name) / (r["ms_per_step"] * 1e-3) / 1e12 / peak,
                                  "kernels_ms": {"scalar_mult": r["main_ms"], "batch_inversion_encode": r["fin_ms"]}}), flush=True)
                del r
                torch.cuda.empty_cache()
            except Exception as e:
                print(json.dumps({"workload": name, "log2_n": lg, "error": repr(e)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
