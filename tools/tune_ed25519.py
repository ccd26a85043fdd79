#!/usr/bin/env python3
"""Launch-shape sweep of Ed25519 mul_base on device-resident inputs (one GPU).

    python tools/tune_ed25519.py [--w 24,26] [--stride 24,32] [--logs 10,12,14,16,17,18,20] > profiles/rNN_tune_ed25519.jsonl

For every (comb width, entry stride, batch size) it times the two-kernel large-batch form
(ed25519_fused = 0), the fused small-batch kernel with every lane count that fits one wave, and the
fused body on the large-batch launch shape (ed25519_fused = 2): CUDA events on the launching stream,
inputs rotating over distinct buffers, >= 0.25 s or 400 steps per point, first 1024 outputs of every
variant compared with the oracle.  One JSON line per point.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", default="24,26")
    ap.add_argument("--stride", default="24")
    ap.add_argument("--logs", default="10,12,14,15,16,17,18,20")
    ap.add_argument("--min-s", type=float, default=0.25)
    args = ap.parse_args()
    import torch

    from eccoxide_b200 import Context
    from oracle import coracle as C

    C.build()
    C.load()
    g = np.random.Generator(np.random.Philox(0xECC00001))
    nmax = 1 << max(int(x) for x in args.logs.split(","))
    wide = g.integers(0, 256, size=(nmax, 32), dtype=np.uint8)
    wide[:, 31] &= 0x0F
    exp = C.ed25519_mul_base(wide[:1024], os.cpu_count() or 1)
    stream = torch.cuda.current_stream().cuda_stream
    for w in [int(x) for x in args.w.split(",")]:
        for stride in [int(x) for x in args.stride.split(",")]:
            ctx = Context()
            ctx.set_option("ed25519_comb_w", w)
            ctx.set_option("ed25519_entry_stride", stride)
            for lg in [int(x) for x in args.logs.split(",")]:
                n = 1 << lg
                nbuf = max(2, min(8, (160 << 20) // (n * 32)))
                bufs = [torch.from_numpy(np.roll(wide[:n], b * 977, axis=0).copy()).cuda() for b in range(nbuf)]
                out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
                variants = [("split", 0, 0)]
                for lanes in (1, 2, 4, 8):
                    if n * lanes <= 148 * 480:
                        variants.append(("fused_l%d" % lanes, 1, lanes))
                if n > 148 * 480:
                    variants.append(("fused_128x4", 2, 0))
                for label, fused, lanes in variants:
                    ctx.set_option("ed25519_fused", fused)
                    ctx.set_option("ed25519_lanes", lanes)
                    try:
                        ctx.warm("ed25519_mul_base", n)
                        for i in range(5):
                            ctx.dev_call("ecb_ed25519_mul_base_dev", 0, bufs[0].data_ptr(), n, out.data_ptr(), stream)
                        torch.cuda.synchronize()
                        ok = bool(np.array_equal(out[:1024].cpu().numpy(), exp[: min(n, 1024)]))
                        steps, total_ms, total_steps = 20, 0.0, 0
                        while total_ms < args.min_s * 1e3 and total_steps < 4000:
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                            for i in range(steps):
                                ctx.dev_call("ecb_ed25519_mul_base_dev", 0, bufs[i % nbuf].data_ptr(), n, out.data_ptr(), stream)
                            e1.record()
                            torch.cuda.synchronize()
                            total_ms += e0.elapsed_time(e1)
                            total_steps += steps
                            steps = min(steps * 2, 1000)
                        us = total_ms / total_steps * 1e3
                        print(json.dumps({"w": w, "stride": stride, "log2_n": lg, "variant": label, "us_per_batch": round(us, 2),
                                          "mops": round(n / us, 1), "parity": ok, "steps": total_steps}), flush=True)
                    except Exception as e:
                        print(json.dumps({"w": w, "stride": stride, "log2_n": lg, "variant": label, "error": repr(e)}), flush=True)
                del bufs, out
                torch.cuda.empty_cache()
            ctx.close()


if __name__ == "__main__":
    main()
