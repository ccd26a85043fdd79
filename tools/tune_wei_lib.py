#!/usr/bin/env python3
"""A/B timing of builds of libeccbatch.so on the Weierstrass variable-base kernel (one GPU).

    python tools/tune_wei_lib.py --libs eccoxide_b200/libeccbatch.so,variants/libeccbatch_x.so [--curve p256r1] [--log 20]

Every library runs in its own process (the loader binds one .so per process): device-resident inputs, CUDA events on
the launching stream around `steps` calls of ecb_wei_mul_dev, 4096 outputs compared with the C oracle and a SHA-256 of
the whole output so that the variants can be compared with one another.  One JSON line per library.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CURVES = {"p256r1": (0, 32, 32, 1), "p384r1": (1, 48, 48, 1), "bls12_381_g1": (2, 48, 32, 2), "p256k1": (3, 32, 32, 1)}


def one(lib, curve, lg, steps):
    import torch

    from eccoxide_b200 import _lib

    _lib.LIB_PATH = os.path.abspath(lib)
    from eccoxide_b200 import Context
    from oracle import coracle as C
    from bench import rand_scalars

    C.build()
    C.load()
    cid, fb, sb, clr = CURVES[curve]
    n = 1 << lg
    g = np.random.Generator(np.random.Philox(0xECC0AB))
    ctx = Context(devices=[0])
    uniq = min(n, 1 << 14)
    pts, inf = ctx.wei_mul_base(curve, rand_scalars(g, uniq, sb, clr, "big"))
    pts = np.ascontiguousarray(np.tile(pts, (n // uniq, 1)))
    ks = [rand_scalars(g, n, sb, clr, "big") for _ in range(2)]
    dk = [torch.from_numpy(k).cuda() for k in ks]
    dp = torch.from_numpy(pts).cuda()
    out = torch.empty((n, 2 * fb), dtype=torch.uint8, device="cuda")
    oinf = torch.empty((n,), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ctx.warm("wei_mul", n, curve=cid)
    for i in range(2):
        ctx.dev_call("ecb_wei_mul_dev", 0, cid, dk[0].data_ptr(), dp.data_ptr(), n, out.data_ptr(), oinf.data_ptr(), stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    m = min(n, 4096)
    exp, einf = C.wei_mul(curve, ks[0][:m], pts[:m], nthreads=os.cpu_count() or 1)
    ok = bool(np.array_equal(got[:m], exp) and not oinf[:m].cpu().numpy().any() and not einf.any())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ctx.dev_call("ecb_wei_mul_dev", 0, cid, dk[i & 1].data_ptr(), dp.data_ptr(), n, out.data_ptr(), oinf.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"lib": os.path.relpath(lib, ROOT), "curve": curve, "log2_n": lg, "ms_per_batch": round(ms, 3), "mops": round(n / ms / 1e3, 2),
                      "parity_4096": ok, "sha256_out": hashlib.sha256(got.tobytes()).hexdigest()[:16], "steps": steps}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default="eccoxide_b200/libeccbatch.so")
    ap.add_argument("--curve", default="p256r1")
    ap.add_argument("--log", type=int, default=20)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--one")
    a = ap.parse_args()
    if a.one:
        return one(a.one, a.curve, a.log, a.steps)
    for lib in a.libs.split(","):
        subprocess.call([sys.executable, os.path.abspath(__file__), "--one", lib, "--curve", a.curve, "--log", str(a.log), "--steps", str(a.steps)])


if __name__ == "__main__":
    main()
