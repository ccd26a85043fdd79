#!/usr/bin/env python3
"""One ecb_wei_msm call per (curve, log2 n) given on the command line — for launch lists under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/msm_once.py bls12_381_g1 20"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eccoxide_b200 import Context

ORD = {"bls12_381_g1": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
       "p256k1": 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141}
curve, lg = sys.argv[1], int(sys.argv[2])
n, period = 1 << lg, 1 << 10
g = np.random.Generator(np.random.Philox(0x3536))
with Context(devices=[0]) as c:
    tb = np.frombuffer(b"".join((int.from_bytes(g.bytes(48), "big") % ORD[curve]).to_bytes(32, "big") for _ in range(period)), dtype=np.uint8).reshape(period, 32)
    base, _ = c.wei_mul_base(curve, tb)
    wide = g.integers(0, 256, size=(min(n, 1 << 14), 64), dtype=np.uint8)
    k = np.frombuffer(b"".join((int.from_bytes(r.tobytes(), "big") % ORD[curve]).to_bytes(32, "big") for r in wide), dtype=np.uint8).reshape(-1, 32)
    k = np.ascontiguousarray(np.tile(k, (n // k.shape[0], 1)))
    pts = np.ascontiguousarray(np.tile(base, (n // period, 1)))
    for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
        out, inf = c.wei_msm(curve, k, pts)
    print(out.tobytes().hex()[:32], bool(inf))
