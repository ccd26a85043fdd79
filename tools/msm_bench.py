#!/usr/bin/env python3
"""Throughput of ecb_wei_msm (bucket method, csrc/msm.cuh) through the host entry point (pinned host buffers; `kernels_ms` = the kernels alone), beside n independent scalar
multiplications (ecb_wei_mul) of the same batch.  One JSON line per (curve, n); every result is checked exactly:
P_i = t_i G, so the sum must equal (sum k_i t_i) G from the generator comb."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eccoxide_b200 import Context

ORD = {"bls12_381_g1": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
       "p256k1": 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141}
g = np.random.Generator(np.random.Philox(0x3535))
with Context() as c:
    for curve in ("bls12_381_g1", "p256k1"):
        n_ord, period = ORD[curve], 1 << 10
        tb = np.frombuffer(b"".join((int.from_bytes(g.bytes(48), "big") % n_ord).to_bytes(32, "big") for _ in range(period)), dtype=np.uint8).reshape(period, 32)
        base, _ = c.wei_mul_base(curve, tb)
        tv = [int.from_bytes(r.tobytes(), "big") for r in tb]
        for lg in (12, 16, 18, 20, 22):
            n = 1 << lg
            # scalars = 64 uniform bytes mod the group order: on p256k1 the top window then holds ONE bit
            wide = g.integers(0, 256, size=(min(n, 1 << 16), 64), dtype=np.uint8)
            k = np.frombuffer(b"".join((int.from_bytes(r.tobytes(), "big") % n_ord).to_bytes(32, "big") for r in wide), dtype=np.uint8).reshape(-1, 32)
            k = np.ascontiguousarray(np.tile(k, (n // k.shape[0], 1)))
            pts = np.ascontiguousarray(np.tile(base, (n // period, 1)))
            k, pts = (torch.from_numpy(a).pin_memory().numpy() for a in (k, pts))   # pinned host buffers, as in bench.py's e2e leg
            out, inf = c.wei_msm(curve, k, pts)
            reps = 3 if lg >= 20 else 10
            t0 = time.perf_counter()
            for _ in range(reps):
                out, inf = c.wei_msm(curve, k, pts)
            dt = (time.perf_counter() - t0) / reps
            c.set_option("profile", 1)                                              # device time of the kernels alone (CUDA events inside the library)
            c.wei_msm(curve, k, pts)
            dev_ms, _, calls = c.profile_collect(0)
            c.set_option("profile", 0)
            ksum = np.zeros(period, dtype=object)
            kv = k.reshape(n // period, period, 32)
            total = 0
            for j in range(period):
                col = kv[:, j, :]
                s = 0
                for b in range(32):
                    s = (s << 8) + int(col[:, b].astype(np.int64).sum())
                total += s * tv[j]
            want, _ = c.wei_mul_base(curve, np.frombuffer((total % n_ord).to_bytes(32, "big"), dtype=np.uint8).reshape(1, 32))
            c.wei_mul(curve, k, pts)          # first call sizes the work buffers
            t1 = time.perf_counter()
            c.wei_mul(curve, k, pts)
            dm = time.perf_counter() - t1
            # worst case for a bucket method: ONE scalar for every point (each window has a single bucket holding all of them)
            kone = np.ascontiguousarray(np.tile(k[:1], (n, 1)))
            c.wei_msm(curve, kone, pts)
            t2 = time.perf_counter()
            c.wei_msm(curve, kone, pts)
            dw = time.perf_counter() - t2
            print(json.dumps({"curve": curve, "log2_n": lg, "msm_ms": round(dt * 1e3, 3), "points_per_s": round(n / dt), "kernels_ms": round(dev_ms / max(calls, 1), 3), "exact": bool(out.tobytes() == want[0].tobytes() and not inf),
                              "independent_mul_ms": round(dm * 1e3, 2), "speedup_vs_independent_mul": round(dm / dt, 1), "all_points_one_scalar_ms": round(dw * 1e3, 3),
                              "scalars": "64 uniform bytes mod the order"}), flush=True)
