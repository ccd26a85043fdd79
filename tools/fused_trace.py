#!/usr/bin/env python3
"""Phase times of the fused small-batch Ed25519 kernel (option "trace"): per block start / comb done / inversion
done / end on %globaltimer; prints medians and the critical path per batch size."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eccoxide_b200 import Context

g = np.random.Generator(np.random.Philox(7))
with Context() as c:
    c.set_option("trace", 1)
    stream = torch.cuda.current_stream().cuda_stream
    for lg, lanes in ((10, 8), (12, 4), (14, 1), (15, 1), (16, 1)):
        n = 1 << lg
        k = g.integers(0, 256, size=(n, 32), dtype=np.uint8); k[:, 31] &= 0x0F
        dk = torch.from_numpy(k).cuda(); out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
        c.set_option("ed25519_lanes", lanes)
        c.warm("ed25519_mul_base", n)
        for _ in range(20):
            c.dev_call("ecb_ed25519_mul_base_dev", 0, dk.data_ptr(), n, out.data_ptr(), stream)
        torch.cuda.synchronize()
        t = c.debug_fused_trace().astype(np.int64)
        t0 = t[:, 0].min()
        rel = (t - t0) / 1e3
        print(json.dumps({"log2_n": lg, "lanes": lanes, "blocks": int(t.shape[0]),
                          "start_spread_us": round(float(rel[:, 0].max()), 2),
                          "comb_us_median": round(float(np.median(rel[:, 1] - rel[:, 0])), 2), "comb_us_max": round(float((rel[:, 1] - rel[:, 0]).max()), 2),
                          "invert_us_median": round(float(np.median(rel[:, 2] - rel[:, 1])), 2), "invert_us_max": round(float((rel[:, 2] - rel[:, 1]).max()), 2),
                          "finish_us_median": round(float(np.median(rel[:, 3] - rel[:, 2])), 2),
                          "kernel_us": round(float(rel[:, 3].max()), 2)}), flush=True)
