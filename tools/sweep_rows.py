#!/usr/bin/env python
"""Fused energy-score fwd+bwd (m=8, D=3072, fp32/bf16) as a function of the number of rows per launch: where a single
launch stops being latency-bound.  One JSON line per point."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddm_b200 import _cabi
from tools.sweep_energy import time_config

L = _cabi.lib()
for dtype in ("f32", "bf16"):
    for B in (32, 64, 128, 148, 256, 296, 512, 1024, 2048, 4096):
        esz = 4 if dtype == "f32" else 2
        nsets = max(3, min(40, (300 * 2**20) // ((2 * B * 8 * 3072 + B * 3072) * esz) + 1))
        us1, gb1 = time_config(L, B, 8, 3072, dtype, nstreams=1, nsets_override=nsets, iters=600)
        us4, gb4 = time_config(L, B, 8, 3072, dtype, nstreams=4, nsets_override=nsets, iters=600)
        print(json.dumps({"dtype": dtype, "B": B, "us_1stream": round(us1, 2), "GBps_1stream": round(gb1, 1),
                          "frac_1stream": round(gb1 / 6452.5, 3), "us_4streams": round(us4, 2), "GBps_4streams": round(gb4, 1),
                          "frac_4streams": round(gb4 / 6452.5, 3), "rows_per_s_4streams": round(B / us4 * 1e6)}), flush=True)
