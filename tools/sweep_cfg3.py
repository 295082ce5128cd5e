#!/usr/bin/env python
"""BASELINE config 3: fused energy-score fwd+bwd over m in {4,8,16,32} x beta in {0.1,1,2} x D in {2,3072,12288}
(B=128), fp32 and bf16.  Prints one JSON line per point: us/launch (4 streams in a CUDA graph, and one stream),
algorithmic GB/s, fp32 GFLOP/s of the direct-difference form, and which kernel variant ran."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddm_b200 import _cabi
from tools.sweep_energy import time_config

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--dtypes", default="f32,bf16")
ap.add_argument("--ms", default="4,8,16,32")
ap.add_argument("--Ds", default="2,3072,12288")
ap.add_argument("--betas", default="0.1,1.0,2.0")
ap.add_argument("--tune", default="")
a = ap.parse_args()
L = _cabi.lib()
for kv in filter(None, a.tune.split(",")):
    _cabi.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
for dtype in a.dtypes.split(","):
  for beta in map(float, a.betas.split(",")):
    for D in map(int, a.Ds.split(",")):
        for m in map(int, a.ms.split(",")):
            esz = 4 if dtype == "f32" else 2
            algo = (2 * a.B * m * D + a.B * D) * esz
            flops = (6 * m + 4.5 * m * (m - 1)) * D * a.B
            rec = {"dtype": dtype, "beta": beta, "B": a.B, "m": m, "D": D, "kernel": _cabi.describe_energy(a.B, m, D, dtype)}
            try:
                nsets = 0 if algo > (1 << 20) else 64
                us4, gbs4 = time_config(L, a.B, m, D, dtype, nstreams=4, nsets_override=nsets, beta=beta)
                us1, gbs1 = time_config(L, a.B, m, D, dtype, nstreams=1, nsets_override=nsets, beta=beta)
                rec.update(us_4streams=round(us4, 3), GBps_4streams=round(gbs4, 1), frac_hbm=round(gbs4 / 6452.5, 3),
                           GFLOPs_4streams=round(flops / us4 / 1e3, 1), us_1stream=round(us1, 3), GBps_1stream=round(gbs1, 1))
            except Exception as e:  # noqa: BLE001
                rec["error"] = str(e)
            print(json.dumps(rec), flush=True)
