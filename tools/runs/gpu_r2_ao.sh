#!/bin/bash
# Round 2, call AO: K2 / K2c / K3 with and without programmatic dependent launch, same box, alternating processes.
mkdir -p gpurun_out
for i in 1 2; do for pdl in 1 0; do timeout 120 python tools/time_elementwise.py --pdl $pdl --only K2,K2c,K3; done; done 2>&1 | grep -v Warning | tee gpurun_out/elementwise_pdl_ab.log
