#!/bin/bash
# Round 2, call AS: small minibatches (the reference's CIFAR recipe is 64 rows per GPU on 4 GPUs, 32 on 8): D-split clusters
# so that B x cluster CTAs cover the SMs, one stream, HBM-cold.
mkdir -p gpurun_out
{
for dt in f32 bf16; do for B in 16 32 64 96; do
  echo "== $dt B=$B one stream"
  timeout 120 python tools/sweep_energy.py --streams 1 --dtype $dt --B $B --configs "variant=3,cluster=1;variant=3,cluster=2;variant=3,cluster=4;variant=3,cluster=8"
done; done
} 2>&1 | grep -v Warning | tee gpurun_out/k1_small_batch_clusters.log
