#!/bin/bash
# Round 2, call AK: K1 on one stream with its inputs L2-resident (1 / 2 / 3 rotating sets: what the launch sees inside a training
# step, where the backbone has just written the draws) against the HBM-cold headline (40 sets).
for s in 1 2 3 4 40; do
  echo "== sets=$s f32"; timeout 200 python tools/sweep_energy.py --streams 1 --sets $s --configs "variant=3"
done
for s in 1 2 4 80; do
  echo "== sets=$s bf16"; timeout 200 python tools/sweep_energy.py --streams 1 --sets $s --dtype bf16 --configs "variant=3"
done
