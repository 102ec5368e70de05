#!/bin/bash
# one rank's share of the config-3 hypothesis split for world = 2, 4, 8 on ONE GPU, planner's item size against forced ones
for w in 2 4 8; do
  R2_WORLD=$w python tools/r2_eighth_probe.py 2>&1 | tr '\n' ' '; echo
  for g in 43 57 64 85 113 128 169 200 256 338 417 695; do
    R2_WORLD=$w RG_FORCE_GPS=$g python tools/r2_eighth_probe.py 2>&1 | tr '\n' ' '; echo
  done
done
