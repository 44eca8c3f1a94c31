#!/usr/bin/env bash
# Tuning sweep of the 2D tiled sweeps on the GPU box: resident CTAs per SM (register budget) x rows per CTA.
# usage (under gpurun): bash tools/sweep_2d.sh > gpurun_out/sweep_2d.txt
set -u
for MINB in 4 5 6 8; do
  FW25_WS_2D_MINB=$MINB python -m fullwave25_b200.build --force > /dev/null 2>&1 || { echo "build failed MINB=$MINB"; continue; }
  for RPT in 4 8; do
    FW25_2D_RPT=$RPT python tools/probe_examples.py --no-ref simple_plane_wave_2d linear_transducer_2d convex_transducer_2d 2>/dev/null |
      python -c "
import sys, json
for l in sys.stdin:
    n, _, j = l.partition(' ')
    try: d = json.loads(j)
    except Exception: continue
    print('MINB=$MINB RPT=$RPT', n, 'gpts=%.2f us/step=%.2f' % (d['engine_gpts'], d['engine_us_per_step']))
"
  done
done
python -m fullwave25_b200.build --force > /dev/null 2>&1
