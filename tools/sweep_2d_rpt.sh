#!/usr/bin/env bash
# rows-per-CTA sweep of the tiled 2D sweeps over grid sizes (B200): which tile height wins where?
for r in 1 2 4 8; do
  FW25_VARIANT=2 FW25_2D_RPT=$r python tools/probe_examples.py --no-ref linear_transducer_2d simple_plane_wave_2d sq1024_2d sq1448_2d convex_transducer_2d 2>/dev/null |
    python -c "
import sys, json
for l in sys.stdin:
    n, _, j = l.partition(' ')
    try: d = json.loads(j)
    except Exception: continue
    print('RPT=$r', n, d['points'], 'gpts=%.2f us/step=%.2f' % (d['engine_gpts'], d['engine_us_per_step']))
"
done
