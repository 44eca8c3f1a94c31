#!/usr/bin/env bash
# tile-shape sweep of the tiled 2D sweeps over grid sizes (B200): marching tiles (rows per thread RPT) and
# one-cell-per-thread tiles (rows per CTA TR)
run() {
  python tools/probe_examples.py --no-ref linear_transducer_2d simple_plane_wave_2d sq1024_2d sq1448_2d convex_transducer_2d 2>/dev/null |
    python -c "
import sys, json
for l in sys.stdin:
    n, _, j = l.partition(' ')
    try: d = json.loads(j)
    except Exception: continue
    print('$1', n, d['points'], 'gpts=%.2f us/step=%.2f' % (d['engine_gpts'], d['engine_us_per_step']))
"
}
export FW25_VARIANT=2
for r in ${RPTS:-1 2 4 8}; do FW25_2D_TR=0 FW25_2D_RPT=$r run "march RPT=$r"; done
for t in ${TRS:-2 4 8}; do FW25_2D_TR=$t run "cell TR=$t"; done
