#!/usr/bin/env bash
# Informational (GPU box): rebuild the warp-specialised sweeps with different tuning knobs and time each.
cd "$(dirname "$0")/.."
GRID=${1:-320x632x632}
for cfg in "8 3 4 24" "8 3 2 24" "8 3 1 24" "16 2 4 24" "16 2 2 24" "8 4 4 24" "8 3 4 12" "8 3 4 6" "16 2 4 12"; do
  set -- $cfg
  FW25_WS_TY=$1 FW25_WS_MINB=$2 FW25_WS_UNROLL=$3 python -m fullwave25_b200.build --force >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  echo "TY=$1 MINB=$2 UNROLL=$3 WAVES=$4: $(FW25_WS_WAVES=$4 python tools/probe_ws.py $GRID 10 3 2>&1 | tail -1)"
done
python -m fullwave25_b200.build --force >/dev/null 2>&1
