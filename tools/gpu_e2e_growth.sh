#!/bin/bash
# streamed (host-buffer) MSM 2^26: chunk growth factor (eighths) x maximum chunk
for g in 16 13 12 11; do
  echo "== growth $g/8"
  P2B_MSM_STREAM_GROWTH=$g python tools/e2e_probe.py 26 24 25 2>&1 | grep -E "max chunk|device-resident [0-9]|plain"
done
