#!/bin/bash
# Round 2, call R: time tiling of the headline configuration (1,024 streams x 250 frames): uniform against a short last tile.
set -x -o pipefail
mkdir -p gpurun_out
P="timeout 120 python tools/probe.py --streams 1024 --frames 250 --reps 3"
$P --tiles 10 2>&1 | tail -1 | cut -c1-260 || exit 1
$P --tiles 5 2>&1 | tail -1 | cut -c1-260
$P --tile-frames 62,124,186,246 2>&1 | tail -1 | cut -c1-260
$P --tile-frames 50,100,150,200,240,248 2>&1 | tail -1 | cut -c1-260
$P --tile-frames 83,166,240,248 2>&1 | tail -1 | cut -c1-260
$P --tile-frames 125,240,248 2>&1 | tail -1 | cut -c1-260
