#!/bin/bash
# compute-sanitizer passes over small banks of every demodulator variant + decode/track/est (development aid)
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for L in 32 64 128; do
  S=$([ $L = 32 ] && echo 48 || echo 160)
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/probe.py --streams $S --frames 2 --reps 1 --lanes $L > gpurun_out/memcheck_$L.log 2>&1
  echo "memcheck lanes=$L: $(grep -E 'ERROR SUMMARY' gpurun_out/memcheck_$L.log)"
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 5 python tools/probe.py --streams $S --frames 1 --reps 1 --lanes $L > gpurun_out/racecheck_$L.log 2>&1
  echo "racecheck lanes=$L: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/racecheck_$L.log)"
done
grep -h -A6 "Race reported\|Invalid\|hazard" gpurun_out/racecheck_*.log gpurun_out/memcheck_*.log | head -60
