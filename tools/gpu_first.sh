set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
nproc; free -g | head -2
./tools/microbench > gpurun_out/microbench.json 2>&1; cat gpurun_out/microbench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 300 python tools/probe.py --streams 1024 --frames 10 2>&1 | tail -4
timeout 300 python tools/probe.py --streams 16384 --frames 4 2>&1 | tail -4
