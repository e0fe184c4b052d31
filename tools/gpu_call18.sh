#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bank_cli" 2>&1 | grep -E "Error|error|opv-demod-bank|assert|passed|failed" | head -20
