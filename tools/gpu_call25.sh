#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16_byte" 2>&1 | tail -12
