#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3 or abi_error" 2>&1 | tail -15
