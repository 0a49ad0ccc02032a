#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -k "fp32_fast_sweep_agrees" 2>&1 | tail -15
