#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 600 -k "full_size" 2>&1 | tail -15
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -4
