#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 -k "culled_engine_posterior" --durations=3 2>&1 | tail -25
