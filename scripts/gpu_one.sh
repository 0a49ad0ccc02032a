#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -k "adversarial or sharded_chain" 2>&1 | tail -30
