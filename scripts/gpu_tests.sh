#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -25
