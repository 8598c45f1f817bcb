#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_vit.log 2>&1; echo "== pytest vit exit $?"; tail -n 40 gpurun_out/pytest_vit.log | cut -c1-300
cat gpurun_out/parity_vit_*.txt
