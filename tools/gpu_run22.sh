#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== adapter e2e benzene"; timeout 900 python tools/adapter_e2e.py > gpurun_out/adapter_e2e.log 2>&1; cat gpurun_out/adapter_e2e.log
