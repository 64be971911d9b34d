#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/probe_mainloop.py > gpurun_out/probe_mainloop.log 2>&1; cat gpurun_out/probe_mainloop.log
