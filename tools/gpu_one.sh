#!/bin/bash
# usage: gpu_one.sh <pytest args...>  -- one pytest invocation on the GPU box, log to gpurun_out/one.log
mkdir -p gpurun_out
timeout 900 python -m pytest "$@" -q -x -m gpu --timeout=600 -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/one.log
