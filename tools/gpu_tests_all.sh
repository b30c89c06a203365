#!/bin/bash
# all gpu tests, no -x: the whole failure list in one call
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/r2_gputests_all.log 2>&1
echo "gputests exit=$?"; tail -n 15 gpurun_out/r2_gputests_all.log
