#!/bin/bash
# pinned caller buffers used in place (pull / mirror) against the copy-engine path: parity test, then the C2 time lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pinned or pipelined or compact" > gpurun_out/r02_pytest_n.log 2>&1; tail -15 gpurun_out/r02_pytest_n.log
timeout 300 python tools/e2e_trace.py > gpurun_out/r02_e2e_trace2.log 2>&1; grep -v "^bbcp_probe" gpurun_out/r02_e2e_trace2.log | cut -c1-200; grep "^bbcp_probe" gpurun_out/r02_e2e_trace2.log | grep -E "pull|flags in place" | cut -c1-250 | head -30
