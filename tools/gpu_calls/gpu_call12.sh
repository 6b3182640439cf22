#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 600 > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest12.log; grep -E "s_smith|s_polar|s_right" gpurun_out/r02_pytest12.log | grep "/tc\]" | head -40
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
