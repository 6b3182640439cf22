#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke29.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke29.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest29.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest29.log
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
timeout 300 python tools/active_tiles_probe.py 2>&1 | tee gpurun_out/r02_active_tiles_service.txt
