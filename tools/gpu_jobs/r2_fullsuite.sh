#!/bin/bash
# round 2: the whole GPU suite (what the driver runs at round end) + smoke on the current tree
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --maxfail=8 --tb=short > gpurun_out/r2_fullsuite.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r2_fullsuite.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
