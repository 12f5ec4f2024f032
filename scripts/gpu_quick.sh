#!/bin/bash
# quick iteration: block-level parity tests, stage clocks of pass A, then the per-shape timeline of one forward
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gated_cab or pass_a_pre or shift_block or stage1 or full_forward_golden" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
timeout 300 python scripts/stage_clocks.py 2>&1 | head -18
GSN_TIMELINE_DETAIL=1 timeout 300 python scripts/timeline_detail.py > gpurun_out/timeline_detail.txt 2>&1
head -8 gpurun_out/timeline_detail.txt
