#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_step.py -x -q -m gpu --timeout 60 2>&1 | tail -2
{
echo "== default lib, tile kernel alone"; DIGS_B200_STEP_NO_WGRAD=1 timeout 120 python scripts/fused_once.py 12
echo "== default lib, full step"; timeout 120 python scripts/fused_once.py 12
} > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log | grep -v Warning
