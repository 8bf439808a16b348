#!/bin/bash
# bench lines + ncu evidence after the round-2b kernel changes
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_r02b_reference.json 2> gpurun_out/bench_r02b_reference.err; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02b.csv \
  python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --no-ref-cuda --grid-res 128 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_tc_step|k_tc_wgrad" -s 4 -c 2 -f -o gpurun_out/step_r02b \
  python scripts/fused_once.py 4 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -30
head -c 600 gpurun_out/bench_r02b.json
