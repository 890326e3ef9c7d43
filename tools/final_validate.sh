#!/bin/bash
# Round-2 final validation on one B200: GPU test suite, smoke, the bench line, ncu launch list and one
# full capture of the drive-mode kernel.  Every step is bounded; the whole script fits the GPU budget.
mkdir -p gpurun_out
echo "== pytest (established suite)"; timeout 200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_map.py -p no:cacheprovider > gpurun_out/r2_final_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_final_pytest.log
echo "== pytest (device Newton, own process)"; timeout 120 python -m pytest tests/test_gpu_map.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_final_pytest_map.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_final_pytest_map.log
echo "== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 150 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "rc=$?"; tail -2 gpurun_out/r2_bench_final.err; cat gpurun_out/r2_bench_final.json
echo "== ncu launch list"; timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench_c3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2_ncu_launch_bench.log 2>&1; echo "rc=$?"
echo "== ncu full (second launch of the drive kernel = the timed one)"; timeout 100 ncu --set full --clock-control none --import-source on -k regex:eval_persist -s 1 -c 1 -f -o gpurun_out/r2_drive_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2_ncu_full_bench.log 2>&1; echo "rc=$?"
timeout 30 ncu -i gpurun_out/r2_drive_full.ncu-rep --page raw --csv > gpurun_out/r2_drive_full_raw.csv 2>/dev/null; echo "rc=$?"; ls -la gpurun_out/r2_drive_full* gpurun_out/r2_launches_bench_c3.csv 2>/dev/null
