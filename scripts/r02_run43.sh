#!/bin/bash
# round 2, GPU call 43: final code — whole GPU suite, launch list of bench.py, ncu of the per-level kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest.log
tail -3 gpurun_out/r02i_pytest.log
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_ncu_bench.log 2>&1
wc -l gpurun_out/r02i_bench_launches.csv
python scripts/prof_sweep.py --iters 3 > gpurun_out/r02i_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:post_level -s 4 -c 2 -f -o gpurun_out/r02i_post_level \
    python scripts/prof_sweep.py --iters 3 > gpurun_out/r02i_ncu_post.log 2>&1
tail -1 gpurun_out/r02i_ncu_post.log
