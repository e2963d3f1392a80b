#!/bin/bash
# round 2, GPU call 11: full parity suite on the final kernels, ncu of C2 (full set) and the launch list of bench.py
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest11.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest11.log
tail -4 gpurun_out/r02_pytest11.log
python scripts/prof_sweep.py --iters 3 > gpurun_out/r02_plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 2 -f -o gpurun_out/r02_c2_final \
    python scripts/prof_sweep.py --iters 3 > gpurun_out/r02_ncu_c2.log 2>&1
tail -2 gpurun_out/r02_ncu_c2.log
python scripts/prof_sweep.py --config C3 --nlam 125000 --iters 3 > gpurun_out/r02_plain_c3.log 2>&1 &&
ncu --set full --clock-control none -k regex:sweep_kernel -s 4 -c 2 -f -o gpurun_out/r02_c3_125k \
    python scripts/prof_sweep.py --config C3 --nlam 125000 --iters 3 > gpurun_out/r02_ncu_c3.log 2>&1
tail -2 gpurun_out/r02_ncu_c3.log
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log; wc -l gpurun_out/r02_bench_launches.csv
