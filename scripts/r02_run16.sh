#!/bin/bash
# round 2, GPU call 16: relay kernel — full bench line, ncu full capture on C2 and on the C3 shard, launch list, e2e phases
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench exit $?"
python scripts/prof_sweep.py --iters 3 > gpurun_out/r02b_plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 2 -f -o gpurun_out/r02b_c2_relay \
    python scripts/prof_sweep.py --iters 3 > gpurun_out/r02b_ncu_c2.log 2>&1
tail -2 gpurun_out/r02b_ncu_c2.log
python scripts/prof_sweep.py --config C3 --nlam 125000 --iters 3 > gpurun_out/r02b_plain_c3.log 2>&1 &&
ncu --set full --clock-control none -k regex:sweep_kernel -s 4 -c 2 -f -o gpurun_out/r02b_c3_125k_relay \
    python scripts/prof_sweep.py --config C3 --nlam 125000 --iters 3 > gpurun_out/r02b_ncu_c3.log 2>&1
tail -2 gpurun_out/r02b_ncu_c3.log
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02b_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02b_ncu_bench.log 2>&1
tail -2 gpurun_out/r02b_ncu_bench.log; wc -l gpurun_out/r02b_bench_launches.csv
python scripts/e2e_phases.py > gpurun_out/r02b_e2e_phases.log 2>&1; tail -3 gpurun_out/r02b_e2e_phases.log
head -c 600 gpurun_out/r02b_bench_n1.json
