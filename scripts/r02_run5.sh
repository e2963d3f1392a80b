#!/bin/bash
# round 2, GPU call 5: parity suite (incl. load-path goldens), e2e phases, then memcheck of every kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest5.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest5.log
tail -6 gpurun_out/r02_pytest5.log
timeout 300 python scripts/e2e_phases.py > gpurun_out/r02_e2e_phases.log 2>&1; tail -5 gpurun_out/r02_e2e_phases.log
timeout 600 python scripts/sanitize_cases.py > gpurun_out/r02_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python scripts/sanitize_cases.py > gpurun_out/r02_memcheck_stdout.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/r02_sanitize_plain.log; tail -5 gpurun_out/r02_memcheck.log; tail -3 gpurun_out/r02_memcheck_stdout.log
