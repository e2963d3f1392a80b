#!/bin/bash
# round 2, GPU call 19: single-sync emission_spectrum — parity of the solves, e2e phases
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "reference or convergence or emission or grid or kat or solve or diagnos or api" 2>&1 | tail -4
python scripts/e2e_phases.py 2>&1 | tail -6
} > gpurun_out/r02_run19.log 2>&1
cat gpurun_out/r02_run19.log
