#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest6.log
grep -v "^\s*$" gpurun_out/r02_pytest6.log | tail -40
