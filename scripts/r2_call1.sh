#!/bin/bash
# round 2, first GPU call: the whole GPU suite (new plan-shape parity tests included) + the pre-processing probe
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 python scripts/prep_probe.py 5000 > gpurun_out/r2a_prep_probe.log 2>&1
tail -40 gpurun_out/r2a_prep_probe.log
