#!/bin/bash
# bash tools/gpu_tests.sh <tag> [pytest args]: the GPU test suite, log kept in gpurun_out/
tag=$1; shift; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q "$@" > $out/${tag}_pytest.log 2>&1; tail -12 $out/${tag}_pytest.log
