#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zz_persist_gpu.py -q > gpurun_out/r2z_persist_tests.log 2>&1
tail -25 gpurun_out/r2z_persist_tests.log
(timeout 200 python tools/persist_probe.py ab_gas; timeout 300 python tools/persist_probe.py graphene_cu) > gpurun_out/r2z_persist_probe.txt 2>&1
cat gpurun_out/r2z_persist_probe.txt | tail -30
