#!/bin/bash
# A few seconds of GPU budget left: the default kernel choice on either side of 256 device slots.
timeout 14 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "small_system_kernel and (256-1 or 300-2 or 512-1 or 330)" > gpurun_out/r1q_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r1q_pytest.log
