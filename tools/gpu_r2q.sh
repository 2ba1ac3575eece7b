#!/bin/bash
timeout 600 python -m pytest tests/test_batch.py tests/test_bayes.py -m gpu -x -q < /dev/null > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2q_pytest.log
python tools/diag_c5.py 128 2>&1 | cut -c1-160 | grep -v kernel_ms
