#!/bin/bash
python tools/diag_c5.py 128 > gpurun_out/r2d_diag_c5.log 2>&1; cat gpurun_out/r2d_diag_c5.log
