#!/bin/bash
python tools/diag_stencil.py > gpurun_out/r2n_stencil.jsonl 2> gpurun_out/r2n_stencil.err; echo "rc=$?"; cat gpurun_out/r2n_stencil.jsonl; tail -3 gpurun_out/r2n_stencil.err
