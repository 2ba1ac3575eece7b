#!/bin/bash
for n in 32 64 128; do python tools/diag_c5.py $n 2>&1 | grep '"lanes": 4' | tail -1 | cut -c1-200; done
