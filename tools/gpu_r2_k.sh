#!/bin/bash
for sm in 0 1; do echo "== MJB_SMAP=$sm"; MJB_SMAP=$sm timeout 300 python tools/ab_kernels.py 2>&1 | tail -1; done
MJB_SPLIT=0 MJB_SMAP=0 timeout 300 python tools/ab_kernels.py 2>&1 | tail -1
MJB_SPLIT=0 MJB_SMAP=1 timeout 300 python tools/ab_kernels.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
