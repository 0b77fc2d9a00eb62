#!/bin/bash
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DVK_STATS -o /tmp/lib_stats.so mjpl_b200/csrc/mjpl_b200.cu 2>&1 | grep error
MJPL_B200_LIB=/tmp/lib_stats.so python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch, ctypes as C
import mjpl_b200 as mj
from mjpl_b200 import models, _abi
from bench import make_rows, MODEL, ALLOWED
model=models.load(MODEL); eng=mj.get_engine(model, ALLOWED)
q=torch.from_numpy(make_rows(model,1_000_000)).cuda()
eng.valid_configs(q,3); torch.cuda.synchronize(); eng.reset_stats()
eng.valid_configs(q,3|8); torch.cuda.synchronize()
import ctypes
st=eng.stats(); print(st)
# raw counters: trips in counters[7] not exposed; uncertain_rows=flushes, queue_overflow=busy
print('trips', st['launches'], 'avg busy groups per trip', st['queue_overflow']/max(1,st['launches']), 'trips per warp-tile', st['launches']/(1e6/32))
print('flushes per warp-tile', st['uncertain_rows']/(1e6/32), 'busy-group-trips', st['queue_overflow'], 'items', st['narrow_items'])
PY
