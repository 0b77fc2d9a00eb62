#!/bin/bash
# two-kernel pipeline on the other workloads (edges, batched planning), same box
for mode in 0 1; do
echo "MJB_SPLIT=$mode"
MJB_SPLIT=$mode timeout 600 python tools/bench_extra.py edges plans --queries 4096 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: d[k] for k in ('case', 'edges_per_s', 'configs_per_s', 'plans_per_s', 'solved', 'seconds', 'replay_failures', 'agreement_on_subsample') if k in d})
"
done
MJB_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_ik.py tests/test_gpu_pose.py -x -q 2>&1 | tail -2
