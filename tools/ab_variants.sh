#!/bin/bash
# A/B of prebuilt library variants on one box: usage tools/ab_variants.sh base pf kt512
mkdir -p gpurun_out
P=uav_motion_planning_b200
for v in "$@"; do
  cp $P/variants/$v.so $P/libuavmp.so
  timeout 300 python -m pytest tests/test_kino_parity.py -m gpu -q -x > gpurun_out/ab_${v}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_${v}_pytest.log
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ab_${v}_bench.log 2>&1
  timeout 200 python tools/prof_search.py 4096 > gpurun_out/ab_${v}_prof.log 2>&1
  echo "== $v: $(tail -n 2 gpurun_out/ab_${v}_pytest.log | tr '\n' ' ')"
  python - <<EOP
import json
for l in open('gpurun_out/ab_${v}_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],1))
EOP
done
