#!/bin/bash
# everything the round-end artefacts need from one 1-GPU box (about 15 minutes); then run tools/make_profiles.py r02 here
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1
timeout 300 python tools/prof_search.py 4096 > gpurun_out/prof_search.log 2>&1
timeout 300 python tools/prof_plan.py 4096 > gpurun_out/prof_plan.log 2>&1
timeout 600 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2.log 2>&1
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/bench_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kino_search -s 1 -c 1 -f -o gpurun_out/prof_search python tools/ncu_plan.py 4096 > gpurun_out/ncu_search.log 2>&1
UAVMP_QP_WARP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:qp_solve -c 1 -f -o gpurun_out/prof_qp python tools/ncu_qp.py 2400 > gpurun_out/ncu_qp.log 2>&1
tail -n 3 gpurun_out/pytest.log
