#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2g; mkdir -p $O
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > $O/pytest_gpu.txt
python profiles/microbench_render.py > $O/microbench_render.json 2> $O/microbench_render.err
NCU=1 ncu --profile-from-start off --set full --import-source on --clock-control none -o $O/prof_r02_render4 python profiles/microbench_render.py > /dev/null 2> $O/ncu_render.err
timeout 900 python bench.py --steps 10 --warmup 3 --skip cpu,ddp > $O/bench.json 2> $O/bench.err
tail -n 6 $O/pytest_gpu.txt; cat $O/microbench_render.json; tail -n 3 $O/bench.err
