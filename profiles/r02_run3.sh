#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c; mkdir -p $O
python -m pytest tests/test_dropin_reference_models.py -m gpu -q --tb=long -k shapehd 2>&1 | grep -v Warning | head -120 > $O/pytest_shapehd.txt
python profiles/microbench_render.py > $O/microbench_render.json 2> $O/microbench_render.err
NCU=1 ncu --profile-from-start off --set full --import-source on --clock-control none -o $O/prof_r02_render python profiles/microbench_render.py > /dev/null 2> $O/ncu_render.err
python -m pytest tests/test_gpu_conv.py tests/test_networks.py -m gpu -q --tb=short -x 2>&1 | tail -40 > $O/pytest_conv.txt
GENRE_B200_CONV_EXACT_IMPL=f16x2 python -m pytest tests/test_networks.py tests/test_dropin_reference_models.py -m gpu -q --tb=short 2>&1 | tail -40 > $O/pytest_x2_nets.txt
GENRE_B200_CONV_EXACT_IMPL=f16x2 timeout 900 python bench.py --steps 10 --warmup 3 --skip cpu,ddp,e2e > $O/bench_x2.json 2> $O/bench_x2.err
tail -n 5 $O/pytest_conv.txt $O/pytest_x2_nets.txt; cat $O/microbench_render.json; tail -n 3 $O/bench_x2.err
