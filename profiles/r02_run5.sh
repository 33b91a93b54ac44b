#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2e; mkdir -p $O
python -m pytest tests/test_gpu_toolbox.py tests/test_golden.py tests/test_dropin_reference_models.py tests/test_abi.py -m gpu -q --tb=short 2>&1 | tail -40 > $O/pytest_gpu.txt
python profiles/microbench_render.py > $O/microbench_render.json 2> $O/microbench_render.err
python profiles/microbench_cam_bp.py > $O/microbench_cam_bp.json 2> $O/microbench_cam_bp.err
GENRE_B200_CAM_BP_PIPELINE=0 python profiles/microbench_cam_bp.py > $O/microbench_cam_bp_nopipe.json 2>> $O/microbench_cam_bp.err
B=16 python profiles/microbench_cam_bp.py > $O/microbench_cam_bp_b16.json 2>> $O/microbench_cam_bp.err
NCU=1 ncu --profile-from-start off --set full --import-source on --clock-control none -o $O/prof_r02_render3 python profiles/microbench_render.py > /dev/null 2> $O/ncu_render.err
timeout 900 python bench.py --steps 10 --warmup 3 --skip cpu,ddp > $O/bench.json 2> $O/bench.err
tail -n 6 $O/pytest_gpu.txt; cat $O/microbench_render.json; cat $O/microbench_cam_bp.json $O/microbench_cam_bp_nopipe.json $O/microbench_cam_bp_b16.json; tail -n 3 $O/bench.err
