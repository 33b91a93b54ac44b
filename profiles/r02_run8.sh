#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2h; mkdir -p $O
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > $O/pytest_gpu.txt
python profiles/microbench_ops.py > $O/microbench_ops.json 2> $O/microbench_ops.err
python profiles/microbench_render.py > $O/microbench_render_occ5.json 2> $O/microbench_render.err
GENRE_B200_RENDER_OCC=4 python profiles/microbench_render.py > $O/microbench_render_occ4.json 2>> $O/microbench_render.err
python profiles/genre_breakdown.py > $O/genre_breakdown.json 2> $O/genre_breakdown.err
timeout 900 python bench.py --steps 10 --warmup 3 --skip cpu,ddp > $O/bench.json 2> $O/bench.err
ncu --set full --clock-control none -k regex:'nnd_forward|calc_prob_forward|calc_prob_backward|sph_project|sph_bp_backward|cam_bp_backward|surface_mask|cam_project|vox_splat' -c 14 -o $O/prof_r02_ops python profiles/microbench_ops.py > /dev/null 2> $O/ncu_ops.err
tail -n 4 $O/pytest_gpu.txt; cat $O/microbench_ops.json; cat $O/microbench_render_occ5.json $O/microbench_render_occ4.json; cat $O/genre_breakdown.json; tail -n 3 $O/bench.err
