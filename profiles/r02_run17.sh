#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2p; mkdir -p $O
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 > $O/pytest_gpu.txt
for w in shapehd wgan genre; do
NCU=$w ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_${w}_launches.csv python profiles/bench_train_ddp.py --which $w > /dev/null 2>> $O/ncu.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_genre_step.csv python bench.py --steps 2 --warmup 1 --no-graph --skip cpu,ddp,e2e,secondary,roofline > $O/bench_under_ncu.log 2>&1
tail -n 4 $O/pytest_gpu.txt; tail -n 3 $O/ncu.err
