#!/bin/bash
# round-2 GPU pass 2: full GPU suite (opt-in paths now default), skip-renderer timing, training-step benches + launch list
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -80 > $O/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 --skip cpu,ddp > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 60 > $O/bench_ref.json 2> $O/bench_ref.err
GENRE_B200_CONV_PRECISION=f16 python profiles/bench_train_unet.py > $O/train_unet_f16.json 2> $O/train_unet_f16.err
python profiles/bench_train_unet.py > $O/train_unet_exact.json 2> $O/train_unet_exact.err
GENRE_B200_CONV_PRECISION=f16 NCU=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_unet_launches_f16.csv python profiles/bench_train_unet.py > /dev/null 2> $O/ncu.err
tail -n 5 $O/pytest_gpu.txt; cat $O/train_unet_f16.json $O/train_unet_exact.json; tail -n 3 $O/bench.err
