#!/bin/bash
# Round-1 measurement pass on one B200 (run through gpurun from the repo root); outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_r01.err
# launch list of the same bench command (per-launch durations are cold-cache/serialised: shares, not absolutes)
GENRE_B200_BENCH_SECONDARY=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
# one warm Unet_3D forward: every kernel (launch list) and a full capture of the tcgen05 kernels
NCU=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/unet_launches_r01.csv python profiles/unet_breakdown.py > gpurun_out/unet_ncu.log 2>&1
NCU=1 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name regex:convt3d_s2_kernel \
    -o gpurun_out/prof_r01_unet_tc python profiles/unet_breakdown.py > gpurun_out/unet_ncu_full.log 2>&1
python profiles/unet_breakdown.py > gpurun_out/unet_breakdown_r01_final.json 2>/dev/null
python profiles/microbench_conv.py > gpurun_out/microbench_conv_r01_final.json 2>/dev/null
python profiles/net_breakdown.py > gpurun_out/net_breakdown_r01.json 2>/dev/null
python profiles/bench_train_unet.py > gpurun_out/train_unet_r01.json 2>/dev/null
python profiles/bench_train_ddp.py > gpurun_out/train_n1_custom.json 2>/dev/null
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench_r01.json
