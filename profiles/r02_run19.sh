#!/bin/bash
# small-volume kernels: per-kernel durations (ncu launch lists), FLAT_MAX=16 variant, launch list of the headline step
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_flat.py "tests/test_gpu_conv.py::test_convt_one_output_channel_vs_torch" -x -q --tb=short 2>&1 | tail -8 > $O/pytest_flat2.txt
for fm in 8 16; do
  GENRE_B200_CONV_FLAT_MAX=$fm python profiles/unet_breakdown.py > $O/unet_breakdown_exact_fm$fm.json 2>> $O/err.txt
  GENRE_B200_CONV_FLAT_MAX=$fm NCU=1 ncu --profile-from-start off --clock-control none \
     --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
     --csv --log-file $O/unet_kernels_exact_fm$fm.csv python profiles/unet_breakdown.py > /dev/null 2>> $O/err.txt
done
GENRE_B200_BENCH_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_genre_step.csv python bench.py --steps 1 --warmup 3 --no-graph --skip cpu,ddp,e2e,secondary,roofline > $O/bench_under_ncu.log 2>&1
tail -n 4 $O/pytest_flat2.txt; cat $O/unet_breakdown_exact_fm8.json $O/unet_breakdown_exact_fm16.json; tail -n 5 $O/err.txt; wc -l $O/*.csv
