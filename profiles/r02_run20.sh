#!/bin/bash
# tap-GEMM + col2im kernel of the 1-channel layer: tests, per-kernel durations, headline step
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_flat.py -x -q --tb=short 2>&1 | tail -30 > $O/pytest_flat3.txt
timeout 1200 python -m pytest tests/test_networks.py tests/test_gpu_conv.py tests/test_dropin_reference_models.py -q --tb=short -m gpu 2>&1 | tail -30 > $O/pytest_nets3.txt
python profiles/unet_breakdown.py > $O/unet_breakdown_exact3.json 2>> $O/err3.txt
GENRE_B200_CONV_PRECISION=f16 python profiles/unet_breakdown.py > $O/unet_breakdown_f163.json 2>> $O/err3.txt
NCU=1 ncu --profile-from-start off --clock-control none \
     --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
     --csv --log-file $O/unet_kernels_exact3.csv python profiles/unet_breakdown.py > /dev/null 2>> $O/err3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --skip cpu,ddp > $O/bench3.json 2>> $O/err3.txt
tail -n 12 $O/pytest_flat3.txt; tail -n 6 $O/pytest_nets3.txt; cat $O/unet_breakdown_exact3.json $O/unet_breakdown_f163.json; head -c 400 $O/bench3.json; echo; tail -n 5 $O/err3.txt
