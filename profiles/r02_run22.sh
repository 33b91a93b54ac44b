#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_flat.py -x -q --tb=short 2>&1 | tail -30 > $O/pytest_flat5.txt
timeout 900 python -m pytest tests/test_networks.py -q --tb=short -m gpu 2>&1 | tail -8 > $O/pytest_nets5.txt
python profiles/unet_breakdown.py > $O/unet_breakdown_exact5.json 2>> $O/err5.txt
NCU=1 ncu --profile-from-start off --clock-control none -k regex:col2im \
     --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__cycles_active.avg \
     --csv --log-file $O/col2im_kernel5.csv python profiles/unet_breakdown.py > /dev/null 2>> $O/err5.txt
tail -n 6 $O/pytest_flat5.txt; tail -n 4 $O/pytest_nets5.txt; cat $O/unet_breakdown_exact5.json; grep col2im $O/col2im_kernel5.csv | cut -d, -f 5,13- | head -12; tail -n 5 $O/err5.txt
