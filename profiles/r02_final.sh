#!/bin/bash
# Round-2 measurement pass on ONE B200 (gpurun): everything profiles/make_r02_summary.py reads, written to gpurun_out/r02/.
#   gpurun --timeout 2400 -- 'bash profiles/r02_final.sh'
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02; mkdir -p $O
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 > $O/pytest_gpu.txt
# the driver's two arms
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --cpu-budget 90 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
# launch list of the headline step (eager launches so that every kernel is a separate record)
GENRE_B200_BENCH_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_genre_step.csv python bench.py --steps 1 --warmup 3 --no-graph --skip cpu,ddp,e2e,secondary,roofline > $O/bench_under_ncu.log 2>&1
# tensor-pipe activity of every conv kernel of one Unet_3D forward, default (exact) mode and the opt-in fp16 mode
for mode in exact f16; do
  GENRE_B200_CONV_PRECISION=$mode NCU=1 ncu --profile-from-start off --clock-control none \
     --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum \
     --csv --log-file $O/unet_tensor_pipe_$mode.csv python profiles/unet_breakdown.py > /dev/null 2>> $O/ncu.err
done
# full captures: the conv kernel of dec5, the voxeliser, the renderer
NCU=1 ncu --profile-from-start off --set full --clock-control none -k "regex:convt3d_s2_kernel|convflat_kernel|col2im_kernel|skinny_n" -c 14 -o $O/prof_unet_convs python profiles/unet_breakdown.py > /dev/null 2>> $O/ncu.err
NCU=1 ncu --profile-from-start off --set full --clock-control none -o $O/prof_render python profiles/microbench_render.py > /dev/null 2>> $O/ncu.err
for k in nnd_forward calc_prob_forward calc_prob_backward sph_project vox_splat cam_project sph_bp_backward cam_bp_backward; do
  ncu --set full --clock-control none -k regex:$k --launch-skip 3 -c 1 -o $O/prof_op_$k python profiles/microbench_ops.py > /dev/null 2>> $O/ncu.err
done
# the reports are too large to travel back (64 MiB cap on gpurun_out/): keep their raw metric tables as CSV
for f in $O/*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.rawcsv 2>/dev/null; rm -f $f; done
# per-layer / per-op / training timings
python profiles/unet_breakdown.py > $O/unet_breakdown_exact.json 2>> $O/misc.err
GENRE_B200_CONV_PRECISION=f16 python profiles/unet_breakdown.py > $O/unet_breakdown_f16.json 2>> $O/misc.err
python profiles/microbench_ops.py > $O/microbench_ops.json 2>> $O/misc.err
python profiles/microbench_render.py > $O/microbench_render.json 2>> $O/misc.err
python profiles/microbench_cam_bp.py > $O/microbench_cam_bp.json 2>> $O/misc.err
python profiles/genre_breakdown.py > $O/genre_breakdown.json 2>> $O/misc.err
python profiles/bench_train_unet.py > $O/train_unet_exact.json 2>> $O/misc.err
GENRE_B200_CONV_PRECISION=f16 python profiles/bench_train_unet.py > $O/train_unet_f16.json 2>> $O/misc.err
GENRE_B200_CONV_PRECISION=f16 NCU=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/train_unet_launches_f16.csv python profiles/bench_train_unet.py > /dev/null 2>> $O/ncu.err
python __graft_entry__.py smoke > $O/smoke.txt 2>&1
tail -n 4 $O/pytest_gpu.txt; tail -n 2 $O/smoke.txt; head -c 600 $O/bench_n1.json; echo; tail -n 3 $O/bench_n1.err $O/misc.err $O/ncu.err
