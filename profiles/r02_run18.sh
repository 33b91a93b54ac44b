#!/bin/bash
# small-volume kernels (convflat.cu, skinny_gemm.cu): tests, per-layer timings, headline step
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_flat.py -x -q --tb=short 2>&1 | tail -40 > $O/pytest_flat.txt
timeout 1200 python -m pytest tests/test_networks.py tests/test_gpu_conv.py tests/test_dropin_reference_models.py -q --tb=short -m gpu 2>&1 | tail -30 > $O/pytest_nets.txt
python profiles/unet_breakdown.py > $O/unet_breakdown_exact.json 2> $O/err.txt
GENRE_B200_CONV_PRECISION=f16 python profiles/unet_breakdown.py > $O/unet_breakdown_f16.json 2>> $O/err.txt
timeout 600 python bench.py --steps 20 --warmup 5 --skip cpu,ddp > $O/bench.json 2>> $O/err.txt
tail -n 12 $O/pytest_flat.txt; tail -n 6 $O/pytest_nets.txt; cat $O/unet_breakdown_exact.json; head -c 400 $O/bench.json; echo; tail -n 5 $O/err.txt
