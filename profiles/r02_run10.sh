#!/bin/bash
# TMA halo producer: conv / network / drop-in tests with the tensor-copy path (default) and the cp.async path, Unet timing both ways
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2j; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_networks.py tests/test_dropin_reference_models.py -m gpu -q --tb=short -x 2>&1 | tail -30 > $O/pytest_tma.txt
GENRE_B200_CONV_TMA=0 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --tb=short -x 2>&1 | tail -8 > $O/pytest_cpasync.txt
python profiles/unet_breakdown.py > $O/unet_tma.json 2> $O/unet_tma.err
GENRE_B200_CONV_TMA=0 python profiles/unet_breakdown.py > $O/unet_cpasync.json 2>> $O/unet_tma.err
GENRE_B200_CONV_PRECISION=f16 python profiles/unet_breakdown.py > $O/unet_tma_f16.json 2>> $O/unet_tma.err
GENRE_B200_CONV_PRECISION=f16 GENRE_B200_CONV_TMA=0 python profiles/unet_breakdown.py > $O/unet_cpasync_f16.json 2>> $O/unet_tma.err
tail -n 6 $O/pytest_tma.txt $O/pytest_cpasync.txt; cat $O/unet_tma.json $O/unet_cpasync.json $O/unet_tma_f16.json $O/unet_cpasync_f16.json; tail -n 5 $O/unet_tma.err
