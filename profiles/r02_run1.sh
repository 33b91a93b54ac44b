#!/bin/bash
# round-2 first GPU pass
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -q -x --deselect tests/test_networks.py 2>&1 | tail -15 > gpurun_out/r2a/pytest_gpu.txt
python -m pytest tests/test_networks.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2a/pytest_networks.txt
GENRE_B200_BN_TRAIN=1 GENRE_B200_CONV_TC_BACKWARD=1 python -m pytest tests/test_gpu_conv.py -q -k "tensor_core_input_gradients or bn_act_train" 2>&1 | tail -30 > gpurun_out/r2a/pytest_optin.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a/bench_ref.json 2> gpurun_out/r2a/bench_ref.err
tail -3 gpurun_out/r2a/pytest_gpu.txt gpurun_out/r2a/pytest_networks.txt gpurun_out/r2a/pytest_optin.txt
head -c 3000 gpurun_out/r2a/bench.json; tail -5 gpurun_out/r2a/bench.err
