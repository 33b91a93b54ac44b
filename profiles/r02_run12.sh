#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2l; mkdir -p $O
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > $O/pytest_gpu.txt
tail -n 12 $O/pytest_gpu.txt
