#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2m; mkdir -p $O
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > $O/pytest_gpu.txt
python profiles/unet_breakdown.py > $O/unet_breakdown.json 2> $O/unet.err
timeout 600 python bench.py --steps 10 --warmup 3 --skip cpu,ddp,e2e > $O/bench.json 2> $O/bench.err
tail -n 6 $O/pytest_gpu.txt; cat $O/unet_breakdown.json; python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
