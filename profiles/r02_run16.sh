#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
python -m pytest tests/test_dropin_reference_models.py -m gpu -q 2>&1 | tail -2
python - <<P
import json
d=json.loads(open("$O/bench_n1.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["nets2d"])
P
