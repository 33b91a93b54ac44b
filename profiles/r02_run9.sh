#!/bin/bash
# 2-GPU pass: the bench line with e2e + DDP blocks, and the stand-alone DDP bench at N=1 for reference
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python bench.py --steps 10 --warmup 3 --skip cpu > $O/bench_n1.json 2> $O/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -n 5 $O/bench_n1.err $O/bench_n2.err
python - <<'P'
import json
for f in ("gpurun_out/r2i/bench_n1.json","gpurun_out/r2i/bench_n2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"] if d["e2e"] else None)
        print(json.dumps(d["secondary_ddp"]))
    except Exception as e:
        print(f, "ERR", e)
P
