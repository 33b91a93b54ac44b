#!/bin/bash
# bench.py on N GPUs of one box:  gpurun --gpus N -- 'NGPU=N bash profiles/r02_multi_gpu.sh'  -> gpurun_out/r02/bench_nN.json
cd $GRAFT_REPO_ROOT
O=gpurun_out/r02; mkdir -p $O
N=${NGPU:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -n 3 $O/bench_n$N.err
python - <<P
import json
d=json.loads(open("$O/bench_n$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
print(json.dumps(d["secondary_ddp"])[:2500])
P
