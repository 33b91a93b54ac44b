#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2n; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_toolbox.py tests/test_golden.py tests/test_dropin_reference_models.py -m gpu -q --tb=short -x 2>&1 | tail -8 > $O/pytest.txt
for ov in 1 0; do for b in 32 16; do
 GENRE_B200_CAM_BP_OVERLAP=$ov B=$b python profiles/microbench_cam_bp.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('overlap',$ov,'B',$b,'python',round(d['forward_python_us'],1),'graph',round(d['forward_graph_us'],1),'GBps',round(d['forward_graph_GBps']))"
done; done > $O/overlap.txt
tail -n 4 $O/pytest.txt; cat $O/overlap.txt
timeout 600 python -m pytest tests/test_networks.py tests/test_gpu_conv.py -m gpu -q --tb=short -x 2>&1 | tail -5 > $O/pytest_nets.txt
python profiles/unet_breakdown.py > $O/unet_breakdown.json 2> $O/unet.err
timeout 600 python bench.py --steps 10 --warmup 3 --skip cpu,ddp,e2e > $O/bench.json 2> $O/bench.err
tail -n 3 $O/pytest_nets.txt; cat $O/unet_breakdown.json; python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['op_us'])"
