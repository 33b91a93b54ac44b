#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2f; mkdir -p $O
for c in 1 2 4 8 16; do GENRE_B200_VOX_CHUNKS=$c python profiles/microbench_cam_bp.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunks',$c, 'python',round(d['forward_python_us'],1),'graph',round(d['forward_graph_us'],1))" ; done > $O/chunks.txt
GENRE_B200_CAM_BP_PIPELINE=0 python profiles/microbench_cam_bp.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nopipe python',round(d['forward_python_us'],1),'graph',round(d['forward_graph_us'],1))" >> $O/chunks.txt
cat $O/chunks.txt
