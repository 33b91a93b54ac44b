#!/bin/bash
# weight multicast across clusters of the conv kernels: identity tests, per-layer timing per cluster size
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_flat.py "tests/test_gpu_conv.py::test_weight_multicast_clusters_give_identical_results" -x -q --tb=short 2>&1 | tail -30 > $O/pytest_cluster.txt
for cl in 1 2 4 8; do
  GENRE_B200_CONV_CLUSTER=$cl timeout 300 python profiles/unet_breakdown.py > $O/unet_breakdown_exact_cl$cl.json 2>> $O/err4.txt
  GENRE_B200_CONV_CLUSTER=$cl GENRE_B200_CONV_PRECISION=f16 timeout 300 python profiles/unet_breakdown.py > $O/unet_breakdown_f16_cl$cl.json 2>> $O/err4.txt
done
GENRE_B200_CONV_CLUSTER=8 timeout 900 python -m pytest tests/test_networks.py tests/test_gpu_conv.py -q --tb=short -m gpu 2>&1 | tail -8 > $O/pytest_nets_cl8.txt
tail -n 12 $O/pytest_cluster.txt; tail -n 4 $O/pytest_nets_cl8.txt
for cl in 1 2 4 8; do python - <<P
import json
for m in ("exact","f16"):
    d=json.load(open("$O/unet_breakdown_%s_cl$cl.json"%m))["custom"]
    print("cl$cl",m,{k:d[k] for k in ("enc1","enc2","enc3","dec4","dec5","total")})
P
done
tail -n 5 $O/err4.txt
