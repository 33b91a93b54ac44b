#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q; mkdir -p $O
NCU=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:col2im -c 1 -o $O/prof_col2im python profiles/unet_breakdown.py > /dev/null 2>> $O/err6.txt
ncu -i $O/prof_col2im.ncu-rep --page details --csv > $O/prof_col2im_details.csv 2>/dev/null
ncu -i $O/prof_col2im.ncu-rep --page source --csv > $O/prof_col2im_source.csv 2>/dev/null
rm -f $O/prof_col2im.ncu-rep
wc -c $O/prof_col2im_*.csv; tail -n 3 $O/err6.txt
