#!/bin/bash
# ncu --set full captures of the other kernels north_star asks evidence for: the IVF list scan
# (tensor regime: phase B of a 250k-query search; HBM regime: the small-batch scan), the k-means
# centroid update and the small-batch flat scan.
O=gpurun_out; mkdir -p $O
timeout 500 ncu --set full --clock-control none --import-source on -k regex:topk_tc3d -s 5 -c 1 -o $O/r02f_prof_ivf_phaseB -f python scripts/ivf_ncu_target.py > $O/r02f_ncu_ivf.log 2>&1; tail -1 $O/r02f_ncu_ivf.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:km_partial -s 1 -c 1 -o $O/r02f_prof_km_partial -f python scripts/bench_kmeans_ncu.py > $O/r02f_ncu_km.log 2>&1; tail -1 $O/r02f_ncu_km.log
NRB_IVF_NQ=64 NRB_WARM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ivf_small_scores -s 1 -c 1 -o $O/r02f_prof_ivf_small -f python scripts/ivf_ncu_target.py > $O/r02f_ncu_ivfs.log 2>&1; tail -1 $O/r02f_ncu_ivfs.log
