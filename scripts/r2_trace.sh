#!/bin/bash
# IVF scan timelines (trace build): phase A and phase B launches of the last query batch
O=gpurun_out; mkdir -p $O
export NRB_LIB=$PWD/newsrecommend_b200/libnrb200_trace.so
for m in 1 2; do
  NRB_TRACE_IVF_MODE=$m timeout 300 python scripts/trace_ivf.py > $O/r02_trace_ivf_mode$m.json 2> $O/r02_trace_ivf_mode$m.err
  tail -c 400 $O/r02_trace_ivf_mode$m.err; cat $O/r02_trace_ivf_mode$m.json
  cp $O/trace_ivf_raw.npz $O/r02_trace_ivf_raw_mode$m.npz
done
timeout 300 python scripts/trace_timeline.py > $O/r02_trace_flat.log 2>&1; tail -40 $O/r02_trace_flat.log
