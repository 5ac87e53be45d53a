#!/bin/bash
# timeline trace + ncu full capture of the replay kernel (one product) at a given scratch budget
O=${1:-gpurun_out/replayprof}; MB=${2:-160}; mkdir -p $O
B200BO_REPLAY_MB=$MB B200BO_TRACE=$O/trace_replay_mb$MB.txt timeout 300 python bench.py --steps 1 --warmup 1 --m-per-gpu 37888 --no-cpu-baseline > $O/trace_run.log 2>&1; echo "trace rc=$?"
B200BO_REPLAY_MB=$MB timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 1 -c 1 -o $O/prof_replay_mb$MB \
  python bench.py --steps 1 --warmup 1 --m-per-gpu 37888 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"

python scripts/ncu_summary.py $O/prof_replay_mb$MB.ncu-rep > $O/replay_mb${MB}_ncu_summary.txt 2>&1; cat $O/replay_mb${MB}_ncu_summary.txt
