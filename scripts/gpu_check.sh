#!/bin/bash
# One gpurun call: probe, smoke, GPU tests, bench, ncu launch list + one full capture of the top kernel.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{
  echo "== probe"; nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
  ls /root/reference 2>&1 | head -2
} > $OUT/probe.log 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
echo "== sanitizer (small)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > $OUT/sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a $OUT/sanitizer.log
tail -5 $OUT/sanitizer.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench C2"; timeout 600 python bench.py --workload C2 --steps 3 --warmup 3 > $OUT/bench_C2.json 2> $OUT/bench_C2.err; echo "bench C2 rc=$?"
tail -c 1500 $OUT/bench_C2.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_fit.csv \
  python bench.py --steps 1 --warmup 3 --m-per-gpu 18944 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kstar|contract|acq_kernel|argmax|mse_kernel" -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --m-per-gpu 75776 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (contract kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract -s 2 -c 1 -o $OUT/prof_contract \
  python bench.py --steps 1 --warmup 3 --m-per-gpu 18944 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
