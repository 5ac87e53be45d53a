#!/bin/bash
# round-2 call H: misc measurements (fast predict errors, gradient throughput), device-path chunking A/B, sanitizer racecheck / synccheck
O=gpurun_out/${1:-r2h}; mkdir -p $O
timeout 900 python scripts/measure_misc.py --out $O/measure_misc.json > $O/measure_misc.log 2>&1; echo "misc rc=$?"; cat $O/measure_misc.log | cut -c1-400
for CT in 0 8 16 32; do B200BO_DEV_CHUNK_TILES=$CT timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_chunk$CT.json 2> $O/bench_chunk$CT.err; python - <<PY
import json
d=json.loads(open('$O/bench_chunk$CT.json').read().strip().splitlines()[-1])
print('chunk_tiles=$CT value %.4e e2e %.4e ms %.2f e2e_ms %.2f kernel %s clocks %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms_per_step'], d['clocks']))
PY
done
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_fast.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $O/sanitizer_racecheck.log
timeout 1200 compute-sanitizer --tool synccheck python scripts/sanitize_fast.py > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -5 $O/sanitizer_synccheck.log
