#!/bin/bash
# round-2 call A: new band pipeline -- fast-path tests, band probe over the four workloads, chunked-launch A/B, scale tests
OUT=gpurun_out/${1:-r2a}
mkdir -p $OUT
python -m pytest tests/test_fast_gpu.py tests/test_candidates.py -x -q -m gpu > $OUT/test_fast.log 2>&1; echo "fast tests rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/test_fast.log
python scripts/band_probe.py C3 C4 C2 C5 --out $OUT/band_probe.json > $OUT/band_probe.log 2>&1; echo "probe rc=$?" | tee -a $OUT/summary.txt
tail -20 $OUT/band_probe.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
B200BO_DEV_CHUNK_TILES=8 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_C3_chunk8.json 2> $OUT/bench_C3_chunk8.err
B200BO_DEV_CHUNK_TILES=16 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_C3_chunk16.json 2> $OUT/bench_C3_chunk16.err
python - <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/r2a/bench_C3*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.3e e2e %.3e ms %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['kernel_ms_per_step'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
python -m pytest tests/test_scale_gpu.py -x -q -m gpu -s > $OUT/test_scale.log 2>&1; echo "scale tests rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/test_scale.log
