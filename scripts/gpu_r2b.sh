#!/bin/bash
# round-2 call B: smoke with the fast path, new rowdot kernel, band-stage launch list, the new default bench (M = 1e7)
OUT=gpurun_out/${1:-r2b}
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -2 $OUT/smoke.log
python -m pytest tests/test_fast_gpu.py -x -q -m gpu -k "argmax or acq or band" > $OUT/test_fast.log 2>&1; echo "fast tests rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/test_fast.log
python -m pytest tests/test_scale_gpu.py -x -q -m gpu -s > $OUT/test_scale.log 2>&1; echo "scale tests rc=$?" | tee -a $OUT/summary.txt; tail -6 $OUT/test_scale.log
python scripts/band_probe.py C3 C2 C4 C5 --out $OUT/band_probe.json > $OUT/band_probe.log 2>&1; echo "probe rc=$?" | tee -a $OUT/summary.txt
grep -E "^C[2-5] " $OUT/band_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_band_C3.csv python scripts/band_probe.py C3 --m 300000 > $OUT/ncu_probe.log 2>&1
python - $OUT/launches_band_C3.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1]))); hdr=None; out=[]
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r
    elif hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); out.append((d['Kernel Name'][:50], d['Metric Value']))
for k,v in out[-40:]: print(k, v)
PY
python bench.py --steps 10 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -c 3000 $OUT/bench_default.json; tail -5 $OUT/bench_default.err
