#!/bin/bash
O=gpurun_out/${1:-r2e}; mkdir -p $O
timeout 300 python scripts/append_debug.py > $O/append_debug.log 2>&1; tail -20 $O/append_debug.log
for DBG in 0 1 2 3; do echo "== OZ_DEBUG=$DBG"; B200BO_OZ_DEBUG=$DBG timeout 300 python scripts/oz_check.py 3968 2>&1 | grep -E "^3968 [78]" | cut -c1-60; done
for G in 2 4 16; do echo "== OZ_G=$G"; B200BO_OZ_G=$G timeout 300 python scripts/oz_check.py 3968 2>&1 | grep -E "^3968 [78]" | cut -c1-60; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_syrk -s 2 -c 1 -o $O/prof_oz python scripts/oz_check.py 3968 > $O/ncu_oz.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py $O/prof_oz.ncu-rep > $O/oz_ncu_summary.txt 2>&1; head -80 $O/oz_ncu_summary.txt
