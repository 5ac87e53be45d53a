#!/bin/bash
# ncu evidence for one bench step: launch list of the step kernels + full captures of the fused kernel (1 and 3 products)
# and of the fit kernels
O=${1:-gpurun_out/r01n}; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused|band_|kstar|rs_|acq_kernel|argmax|gather|contract" -c 80 --csv --log-file $O/launches_step.csv \
  python bench.py --steps 2 --warmup 2 --m-per-gpu 303104 --no-cpu-baseline > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
for P in 1 3; do
B200BO_FAST_PRODUCTS=$P timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_fused -s 1 -c 1 -o $O/prof_pair_p$P \
  python bench.py --steps 1 --warmup 1 --m-per-gpu 37888 --no-cpu-baseline > $O/ncu_full_p$P.log 2>&1; echo "ncu full p$P rc=$?"
done
timeout 600 ncu --set full --clock-control none -k regex:"kmat_assemble|chol_diag|dgemm|llf_grad|linv_split" -c 10 -o $O/prof_fit \
  python bench.py --steps 1 --warmup 1 --m-per-gpu 18944 --no-cpu-baseline > $O/ncu_fit.log 2>&1; echo "ncu fit rc=$?"
