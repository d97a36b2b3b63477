#!/bin/bash
# Measurement visit on ONE GPU: bench lines (MLP and grid conf), launch list of one training step, ncu --set full
# captures of the tcgen05 kernels.  Everything lands in gpurun_out/${TAG}_*; summaries are copied to profiles/ by hand.
TAG=${1:-r1v}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.log | head -c 300; echo
python bench.py --config grid --rays 32768 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_grid.log 2>&1; echo "bench grid rc=$?"
python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/${TAG}_prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_bf16.csv \
    python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/${TAG}_ncu.log 2>&1
for spec in "k_tc_wgrad:5:12" "EpiTan:2:1" "EpiFwdAct:2:1" "EpiBwd<:2:1" "EpiRev:2:1"; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  name=$(echo $k | tr -d '<')
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$k" -s $skip -c $cnt -f \
      -o gpurun_out/${TAG}_${name} python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name rc=$?"
done
ls -la gpurun_out/${TAG}_*.ncu-rep
