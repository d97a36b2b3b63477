#!/bin/bash
# ncu --set full captures of the round-2 backward-side kernels inside one training step (32768 rays: 262144-row chunks).
TAG=${1:-r2m}
python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
for spec in "k_tc_wgrad:60:2" "EpiBwdS:10:1" "EpiTanS:10:1" "EpiRevS:10:1"; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"$k" -s $skip -c $cnt -f \
      -o gpurun_out/${TAG}_${k} python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/${TAG}_ncu_${k}.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/${TAG}_*.ncu-rep
