#!/bin/bash
# ncu --set full captures of the tcgen05 sweep kernels inside one training step (32768 rays: 262144-row chunks).
TAG=${1:-r1x}
for spec in "EpiTan:2:1" "EpiFwdAct:2:1" "EpiBwd<:2:1" "EpiRev:2:1"; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  name=$(echo $k | tr -d '<')
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"$k" -s $skip -c $cnt -f \
      -o gpurun_out/${TAG}_${name} python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name rc=$?"
done
ls -la gpurun_out/${TAG}_*.ncu-rep
