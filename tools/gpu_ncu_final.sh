for spec in "EpiTan:2:1" "EpiRev:2:1"; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"$k" -s $skip -c $cnt -f \
      -o gpurun_out/r2q_${k} python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/r2q_ncu_${k}.log 2>&1
  echo "ncu $k rc=$?"
done
