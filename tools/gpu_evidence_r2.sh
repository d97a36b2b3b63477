#!/bin/bash
# Round-2 evidence visit: launch lists (MLP and grid conf), ncu --set full of the backward sweep kernel, driver-shaped bench of both arms.
mkdir -p gpurun_out
python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/ev_ps_mlp.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ev_launches_mlp.csv \
    python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/ev_ncu_mlp.log 2>&1; echo "launch list mlp rc=$?"
python tools/profile_step.py --rays 32768 --precision bf16 --config grid > gpurun_out/ev_ps_grid.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ev_launches_grid.csv \
    python tools/profile_step.py --rays 32768 --precision bf16 --config grid > gpurun_out/ev_ncu_grid.log 2>&1; echo "launch list grid rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:EpiBwdS -s 10 -c 1 -f \
    -o gpurun_out/ev_EpiBwdS python tools/profile_step.py --rays 32768 --precision bf16 > gpurun_out/ev_ncu_bwds.log 2>&1; echo "ncu BwdS rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/ev_bench_ref.json 2> gpurun_out/ev_bench_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err; echo "bench rc=$?"
