#!/bin/sh
# evaluate_d experiment: bench line + launch list of the Jacobian path (config 1d)
timeout 300 python bench.py --config 1d --no-cpu --no-e2e --steps 5 > gpurun_out/r2_1d_env.json 2> gpurun_out/r2_1d_env.err; tail -3 gpurun_out/r2_1d_env.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_1d_launches.csv python bench.py --config 1d --no-cpu --no-e2e --steps 1 --warmup 1 --envs 20000 > /dev/null 2>&1
if [ -n "$NCU_FULL" ]; then timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dB_env -c 1 -f -o gpurun_out/r2_dbenv python bench.py --config 1d --no-cpu --no-e2e --steps 1 --warmup 1 --envs 20000 > /dev/null 2>&1; fi
