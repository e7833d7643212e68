#!/bin/bash
# 8-GPU pass: config 3 (north_star's 8-GPU configuration) and config 2 under torchrun, and one handle driving 8 GPUs
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r2_topo_8gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config 3 --steps 5 --warmup 3 --no-cpu > $OUT/r2_bench_c3_n8.json 2> $OUT/r2_bench_c3_n8.err; tail -c 300 $OUT/r2_bench_c3_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > $OUT/r2_bench_n8.json 2> $OUT/r2_bench_n8.err; tail -c 300 $OUT/r2_bench_n8.json
python benchmarks/multi_device.py --envs 500000 > $OUT/r2_multi_device_8.txt 2>&1; tail -2 $OUT/r2_multi_device_8.txt | cut -c1-900
