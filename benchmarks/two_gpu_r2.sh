#!/bin/bash
# 2-GPU pass: multi-device test, set_devices e2e scaling, torchrun bench at N = 2 (both arms)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r2_topo.txt 2>&1
python -m pytest tests/test_multi_device.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/abh_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/abh_tests.log
python benchmarks/multi_device.py > $OUT/r2_multi_device.txt 2>&1; tail -4 $OUT/r2_multi_device.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $OUT/abh_n2.log 2>&1; tail -1 $OUT/abh_n2.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 3 --steps 5 --warmup 3 --no-cpu --no-e2e > $OUT/abh_c3_n2.log 2>&1; tail -1 $OUT/abh_c3_n2.log | cut -c1-400
