set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload ngdbfhw_8023 --no-extras > gpurun_out/r2l_bench_2gpu_hw.json 2> gpurun_out/r2l_bench_2gpu_hw.err
