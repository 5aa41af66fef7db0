python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/r2x_bench8.json 2> gpurun_out/r2x_bench8.err
(numactl -H; nvidia-smi topo -m; lscpu | head -30) > gpurun_out/r2x_numa.log 2>&1
