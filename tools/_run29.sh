python tools/time_e2e_dev.py > gpurun_out/r2ac_e2e.log 2>&1
for mb in 8 32; do LDPC_GPU_CHUNK_MB=$mb python tools/time_e2e_dev.py >> gpurun_out/r2ac_e2e.log 2>&1; done
