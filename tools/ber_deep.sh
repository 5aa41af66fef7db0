#!/bin/bash
set -e
mkdir -p gpurun_out/ber_deep; rm -f gpurun_out/ber_deep/*
H=codes/802_3/802_3_H.alist
export LDPC_SEED=20251019
run() { # name precision snr frames
  LDPC_PRECISION=$2 LDPC_FRAMES=$4 ./bin/decodeNormalizedMinSum $H 0.8413 $3 10 2.0 6 1.25 gpurun_out/ber_deep/$1.tsv | grep "Final result" | sed "s/^/$1 $3 /" >> gpurun_out/ber_deep/final.log
}
run nms_f64 f64 4.5 100000000
run nms_f32 f32 4.5 1000000000
run nms_f32 f32 4.75 1000000000
run nms_f64 f64 4.75 200000000
cat gpurun_out/ber_deep/final.log; cat gpurun_out/ber_deep/*.tsv
