#!/bin/bash
# BASELINE.json configs[1]: IEEE 802.3an (2048,1723), normalised / offset min-sum and sum-product, Eb/N0 sweep
# 3-5 dB, 1e8 frames per point on one B200, through the reference's own command lines (bin/decode*).
# Usage: tools/ber_sweep.sh OUT_DIR [FRAMES]      (run from the repository root on a GPU box)
set -e
OUT=${1:-gpurun_out/ber_sweep}; FRAMES=${2:-100000000}
mkdir -p "$OUT"; rm -f "$OUT"/*.tsv "$OUT"/final.log
H=codes/802_3/802_3_H.alist
export LDPC_SEED=20251018 LDPC_PRECISION=f32
for snr in 3.0 3.25 3.5 3.75 4.0 4.25 4.5 4.75 5.0; do
    LDPC_FRAMES=$FRAMES ./bin/decodeNormalizedMinSum $H 0.8413 $snr 10 2.0 6 1.25 "$OUT/nms_f32.tsv" | grep "Final result" | sed "s/^/nms_f32 $snr /" >> "$OUT/final.log"
    LDPC_FRAMES=$FRAMES ./bin/decodeOffsetMinSum     $H 0.8413 $snr 10 1.9375 5 0.125 "$OUT/oms_f32.tsv" | grep "Final result" | sed "s/^/oms_f32 $snr /" >> "$OUT/final.log"
    LDPC_FRAMES=$((FRAMES / 4)) ./bin/decodeBP       $H 0.8413 $snr 10 "$OUT/bp_f32.tsv" | grep "Final result" | sed "s/^/bp_f32 $snr /" >> "$OUT/final.log"
    LDPC_PRECISION=f64 LDPC_FRAMES=$((FRAMES / 20)) ./bin/decodeNormalizedMinSum $H 0.8413 $snr 10 2.0 6 1.25 "$OUT/nms_f64.tsv" | grep "Final result" | sed "s/^/nms_f64 $snr /" >> "$OUT/final.log"
done
cat "$OUT/final.log"
