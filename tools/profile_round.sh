#!/bin/bash
# Round-2 profiling captures (run on the GPU box through gpurun; small outputs only under gpurun_out/: the .ncu-rep files
# are summarised with tools/ncu_summary.py and deleted — gpurun copies back at most 64 MiB).
#   1. ncu launch list of part of one bench step (kernel names + durations; cold-cache, serialised: compare SHARES)
#   2. ncu --set full of two consecutive DiT layers (qkv, attention, wo, cross.wq+attn, cross.wo, w13, w2) in ONE run
#   3. ncu --set full of the codec kernels (tools/codec_prof.py)
set -x
mkdir -p gpurun_out
export SAB_NO_GRAPH=1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --kernel-name-base mangled -k regex:"gemm_tc_kernelILi256|attention_tc2" \
  --launch-skip 3000 -c 16 -o gpurun_out/ncu_r2_layers $B > gpurun_out/ncu_r2_layers.log 2>&1
timeout 400 ncu --set full --clock-control none --kernel-name-base mangled -k regex:"gemm_tc_kernel|enc_conv0|dec_last" \
  --launch-skip 66 -c 66 -o gpurun_out/ncu_r2_codec python tools/codec_prof.py > gpurun_out/ncu_r2_codec.log 2>&1
for f in gpurun_out/ncu_r2_*.ncu-rep; do
  ncu -i $f --page raw --csv 2>/dev/null | python tools/ncu_summary.py > ${f%.ncu-rep}_summary.csv
  rm -f $f
done
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 57500 -c 2500 --csv \
  --log-file gpurun_out/launches_r2_step.csv $B > gpurun_out/launches_r2.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -12
