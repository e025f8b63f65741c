#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE GPU; outputs land in gpurun_out/).
#   1. launch list of one short bench run (device time per launch, cold-cache / serialised: compare SHARES)
#   2. --set full captures of the prefill GEMMs (one layer: qkv, o, gate/up, down), the tcgen05 prompt attention,
#      the prefill RMSNorm, the decode GEMV and the decode attention
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --gen 4 --sections none --no-cpu-baseline --traffic off"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $OUT/r02_ncu_launches.csv $CMD > $OUT/r02_ncu_launches.out 2>&1
FULL="$NCU --set full --import-source on"
# second prefill of the run (the first is the warm-up round): 4 GEMMs x 32 layers = 128 launches per prefill
$FULL -k regex:gemm_tc_kernel -s 132 -c 4 -o $OUT/r02_prof_gemm -f $CMD > $OUT/r02_prof_gemm.out 2>&1
$FULL -k regex:attn_prefill_tc_kernel -s 40 -c 2 -o $OUT/r02_prof_attn_prefill -f $CMD > $OUT/r02_prof_attn_prefill.out 2>&1
$FULL -k regex:rmsnorm_vec_kernel -s 70 -c 2 -o $OUT/r02_prof_rmsnorm -f $CMD > $OUT/r02_prof_rmsnorm.out 2>&1
$FULL -k regex:gemv_mma_kernel -s 140 -c 5 -o $OUT/r02_prof_gemv -f $CMD > $OUT/r02_prof_gemv.out 2>&1
$FULL -k regex:attn_decode_mma_kernel -s 40 -c 2 -o $OUT/r02_prof_attn_decode -f $CMD > $OUT/r02_prof_attn_decode.out 2>&1
for f in gemm attn_prefill rmsnorm gemv attn_decode; do
  ncu -i $OUT/r02_prof_$f.ncu-rep --page raw --csv > $OUT/r02_prof_$f.raw.csv 2>/dev/null
done
ls -la $OUT | grep r02_prof | head -20
