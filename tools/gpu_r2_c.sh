#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn_variants.log; : > $L
for sc in 1.0 1.5; do
  ATTN_AB_SDPA=1 python tools/attn_ab.py $sc >> $L 2>&1
  CVIT_FA_EXACT=1 python tools/attn_ab.py $sc >> $L 2>&1
  CRYOVIT_B200_LIB=$PWD/tools/_variants/lib_early_sfree.so python tools/attn_ab.py $sc >> $L 2>&1
  CVIT_FA_EXACT=1 CRYOVIT_B200_LIB=$PWD/tools/_variants/lib_exact16.so python tools/attn_ab.py $sc >> $L 2>&1
done
cat $L
