#!/bin/bash
# A/B of attention library variants built with tools/build_variant.py: tools/ab_attn_variants.sh base tma15 ...
for v in "$@"; do
  CRYOVIT_B200_LIB=$PWD/tools/_variants/lib_$v.so timeout 120 python tools/attn_ab.py 1.0 fp16 2>&1 | tail -1
done
for v in "$@"; do
  CRYOVIT_B200_LIB=$PWD/tools/_variants/lib_$v.so timeout 120 python tools/attn_ab.py 1.0 fp16 2>&1 | tail -1
done
