#!/bin/bash
# 1-GPU call: general-p kernel variants (inline / unroll / register cap) on the config-4 shape + profiler view of DESeq_device
cd "$(dirname "$0")/.."; O=gpurun_out/r02_k; mkdir -p $O
for lib in libb200nb.so libb200nb_exp_gen_noinline.so libb200nb_exp_gen_noinline_u2.so libb200nb_exp_gen_u2.so libb200nb_exp_gen_u1.so libb200nb_exp_gen_u8.so libb200nb_exp_gen_mb2.so; do
  echo "== $lib"; B200NB_LIB=$PWD/deseq2_b200/$lib B200NB_GENERIC_ROWS=smem C4_AB_MODES=smem timeout 300 python scripts/c4_ab.py 20000 2>&1 | tail -1
done > $O/gen_variants.txt 2>&1
cat $O/gen_variants.txt
timeout 600 python scripts/c4_diag.py 50000 > $O/c4_diag.txt 2>&1; grep -v "^---" $O/c4_diag.txt | cut -c1-260 | head -90
