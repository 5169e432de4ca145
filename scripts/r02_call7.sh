mkdir -p gpurun_out/r02_g; O=gpurun_out/r02_g
echo "== one gene per warp, product (32 lanes forced) vs TMA row prefetch (NB_EXP_TMA_ROWS)" > $O/tma_ab.txt
for shape in "50000 100" "20000 1000" "20000 500"; do set -- $shape
  B200NB_GROUP_LANES=32 scripts/microbench --genes $1 --samples $2 deseq2_b200/libb200nb.so deseq2_b200/libb200nb_exp_tma.so >> $O/tma_ab.txt 2>&1
done
echo "== default product (grouped kernels where they fit)" >> $O/tma_ab.txt
scripts/microbench --genes 50000 --samples 100 deseq2_b200/libb200nb.so >> $O/tma_ab.txt 2>&1
timeout 600 python scripts/c4_ab.py 20000 > $O/c4_ab.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > $O/b_ncu.log 2>&1
scripts/ncu_capture.sh r02g_fit_disp fit_disp_grp_kernel > $O/ncu_disp.txt 2>&1
cat $O/tma_ab.txt $O/c4_ab.txt; tail -5 $O/ncu_disp.txt | cut -c1-200
