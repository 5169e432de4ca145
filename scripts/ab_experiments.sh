#!/bin/bash
# Round-2 A/B of the kernel experiments that were written (and parity-checked under the SIMT emulator) without GPU
# access.  Step 1, on the build machine (no GPU needed):   scripts/ab_experiments.sh build
# Step 2, on the GPU box (one gpurun call):                scripts/ab_experiments.sh quick > gpurun_out/ab_quick.txt
#                                                          scripts/ab_experiments.sh run > gpurun_out/ab.txt
# Each line of the output is the bench.py JSON line of one library; compare "value" and config.kernel_ms.
set -e
cd "$(dirname "$0")/.."
declare -A EXP=(
  [estrin_rcp3]="-DNB_EXP_LOG_ESTRIN -DNB_EXP_RCP_CUBIC"
  [estrin]="-DNB_EXP_LOG_ESTRIN"
  [rcp3]="-DNB_EXP_RCP_CUBIC"
  [split]="-DNB_EXP_SPLIT_MODES"
  [split_estrin_rcp3]="-DNB_EXP_SPLIT_MODES -DNB_EXP_LOG_ESTRIN -DNB_EXP_RCP_CUBIC"
  [lb3]="-DNB_LB_THREADS=256 -DNB_LB_CTAS=3"
  [beta3]="-DNB_EXP_BETA_CTAS3"
  [tab4]="-DNB_EXP_TAB_UNROLL4"
  [heavy_first]="-DNB_EXP_HEAVY_FIRST"
  [lfact]="-DNB_EXP_LFACT_TABLE"
  [half_warp]="-DNB_EXP_HALF_WARP"
  [half_warp_buckets]="-DNB_EXP_HALF_WARP -DNB_EXP_TAB_BUCKETS"
  [buckets]="-DNB_EXP_TAB_BUCKETS"
  [half_warp_buckets_estrin_rcp3]="-DNB_EXP_HALF_WARP -DNB_EXP_TAB_BUCKETS -DNB_EXP_LOG_ESTRIN -DNB_EXP_RCP_CUBIC"
  [all]="-DNB_EXP_SPLIT_MODES -DNB_EXP_LOG_ESTRIN -DNB_EXP_RCP_CUBIC -DNB_EXP_BETA_CTAS3 -DNB_EXP_TAB_UNROLL4 -DNB_EXP_LFACT_TABLE"
)
if [ "$1" = build ]; then
  make -C deseq2_b200/csrc -s
  for name in "${!EXP[@]}"; do
    make -C deseq2_b200/csrc exp EXPFLAGS="${EXP[$name]}" EXPNAME="$name" -s
  done
  nvcc -O2 -std=c++17 -Wno-deprecated-gpu-targets -o scripts/microbench scripts/microbench.cu -ldl
  ls -la deseq2_b200/libb200nb*.so scripts/microbench
elif [ "$1" = quick ]; then
  # seconds instead of minutes: kernel times of every library on one resident workload, no Python (scripts/microbench.cu)
  scripts/microbench --genes 50000 --samples 100 deseq2_b200/libb200nb.so deseq2_b200/libb200nb_exp_*.so
  for gl in 8 16; do
    echo "== B200NB_GROUP_LANES=$gl (libraries built with NB_EXP_HALF_WARP)"
    B200NB_GROUP_LANES=$gl scripts/microbench --genes 50000 --samples 100 deseq2_b200/libb200nb.so deseq2_b200/libb200nb_exp_half_warp*.so
  done
  echo "== 20 000 genes x 12 samples (a typical small experiment)"
  scripts/microbench --genes 20000 --samples 12 deseq2_b200/libb200nb.so deseq2_b200/libb200nb_exp_half_warp*.so
elif [ "$1" = host ]; then
  # the end-to-end path (R-layout pageable buffers through the host entry points) under the run-time knobs of capi.cu;
  # one process per setting because the knobs are read once; `B200NB_HOST_TIMING=1` adds the per-phase breakdown
  mb="scripts/microbench --host --genes 50000 --samples 100 --reps 9 deseq2_b200/libb200nb.so"
  $mb
  B200NB_HOST_TIMING=1 scripts/microbench --host --genes 50000 --samples 100 --reps 1 deseq2_b200/libb200nb.so 2>&1 | tail -4
  for s in "B200NB_DETECT_SF=1" "B200NB_D2H_POPULATE=1" "B200NB_D2H_HUGEPAGE=1" "B200NB_D2H_THREADS=16" "B200NB_D2H_THREADS=32" \
           "B200NB_STAGE_THREADS=16" "B200NB_STAGE_CHUNK_MB=4" "B200NB_STAGE_CHUNK_MB=32" \
           "B200NB_CHUNK_GENES=12500 B200NB_CHUNK_WORKERS=2" "B200NB_CHUNK_GENES=12500 B200NB_CHUNK_WORKERS=3" \
           "B200NB_CHUNK_GENES=6250 B200NB_CHUNK_WORKERS=3" "B200NB_CHUNK_GENES=25000 B200NB_CHUNK_WORKERS=2" \
           "B200NB_CHUNK_GENES=12500 B200NB_CHUNK_WORKERS=3 B200NB_DETECT_SF=1 B200NB_D2H_POPULATE=1" \
           "B200NB_CHUNK_GENES=12500 B200NB_CHUNK_WORKERS=3 B200NB_DETECT_SF=1 B200NB_D2H_POPULATE=1 B200NB_D2H_THREADS=16"; do
    echo "$s | $(env $s $mb | tail -1)"
  done
elif [ "$1" = run ]; then
  echo "default $(python bench.py --no-e2e --no-cpu-baseline --steps 30 | tail -1)"
  for name in "${!EXP[@]}"; do
    lib="$PWD/deseq2_b200/libb200nb_exp_$name.so"
    [ -f "$lib" ] || continue
    # correctness first: the parity suite must stay green with the experiment library
    if B200NB_LIB="$lib" python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "mle_parity or map_parity or fit_beta_parity or big_and_mixed" > /dev/null 2>&1; then ok=parity-ok; else ok=PARITY-FAILED; fi
    echo "$name $ok $(B200NB_LIB="$lib" python bench.py --no-e2e --no-cpu-baseline --steps 30 | tail -1)"
    if [[ "$name" == half_warp* ]]; then   # the grouped kernels pick 8 lanes for m <= 40 and 16 above; force each width
      for gl in 8 16; do
        echo "$name lanes=$gl $(B200NB_GROUP_LANES=$gl B200NB_LIB="$lib" python bench.py --no-e2e --no-cpu-baseline --steps 30 | tail -1)"
      done
      # one kernel grouped, the other as in the product (config.kernel_ms tells the two apart anyway)
      echo "$name disp-only $(B200NB_GROUP_LANES_BETA=32 B200NB_LIB="$lib" python bench.py --no-e2e --no-cpu-baseline --steps 30 | tail -1)"
      echo "$name beta-only $(B200NB_GROUP_LANES_DISP=32 B200NB_LIB="$lib" python bench.py --no-e2e --no-cpu-baseline --steps 30 | tail -1)"
    fi
  done
else
  echo "usage: $0 build|quick|host|run"; exit 2
fi
