#!/bin/bash
# K concurrent single-GPU processes (K = 1, 2, 4, 8) running the three host-buffer calls of one C2 step: where does the
# per-step time go when ranks share the host?  usage (under gpurun --gpus 8): scripts/contention_probe.sh [out dir]
cd "$(dirname "$0")/.."; O=${1:-gpurun_out/contention}; mkdir -p $O
lscpu | grep -i "model name\|socket\|numa\|^cpu(s)\|l3\|l2" > $O/lscpu.txt
group() {   # group <tag> <K> [env...]
  local tag=$1 K=$2; shift 2
  local bd; bd=$(mktemp -d)
  for ((i = 0; i < K; i++)); do
    env "$@" PROBE_PCI=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $i) CUDA_VISIBLE_DEVICES=$i LOCAL_WORLD_SIZE=$K PROBE_WORLD=$K PROBE_BARRIER_DIR=$bd PROBE_REPS=25 \
      timeout 300 python scripts/e2e_probe2.py > $O/$tag.$i.txt 2>&1 &
  done
  wait
  echo "== $tag (K=$K) $*"; for ((i = 0; i < K; i++)); do tail -1 $O/$tag.$i.txt | cut -c1-330; done
  rm -rf $bd
}
group k1 1
group k2 2
group k4 4
group k8 8
group k8_ring1m 8 B200NB_STAGE_KB=1024 B200NB_STAGE_BLOCK_KB=64
group k8_ring2m 8 B200NB_STAGE_KB=2048 B200NB_STAGE_BLOCK_KB=128
group k8_thr4 8 B200NB_HOST_THREADS=4
group k8_thr16 8 B200NB_HOST_THREADS=16
group k8_nospec 8 B200NB_SPECULATE=0
group k8_chunks1 8 B200NB_CHUNKS=1
group k8_timing 8 B200NB_HOST_TIMING=1
grep "b200nb timing" $O/k8_timing.0.txt | tail -6
group k1_timing 1 B200NB_HOST_TIMING=1
grep "b200nb timing" $O/k1_timing.0.txt | tail -6
