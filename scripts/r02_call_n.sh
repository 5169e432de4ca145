#!/bin/bash
# multi-GPU call: scripts/r02_call_n.sh <N>   (run under gpurun --gpus N)
N=$1; cd "$(dirname "$0")/.."; O=gpurun_out/r02_n$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
(timeout 900 python -m pytest tests/test_sharded_nccl_gpu.py -m gpu -x -q 2>&1 | tail -6) > $O/pytest_nccl.txt
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 3 > $O/bench.json 2> $O/bench.err) 2> $O/bench_time.txt
cat $O/pytest_nccl.txt $O/bench_time.txt; tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "kernels", d["roofline"]["kernel_ms"])
print(json.dumps(d["configs"], indent=1)[:3000])
PY
grep -i "NCCL INFO.*nranks\|comm.*nranks\|Init COMPLETE" $O/bench.err | head -3
