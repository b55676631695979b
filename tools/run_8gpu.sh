#!/bin/bash
# One 8-GPU session: box layout, host-to-device ceiling at 1/4/8 ranks, NCCL sharded parity, bench at 8 GPUs, sweep.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
nvidia-smi topo -m > $O/r02_topo_8gpu.txt 2>&1
(numactl -H; lscpu | grep -i -E 'numa|socket|model name|^CPU\(s\)'; grep -E 'MemTotal|MemAvailable' /proc/meminfo) >> $O/r02_topo_8gpu.txt 2>&1
for n in 8 4 2 1; do
  $TR --nproc-per-node $n --master-port 2951$n tools/h2d_probe.py 2>$O/probe$n.err | grep '"probe": "h2d"' > $O/r02_h2d_probe_${n}gpu.jsonl
done
python -m pytest tests/test_gpu_sharded.py -m gpu -q -k nccl 2>&1 | tail -3 > $O/r02_sharded_nccl_8gpu.txt
$TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 5 --warmup 3 --no-variants 2>$O/bench8.err | tail -1 > $O/r02_bench_8gpu.json
$TR --nproc-per-node 8 --master-port 29521 tools/sweep.py 2>$O/sweep8.err > $O/r02_sweep_8gpu.md
cat $O/r02_h2d_probe_*gpu.jsonl; cat $O/r02_sharded_nccl_8gpu.txt; python - <<PY
import json
d=json.load(open("$O/r02_bench_8gpu.json"))
for k in ("value","ms_per_step","e2e","e2e_u16","strong"): print(k, d[k])
PY
tail -5 $O/r02_sweep_8gpu.md
