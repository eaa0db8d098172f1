#!/bin/bash
# 2-GPU visit: bench at N = 2 with the per-rank diagnostics (which GPU is the slow one, at which clock).
set -u
TAG=${1:-v2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"; echo "rc=$?"
python - "$OUT/bench_n2.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
print("N=%d value %.0f ms/step %.4f" % (d["n_gpus"], d["value"], d["ms_per_step"])); print("per_rank", d["per_rank"]); print("clocks", d["clocks"]); print("e2e", d["e2e"]["value"], d["e2e"]["regions"])
PY
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline --no-gnet 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 value %.0f ms/step %.4f per_rank %s'%(d['value'],d['ms_per_step'],d['per_rank']))"
