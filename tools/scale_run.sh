#!/bin/bash
# tools/scale_run.sh N OUTDIR : the data-parallel measurements of one box size (run under `gpurun --gpus N`)
N=$1; OUT=$2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for c in mim_club clear64; do
  timeout -s KILL 240 $TR bench.py --gpus $N --steps 30 --config $c > $OUT/bench_${c}_${N}gpu.json 2> $OUT/bench_${c}_${N}gpu.err
done
timeout -s KILL 240 $TR tools/dp_check.py > $OUT/dp_check_${N}gpu.log 2>&1
timeout -s KILL 300 $TR tools/latent_sweep.py > $OUT/latent_sweep_${N}gpu.jsonl 2> $OUT/latent_sweep_${N}gpu.err
python - <<PY
import json
for n in ("mim_club","clear64"):
    try:
        d=json.loads([l for l in open("$OUT/bench_%s_${N}gpu.json"%n) if l.startswith("{")][-1])
        print(n, "N=$N value %.0f (%.3f ms) e2e %.0f (%.3f ms)"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["impl_config"]["collectives"], d["impl_config"]["peer_error"], d["impl_config"].get("peer_phases_rank0"))
    except Exception as e: print(n, "ERR", e)
PY
grep "trainer DP\|overlapped\|model DP\|peer kernels" $OUT/dp_check_${N}gpu.log | cut -c1-260
tail -3 $OUT/latent_sweep_${N}gpu.jsonl | cut -c1-250
