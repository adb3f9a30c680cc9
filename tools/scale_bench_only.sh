#!/bin/bash
# tools/scale_bench_only.sh N OUTDIR : bench.py for configs[1] and configs[3] on N GPUs (run under `gpurun --gpus N`)
N=$1; OUT=$2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for c in mim_club clear64; do
  if [ "$N" = "1" ]; then python bench.py --steps 30 --config $c --no-configs --no-latent > $OUT/bench_${c}_${N}gpu.json 2> $OUT/bench_${c}_${N}gpu.err
  else timeout -s KILL 240 $TR bench.py --gpus $N --steps 30 --config $c > $OUT/bench_${c}_${N}gpu.json 2> $OUT/bench_${c}_${N}gpu.err; fi
done
python - <<PY
import json
for n in ("mim_club","clear64"):
    try:
        d=json.loads([l for l in open("$OUT/bench_%s_${N}gpu.json"%n) if l.startswith("{")][-1])
        print(n, "N=$N value %.0f (%.3f ms) e2e %.0f (%.3f ms)"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["impl_config"]["collectives"], d["impl_config"]["peer_error"])
    except Exception as e: print(n, "ERR", e)
PY
