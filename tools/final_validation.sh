#!/bin/bash
# tools/final_validation.sh OUTDIR : the single-GPU artefacts behind DESIGN.md / profiles (run under gpurun)
OUT=${1:-gpurun_out/final}; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/gpu_tests.log 2>&1; tail -1 $OUT/gpu_tests.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 300 $OUT/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_mim_club.csv python tools/one_step.py > $OUT/launches.log 2>&1
timeout 120 python tools/latent_timeline.py 16384 8 > $OUT/latent_bwd_timeline.txt 2>&1
timeout 120 python tools/latent_timeline.py 16384 32 >> $OUT/latent_bwd_timeline.txt 2>&1
timeout 120 python tools/persist_timeline.py > $OUT/conv_tma_timeline.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:snn_bwd_tc -c 1 -o $OUT/latent_bwd_tc_d8 -f python tools/latent_once.py 16384 8 > $OUT/ncu_latent.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_default.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["gpu_launches_per_step"], d["parity"].get("ok") if isinstance(d.get("parity"), dict) else d.get("parity"))
l=d["latent_roofline"]; print("latent", l["forward"], l["backward"], l["d32"])
for c in d.get("configs", []): print(c.get("config",{}).get("baseline_config"), c.get("value"), c.get("ms_per_step"), (c.get("eager_gpu_baseline") or {}).get("value"))
PY
