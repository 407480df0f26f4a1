#!/bin/bash
# A/B on one box: tap-packed last conv on / off, 4K frame, alternating
mkdir -p gpurun_out
F="--steps 40 --warmup 5 --no-8k --no-kernel-to-beat --no-e2e --no-cpu-baseline"
for i in 1 2; do
  OIDN_B200_NO_TAP_PACK=1 timeout 200 python bench.py $F 2>/dev/null | tail -1 > gpurun_out/ab_off_$i.json
  timeout 200 python bench.py $F 2>/dev/null | tail -1 > gpurun_out/ab_on_$i.json
done
python - <<EOF
import json
for n in ("off_1","on_1","off_2","on_2"):
  d=json.loads(open("gpurun_out/ab_%s.json"%n).read())
  r=d["roofline"]
  print(n, d["ms_per_step"], r["conv_ms_per_frame"], r["frac"], r["stamped_pass"], r["sustained"]["ms_per_step"], r["sustained"]["frac"], d["passes"]["conv_layers_ms"]["dec_conv0"])
EOF
