#!/bin/bash
# A/B of library variants on one box: per-kernel timing probe for the default build and every sfm-nerf_b200/variants/*.so
mkdir -p gpurun_out
python tools/perf_probe.py > gpurun_out/probe_default.json 2> gpurun_out/probe_default.err; python - <<'PY'
import json; d=json.load(open("gpurun_out/probe_default.json")); print("default", {k: round(v, 4) for k, v in d.items() if k.endswith("_ms")})
PY
for f in sfm-nerf_b200/variants/*.so; do
  n=$(basename $f .so)
  NERF_B200_LIB=$PWD/$f python tools/perf_probe.py > gpurun_out/probe_$n.json 2> gpurun_out/probe_$n.err
  N=$n python - <<'PY'
import json, os; n=os.environ["N"]; d=json.load(open(f"gpurun_out/probe_{n}.json")); print(n, {k: round(v, 4) for k, v in d.items() if k.endswith("_ms")})
PY
done
