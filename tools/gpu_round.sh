#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench (plain), then optionally the ncu launch list / full captures.
#   tools/gpu_round.sh [ncu] [full] [ref]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "timing", "e2e", "step_launch", "kernels", "roofline", "clocks", "train_c3", "render_c3", "cpu_baseline"):
    print(k, d.get(k))
PY
for a in "$@"; do
if [ "$a" = "ref" ]; then
  python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
fi
if [ "$a" = "ncu" ]; then
  # only this library's kernels: the synthetic-scene generation in front of the bench is thousands of torch launches
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
      --kernel-name 'regex:mlp_fused|wgrad|head_grad|pack_|adam_|composite_|train_batch|sample_z|raygen_|draw_pixels|mse_loss' \
      python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -c 300 gpurun_out/ncu_bench.log
fi
if [ "$a" = "full" ]; then
  # one full-section capture of each hot kernel (inference fwd, training fwd, dgrad, wgrad, head-grad) at the C2 shape
  python tools/profile_kernels.py > gpurun_out/profile_plain.log 2>&1 && \
  ncu --set full --import-source on --clock-control none --kernel-name 'regex:mlp_fused|wgrad' \
      --launch-skip 8 -c 4 -f -o gpurun_out/prof_r2 python tools/profile_kernels.py > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
done
