#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench (plain), then the ncu launch list of the same bench command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["step_launch"], d["kernels"], d["clocks"])
PY
if [ "$1" = "ncu" ]; then
  # only this library's kernels: the synthetic-scene generation in front of the bench is thousands of torch launches
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      --kernel-name 'regex:mlp_fused|wgrad_kernel|head_grad|pack_|adam_|composite_|train_batch|sample_z|raygen_|draw_pixels|mse_loss' \
      python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -c 300 gpurun_out/ncu_bench.log
fi
if [ "$2" = "full" ]; then
  # one full-section capture of each hot kernel (inference fwd, training fwd, dgrad, wgrad, head-grad) at the C2 shape
  python tools/profile_kernels.py > gpurun_out/profile_plain.log 2>&1 && \
  ncu --set full --import-source on --clock-control none --kernel-name 'regex:mlp_fused|wgrad_kernel|head_grad' \
      --launch-skip 10 -c 5 -f -o gpurun_out/prof_r1f python tools/profile_kernels.py > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
