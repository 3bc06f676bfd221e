mkdir -p gpurun_out/r02k
python -m pytest tests -m gpu -q > gpurun_out/r02k/pytest_gpu.log 2>&1; tail -12 gpurun_out/r02k/pytest_gpu.log
BLCD_PIPELINE=1 python -m pytest tests -m gpu -q > gpurun_out/r02k/pytest_gpu_pipeline.log 2>&1; tail -8 gpurun_out/r02k/pytest_gpu_pipeline.log
python tools/pipe_time.py Urchin 262144 20 2>&1 | tail -1
python bench.py --steps 2 --warmup 3 > gpurun_out/r02k/bench.json 2> gpurun_out/r02k/bench.err; tail -3 gpurun_out/r02k/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02k/bench.json"))
print({k:d[k] for k in ["value","ms_per_step","gpu_launches"]}, d["e2e"]["value"], d["e2e"]["pipelined_value"], d["roofline_solver"].get("frac"), d["roofline_solver"].get("counters",{}).get("active_lanes_per_warp_inst"), d["roofline"]["traffic"], d["render_roofline"])
PY
