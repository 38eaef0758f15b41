#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/pytest_r3g.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r3g.log
timeout 600 python bench.py > gpurun_out/bench_r3g.json 2> gpurun_out/bench_r3g.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r3g.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_r3g.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step")}, r["roofline"]["frac"], r["e2e"]["ms_per_step"], r["e2e_full"], r["cg"]["pcg_iters_per_s"], r["parity_checked"])
PY
