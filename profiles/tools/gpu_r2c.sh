#!/bin/bash
# GPU box: panel LU vs rank-1 LU (bitwise test + timing A/B), guard-band tests, config-4 test
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -x -k "panel_lu or guard_bands or config4 or lu_solve or stepwise" > gpurun_out/pytest_sel2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sel2.log
python bench.py --no-cpu --no-other --steps 5 > gpurun_out/bench_lu_panel.json 2> gpurun_out/bench_lu_panel.err
HPF_LU_CLASSIC=1 python bench.py --no-cpu --no-other --steps 5 > gpurun_out/bench_lu_classic.json 2> gpurun_out/bench_lu_classic.err
tail -12 gpurun_out/pytest_sel2.log
python - <<'PY'
import json
for n in ("panel","classic"):
    try:
        d=json.load(open("gpurun_out/bench_lu_%s.json"%n))
        print(n, "lu_solve_kernel", d["roofline_kernels"][2]["ms"], d["roofline_kernels"][2]["frac"], "dense path", d["dense_lu_path"]["ms_per_step"], d["dense_lu_path"]["roofline"]["frac"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/bench_lu_panel.err
