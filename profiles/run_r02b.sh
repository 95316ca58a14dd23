# round 2, session b: loss read-back after backward + programmatic dependent launch (A/B), tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r02b_pytest.log
CNX_PDL=0 python bench.py --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02b_kernels_nopdl.json > gpurun_out/r02b_bench_nopdl.json 2> gpurun_out/r02b_bench_nopdl.err; echo "bench nopdl rc=$?"
python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02b_kernels_pdl.json > gpurun_out/r02b_bench_pdl.json 2> gpurun_out/r02b_bench_pdl.err; echo "bench pdl rc=$?"
python - <<'PY'
import json
for t in ("nopdl","pdl"):
    try:
        d=json.loads(open(f"gpurun_out/r02b_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["cnx_kernels_ms_per_step"], d.get("variants"))
    except Exception as e:
        print(t, "failed", e); print(open(f"gpurun_out/r02b_bench_{t}.err").read()[-2000:])
PY
