set -x
export DFM_KSLICE_S1=1
python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -4
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json
python - <<PY
import json
j=json.load(open("gpurun_out/bench_final.json"))
print("final", j["value"], j["ms_per_step"], j["conv_ms_per_step"], j["e2e"]["value"], j["gpu_launches"])
for k,v in j["kernels"].items():
    if "64->64,s1" in k: print(k, round(v["ms"]/v["launches"],4))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 480 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -4
