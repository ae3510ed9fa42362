export DFM_KSLICE_S1=1
python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_ks1.json
python - <<PY
import json
j=json.load(open("gpurun_out/bench_ks1.json"))
print("ks1", j["value"], j["ms_per_step"], j["conv_ms_per_step"], j["e2e"]["value"])
for k,v in j["kernels"].items():
    if "64->64,s1" in k: print(k, round(v["ms"]/v["launches"],4))
PY
