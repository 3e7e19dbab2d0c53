python - <<'PY'
from lio_mapping_b200 import estimator
for n in (1<<24, 1<<21, 1<<20, 1<<19, 1<<17):
    r = estimator.asm_stream_bench(n, 30)
    print(n, n*32/2**20, "MB", {k: (round(float(v),4) if k!='launches' else v) for k,v in r.items()})
PY
