# pipeline experiments with the alternative BGR pixel kernel (run on the GPU box): bench.py short runs
run() { tag=$1; shift; env "$@" python bench.py --steps 6 --warmup 3 --batch 1024 --no-cpu --no-extras --e2e-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, round(d['roofline']['frac'],3))"; }
run band A=1
run bandstrip RMCV_BGR_STRIP=1
run bandstrip_rc4 RMCV_BGR_STRIP=1 RMCV_BANDSTRIP_RC=4
