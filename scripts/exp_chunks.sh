# experiments: scratch slots x chunk size x pixel kernel (run on the GPU box)
run() { tag=$1; shift; args=$1; shift; env "$@" python bench.py --steps 8 --warmup 3 --batch 1024 --no-cpu --no-extras --e2e-steps 1 $args 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value']), {k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"; }
for s in 4 6; do for c in 128 256; do run band_s${s}_c$c "--chunk $c" RMCV_SLOTS=$s; run bstrip_s${s}_c$c "--chunk $c" RMCV_SLOTS=$s RMCV_BGR_STRIP=1; done; done
run bstrip_s3_c128 "--chunk 128" RMCV_BGR_STRIP=1
run bstrip_s8_c128 "--chunk 128" RMCV_SLOTS=8 RMCV_BGR_STRIP=1
run band_s8_c128 "--chunk 128" RMCV_SLOTS=8
