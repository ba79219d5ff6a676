#!/bin/bash
# Runs on a multi-GPU box (gpurun --gpus N): partition tests, the strong/weak bench lines under torchrun, one-process multi.
N=${1:-2}; TAG=${2:-m}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_detect.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err; echo "bench N=$N rc=$?"; tail -2 $OUT/bench_${N}gpu_$TAG.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/shard_parity_gpu.py 67 2>&1 | tail -2
python - <<PY
import json
d=json.load(open("$OUT/bench_${N}gpu_$TAG.json"))
print("N=$N weak value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "nomask", round(d["e2e"]["without_mask_download"]))
print("strong", d["strong"])
PY
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, ".")
import rmcv_b200 as rb
from rmcv_b200 import synth
B=1024
frames0=np.stack([synth.make_frame(s,1280,1024,synth.plates_for_seed(s)) for s in range(32)])
c=rb.Context(max_batch=1)
pin=c.pinned((B,1024,1280,3)); pin.array[:]=np.tile(frames0,(B//32,1,1,1))
pm=c.pinned((B,1024,1280))
prm=rb.default_params()
for devs in ([0], None):
    with rb.MultiContext(devs, max_batch=B) as m:
        for _ in range(2): m.detect_batch_host(pin.array, prm, pm.array)
        t=time.perf_counter()
        for _ in range(5): r=m.detect_batch_host(pin.array, prm, pm.array)
        dt=(time.perf_counter()-t)/5
        print("one process, %d device(s): %.0f frames/s e2e (1024 host frames in, masks + records out), %d armours" % (m.n_devices, B/dt, r.total_armours))
PY
