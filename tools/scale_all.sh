# 1/2/4/8-GPU strong-scaling evidence on one box: bench.py under torchrun, every workload, peer-memory combine
# (plus the NCCL route at 8 GPUs for comparison).  Writes gpurun_out/scale_*.json.
set -x
run() { # n combine tag extra...
  n=$1; c=$2; tag=$3; shift 3
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --combine $c "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  fi
  tail -c 300 gpurun_out/scale_$tag.err | tail -2
}
run 8 peer 8
if [ -z "$QUICK" ]; then  # QUICK=1: only the 8-GPU and the 1-GPU run (8x box time is charged)
run 8 nccl 8_nccl
run 4 peer 4
run 2 peer 2
fi
run 1 peer 1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi8.log 2>&1; tail -3 gpurun_out/pytest_multi8.log
python - <<'PY'
import json
for tag in ("1", "2", "4", "8", "8_nccl"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/scale_{tag}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(tag, "FAILED", e); continue
    print(tag, d["n_gpus"], d["config"]["collective"][:40], "|", d["metric"], "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], d["price"])
    for k, v in d["also"].items():
        print("    ", k, "%.4g" % v["value"], "%.3f ms" % v["ms_per_step"], v["price"])
PY
