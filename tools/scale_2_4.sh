set -x
run() { n=$1; tag=$2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --combine peer > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  tail -c 300 gpurun_out/scale_$tag.err | tail -2
}
run 4 4
run 2 2
python - <<'PY'
import json
for tag in ("2", "4"):
    d = json.loads([l for l in open(f"gpurun_out/scale_{tag}.json") if l.startswith("{")][-1])
    print(tag, d["n_gpus"], "|", d["metric"], "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], d["price"])
    for k, v in d["also"].items():
        print("    ", k, "%.4g" % v["value"], "%.3f ms" % v["ms_per_step"], v["price"])
PY
