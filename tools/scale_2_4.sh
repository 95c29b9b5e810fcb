# multi-GPU check on a 2- (or 4-) GPU box: real-rank bit-identity tests, then bench.py under torchrun in the combine modes.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/scale_2_4.sh 2'
set -x
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -5
run() { # tag extra...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale${N}_$tag.json 2> gpurun_out/scale${N}_$tag.err
  tail -c 300 gpurun_out/scale${N}_$tag.err | tail -2
}
run push
run wait --peer-mode wait
run nccl --combine nccl
run push_nooverlap --no-overlap
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/scale${N}_one.json 2> gpurun_out/scale${N}_one.err
python - <<PY
import json
rows = {}
for tag in ("one", "push", "wait", "nccl", "push_nooverlap"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/scale${N}_{tag}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(tag, "FAILED", e); continue
    rows[tag] = {d["config"]["workload"]: (d["ms_per_step"], d["price"], d["limbs"], d["e2e"]["ms_per_call"]), **{k: (v["ms_per_step"], v["price"], v.get("limbs"), v.get("e2e_ms")) for k, v in d["also"].items() if "ms_per_step" in v}}
for w in rows.get("one", {}):
    t1 = rows["one"][w][0]
    print(f"{w:20s} 1 GPU {t1:9.4f} ms |", "  ".join(f"{tag}: {rows[tag][w][0]:8.4f} ms x{t1 / rows[tag][w][0]:.3f} e2e {rows[tag][w][3]:.4f}" for tag in rows if tag != "one" and w in rows[tag]),
          "| bits", "same" if all(rows[tag][w][2] == rows["one"][w][2] for tag in rows if w in rows[tag]) else "DIFFER")
PY
