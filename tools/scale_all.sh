# 1/2/4/8-GPU strong-scaling evidence on ONE 8-GPU box: bench.py under torchrun, every workload, split-phase peer combine
# (plus the single-phase variant at 8 GPUs for comparison).  Writes gpurun_out/scale_*.json.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/scale_all.sh'
set -x
run() { # n tag extra...
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  fi
  tail -c 300 gpurun_out/scale_$tag.err | tail -2
}
run 8 8
if [ -z "$NOWAIT" ]; then run 8 8_wait --peer-mode wait; fi
if [ -z "$QUICK" ]; then
run 4 4
run 2 2
fi
run 1 1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi8.log 2>&1; tail -3 gpurun_out/pytest_multi8.log
python - <<'PY'
import json
rows = {}
for tag in ("1", "2", "4", "8", "8_wait"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/scale_{tag}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(tag, "FAILED", e); continue
    rows[tag] = {d["config"]["workload"]: (d["ms_per_step"], d["limbs"], d["e2e"]["ms_per_call"]), **{k: (v["ms_per_step"], v.get("limbs"), v.get("e2e_ms")) for k, v in d["also"].items() if "ms_per_step" in v}}
for w in rows.get("1", {}):
    t1 = rows["1"][w][0]
    print(f"{w:20s} 1 GPU {t1:9.4f} ms |", "  ".join(f"{tag}: {rows[tag][w][0]:8.4f} ms x{t1 / rows[tag][w][0]:.3f} (e2e {rows[tag][w][2]:.4f})" for tag in rows if tag != "1" and w in rows[tag]),
          "| bits", "same" if all(rows[tag][w][1] == rows["1"][w][1] for tag in rows if w in rows[tag]) else "DIFFER")
PY
