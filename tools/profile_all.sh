# ncu captures of one launch of every workload's kernel at the full path count, plus the launch list of bench.py.
# Run on the GPU box: /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/profile_all.sh r02a'; then, HERE and
# without rebuilding in between, tools/profile_summaries.sh r02a turns the reports into profiles/<tag>_*.txt and
# profiles/kernel_work.json (the build manifest of the profiled library travels back next to the reports).
set -x
TAG=${1:-prof}
PIPES=sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_tensor.sum
cp montecarlocuda_b200/lib/libmcb200.manifest.json gpurun_out/${TAG}_manifest.json
for w in vanilla_f64_2p32 vanilla_f32_2p32 basket10_f64_2p28 cva50_f64_2p26 basket64_f32_2p30; do
  timeout 300 ncu --set full --metrics $PIPES --clock-control none --import-source on -k regex:"accumulate" -c 1 -f -o gpurun_out/${TAG}_$w python bench.py --workload $w --also none --no-cpu-baseline --no-gpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_$w.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep
