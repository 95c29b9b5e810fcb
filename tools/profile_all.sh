# ncu captures of one launch of every workload's kernel at the full path count, plus the launch list of bench.py.
# Run on the GPU box: /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/profile_all.sh'; read the reports here
# with tools/ncu_summary.py.
set -x
TAG=${1:-prof}
for w in vanilla_f64_2p32 vanilla_f32_2p32 basket10_f64_2p28 cva50_f64_2p26 basket64_f32_2p30; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"accumulate" -c 1 -f -o gpurun_out/${TAG}_$w python bench.py --workload $w --also none --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_$w.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep
