set -x
for w in vanilla_f64_2p32 vanilla_f32_2p32 basket10_f64_2p28 cva50_f64_2p26 basket64_f32_2p30; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"accumulate" -c 1 -f -o gpurun_out/prof7_$w python bench.py --workload $w --also none --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_$w.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01j.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -6
