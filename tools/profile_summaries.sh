# gpurun_out/<tag>_<workload>.ncu-rep -> profiles/<tag>_<workload>.txt + profiles/kernel_work.json (see profile_all.sh)
set -e
TAG=${1:-prof}
for w in vanilla_f64_2p32 vanilla_f32_2p32 basket10_f64_2p28 cva50_f64_2p26 basket64_f32_2p30; do
  python tools/ncu_summary.py gpurun_out/${TAG}_$w.ncu-rep --work $w --manifest gpurun_out/${TAG}_manifest.json --capture profiles/${TAG}_$w.txt > profiles/${TAG}_$w.txt
done
cp gpurun_out/${TAG}_launches_bench.csv profiles/${TAG}_launches_bench.csv
