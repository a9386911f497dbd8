# The round's measurement records in one GPU call (outputs under gpurun_out/, copied to profiles/ afterwards).
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --workload gan784 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_gan784.json 2> gpurun_out/r02_bench_n1_gan784.err
python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err
python tools/timeline.py --workload lenet_b > gpurun_out/r02_timeline_lenet_b.txt 2>/dev/null
python tools/timeline.py --workload gan784 > gpurun_out/r02_timeline_gan784.txt 2>/dev/null
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
for w in lenet_b gan784; do
  python tools/one_step.py $w 1 && ncu --set full --clock-control none -o /tmp/r02_$w python tools/one_step.py $w 1 > gpurun_out/ncu_$w.log 2>&1
  python tools/ncu_summary.py full /tmp/r02_$w.ncu-rep > gpurun_out/r02_${w}_kernels.md
  ncu -i /tmp/r02_$w.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum > gpurun_out/r02_${w}_traffic.csv
done
du -sh gpurun_out
