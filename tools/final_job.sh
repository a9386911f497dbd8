set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench32.json 2> gpurun_out/bench32.err
python bench.py --impl reference > gpurun_out/bench32_ref.json 2> gpurun_out/bench32_ref.err
python tools/timeline.py > gpurun_out/timeline32.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r1i.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_relu_pool_fwd|conv_gather_mma|conv_dgrad_rows|conv_wgrad_mma|sparse|maxpool2x2_fwd|conv_bgrad" --launch-skip 30 -c 8 -f -o gpurun_out/r1i_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1i.log 2>&1
ls -la gpurun_out/r1i_full.ncu-rep
