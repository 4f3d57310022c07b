set -x
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r01g.json 2>>gpurun_out/bench_r01g.err
python bench.py --steps 20 --warmup 3 --no-synthetic > gpurun_out/plain_g.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 20 --warmup 3 --no-synthetic > gpurun_out/ncu_launch_g.log 2>&1
python tools/profile_sweep.py 10000 512 3 > gpurun_out/plain_gk.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 2 -c 1 -o gpurun_out/prof_r01g_kin40k python tools/profile_sweep.py 10000 512 3 > gpurun_out/ncu_gk.log 2>&1
python tools/profile_sweep.py 400000 1024 2 > gpurun_out/plain_gs.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 1 -c 1 -o gpurun_out/prof_r01g python tools/profile_sweep.py 400000 1024 2 > gpurun_out/ncu_gs.log 2>&1
tail -n 2 gpurun_out/ncu_gs.log; tail -n 2 gpurun_out/ncu_gk.log; cut -c1-300 gpurun_out/bench_r01g.json
