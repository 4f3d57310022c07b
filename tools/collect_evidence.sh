# round-2 evidence on ONE B200: bench (both arms), launch list, ncu --set full captures of the sweep kernel (both shapes) and the M x M job kernel
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02.json 2>> gpurun_out/bench_r02.err
python bench.py --steps 5 --warmup 3 --no-synthetic > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --no-synthetic > gpurun_out/ncu_launch.log 2>&1
python tools/profile_sweep.py 10000 512 3 > gpurun_out/plain_k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 2 -c 1 -o gpurun_out/prof_r02_kin40k python tools/profile_sweep.py 10000 512 3 > gpurun_out/ncu_k.log 2>&1
python tools/profile_sweep.py 400000 1024 2 > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 1 -c 1 -o gpurun_out/prof_r02_syn400k python tools/profile_sweep.py 400000 1024 2 > gpurun_out/ncu_s.log 2>&1
python tools/profile_dense.py 512 > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_job_kernel -s 6 -c 2 -o gpurun_out/prof_r02_dense512 python tools/profile_dense.py 512 > gpurun_out/ncu_d.log 2>&1
python tools/profile_dense.py 512 1024 > gpurun_out/r02_profile_dense.log 2>&1
SGP_DENSE_CLOCKS=1 python tools/dense_clocks.py 512 1024 2>&1 | tail -12 > gpurun_out/r02_dense_clocks.log
tail -n 2 gpurun_out/ncu_s.log gpurun_out/ncu_k.log gpurun_out/ncu_d.log; cut -c1-400 gpurun_out/bench_r02.json; cut -c1-300 gpurun_out/bench_ref_r02.json
