# round-2 evidence after the dense / theta rework, ONE B200: GPU tests, bench (both arms), launch list, M x M job kernel ncu capture, dense and theta timings
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r02.json 2>> gpurun_out/bench_r02.err
python bench.py --steps 5 --warmup 3 --no-synthetic > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --no-synthetic > gpurun_out/ncu_launch.log 2>&1
python tools/profile_dense.py 512 > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_job_kernel -s 6 -c 2 -o gpurun_out/prof_r02_dense512 python tools/profile_dense.py 512 > gpurun_out/ncu_d.log 2>&1
python tools/profile_dense.py 512 1024 > gpurun_out/r02_profile_dense.log 2>&1
SGP_DENSE_CLOCKS=1 python tools/dense_clocks.py 512 1024 2>&1 | grep -A1 "dense job" | tail -12 > gpurun_out/r02_dense_clocks.log
for m in host resident; do python tools/theta_profile.py 512 500 $m; done > gpurun_out/r02_theta_profile.log 2>&1
SGP_THETA_TIMING=1 REPS=2 python tools/theta_profile.py 512 500 host 2>&1 | tail -3 >> gpurun_out/r02_theta_profile.log
cat gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/ncu_d.log; cut -c1-600 gpurun_out/bench_r02.json; cut -c1-300 gpurun_out/bench_ref_r02.json; cat gpurun_out/r02_theta_profile.log
