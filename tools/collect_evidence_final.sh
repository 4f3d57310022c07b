# evidence at the end of round 2 on ONE B200 (after the phase-2 plan / host-step changes): smoke, bench (both arms), launch list, ncu --set full of the
# sweep kernel at both shapes.  Every ncu command runs only after the same command has exited 0 without ncu.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02z_smoke.log
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
python bench.py --impl reference > gpurun_out/r02z_bench_reference.json 2>> gpurun_out/r02z_bench.err
python bench.py --steps 5 --warmup 3 --no-synthetic --no-dense > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 5 --warmup 3 --no-synthetic --no-dense > gpurun_out/ncu_launch.log 2>&1
python tools/profile_sweep.py 10000 512 3 > gpurun_out/plain_k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02z_kin40k python tools/profile_sweep.py 10000 512 3 > gpurun_out/ncu_k.log 2>&1
python tools/profile_sweep.py 400000 1024 2 > gpurun_out/plain_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02z_syn400k python tools/profile_sweep.py 400000 1024 2 > gpurun_out/ncu_s.log 2>&1
tail -n 2 gpurun_out/ncu_s.log gpurun_out/ncu_k.log gpurun_out/r02z_smoke.log; cut -c1-300 gpurun_out/r02z_bench.json; cut -c1-300 gpurun_out/r02z_bench_reference.json
