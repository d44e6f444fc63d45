set -x
python bench.py > gpurun_out/r01d_bench.json 2> gpurun_out/r01d_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01d_bench_reference.json 2>> gpurun_out/r01d_bench.err
python bench.py --batch 8192 --gaits trot,pronk,amble,pseudo_gallop --no-cpu-baseline > gpurun_out/r01d_bench_config3_mixed_1gpu.json 2>> gpurun_out/r01d_bench.err
python bench.py --batch 16384 --horizon 30 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_bench_config4_n30.json 2>> gpurun_out/r01d_bench.err
python scripts/rollout_bench.py > gpurun_out/r01d_rollout_config5.jsonl 2>> gpurun_out/r01d_bench.err
python scripts/profile_target.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1f.csv python scripts/profile_target.py > gpurun_out/ncu1.log 2>&1
python scripts/profile_target.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 -o gpurun_out/prof_r1f -f python scripts/profile_target.py > gpurun_out/ncu2.log 2>&1
tail -c 600 gpurun_out/r01d_bench.json; echo; cat gpurun_out/r01d_bench.err | tail -5
