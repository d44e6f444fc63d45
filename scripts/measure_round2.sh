# round 2 measurement suite (run through gpurun on one B200); everything lands in gpurun_out/
set -x
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
python scripts/config1_replay.py > gpurun_out/r02_config1_replay.jsonl 2>> gpurun_out/r02_bench.err
# launch list (cold-cache, serialised: compare shares) of a short bench run
python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
# full capture of the config-2 kernel and of the config-4 (Riccati) kernel
python scripts/profile_target.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 -o gpurun_out/prof_r2_c2 -f python scripts/profile_target.py > gpurun_out/ncu2.log 2>&1
python scripts/profile_target.py 4096 30 > gpurun_out/plain30.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:solve_riccati -s 2 -c 1 -o gpurun_out/prof_r2_c4 -f python scripts/profile_target.py 4096 30 > gpurun_out/ncu3.log 2>&1
tail -c 400 gpurun_out/r02_bench.json; echo; tail -5 gpurun_out/r02_bench.err; tail -n 3 gpurun_out/ncu2.log; tail -n 3 gpurun_out/ncu3.log
