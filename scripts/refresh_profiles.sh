# Turn the files a measure_round2.sh run left in gpurun_out/ into the committed summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
SHA=$(python -c "import bench; print(bench.kernel_source_sha())")
python scripts/ncu_summary.py gpurun_out/prof_r2_c2.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 2 -c 1 python scripts/profile_target.py (config 2: 4096 trot problems, N=10, LPT order, tensor-core sweep, 6 CTAs/SM); gpurun_out/prof_r2_c2.ncu-rep; kernel source sha $SHA" > profiles/r02_solve_kernel_ncu_full.json
python scripts/ncu_summary.py gpurun_out/prof_r2_c4.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:solve_riccati -s 2 -c 1 python scripts/profile_target.py 4096 30 (N=30 trot, 4096 problems: a quarter of config 4); gpurun_out/prof_r2_c4.ncu-rep; kernel source sha $SHA" > profiles/r02_riccati_kernel_n30_ncu_full.json
for r in c2 c4; do ncu -i gpurun_out/prof_r2_$r.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/${r}_src.csv; done
python scripts/ncu_lines.py /tmp/c2_src.csv 44 > profiles/r02_solve_kernel_hot_lines.txt
python scripts/ncu_lines.py /tmp/c4_src.csv 40 > profiles/r02_riccati_kernel_n30_hot_lines.txt
for f in r02_bench_reference.json r02_config1_replay.jsonl r02_launches.csv; do cp gpurun_out/$f profiles/$f; done
python - <<'PY'
import json, subprocess
j = json.load(open('profiles/r02_solve_kernel_ncu_full.json'))
mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6}
tot = sum(float(j[k].split()[0]) * mul[j[k].split()[1]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
d = json.load(open('profiles/traffic.json'))
d["solve_kernel_dram_bytes_per_launch"] = int(tot)
d["smem_pipe_pct_of_peak"] = float(j['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'].split()[0])
d["kernel_source_sha"] = subprocess.run(['python', '-c', 'import bench; print(bench.kernel_source_sha())'], capture_output=True, text=True).stdout.strip()
json.dump(d, open('profiles/traffic.json', 'w'), indent=1)
txt = [l for l in open('gpurun_out/r02_bench.json') if l.startswith('{')][-1]
open('profiles/r02_bench.json', 'w').write(txt)
b = json.loads(txt)
print({k: b[k] for k in ('value', 'ms_per_step')}, 'e2e', b['e2e']['value'], 'frac', b['roofline']['frac'], 'kernel_ms', b['roofline']['kernel_ms'], 'traffic', b['roofline']['traffic'])
for k, v in b.get('configs', {}).items(): print(k, {a: c for a, c in v.items() if a in ('value', 'ms_per_step', 'ms_per_tick')})
print('ncu:', j['gpu__time_duration.sum'], 'active/elapsed', float(j['smsp__cycles_active.avg'].split()[0]) / float(j['sm__cycles_elapsed.avg'].split()[0]), 'issue', j['smsp__issue_active.avg.pct_of_peak_sustained_active'], 'grid', j['launch__grid_size'])
PY
