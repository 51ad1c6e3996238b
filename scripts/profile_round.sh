# The round's profiling pass, run as ONE gpurun call after bench.py has exited 0 without ncu: source hash, full-set captures of the two hot
# kernels, launch list of a short bench.  Reports land in gpurun_out/ (scratch); scripts/ncu_counters.py, ncu_summary.py, ncu_by_line.py,
# launch_list.py turn them into profiles/*.  Usage: gpurun --timeout 1500 -- "bash scripts/profile_round.sh"
set -x
python -c "from cylinder_b200.build import source_sha; print(source_sha())" > gpurun_out/s5_sha.txt
python bench.py --steps 2 --warmup 3 --no-cpu --no-aux --no-sustained --sweeps-per-step 20 --hbm-sweeps 10 > gpurun_out/s5_bench_short.json 2> gpurun_out/s5_bench_short.err
ncu --set full --import-source on --clock-control none -k regex:sweep_smem -o gpurun_out/s5_smem_full -f python scripts/inst_count_smem.py 8192 200 > gpurun_out/s5_ncu_smem.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:sweep_hbm -o gpurun_out/s5_hbm_full -f python scripts/profile_target.py hbm > gpurun_out/s5_ncu_hbm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aux --no-sustained --sweeps-per-step 20 --hbm-sweeps 10 > gpurun_out/s5_ncu_bench.log 2>&1
ls -la gpurun_out/ | tail -8
