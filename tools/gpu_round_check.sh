mkdir -p gpurun_out
# species rules + non-periodic goldens of the reference (written at the end of round 1, never run on a GPU yet)
timeout 120 python tools/species_check.py > gpurun_out/species_check.log 2>&1; echo "species rc=$?"; tail -9 gpurun_out/species_check.log
(timeout 600 python -m pytest tests/test_gpu_stats.py tests/test_gpu_synth_golden.py tests/test_gpu_host_driver.py tests -m gpu -q --durations=20 -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log)
tail -8 gpurun_out/tests.log
timeout 300 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python tools/next_rows_bench.py > gpurun_out/next_rows.json 2> gpurun_out/next_rows.err; echo "next rc=$?"
python -c "import sys; sys.path.insert(0,'.'); from skid_b200 import synth; s=synth.make_box(1<<20,seed=7,kind='gasdark'); synth.write_std(s,'/tmp/in20.std'); print(' '.join(s['ref_args']))" > /tmp/args.txt && timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_hostskid_2e20_stats.csv ./host/skid $(cat /tmp/args.txt) -den -ray -stats -o /tmp/o20 < /tmp/in20.std > gpurun_out/hostskid_ncu.log 2>&1; echo "ncu rc=$?"
tail -c 400 gpurun_out/bench_r01_final.json; tail -c 1500 gpurun_out/next_rows.json
