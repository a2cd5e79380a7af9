# One call on the GPU box (gpurun -- bash tools/gpu_round_check.sh): the GPU test suite, the default bench line, the
# reference arm, the config table and a launch list of one pass of the bench workload.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log)
tail -6 gpurun_out/tests.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm rc=$?"
timeout 300 python tools/config_table.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_C3.csv python tools/ncu_driver.py 24 gasdark > gpurun_out/ncu_c3.log 2>&1; echo "launch list rc=$?"
tail -c 600 gpurun_out/bench_n1.json
