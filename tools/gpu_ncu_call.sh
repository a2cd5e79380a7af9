mkdir -p gpurun_out
timeout 140 ncu --set full --clock-control none --import-source on -k regex:"k_tile_walk|k_tile_step" --launch-skip 2 --launch-count 2 \
  -f -o gpurun_out/r01_v7_tilewalk_tilestep_2e24 python tools/ncu_driver.py 24 gasdark > gpurun_out/ncu_a.log 2>&1; echo "ncu A rc=$?"
timeout 110 ncu --set full --clock-control none --import-source on -k regex:"k_stat_groups|k_stat_keys" --launch-count 2 \
  -f -o gpurun_out/r01_v7_stats_2e24 python tools/ncu_driver.py 24 gasdark > gpurun_out/ncu_b.log 2>&1; echo "ncu B rc=$?"
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log; ls -la gpurun_out/*.ncu-rep
