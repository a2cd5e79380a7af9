mkdir -p gpurun_out
timeout 300 python tools/ab_probe.py --log2n 24 --kind gasdark --passes 3 \
  "" "SKIDGPU_TILE_OVERLAP=1" "" "SKIDGPU_TILE_OVERLAP=1" \
  > gpurun_out/ab_gasdark.jsonl 2> gpurun_out/ab_gasdark.err; echo "rc=$?"
timeout 300 python tools/ab_probe.py --log2n 24 --kind massive --passes 2 \
  "" "SKIDGPU_TILE_OVERLAP=1" "" "SKIDGPU_TILE_OVERLAP=1" \
  > gpurun_out/ab_massive.jsonl 2> gpurun_out/ab_massive.err; echo "rc=$?"
cut -c1-330 gpurun_out/ab_gasdark.jsonl; cut -c1-330 gpurun_out/ab_massive.jsonl
(SKIDGPU_TILE_OVERLAP=1 timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests_overlap.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_overlap.log); tail -4 gpurun_out/tests_overlap.log
(timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log); tail -4 gpurun_out/tests.log
