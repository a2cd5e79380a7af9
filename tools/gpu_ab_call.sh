mkdir -p gpurun_out
(timeout 60 python -m pytest tests/test_gpu_demo.py tests/test_gpu_synth_golden.py tests/test_gpu_move_kernels.py tests/test_gpu_properties.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2) 
(SKIDGPU_TILE_WINDOW=10 timeout 30 python -m pytest tests/test_gpu_demo.py tests/test_gpu_synth_golden.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2)
timeout 60 python tools/ab_probe.py --log2n 24 --kind gasdark --passes 2 "" "SKIDGPU_TILE_WINDOW=10" > gpurun_out/ab_w10_gasdark.jsonl 2>/dev/null; cut -c1-330 gpurun_out/ab_w10_gasdark.jsonl
timeout 50 python tools/ab_probe.py --log2n 24 --kind massive --passes 2 "" "SKIDGPU_TILE_WINDOW=10" > gpurun_out/ab_w10_massive.jsonl 2>/dev/null; cut -c1-330 gpurun_out/ab_w10_massive.jsonl
