mkdir -p gpurun_out
timeout 420 python tools/ab_probe.py --log2n 24 --kind gasdark --passes 3 \
  "" "SKIDGPU_TILEWALK_OCC=6" "SKIDGPU_TILEWALK_OCC=8" "SKIDGPU_SUPER_CAP=3072" "SKIDGPU_SUPER_CAP=4096" \
  "SKIDGPU_SUPER_CAP=4096,SKIDGPU_TILEWALK_OCC=6" "SKIDGPU_SUPER_CAP=6144" "SKIDGPU_TILE_SORT_EVERY=2" \
  "SKIDGPU_TILE_SORT_EVERY=8" "SKIDGPU_TILE_SORT_EVERY=1" "" "SKIDGPU_TILE_DIAG=1" \
  --host "" --host "SKID_PREALLOC_GB=12" > gpurun_out/ab_gasdark.jsonl 2> gpurun_out/ab_gasdark.err; echo "rc=$?"
timeout 300 python tools/ab_probe.py --log2n 24 --kind massive --passes 2 \
  "" "SKIDGPU_TILEWALK_OCC=6" "SKIDGPU_SUPER_CAP=4096" "SKIDGPU_SUPER_CAP=4096,SKIDGPU_TILEWALK_OCC=6" "" \
  > gpurun_out/ab_massive.jsonl 2> gpurun_out/ab_massive.err; echo "rc=$?"
cut -c1-330 gpurun_out/ab_gasdark.jsonl; cut -c1-330 gpurun_out/ab_massive.jsonl
