for pol in "4,2.5,128,0.3" "3,1.5,96,0.3" "3,2,128,0.2" "5,3,192,0.3" "6,4,256,0.4" "8,5,320,0.5" "3,1.2,64,0.15" "100,0,0,0.3" "100,0,0,0.5" "100,0,0,0.8"; do
  echo "POLICY $pol"; SKIDGPU_LIST_POLICY=$pol python tools/scale_probe.py --log2n 21 --ref-max 0 --repeat 2 2>&1 | tail -1 | grep -o "move=[0-9.]*\|groupsBefore=[0-9]*\|gpu_ms=[0-9.]*" | tr '\n' ' '; echo
done
