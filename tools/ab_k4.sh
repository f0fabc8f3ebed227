# K4 (interactive frames, < 8 samples per pixel) across library builds: bash tools/ab_k4.sh "cur k4c6 k4c8"
for lib in ${1:-cur}; do
  if [ $lib == cur ]; then L=cpuvoxelraycaster_b200/libvrt.so; else L=tools/libvrt_$lib.so; fi
  echo "== $lib"
  VRT_LIBRARY=$L python tools/measure_configs.py --configs 1 --iters 20 | cut -c1-200
  VRT_LIBRARY=$L PROBE_CASES="1,0" PROBE_SPP=4 python tools/probe_slice.py
  VRT_LIBRARY=$L PROBE_CASES="1,0" PROBE_SPP=1 python tools/probe_slice.py
  VRT_LIBRARY=$L python tools/flythrough.py --ticks 240 2>/dev/null | cut -c1-260
done
