for lib in g1 cur g10 g12 g1 cur; do
  if [ $lib == cur ]; then L=cpuvoxelraycaster_b200/libvrt.so; else L=tools/libvrt_$lib.so; fi
  echo "== $lib"; VRT_LIBRARY=$L python tools/measure_configs.py --configs 3 --iters 5 | cut -c1-420
done
for lib in cur k4c10; do
  if [ $lib == cur ]; then L=cpuvoxelraycaster_b200/libvrt.so; else L=tools/libvrt_$lib.so; fi
  echo "== $lib"; VRT_LIBRARY=$L python tools/measure_configs.py --configs 1 --iters 20 | cut -c1-200
  VRT_LIBRARY=$L PROBE_CASES="1,0" PROBE_SPP=4 python tools/probe_slice.py
done
