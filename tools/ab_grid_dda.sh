for lib in h3 cur h3 cur; do
  if [ $lib == cur ]; then L=cpuvoxelraycaster_b200/libvrt.so; else L=tools/libvrt_$lib.so; fi
  echo "== $lib"; VRT_LIBRARY=$L python tools/measure_configs.py --configs 2,3 --iters 5 | cut -c1-260
done
