for lib in cur p10 p12 cur; do
  if [ $lib == cur ]; then L=cpuvoxelraycaster_b200/libvrt.so; else L=tools/libvrt_$lib.so; fi
  echo "== $lib"; VRT_LIBRARY=$L python tools/measure_configs.py --configs 5 --iters 5 | cut -c1-330
done
export PROBE_CASES="1,0;8,3"
for o in "beam_tile=8" "beam_tile=4" "beam_tile=16" "beam_tile=8"; do echo "== $o"; PROBE_OPTS=$o python tools/probe_slice.py; done
