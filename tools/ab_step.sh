# Same-box A/B of two library builds over the headline frame, the interactive frame, cfg 1 and cfg 5:
#   bash tools/ab_step.sh "h5 cur"      (cur = the in-tree build, X = tools/libvrt_X.so)
mkdir -p gpurun_out
bash tools/ab_libs.sh "${1:-h5 cur} ${1:-h5 cur}"
for lib in ${1:-h5 cur}; do
  if [ $lib == cur ]; then L=cpuvoxelraycaster_b200/libvrt.so; else L=tools/libvrt_$lib.so; fi
  echo "== $lib"
  VRT_LIBRARY=$L python tools/measure_configs.py --configs 1,5 --iters 10 | cut -c1-220
  VRT_LIBRARY=$L PROBE_CASES="8,0" python tools/probe_slice.py 2>&1 | tail -3
done
