# A/B of library builds on one box: bash tools/ab_libs.sh "prev c8 cur ..."   (cur = the in-tree build, X = tools/libvrt_X.so)
mkdir -p gpurun_out
for lib in ${1:-prev cur prev cur}; do
  if [ $lib == cur ]; then L=cpuvoxelraycaster_b200/libvrt.so; else L=tools/libvrt_$lib.so; fi
  VRT_LIBRARY=$L python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', d['ms_per_step'], d['e2e']['ms_per_step'], d['frame_identity']['sha256'])
"
done
