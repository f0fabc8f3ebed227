set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_final_pytest.txt; cat gpurun_out/r02_final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 > gpurun_out/r02_final_smoke.txt; cat gpurun_out/r02_final_smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -c 400 gpurun_out/r02_final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err; tail -c 300 gpurun_out/r02_final_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_final_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"render_rounds|sort_samples|beam_floor" -c 3 -f -o gpurun_out/r02_k6_final4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_k6_final4.log 2>&1
tail -c 300 gpurun_out/r02_ncu_k6_final4.log
