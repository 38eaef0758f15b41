set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 600 gpurun_out/bench_r01_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_final_reference.json 2>/dev/null
timeout 300 python bench.py --steps 3 --warmup 3 --cg-iters 20 --no-cpu > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --cg-iters 20 --no-cpu > gpurun_out/ncu_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_assemble_tiled|k_spmv" -c 6 -o gpurun_out/full_final -f python bench.py --steps 1 --warmup 3 --cg-iters 2 --no-cpu > gpurun_out/ncu_full_final.log 2>&1
ls -la gpurun_out | tail -8
