for l in 1 2 4; do echo "FE_NODAL_LANES=$l"; FE_NODAL_LANES=$l python bench.py --steps 3 --warmup 3 --no-cpu --cg-iters 300 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pcg it/s', round(d['cg']['pcg_iters_per_s'],1), 'spmv(csr) ms', round(d['cg']['spmv_ms'],3), 'asm ms', round(d['roofline']['ms_per_launch'],3))"; done
