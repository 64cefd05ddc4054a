# Proxy fence before the aux-slot release: (1) the experimental two-group variant with and without it, (2) the product library through
# the diagnostic, the GPU test suite and smoke, (3) a C4 bench line (the fence adds one FENCE.VIEW.ASYNC per aux chunk per epilogue warp).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DIAG_REPS=10 bash scripts/experimental/run_variants.sh
DIAG_REPS=10 timeout 100 python scripts/tc_diag.py 2>&1 | grep "bad frac" > gpurun_out/r02c_fix_product_diag.txt; echo "product: $(grep -vc 'bad frac 0.0 ' gpurun_out/r02c_fix_product_diag.txt) failing of $(wc -l < gpurun_out/r02c_fix_product_diag.txt)"
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r02c_fix_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02c_fix_pytest_gpu.log | cut -c1-200
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sweep > gpurun_out/r02c_fix_bench_c4.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02c_fix_bench_c4.json')); print('c4', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gemm class ms', d['kernel_breakdown_ms']['gemm'], 'fused_fwd', d['kernel_breakdown_ms']['fused_fwd'])"
