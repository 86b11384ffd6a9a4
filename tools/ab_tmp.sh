set -x
mkdir -p gpurun_out
for m in 1 2; do
  ZENU_B200_BN_ORDER=$m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bn" > gpurun_out/r2q_bn_tests_m$m.log 2>&1; tail -2 gpurun_out/r2q_bn_tests_m$m.log
done
for m in 0 1 2 0 2; do
  ZENU_B200_BN_ORDER=$m timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/r2q_err_$m.log | tee -a gpurun_out/r2q_bench_order_$m.json | cut -c1-200
done
ZENU_B200_BN_ORDER=0 timeout 300 python tools/profile_step.py --out gpurun_out/r2q_step_profile_m0.tsv > /dev/null 2>&1
ZENU_B200_BN_ORDER=2 timeout 300 python tools/profile_step.py --out gpurun_out/r2q_step_profile_m2.tsv > /dev/null 2>&1
