mkdir -p gpurun_out
b() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 $EXTRA 2>gpurun_out/r2n_err_$n.log | tee -a gpurun_out/r2n_bench_$n.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$n', d['metric'], round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['clocks']['sm_mhz'], d.get('step_graphs'), d['tape_nodes_by_class']['bn']['ms_per_step'], d['tape_nodes_by_class']['conv']['ms_per_step'], d['final_loss'])
"
}
b sweep0 ZENU_B200_WGRAD_SWEEP=0
b sweep1 ZENU_B200_WGRAD_SWEEP=1
b sweep0 ZENU_B200_WGRAD_SWEEP=0
b sweep1 ZENU_B200_WGRAD_SWEEP=1
ZENU_B200_WGRAD_SWEEP=0 timeout 200 python tools/profile_step.py --out gpurun_out/r2n_step_profile_sweep0.tsv > /dev/null 2>&1
ZENU_B200_WGRAD_SWEEP=1 timeout 200 python tools/profile_step.py --out gpurun_out/r2n_step_profile_sweep1.tsv > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv_vs_oracle or conv_golden or conv_layers" > gpurun_out/r2n_parity_tests.log 2>&1; tail -n 3 gpurun_out/r2n_parity_tests.log
