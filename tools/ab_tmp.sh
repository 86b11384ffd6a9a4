mkdir -p gpurun_out
b() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 $EXTRA 2>gpurun_out/r2l_err_$n.log | tee -a gpurun_out/r2l_bench_$n.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$n', d['metric'], round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['clocks']['sm_mhz'], d.get('step_graphs'), d['tape_nodes_by_class']['bn']['ms_per_step'], d['tape_nodes_by_class']['conv']['ms_per_step'], d['final_loss'])
"
}
b pdl0 ZENU_B200_BN_PDL=0
b pdl1 ZENU_B200_BN_PDL=1
b pdl0 ZENU_B200_BN_PDL=0
b pdl1 ZENU_B200_BN_PDL=1
EXTRA="--arch resnet18"
b r18_pdl0 ZENU_B200_BN_PDL=0
b r18_pdl1 ZENU_B200_BN_PDL=1
EXTRA="--arch small_cnn"
b s_pdl0 ZENU_B200_BN_PDL=0
b s_pdl1 ZENU_B200_BN_PDL=1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bn" > gpurun_out/r2l_bn_tests.log 2>&1; tail -n 3 gpurun_out/r2l_bn_tests.log
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "graph" > gpurun_out/r2l_model_tests.log 2>&1; tail -n 3 gpurun_out/r2l_model_tests.log
