mkdir -p gpurun_out
b() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 $EXTRA 2>gpurun_out/r2m_err_$n.log | tee -a gpurun_out/r2m_bench_$n.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$n', d['metric'], round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['clocks']['sm_mhz'], d.get('step_graphs'), d['tape_nodes_by_class']['bn']['ms_per_step'], d['tape_nodes_by_class']['conv']['ms_per_step'], d['final_loss'])
"
}
b base ZENU_B200_BN_SWEEP_ROWS=8
b rows4 ZENU_B200_BN_SWEEP_ROWS=4
b rows16 ZENU_B200_BN_SWEEP_ROWS=16
b rows32 ZENU_B200_BN_SWEEP_ROWS=32
b ctas4 ZENU_B200_BN_CTAS=4
b ctas6 ZENU_B200_BN_CTAS=6
b base ZENU_B200_BN_SWEEP_ROWS=8
