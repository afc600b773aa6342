# full bench under torchrun at N ranks: bash tools/exp/run_n.sh N tag
N=$1; T=$2
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; tail -c 200 gpurun_out/${T}_bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N', d['value'], 'e2e', d['e2e']['value'], 'ring', d['e2e'].get('ring',{}).get('value'), 'ceil', d['e2e'].get('frac_of_copy_ceiling'))
for k,v in d['sharded'].items(): print(' ', k, v.get('value'), v.get('parity_rel_l2'), v.get('ms_per_step'), v.get('nvlink',{}).get('achieved_gbs_per_direction'), v.get('error'))"
