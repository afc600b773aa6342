# sharded C4 only (fused + nccl), N ranks under torchrun: bash tools/exp/run_c4.sh N tag
N=$1; T=$2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c4 --buffers ${3:-8} > gpurun_out/${T}_c4_n$N.json 2> gpurun_out/${T}_c4_n$N.err; tail -c 300 gpurun_out/${T}_c4_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/${T}_c4_n$N.json').read().strip().splitlines()[-1]); print('value', d['value'], d.get('ms_per_step'), d.get('parity_rel_l2'), d.get('nvlink',{}).get('achieved_gbs_per_direction')); print({k:(v.get('value'),v.get('parity_rel_l2')) for k,v in d.get('extra',{}).items()} if isinstance(d.get('extra'),dict) else '')"
