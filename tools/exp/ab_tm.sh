for v in default w10c2; do
  if [ $v != default ]; then export HZSDR_LIB=/root/repo/go-sdr_b200/variants/$v/libhzsdrcuda.so; else unset HZSDR_LIB; fi
  for k in 1 64; do K=$k timeout 200 python tools/exp/exp_c2_quick.py 2>&1 | tail -1; done
  timeout 300 python bench.py --workload c5 --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5', d['value'])"
done
HZSDR_LIB=/root/repo/go-sdr_b200/variants/w10c2/libhzsdrcuda.so timeout 600 python -m pytest tests/test_gpu_steady_state.py tests/test_gpu_parity.py -m gpu -x -q -k "chain or channelizer or dependent or convolve" 2>&1 | tail -3
