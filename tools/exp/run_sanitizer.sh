# compute-sanitizer over the kernels that are new in round 2 (small cases; slow under the tool)
T=${1:-r02}
{
echo "== memcheck: polyphase, batched chain, ring path, channelizer (parameter descriptors)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_polyphase.py tests/test_gpu_steady_state.py -m gpu -x -q -k "polyphase_parity or exec_batch_orders or ring_to_chain or channelizer_at_steady" 2>&1 | tail -6
echo "== racecheck: polyphase (pair barriers, shared rows), batched chain"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_polyphase.py tests/test_gpu_steady_state.py -m gpu -x -q -k "test_polyphase_pluto_lsb or exec_batch_orders" 2>&1 | tail -6
} > gpurun_out/${T}_sanitizer.txt 2>&1
tail -20 gpurun_out/${T}_sanitizer.txt
