# re-captures of the kernels that changed after run_profiles.sh (tag = $1)
T=${1:-r02}
cap() {
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --kernel-name-base demangled --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -f \
      -o gpurun_out/${T}_$name python tools/prof_run.py "$@" > gpurun_out/${T}_ncu_$name.log 2>&1
  tail -1 gpurun_out/${T}_ncu_$name.log
}
cap c1 'k_shift_batch' 2 c1 16
cap c3os 'k_chain16k<\(int\)3, \(bool\)0, \(bool\)1' 6 c3os 12
cap c4 'k_beamform<' 6 c4 12
cap c4rs 'k_beamform_rs' 1 c4rs 12
cap poly 'k_polyphase_chain' 2 poly 16
