# round-end ncu captures (one GPU): tag = $1.  Numbers printed under a profiler are never bench values.
T=${1:-r02}
cap() {  # name, kernel regex (demangled), skip, prof_run args...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --kernel-name-base demangled --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -f \
      -o gpurun_out/${T}_$name python tools/prof_run.py "$@" > gpurun_out/${T}_ncu_$name.log 2>&1
  tail -1 gpurun_out/${T}_ncu_$name.log
}
cap c2 'k_chain1024<\(int\)4, \(bool\)1' 2 c2b 16
cap c2single 'k_chain1024<\(int\)4, \(bool\)0' 6 c2 12
cap c5 'k_chain1024<\(int\)3, \(bool\)1' 1 c5 12
cap c3 'k_chain16k<\(int\)3, \(bool\)0, \(bool\)0' 6 c3 12
cap c3os 'k_chain16k<\(int\)3, \(bool\)0, \(bool\)1' 6 c3os 12
cap c4 'k_beamform<' 6 c4 12
cap c4rs 'k_beamform_rs' 1 c4rs 12
cap poly 'k_polyphase_chain' 2 poly 16
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_b_under_ncu.log 2>&1
ls -la gpurun_out | grep ${T}_
