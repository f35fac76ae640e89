set +e
out=gpurun_out/sanitizer_r2.txt
echo "compute-sanitizer on B200, round 2 (register-resident path; python tools/profile_target.py 3 2 rodent unless noted)" > $out
for mode in 1 3 4 0 2; do
  for tool in memcheck racecheck synccheck initcheck; do
    extra=""; [ $tool = racecheck ] && extra="--racecheck-report all"
    r=$(STACB_MODE=$mode timeout 600 compute-sanitizer --tool $tool $extra python tools/profile_target.py 3 2 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)
    echo "  mode $mode $tool: $r" >> $out
  done
done
for tool in memcheck racecheck; do
  extra=""; [ $tool = racecheck ] && extra="--racecheck-report all"
  r=$(STACB_MODE=1 timeout 600 compute-sanitizer --tool $tool $extra python tools/profile_target.py 3 2 celegans 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)
  echo "  celegans mode 1 $tool: $r" >> $out
done
r=$(timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "epilogues or m_opt or look_ahead" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tail -2 | tr '\n' ' ')
echo "  memcheck over the epilogue / m-phase / session-staging GPU tests: $r" >> $out
cat $out
