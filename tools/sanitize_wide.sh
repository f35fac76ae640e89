set +e
out=gpurun_out/sanitizer_r2_wide.txt
echo "compute-sanitizer on B200, multi-warp register-resident path (python tools/profile_target.py 2 1 mouse: 6 warps per chain; STACB_MODE=4: pair mode, two groups of 6 warps) and folded fruitfly (3 2 fly_treadmill)" > $out
for tool in memcheck racecheck synccheck initcheck; do
  extra=""; [ $tool = racecheck ] && extra="--racecheck-report all"
  r=$(timeout 1500 compute-sanitizer --tool $tool $extra python tools/profile_target.py 2 1 mouse 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)
  echo "  mouse (auto) $tool: $r" >> $out
done
r=$(STACB_MODE=2 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python tools/profile_target.py 2 1 mouse 2>&1 | grep -E "RACECHECK SUMMARY" | tail -1)
echo "  mouse (register cap for two CTAs per SM) racecheck: $r" >> $out
for tool in memcheck racecheck synccheck; do
  extra=""; [ $tool = racecheck ] && extra="--racecheck-report all"
  r=$(STACB_MODE=4 timeout 1500 compute-sanitizer --tool $tool $extra python tools/profile_target.py 2 1 mouse 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)
  echo "  mouse (pair mode: two groups of 6 warps, named barriers) $tool: $r" >> $out
done
[ -n "$SKIP_FLY" ] && { cat $out; exit 0; }
for mode in 1 4 0; do
  for tool in memcheck racecheck; do
    extra=""; [ $tool = racecheck ] && extra="--racecheck-report all"
    r=$(STACB_MODE=$mode timeout 900 compute-sanitizer --tool $tool $extra python tools/profile_target.py 3 2 fly_treadmill 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | tail -1)
    echo "  fly_treadmill mode $mode $tool: $r" >> $out
  done
done
cat $out
