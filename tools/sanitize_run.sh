for tool in memcheck racecheck; do
  for cap in 0 14; do
    if [ $cap = 0 ]; then unset FSD_TEST_CAP; else export FSD_TEST_CAP=$cap; fi
    echo "== $tool FSD_TEST_CAP=$cap"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_target.py 192 2>&1 | grep -E "outputs identical|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -8
  done
done
