# compute-sanitizer passes over the kernels added in the last session of round 2 (outputs under gpurun_out/)
S="compute-sanitizer --tool memcheck --print-limit 5"
( $S python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke:|ERROR SUMMARY" ) | tee gpurun_out/san_smoke.txt
( timeout 500 $S python -m pytest tests/test_gpu_parity.py -x -q -k "range_and_sun_rows_match_oracle and slam_only_no_qr or imu_batch" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" ) | tee gpurun_out/san_sensors_imu.txt
( timeout 500 $S python -m pytest tests/test_gpu_gemm.py -x -q -k "2811 or 1301" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" ) | tee gpurun_out/san_tma.txt
( timeout 500 $S python -m pytest tests/test_gpu_parity.py -x -q -k "multi_uav_msckf" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" ) | tee gpurun_out/san_mm.txt
( timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke:|RACECHECK SUMMARY" ) | tee gpurun_out/san_race.txt
