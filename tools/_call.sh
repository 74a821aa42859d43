mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
(timeout 500 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_thermal.py -x -q -k "test_thermal_extreme_schedules or test_external_setters or test_reference_test_interactions" > gpurun_out/san_race_thermal.log 2>&1; echo rc=$? >> gpurun_out/san_race_thermal.log)
tail -4 gpurun_out/san_race_thermal.log
(timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -x -q -k "test_fixed_steps_line2d or test_minimise_and_event_driven_match_oracle or test_tiny_lines" > gpurun_out/san_race_parity.log 2>&1; echo rc=$? >> gpurun_out/san_race_parity.log)
tail -4 gpurun_out/san_race_parity.log
grep -c "Error: Race" gpurun_out/san_race_parity.log
