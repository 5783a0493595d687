python -m pytest tests -m gpu -x -q > gpurun_out/t7.log 2>&1; tail -3 gpurun_out/t7.log
python tools/profile_misc.py rollout_c2 > gpurun_out/roll7.json 2>&1; tail -1 gpurun_out/roll7.json
python tools/profile_misc.py obs_c2 2>&1 | tail -1 | cut -c1-600
HK_GEOS=0 python tools/time_census.py c2 4 2>&1 | tail -1 | cut -c1-900
