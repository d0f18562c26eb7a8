python -m pytest tests/test_device.py tests/test_engine.py -q -m gpu -x 2>&1 | tail -2
python tools/fuzz_faithful.py 30 43 2>&1 | tail -1
python tools/step_api_cost.py 2>&1 | grep -E "run"
