#!/bin/bash
# Call 38: after the comment-only rebuild of the library — smoke() and the feature / criterion GPU tests once more.
O=gpurun_out/r02c38; mkdir -p $O
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 170 python -m pytest tests/test_features_gpu.py tests/test_criterion_metrics_gpu.py tests/test_boundary.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/smoke.log; tail -3 $O/pytest.log
