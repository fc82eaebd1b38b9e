#!/bin/bash
# Call 39: the whole -m gpu suite on the final tree (library rebuilt after the comment-only edits).
O=gpurun_out/r02c39; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest.log
