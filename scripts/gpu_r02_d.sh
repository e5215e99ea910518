#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -m pytest tests -m gpu -q --deselect "tests/test_imagenet_gpu.py::test_fooling_rate_within_half_a_point_of_the_reference[densenet121]" > $OUT/d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/d_summary.log
tail -25 $OUT/d_pytest.log | tee -a $OUT/d_summary.log
