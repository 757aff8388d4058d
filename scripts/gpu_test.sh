#!/bin/bash
OUT=gpurun_out/${1:-test}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest.log
