#!/bin/bash
TAG=$1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
