#!/bin/bash
timeout 300 python -m pytest "tests/test_gpu_closure.py::test_closure_matches_reference_fixtures" -x -q -k "True-gmm8_gt" 2>&1 | grep -E "^E|assert|Error" | head -30
