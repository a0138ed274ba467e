#!/bin/bash
# round 2, call 57: smoke + block-cache / lazy-row tests on the last build
timeout -k 5 45 python -m pytest tests/test_block_cache_gpu.py tests/test_lazy_rows_gpu.py -m gpu -x -q 2>&1 | tail -2
