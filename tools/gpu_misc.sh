#!/bin/bash
timeout 300 python tools/netvlad_time.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_aggregate.py -q -x -k "tensor_core_assignment or 17places" 2>&1 | tail -2
