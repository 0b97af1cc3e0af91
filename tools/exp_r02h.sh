#!/bin/bash
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python tools/timeline.py 2>&1 | tail -3
