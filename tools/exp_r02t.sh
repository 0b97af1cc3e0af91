#!/bin/bash
set -x
timeout 300 python tools/unet_fp32_profile.py 2>&1 | tail -70
