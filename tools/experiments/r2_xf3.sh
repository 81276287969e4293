#!/bin/bash
KDIP_FUSE_GNAPPLY=0 timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
KDIP_FUSE_GNAPPLY=0 timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
timeout 300 python tools/time_unet.py 32 20 2>&1 | tail -n 1
