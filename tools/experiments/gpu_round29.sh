#!/bin/bash
for c in 8 4 6 12 16 32; do echo "== GN CTAs/SM $c"; KDIP_GN_CTAS=$c timeout 200 python tools/time_unet.py 32 30 2>&1 | tail -1; done
echo "== again 8"; timeout 200 python tools/time_unet.py 32 30 2>&1 | tail -1
