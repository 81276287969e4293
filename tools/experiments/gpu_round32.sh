#!/bin/bash
run() { echo "== $*"; env "$@" timeout 200 python tools/time_unet.py 32 50 2>&1 | tail -1; }
run A=0
run KDIP_CONV_HALO=1
run KDIP_CONV_HALO=1 KDIP_HALO_PAIR=1
run KDIP_CONV_WS=0
run KDIP_CONV_PAIRMT=1
run KDIP_CONV_PAIR=0
run KDIP_CONV_MT=1
run A=0
