#!/bin/bash
# Runs ON THE GPU BOX (gpurun): ncu --set full capture + launch list of the SURVEY 8(f) / proposal / detection kernels
# (tools/prof_next.py).  Kept apart from tools/collect_profiles.sh: the two together exceed gpurun's 64 MiB return limit.
set -u
R=${1:-r01}
O=gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"proposal_|nms_|rpn_pack|full_masks|paste_prepare|decode_|detection_" -c 28 \
    -o $O/${R}_next python tools/prof_next.py > $O/${R}_next.log 2>&1
tail -1 $O/${R}_next.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_next_launches.csv \
    python tools/prof_next.py > $O/${R}_next_launches.log 2>&1
ls -la $O | tail -8
