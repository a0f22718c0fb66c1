#!/bin/bash
# Runs ON THE GPU BOX (gpurun): the bench without ncu first, then the ncu launch list of the same command and one
# `ncu --set full` capture per RoIAlign kernel.  Outputs land in gpurun_out/; tools/summarise_profiles.py turns
# them into the text summaries committed under profiles/.
set -u
R=${1:-r01}
O=gpurun_out
python bench.py --steps 20 --warmup 3 --no-extras > $O/${R}_bench_noextras.json 2> $O/${R}_bench_noextras.err || { echo "bench failed"; tail -5 $O/${R}_bench_noextras.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-extras > $O/${R}_launches.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"roialign_fwd_nhwc_pair" -c 2 \
    -o $O/${R}_fwd_pair_nhwc python tools/prof_one.py pair 14 nhwc > $O/${R}_fwd_pair_nhwc.log 2>&1
for spec in "fwd 7" "fwd 14" "bwd 7" "bwd 14"; do
    set -- $spec
    extra=""; [ "$1" = "bwd" ] && extra="gather"
    timeout 400 ncu --set full --import-source on --clock-control none -k regex:"roialign_|bwd_items|bwd_alloc" -c 4 \
        -o $O/${R}_$1$2_nhwc python tools/prof_one.py $1 $2 nhwc $extra > $O/${R}_$1$2_nhwc.log 2>&1
    tail -1 $O/${R}_$1$2_nhwc.log
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"roialign_bwd_nhwc|zero_levels" -c 2 \
    -o $O/${R}_bwd14_nhwc_scatter python tools/prof_one.py bwd 14 nhwc > $O/${R}_bwd14_scatter.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"crop_plane_fwd" -s 2 -c 1 \
    -o $O/${R}_mask_targets python tools/prof_mask_targets.py > $O/${R}_mask_targets.log 2>&1
# the 8(f) / proposal / detection kernels: tools/collect_next.sh (a separate gpurun call: 64 MiB return limit)
ls -la $O | tail -20
