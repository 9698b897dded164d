set -e
P="python tools/frame_probe.py --world 8 --col-block 32 --reps 7"
for hf in 64 128 256 512; do for hb in 4 6 10; do RTB_HEAVY_FRACTION_SMALL=$hf RTB_HEAVY_BUCKETS_SMALL=$hb $P --workloads p5_sah_4k --tag "hf=$hf hb=$hb" | sed 's/first-frame.*| steady/steady/'; done; done
RTB_OCT_TIER=0 $P --workloads p5_sah_4k --tag "oct=0" | sed 's/first-frame.*| steady/steady/'
for cap in 148 296 592 1184 2368; do RTB_WIDE_CAP=$cap RTB_WIDE_FRACTION=16 $P --workloads p5_rgrid_4k --tag "widecap=$cap" | sed 's/first-frame.*| steady/steady/'; done
for hf in 32 64 128 256; do RTB_HEAVY_FRACTION_SMALL=$hf $P --workloads p5_kd_4k --tag "hf=$hf" | sed 's/first-frame.*| steady/steady/'; done
