run() { echo "== $*"; env "$@" python scratch/perf4.py; }
RTB_OCT_RGRID=1 RTB_OCT_TIER=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run TAG=oct_rgrid SCENES=5rgrid,4rgrid RTB_OCT_RGRID=1
run TAG=oct_rgrid_nowide SCENES=5rgrid RTB_OCT_RGRID=1 RTB_WIDE_CAP=0
run TAG=oct_tier SCENES=5sah,5kd,5fgrid,4sah RTB_OCT_TIER=1
run TAG=oct_tier_f32 SCENES=5sah,5kd RTB_OCT_TIER=1 RTB_HEAVY_FRACTION_SMALL=32 RTB_HEAVY_BUCKETS_SMALL=12
run TAG=oct_tier_f16 SCENES=5sah,5kd RTB_OCT_TIER=1 RTB_HEAVY_FRACTION_SMALL=16 RTB_HEAVY_BUCKETS_SMALL=16
