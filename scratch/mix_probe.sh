# sweep of the "light tiles stay in raster order" threshold for frames stored straight into host memory:
# quarter-octaves above the median tile cost (k_cost_offsets); 0 = pure heaviest-first.  See profiles/r01_e2e_floor_bucket.log
for d in 0 3 6 9 12 15; do RTB_FLOOR_DELTA_HOST=$d TAG=floordelta$d python scratch/e2e_probe2.py 2>&1 | head -2; done
