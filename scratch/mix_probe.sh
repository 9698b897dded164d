for d in 0 32 36 40 44 48 52; do RTB_FLOOR_BUCKET=$d TAG=floor$d python scratch/e2e_probe2.py 2>&1 | head -2; done
