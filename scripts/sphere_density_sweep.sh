timeout 400 python -m pytest tests -m gpu -x -q -k "exact_grid or full_sky or random or sphere" 2>&1 | tail -2
for D in 8 4 3 2 1.5; do echo "density $D"; XM_SPHERE_DENSITY=$D ITERS=3 timeout 400 python scripts/c5_probe.py 1e9 1e8 2>&1 | grep "^C5" | tail -1; done
