python tools/variants.py bench --steps 200 --warmup 20 --no-cpu-baseline
