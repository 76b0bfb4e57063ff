for k in tile pencil; do for c in 8 16 32 64; do
echo -n "$k chunk=$c: "; DIND_GRID_KERNEL=$k DIND_CHUNK=$c python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['kernel'], d['roofline']['frac'])"
done; done
