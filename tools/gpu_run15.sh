mkdir -p gpurun_out
for ch in 2 4; do
ESLAM_HOST_CHUNKS=$ch timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample 16 2>/dev/null | grep -o '"e2e": {[^}]*}' | cut -c1-60,200-400 | sed "s/^/chunks $ch: /"
done
