for m in 0 2 16 18 1 0 2 16; do echo "VAE mask=$m"; LTXV_NO_PDL=$m python tools/vae_time.py 20; done
for m in 0 2 4 8 1 0 14; do echo "DIT mask=$m"; LTXV_NO_PDL=$m python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-c1 --no-other-configs --no-encode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step_ms', d['ms_per_step'], 'stg', d['stg_preset_step']['ms_per_step'])"; done
