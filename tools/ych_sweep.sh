for y in 1024 256 128 96 64 32; do FT8_Y_CYCLES=$y python bench.py --device-only --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$y', round(j['value']), [(s['kernel'], s['ms']) for s in j['roofline_stages'][:3]], j['gpu_launches'])"; done
