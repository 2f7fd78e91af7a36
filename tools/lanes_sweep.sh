for n in 2 3 4; do SIFT_B200_LANES=$n python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lanes',d['config']['frames_in_flight_per_gpu'],'value',round(d['value']),'e2e',round(d['e2e']['value']))
"; done
