#!/bin/bash
# usage: tools/variants.sh  -- bench each library variant under _build/variants (stage timings)
for lib in _build/variants/*.so; do
    BNFLAC_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('$lib', 'total %.3f scan %.3f crc %.3f parse %.3f decode %.3f' % (s['total'], s['scan'], s['crc'], s['parse'], s['decode']))"
done
