#!/bin/bash
# usage: tools/variants.sh  -- bench each library variant under _build/variants (decode-stage timing)
for lib in _build/variants/*.so; do
  for bal in 1 0; do
    BNFLAC_LIB=$PWD/$lib BNFLAC_DEC_BALANCE=$bal python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-verify 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('$lib balance=$bal', 'total %.3f scan %.3f crc %.3f parse %.3f decode %.3f' % (s['total'], s['scan'], s['crc'], s['parse'], s['decode']))"
  done
done
