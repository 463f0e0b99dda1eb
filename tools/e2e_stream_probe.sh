python - <<'P'
import sys; sys.path[:0]=['corpus','oracle']
import pycorpus
s=pycorpus.make(ch=2, bps=24, sr=96000, seconds=60, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, tile=60, seed=2026)
open('/dev/shm/cfg2.flac','wb').write(s.flac)
P
BNFLAC_TRACE=1 birdnest/audio_b200/flacdecoder_demo --bench /dev/shm/cfg2.flac 0 3 2>&1 | grep "streamed Read\|^rep" | tail -8
