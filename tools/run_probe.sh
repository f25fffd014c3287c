timeout 200 python tools/dev_skinny.py parity 2>&1 | grep -v "^ok\|^skip"
for T in 16 8; do
timeout 30 python tools/dev_time.py q4_k 128256 4096 $T
timeout 30 python tools/dev_time.py q8_0 28672 8192 $T
timeout 30 python tools/dev_time.py q6_k 128256 4096 $T
done
for T in 32 64; do
timeout 30 python tools/dev_time.py q4_k 128256 4096 $T
done
GGQ_SKINNY_PROBE=2 timeout 30 python tools/dev_time.py q4_k 128256 4096 16
GGQ_SKINNY_PROBE=1 timeout 30 python tools/dev_time.py q4_k 128256 4096 16
timeout 30 python tools/dev_time.py q4_k 14336 4096 16
timeout 30 python tools/dev_prof.py q4_k 128256 4096 16 2>&1 | tail -66
