for w in 12 8 6 4; do
echo "== tma ring, $w warps"; SF_DEC_WARPS=$w SF_DECODE_TMA=1 timeout 60 python tools/decode_bench.py | grep -E "seen=(33|47|63)"
done
echo "== direct"; timeout 60 python tools/decode_bench.py | grep -E "seen=(33|47|63)"
