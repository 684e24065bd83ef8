cd tests/native/variants
run() { echo "== $*"; timeout 90 "$@" | sed 's/max|tc-mma|=\([0-9.e-]*\) max|tc-fp32|=\([0-9.e-]*\)/d=\1 \2/'; }
run ./attn_P0.bin
CLIPB200_ATTN_ONE_TILE=1 run ./attn_P0.bin 11
for v in P3 PN P2; do for c in 11 13 12; do run ./attn_$v.bin $c | grep -v TEST; done; done
