cd tests/native/variants
run() { echo "== $*"; timeout 90 "$@" | grep -v "mismatch\|TEST" | sed 's/max|tc-mma.*bad=[0-9]*//'; }
for d in 0 1 2 3 4 5 7; do run ./attn_Y$d.bin 11; run ./attn_Y$d.bin 13; done
ATTN_HALF_GRID=1 run ./attn_Y0.bin 11
ATTN_HALF_GRID=1 run ./attn_Y7.bin 11
