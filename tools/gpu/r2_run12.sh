cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export FNSSL_TC_PAIR=0
F="full_in16_H128x2 full_in256_H128x2_add narrow_in256_H128x2"
echo "== tc4, bounded waits (FNSSL_TC_WAIT_TIMEOUT=1)"; FNSSL_TC_WAIT_TIMEOUT=1 python tools/lstm_time.py $F | grep -v "_B1\|_B2"
echo "== tc4, unbounded waits"; FNSSL_TC_WAIT_TIMEOUT=0 python tools/lstm_time.py $F | grep -v "_B1\|_B2"
echo "== tc4, no gate math (debug 1), unbounded"; FNSSL_TC_WAIT_TIMEOUT=0 FNSSL_TC_DEBUG=1 python tools/lstm_time.py $F | grep -v "_B1\|_B2"
echo "== tc4, ex2/rcp gates (debug 32), unbounded"; FNSSL_TC_WAIT_TIMEOUT=0 FNSSL_TC_DEBUG=32 python tools/lstm_time.py $F | grep -v "_B1\|_B2"
echo "== tc5 forced, unbounded"; FNSSL_TC_PAIR=1 FNSSL_TC_PAIR_MIN=1 FNSSL_TC_WAIT_TIMEOUT=0 python tools/lstm_time.py $F | grep -v "_B1\|_B2"
